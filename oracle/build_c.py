"""Builds oracle/rbd_oracle.c into oracle/_build/liboracle.so (gcc; test infrastructure only)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def build_c_oracle(force: bool = False) -> str:
    src = os.path.join(HERE, "rbd_oracle.c")
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "liboracle.so")
    if not force and os.path.exists(so) and os.path.getmtime(so) >= os.path.getmtime(src):
        return so
    subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-pthread", "-o", so + ".tmp", src, "-lm"], check=True)
    os.replace(so + ".tmp", so)
    return so
