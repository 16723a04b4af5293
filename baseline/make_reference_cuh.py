#!/usr/bin/env python3
"""Runs the UNMODIFIED reference generator (/root/reference, read-only) on our robot stand-in and
writes its emitted header to baseline/_ref/<robot>/grid.cuh (git-ignored, travels to the GPU box).
The emitted CUDA is the reference's own GPU implementation of the hot path: compiled for sm_100a
it is the GPU baseline our kernels are timed against (tools/ref_gpu_bench.py).  Nothing from the
reference is copied into the repository.  Build container only (needs /root/reference + sympy).
  python baseline/make_reference_cuh.py [robot ...]
"""
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root")

from gridcodegenerator_b200 import load_named_robot   # noqa: E402


def main(names):
    from reference import GRiDCodeGenerator as RefGen
    for name in names:
        out = os.path.join(HERE, "_ref", name)
        os.makedirs(out, exist_ok=True)
        robot = load_named_robot(name)
        t = time.time()
        cwd = os.getcwd()
        os.chdir(out)
        try:
            RefGen(robot).gen_all_code()
        finally:
            os.chdir(cwd)
        print("reference gen_all_code(%s): %.1fs -> %s" % (name, time.time() - t, os.path.join(out, "grid.cuh")))
        for tag, extra in (("", []), ("_r168", ["-maxrregcount=168"])):
            exe = os.path.join(out, "ref_harness" + tag)
            t = time.time()
            cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-w", "-I", out,
                   "-o", exe, os.path.join(HERE, "ref_harness.cu")] + extra
            p = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
            print("  nvcc%s: %.1fs rc=%d %s" % (tag, time.time() - t, p.returncode, p.stderr[-500:] if p.returncode else ""))


if __name__ == "__main__":
    main(sys.argv[1:] or ["iiwa14"])
