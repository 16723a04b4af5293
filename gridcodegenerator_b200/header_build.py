"""Builds the emitted-header test harness: gen_all_code() -> grid.cuh -> nvcc tests/header_harness.cu.
The binary lives in _lib/ (git-ignored, travels to the GPU box)."""
import hashlib
import os
import subprocess

from .build import GEN_DIR, LIB_DIR, ROOT, find_nvcc, _static_hash
from .facade import GRiDCodeGenerator
from .urdf import load_named_robot

HARNESS_SRC = os.path.join(ROOT, "tests", "header_harness.cu")


def build_header_harness(robot_name: str = "iiwa14", force: bool = False, device_fns: bool = None,
                         timeout_s: int = 900) -> str:
    robot = load_named_robot(robot_name)
    from .codegen import KernelPlan
    kinds = KernelPlan(robot).kind
    wide_fns = False
    if device_fns is None:          # single-thread _inner/_device bodies, or the wide bodies behind the same names
        device_fns = all("tps" in k for k in kinds.values())
        wide_fns = not device_fns and all("tps" in k or "wps" in k for k in kinds.values())
    h = hashlib.sha256((_static_hash() + robot.param_hash()).encode())
    for fn in (HARNESS_SRC, os.path.join(os.path.dirname(__file__), "facade.py")):
        with open(fn, "rb") as f:
            h.update(f.read())
    exe = os.path.join(LIB_DIR, "header_harness_%s_%s" % (robot_name, h.hexdigest()[:10]))
    if os.path.exists(exe) and not force:
        return exe
    hdr_dir = os.path.join(GEN_DIR, "header_" + robot_name)
    os.makedirs(hdr_dir, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    cwd = os.getcwd()
    try:
        os.chdir(hdr_dir)                       # gen_all_code writes into the CWD, like the reference
        GRiDCodeGenerator(robot).gen_all_code()
    finally:
        os.chdir(cwd)
    cmd = [find_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-w",
           "-I", hdr_dir, "-o", exe + ".tmp", HARNESS_SRC] + (["-DHARNESS_DEVICE_FNS"] if device_fns else []) + (
               ["-DHARNESS_WIDE_DEVICE_FNS"] if wide_fns else [])
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed on the emitted header:\n" + proc.stderr[-6000:])
    os.replace(exe + ".tmp", exe)
    for fn in os.listdir(LIB_DIR):
        if fn.startswith("header_harness_%s_" % robot_name) and os.path.join(LIB_DIR, fn) != exe:
            os.remove(os.path.join(LIB_DIR, fn))
    return exe
