"""nvcc driver: robot -> generated .cu -> in-tree libgrid_<robot>_<hash>.so.

Artefacts live under ``gridcodegenerator_b200/_generated/`` (sources) and
``gridcodegenerator_b200/_lib/`` (shared objects); both are git-ignored but travel to
the GPU box with the repo snapshot.  The cache key is robot hash + codegen version +
hash of the static csrc/ and include/ files, so a stale library is never loaded.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import subprocess
import time
from typing import Dict, Optional, Tuple

from .codegen import CODEGEN_VERSION, KernelPlan, generate_translation_unit, plan_signature
from .robot import Robot

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
GEN_DIR = os.path.join(PKG, "_generated")
LIB_DIR = os.path.join(PKG, "_lib")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-gnu-unique", "-shared"]
# (-split-compile=0 halves the build time of the Atlas library but costs ~4 % more instructions in the
# straight-line kernels - 8 896 vs 8 544 for the iiwa14 FD gradient - so it is not used.)


def _static_hash() -> str:
    h = hashlib.sha256()
    h.update(CODEGEN_VERSION.encode())
    h.update(" ".join(NVCC_FLAGS).encode())
    for d in (CSRC, INCLUDE):
        for fn in sorted(os.listdir(d)):
            if fn.endswith((".cuh", ".h", ".cu")):
                with open(os.path.join(d, fn), "rb") as f:
                    h.update(fn.encode())
                    h.update(f.read())
    for fn in ("codegen.py", "algorithms.py", "ir.py", "pipeline.py"):
        with open(os.path.join(PKG, fn), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:10]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the B200 kernels cannot be built (there is no CPU fallback)")


def lib_path(robot: Robot, tag: str = "", plan: Optional[KernelPlan] = None) -> str:
    """<robot>_<robot hash>_<static hash>[_p<plan hash>][tag]: a non-default KernelPlan is part of the key."""
    sig = plan_signature(plan)
    return os.path.join(LIB_DIR, "libgrid_%s_%s_%s%s%s.so" % (robot.name, robot.param_hash(), _static_hash(),
                                                             "_p" + sig if sig else "", tag))


def build_robot_library(robot: Robot, plan: Optional[KernelPlan] = None, force: bool = False, tag: str = "",
                        verbose: bool = False, extra_flags=()) -> Tuple[str, Dict[str, dict]]:
    """Returns (path to .so, build info).  Rebuilds only when the cache key changed."""
    os.makedirs(GEN_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    so = lib_path(robot, tag, plan)
    info: Dict[str, dict] = {}
    if os.path.exists(so) and not force:
        return so, info
    t0 = time.time()
    sig = plan_signature(plan)
    src, stats = generate_translation_unit(robot, plan, ns_tag=tag + ("_p" + sig if sig else ""))
    cu = os.path.join(GEN_DIR, "grid_%s_%s%s%s.cu" % (robot.name, robot.param_hash(), "_p" + sig if sig else "", tag))
    # concurrent builders (one rank per GPU) must not share a temporary: per-process names, atomic rename
    cu_tmp = "%s.%d.tmp.cu" % (cu[:-3], os.getpid())
    so_tmp = "%s.%d.tmp" % (so, os.getpid())
    with open(cu_tmp, "w") as f:
        f.write(src)
    os.replace(cu_tmp, cu)
    t1 = time.time()
    cmd = [find_nvcc()] + NVCC_FLAGS + ["-Xptxas", "-v", "-I", CSRC, "-I", INCLUDE, "-o", so_tmp, cu] + list(extra_flags)
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        if os.path.exists(so_tmp):
            os.remove(so_tmp)
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (cu, proc.stdout[-4000:], proc.stderr[-8000:]))
    os.replace(so_tmp, so)
    # drop stale libraries of the same robot
    prefix = "libgrid_%s_" % robot.name
    keep = "_%s_%s" % (robot.param_hash(), _static_hash())
    for fn in os.listdir(LIB_DIR):
        if fn.startswith(prefix) and not tag and keep not in fn:
            try:
                os.remove(os.path.join(LIB_DIR, fn))
            except OSError:
                pass
    info = {"stats": stats, "codegen_s": t1 - t0, "nvcc_s": time.time() - t1, "ptxas": proc.stderr}
    with open(so[:-3] + ".ptxas.txt", "w") as f:
        f.write(proc.stderr)
    with open(so[:-3] + ".stats.json", "w") as f:
        json.dump(stats, f)
    if verbose:
        print(proc.stderr)
    return so, info
