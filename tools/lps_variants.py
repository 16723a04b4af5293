#!/usr/bin/env python3
"""Chain-kernel tuning: column-kernel software pipelining / register cap variants of the chain-64 library.
  python tools/lps_variants.py build      (CPU)        python tools/lps_variants.py run   (GPU) -> JSON lines
"""
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gridcodegenerator_b200 import load_named_robot                     # noqa: E402
from gridcodegenerator_b200.build import build_robot_library            # noqa: E402
from gridcodegenerator_b200.codegen import KernelPlan                   # noqa: E402

VARIANTS = {"p0_b3": (0, 3), "bulk_s2": ("bulk", 2), "bulk_s3": ("bulk", 3), "bulk_s4": ("bulk", 4)}


def _build(name):
    robot = load_named_robot("chain64")
    pipe, minb = VARIANTS[name]
    plan = KernelPlan(robot, only_algs=("fd_grad", "id_grad"))
    flags = (["-DGRID_LPS_BULK=1", "-DGRID_LPS_BULK_STAGES=%d" % minb] if pipe == "bulk" else
             ["-DGRID_LPS_PIPELINE=%d" % pipe, "-DGRID_LPS_MINB=%d" % minb])
    so, info = build_robot_library(robot, plan, tag="_y" + name, extra_flags=flags)
    regs = [l.strip() for l in info.get("ptxas", "").splitlines() if "grad_columns_kernel" in l or "Used" in l]
    out = []
    for i, l in enumerate(regs):
        if "grad_columns_kernel" in l and i + 1 < len(regs):
            pass
    import re
    txt = info.get("ptxas", "")
    found = re.findall(r"grad_columns(?:_bulk)?_kernelILi8ELi(\d).*?\n.*?\n.*?Used (\d+) registers", txt)
    return name, found


def build():
    with ProcessPoolExecutor(6) as ex:
        for r in ex.map(_build, VARIANTS):
            print(r, flush=True)


def run():
    import numpy as np
    import torch
    from gridcodegenerator_b200.runtime import GridEngine
    from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u
    from oracle import c_oracle as C
    robot = load_named_robot("chain64")
    n, N = robot.n, 16384
    q, qd, u, _ = make_states(n, N, 3)
    x = torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda()
    out = torch.empty(N, 2 * n * n, device="cuda")
    ref = C.batch(robot, "fd_grad", q[:128], qd[:128], u[:128])
    for name in VARIANTS:
        try:
            eng = GridEngine(robot, plan=KernelPlan(robot, only_algs=("fd_grad", "id_grad")), tag="_y" + name)
            eng.forward_dynamics_gradient_device(out, x)
            torch.cuda.synchronize()
            err = float(np.abs(out[:128].cpu().numpy() - ref).max() / np.abs(ref).max())
            res = {"variant": name, "pipeline": VARIANTS[name][0], "min_blocks": VARIANTS[name][1], "relerr": err}
            for alg in ("fd_grad", "id_grad"):
                for T in (16384, 4096):
                    us = eng.time_launches(alg, out, x, num_timesteps=T, stride=3 * n, reps=5)
                    res["%s_us_N%d" % (alg, T)] = float(np.median(us))
            print(json.dumps(res), flush=True)
        except Exception as e:
            print(json.dumps({"variant": name, "error": str(e)[:300]}), flush=True)


if __name__ == "__main__":
    (build if sys.argv[1] == "build" else run)()
