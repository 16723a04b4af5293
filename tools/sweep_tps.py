#!/usr/bin/env python3
"""Builds tagged variants of a robot library (register cap / warps per CTA) and, on a GPU box,
times one algorithm for each.  Build step runs anywhere; timing needs CUDA.
  python tools/sweep_tps.py build            (here, CPU)
  python tools/sweep_tps.py run              (on the B200)
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gridcodegenerator_b200 import load_named_robot                     # noqa: E402
from gridcodegenerator_b200.build import build_robot_library, lib_path  # noqa: E402
from gridcodegenerator_b200.codegen import KernelPlan                   # noqa: E402

ROBOT = os.environ.get("SWEEP_ROBOT", "iiwa14")
ALG = os.environ.get("SWEEP_ALG", "fd_grad")
# (warps per CTA, CTAs per SM, __syncthreads every K traced ops)
VARIANTS = [(1, 16, 0), (1, 12, 0), (1, 8, 0),
            (4, 2, 0), (4, 2, 256), (4, 2, 1024),
            (8, 1, 0), (8, 1, 256), (8, 1, 1024),
            (14, 1, 0), (14, 1, 256), (14, 1, 1024),
            (16, 1, 0), (16, 1, 256)]
if os.environ.get("SWEEP_VARIANTS"):
    VARIANTS = [tuple(int(x) for x in v.split(":")) for v in os.environ["SWEEP_VARIANTS"].split(",")]


def tag(w, mb, sync):
    return "_w%d_mb%d_s%d" % (w, mb, sync)


def build():
    robot = load_named_robot(ROBOT)
    for w, mb, sync in VARIANTS:
        plan = KernelPlan(robot, tps_warps=w, tps_sync_every=sync,
                          tps_min_blocks={a: mb for a in ("id", "minv", "fd", "id_grad", "fd_grad")})
        t = time.time()
        so, info = build_robot_library(robot, plan, tag=tag(w, mb, sync))
        spills = [l for l in info.get("ptxas", "").splitlines() if "spill" in l][:1]
        print(tag(w, mb, sync), "%.1fs" % (time.time() - t), spills, flush=True)


def run():
    import numpy as np
    import torch
    from gridcodegenerator_b200.runtime import GridEngine
    from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u
    robot = load_named_robot(ROBOT)
    n, N = robot.n, int(os.environ.get("SWEEP_N", "65536"))
    q, qd, u, _ = make_states(n, N, 3)
    x = torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda()
    outw = {"id": n, "minv": n * n, "fd": n, "id_grad": 2 * n * n, "fd_grad": 2 * n * n}[ALG]
    out = torch.empty(N, outw, device="cuda")
    res = []
    for w, mb, sync in VARIANTS:
        so = lib_path(robot, tag(w, mb, sync))
        if not os.path.exists(so):
            print("missing", so, flush=True)
            continue
        eng = GridEngine(robot, lib_path=so)
        call = {"fd_grad": eng.forward_dynamics_gradient_device, "id_grad": eng.inverse_dynamics_gradient_device,
                "fd": eng.forward_dynamics_device, "minv": eng.direct_minv_device, "id": eng.inverse_dynamics_device}[ALG]
        for _ in range(20):
            call(out, x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100):
            call(out, x)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 10.0
        res.append({"warps": w, "ctas_per_sm": mb, "sync_every": sync, "us": us, "evals_per_s": N / us * 1e6})
        print(json.dumps(res[-1]), flush=True)
    return res


if __name__ == "__main__":
    build() if sys.argv[1:] == ["build"] else run()
