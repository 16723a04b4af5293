#!/usr/bin/env python3
"""Atlas Minv (phase-split, single-stage component programs) under different CTA shapes.
  python tools/minv_variants.py build|run"""
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gridcodegenerator_b200 import load_named_robot                     # noqa: E402
from gridcodegenerator_b200.build import build_robot_library            # noqa: E402
from gridcodegenerator_b200.codegen import KernelPlan                   # noqa: E402

VARIANTS = {
    "m_w8": dict(),
    "m_w4": dict(pipe_warps=4, pipe_min_blocks=(2, 2)),
    "m_w2": dict(pipe_warps=2, pipe_min_blocks=(4, 4)),
    "m_w1": dict(pipe_warps=1, pipe_min_blocks=(8, 8)),
    "m_w1_s0": dict(pipe_warps=1, pipe_min_blocks=(8, 8), pipe_sync_every=0),
}


def plan_for(robot, name):
    return KernelPlan(robot, only_algs=("minv",), **VARIANTS[name])


def _build(name):
    robot = load_named_robot("atlas")
    so, info = build_robot_library(robot, plan_for(robot, name), tag="_x" + name)
    return name, os.path.basename(so)


if sys.argv[1] == "build":
    with ProcessPoolExecutor(max_workers=5) as ex:
        for r in ex.map(_build, list(VARIANTS)):
            print(r, flush=True)
else:
    import numpy as np
    import torch
    from gridcodegenerator_b200.runtime import GridEngine
    from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u
    from oracle import c_oracle as C
    robot = load_named_robot("atlas")
    n, N = robot.n, 65536
    q, qd, u, _ = make_states(n, N, 3)
    x = torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda()
    out = torch.empty(N, n * n, device="cuda")
    ref = C.batch(robot, "minv", q[:256], qd[:256], None)
    for name in VARIANTS:
        eng = GridEngine(robot, plan=plan_for(robot, name), tag="_x" + name)
        eng.direct_minv_device(out, x)
        torch.cuda.synchronize()
        res = {"variant": name, "plan": VARIANTS[name], "kind": eng.kernel_kind("minv"),
               "relerr": float(np.abs(out[:256].cpu().numpy() - ref).max() / np.abs(ref).max())}
        for M in (65536, 8192, 128):
            res["us_N%d" % M] = float(np.median(eng.time_launches("minv", out, x, num_timesteps=M, stride=3 * n, reps=20)))
        print(json.dumps(res), flush=True)
