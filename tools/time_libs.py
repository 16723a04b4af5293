#!/usr/bin/env python3
"""Times tagged library variants of one robot (built with build_robot_library(..., tag=...)).
  python tools/time_libs.py iiwa14 fd_grad 65536 "" _loop_mb8 _loop_mb12 ...
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gridcodegenerator_b200 import load_named_robot                      # noqa: E402
from gridcodegenerator_b200.build import lib_path                        # noqa: E402
from gridcodegenerator_b200.runtime import GridEngine                    # noqa: E402
from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u    # noqa: E402

name, alg, N = sys.argv[1], sys.argv[2], int(sys.argv[3])
robot = load_named_robot(name)
n = robot.n
q, qd, u, _ = make_states(n, N, 3)
x = torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda()
outw = {"id": n, "minv": n * n, "fd": n, "id_grad": 2 * n * n, "fd_grad": 2 * n * n}[alg]
out = torch.empty(N, outw, device="cuda")
os.environ["GRID_FORCE_KERNEL"] = "tps"
ref = None
for tag in sys.argv[4:]:
    so = lib_path(robot, tag)
    if not os.path.exists(so):
        print(json.dumps({"tag": tag, "missing": so}))
        continue
    eng = GridEngine(robot, lib_path=so)
    us = eng.time_launches(alg, out, x, reps=50)
    res = out.cpu().numpy().copy()
    if ref is None:
        ref = res
    print(json.dumps({"tag": tag, "p50_us": float(np.median(us)), "min_us": float(us.min()),
                      "evals_per_s": N / float(np.median(us)) * 1e6,
                      "rel_diff_vs_first": float(np.abs(res - ref).max() / np.abs(ref).max())}), flush=True)
