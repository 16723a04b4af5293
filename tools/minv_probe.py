#!/usr/bin/env python3
"""Atlas Minv (phase-split, single-stage): per-program times (GRID_PIPE_ONLY_TASK) - what the 200 us are made of."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                                                       # noqa: E402
import torch                                                             # noqa: E402
from gridcodegenerator_b200 import load_named_robot                     # noqa: E402
from gridcodegenerator_b200.runtime import get_engine                    # noqa: E402
from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u    # noqa: E402

robot = load_named_robot(sys.argv[1] if len(sys.argv) > 1 else "atlas")
alg = sys.argv[2] if len(sys.argv) > 2 else "minv"
eng = get_engine(robot)
n, N = robot.n, 65536
q, qd, u, _ = make_states(n, N, 3)
x = torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda()
out = torch.empty(N, 2 * n * n, device="cuda")
eng.set_option("GRID_FORCE_KERNEL", "pipe")
res = {"robot": robot.name, "alg": alg, "N": N, "us_all": float(np.median(eng.time_launches(alg, out, x, num_timesteps=N, stride=3 * n, reps=20)))}
for k in range(8):
    eng.set_option("GRID_PIPE_ONLY_TASK", str(k))
    us = float(np.median(eng.time_launches(alg, out, x, num_timesteps=N, stride=3 * n, reps=10)))
    if k > 0 and us < 6.5:
        break
    res["us_s0_t%d" % k] = us
eng.set_option("GRID_PIPE_ONLY_TASK", None)
print(json.dumps(res), flush=True)
