#!/usr/bin/env python3
"""Per-task times of a phase-split library variant (GRID_PIPE_ONLY_TASK): which program of the split costs what.
  python tools/split_tasks.py <robot> <split|x2|default> [N]          (GPU) -> JSON lines"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                                                       # noqa: E402
import torch                                                             # noqa: E402
import __graft_entry__ as G                                              # noqa: E402
from gridcodegenerator_b200 import load_named_robot                     # noqa: E402
from gridcodegenerator_b200.runtime import GridEngine                    # noqa: E402
from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u    # noqa: E402

name, kind = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
robot = load_named_robot(name)
n = robot.n
if kind == "split":
    eng = GridEngine(robot, plan=G.split_test_plan(robot), tag=G.SPLIT_TEST_TAG)
elif kind == "x2":
    eng = GridEngine(robot, plan=G.x2_test_plan(robot), tag=G.X2_TEST_TAG)
else:
    eng = GridEngine(robot)
q, qd, u, _ = make_states(n, N, 3)
x = torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda()
out = torch.empty(N, 2 * n * n, device="cuda")
res = {"robot": name, "kind": kind, "N": N}
for fam in ("tps", "pipe"):
    if fam not in eng.kernel_kind("fd_grad"):
        continue
    eng.set_option("GRID_FORCE_KERNEL", fam)
    res["us_" + fam] = float(np.median(eng.time_launches("fd_grad", out, x, num_timesteps=N, stride=3 * n, reps=30)))
    if fam == "pipe":
        for w in (8, 4, 2, 1):
            eng.set_option("GRID_PIPE_WARPS", str(w))
            res["us_pipe_w%d" % w] = float(np.median(eng.time_launches("fd_grad", out, x, num_timesteps=N, stride=3 * n, reps=30)))
        eng.set_option("GRID_PIPE_WARPS", None)
        per = {}
        for stage in (0, 1):
            for k in range(40):
                eng.set_option("GRID_PIPE_ONLY_TASK", str(100 * stage + k))
                us = float(np.median(eng.time_launches("fd_grad", out, x, num_timesteps=N, stride=3 * n, reps=10)))
                if us < 6.5 and k > 0:
                    break
                per["s%d_t%d" % (stage, k)] = us
        eng.set_option("GRID_PIPE_ONLY_TASK", "999")          # neither stage: memset + alloc only
        res["us_pipe_empty"] = float(np.median(eng.time_launches("fd_grad", out, x, num_timesteps=N, stride=3 * n, reps=30)))
        eng.set_option("GRID_PIPE_ONLY_TASK", None)
        res["task_us"] = per
eng.set_option("GRID_FORCE_KERNEL", None)
print(json.dumps(res), flush=True)
