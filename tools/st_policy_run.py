#!/usr/bin/env python3
"""Atlas FD gradient with different cache policies on the output stores of the phase-split kernels (libraries built with
-DGRID_PIPE_ST_POLICY=".cs" etc.): python tools/st_policy_run.py tag [tag ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np                                                       # noqa: E402
import torch                                                             # noqa: E402
import atlas_variants as AV                                              # noqa: E402
from gridcodegenerator_b200 import load_named_robot                     # noqa: E402
from gridcodegenerator_b200.runtime import GridEngine                    # noqa: E402
from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u    # noqa: E402
from oracle import c_oracle as C                                         # noqa: E402

robot = load_named_robot("atlas")
n, N = robot.n, 65536
q, qd, u, _ = make_states(n, N, 3)
x = torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda()
out = torch.empty(N, 2 * n * n, device="cuda")
ref = C.batch(robot, "fd_grad", q[-256:], qd[-256:], u[-256:])
for tag in sys.argv[1:]:
    eng = GridEngine(robot, plan=AV.plan_for(robot, "base"), tag="_x" + tag)
    eng.forward_dynamics_gradient_device(out, x)
    torch.cuda.synchronize()
    res = {"variant": tag, "relerr": float(np.abs(out[-256:].cpu().numpy() - ref).max() / np.abs(ref).max())}
    for M in (65536, 16384, 8192, 128):
        res["us_N%d" % M] = float(np.median(eng.time_launches("fd_grad", out, x, num_timesteps=M, stride=3 * n, reps=20)))
    for alg in ("id_grad",):
        pass
    print(json.dumps(res), flush=True)
