#!/bin/bash
mkdir -p gpurun_out
python - > gpurun_out/g6_no_store.jsonl 2> gpurun_out/g6_no_store.err <<'PY'
import sys, json
sys.path.insert(0, ".")
import numpy as np, torch
sys.path.insert(0, "tools")
import atlas_variants as AV
from gridcodegenerator_b200 import load_named_robot
from gridcodegenerator_b200.runtime import GridEngine
from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u
robot = load_named_robot("atlas"); n = robot.n; N = 65536
eng = GridEngine(robot, plan=AV.plan_for(robot, "base"), tag="_xbase")
q, qd, u, _ = make_states(n, N, 3)
x = torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda()
out = torch.empty(N, 2 * n * n, device="cuda")
res = {}
for M in (65536, 8192):
    eng.set_option("GRID_PIPE_ONLY_TASK", None)
    res["us_N%d" % M] = float(np.median(eng.time_launches("fd_grad", out, x, num_timesteps=M, stride=3 * n, reps=20)))
    eng.set_option("GRID_PIPE_ONLY_TASK", "1000")
    res["us_N%d_no_output_stores" % M] = float(np.median(eng.time_launches("fd_grad", out, x, num_timesteps=M, stride=3 * n, reps=20)))
eng.set_option("GRID_PIPE_ONLY_TASK", None)
print(json.dumps(res))
PY
cat gpurun_out/g6_no_store.jsonl; tail -3 gpurun_out/g6_no_store.err
