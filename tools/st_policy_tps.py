#!/usr/bin/env python3
"""Thread-per-state kernels with streaming (__stcs) vs write-back output stores, buffers rotating over > L2 like bench.py:
  python tools/st_policy_tps.py robot tag [tag ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                                                       # noqa: E402
import torch                                                             # noqa: E402
from gridcodegenerator_b200 import load_named_robot                     # noqa: E402
from gridcodegenerator_b200.codegen import KernelPlan                   # noqa: E402
from gridcodegenerator_b200.runtime import GridEngine                    # noqa: E402
from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u    # noqa: E402

name = sys.argv[1]
robot = load_named_robot(name)
n = robot.n
for tag in sys.argv[2:]:
    eng = GridEngine(robot, plan=KernelPlan(robot, only_algs=("fd_grad", "fd", "minv", "id_grad")), tag="_" + tag)
    res = {"robot": name, "variant": tag}
    for N in (65536, 262144):
        sets = max(2, int(400e6 // (4 * N * (3 * n + 2 * n * n))) + 1)
        q, qd, u, _ = make_states(n, N, 3)
        xs = [torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda() for _ in range(sets)]
        outs = [torch.empty(N, 2 * n * n, device="cuda") for _ in range(sets)]
        for alg, call in (("fd_grad", eng.forward_dynamics_gradient_device), ("id_grad", eng.inverse_dynamics_gradient_device)):
            for i in range(20):
                call(outs[i % sets], xs[i % sets])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 200 if N <= 65536 else 60
            e0.record()
            for i in range(reps):
                call(outs[i % sets], xs[i % sets])
            e1.record()
            torch.cuda.synchronize()
            res["us_%s_N%d" % (alg, N)] = e0.elapsed_time(e1) * 1e3 / reps
        del xs, outs
    print(json.dumps(res), flush=True)
