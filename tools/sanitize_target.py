#!/usr/bin/env python3
"""Small run of every kernel family for compute-sanitizer (racecheck / memcheck)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gridcodegenerator_b200 import load_named_robot                      # noqa: E402
from gridcodegenerator_b200.runtime import get_engine                    # noqa: E402
from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u    # noqa: E402

for name, N in (("atlas", 5), ("mixed5", 40), ("iiwa14", 70)):
    robot = load_named_robot(name)
    eng = get_engine(robot)
    n = robot.n
    q, qd, u, qdd = make_states(n, N, 1)
    x = torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda()
    for fam in ("tps", "wps", "cps"):
        os.environ["GRID_FORCE_KERNEL"] = fam
        for alg, words, call in (("minv", n * n, eng.direct_minv_device), ("fd", n, eng.forward_dynamics_device),
                                 ("id_grad", 2 * n * n, eng.inverse_dynamics_gradient_device),
                                 ("fd_grad", 2 * n * n, eng.forward_dynamics_gradient_device)):
            if fam not in eng.kernel_kind(alg):
                continue
            out = torch.empty(N, words, device="cuda")
            call(out, x)
            torch.cuda.synchronize()
            assert torch.isfinite(out).all(), (name, fam, alg)
    print("ran", name)
