#!/usr/bin/env python3
"""Measured FP32 error of every kernel family against the float64 C oracle (which is pinned to the
reference's own outputs): max|x - ref| / max|ref| per output tensor and the worst single state.
JSON lines on stdout."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gridcodegenerator_b200 import load_named_robot                                    # noqa: E402
from gridcodegenerator_b200.runtime import get_engine                                  # noqa: E402
from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u, seed_for        # noqa: E402
from oracle import c_oracle as C                                                       # noqa: E402

SIZES = {"iiwa14": 65536, "hyq": 16384, "atlas": 4096, "chain64": 512, "mixed5": 16384}
for name, N in SIZES.items():
    robot = load_named_robot(name)
    eng = get_engine(robot)
    n = robot.n
    q, qd, u, _ = make_states(n, N, seed_for(name) + 7)
    x = torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda()
    q64, qd64, u64 = (a.astype(np.float64) for a in (q, qd, u))
    for alg, words, call in (("id", n, eng.inverse_dynamics_device), ("minv", n * n, eng.direct_minv_device),
                             ("fd", n, eng.forward_dynamics_device),
                             ("id_grad", 2 * n * n, eng.inverse_dynamics_gradient_device),
                             ("fd_grad", 2 * n * n, eng.forward_dynamics_gradient_device)):
        ref = C.batch(robot, alg, q64, qd64, u64 if alg in ("fd", "fd_grad") else None)
        for fam in ("tps", "wps", "cps", "pipe"):
            if fam not in eng.kernel_kind(alg):
                continue
            os.environ["GRID_FORCE_KERNEL"] = fam
            out = torch.empty(N, words, device="cuda")
            call(out, x)
            torch.cuda.synchronize()
            o = out.cpu().numpy().astype(np.float64)
            per = np.abs(o - ref).max(axis=1) / np.abs(ref).max(axis=1)
            print(json.dumps({"robot": name, "alg": alg, "kernel": fam, "states": N,
                              "rel_err_tensor": float(np.abs(o - ref).max() / np.abs(ref).max()),
                              "rel_err_worst_state": float(per.max()), "rel_err_median_state": float(np.median(per)),
                              "tolerance": 1e-4 if alg in ("id", "minv", "fd") else 1e-3}), flush=True)
os.environ.pop("GRID_FORCE_KERNEL", None)
