#!/usr/bin/env python3
"""Times the USE_QDD_MINV_FLAG overload of the FD gradient (qdd and Minv given) on the kernel families that serve it:
Atlas phase-split vs wide, 64-link chain chain-kernels (+ tensor-core product) vs wide.  GPU; JSON lines."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gridcodegenerator_b200 import load_named_robot                      # noqa: E402
from gridcodegenerator_b200.runtime import get_engine                    # noqa: E402
from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u    # noqa: E402


def main():
    for name, N, fams in (("atlas", 65536, ("pipe", "wps")), ("atlas", 8192, ("pipe", "wps")), ("chain64", 16384, ("lps", "wps")),
                          ("hyq", 65536, ("pipe", "tps"))):
        robot = load_named_robot(name)
        eng = get_engine(robot)
        n = robot.n
        q, qd, u, qdd = make_states(n, N, 3)
        x = torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda()
        d_qdd = torch.from_numpy(qdd).cuda()
        Minv = torch.empty(N, n * n, device="cuda")
        eng.direct_minv_device(Minv, x)
        out = torch.empty(N, 2 * n * n, device="cuda")
        stream = torch.cuda.current_stream()
        for fam in fams:
            if fam not in eng.kernel_kind("fd_grad"):
                continue
            eng.set_option("GRID_FORCE_KERNEL", fam)
            for _ in range(3):
                eng.forward_dynamics_gradient_device(out, x, d_qdd, Minv)
            torch.cuda.synchronize()
            reps = 3 if (fam == "wps" and N > 10000) else 10
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                eng.forward_dynamics_gradient_device(out, x, d_qdd, Minv)
            e1.record(stream)
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / reps * 1e3
            print(json.dumps({"robot": name, "N": N, "overload": "fd_grad(q, qd, qdd, Minv)", "kernel": fam, "us": us,
                              "evals_per_s": N / us * 1e6}), flush=True)
        eng.set_option("GRID_FORCE_KERNEL", None)


if __name__ == "__main__":
    main()
