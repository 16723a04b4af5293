#!/usr/bin/env python3
"""Times (robot, algorithm, batch, kernel family) combinations with CUDA events; JSON lines out.
  python tools/bench_matrix.py iiwa14:fd_grad:128:wps iiwa14:fd_grad:128:tps atlas:fd_grad:65536:auto ...
A robot name may carry a library tag (iiwa14@_pipe16: an experimental build made beforehand with
build_robot_library(..., tag="_pipe16")).  Timing only: parity lives in tests/ (tests/parity_report.py,
tests/fuzz_parity.py).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gridcodegenerator_b200 import load_named_robot                      # noqa: E402
from gridcodegenerator_b200.algorithms import algorithmic_flops         # noqa: E402
from gridcodegenerator_b200.runtime import get_engine                    # noqa: E402
from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u    # noqa: E402


def main():
    for spec in sys.argv[1:]:
        name, alg, N, family = spec.split(":")
        name, _, tag = name.partition("@")
        N = int(N)
        robot = load_named_robot(name)
        eng = get_engine(robot, tag=tag)
        n = robot.n
        if family in ("tps", "wps", "cps", "pipe", "lps"):
            os.environ["GRID_FORCE_KERNEL"] = family
        else:
            os.environ.pop("GRID_FORCE_KERNEL", None)
        if family in ("tps", "wps", "cps", "pipe", "lps") and family not in eng.kernel_kind(alg):
            print(json.dumps({"spec": spec, "skipped": "no %s kernel" % family}), flush=True)
            continue
        q, qd, u, _ = make_states(n, N, 3)
        x = torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda()
        outw = {"id": n, "minv": n * n, "fd": n, "id_grad": 2 * n * n, "fd_grad": 2 * n * n}[alg]
        out = torch.empty(N, outw, device="cuda")
        call = {"fd_grad": eng.forward_dynamics_gradient_device, "id_grad": eng.inverse_dynamics_gradient_device,
                "fd": eng.forward_dynamics_device, "minv": eng.direct_minv_device, "id": eng.inverse_dynamics_device}[alg]
        for _ in range(5):
            call(out, x)
        torch.cuda.synchronize()
        reps = 100 if N <= 4096 else 20
        us = eng.time_launches(alg, out, x, reps=reps)      # event pairs recorded in C, launches queued back to back
        p50 = float(np.median(us))
        fl = algorithmic_flops(robot)[alg]
        print(json.dumps({"spec": spec, "p50_us": p50, "min_us": float(us.min()), "evals_per_s": N / p50 * 1e6,
                          "alg_tflops": fl * N / p50 / 1e6}), flush=True)


if __name__ == "__main__":
    main()
