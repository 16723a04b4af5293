#!/usr/bin/env python3
"""Randomised parity sweep: random batch sizes, seeds, strides and kernel families against the float64
C oracle (pinned to the reference's outputs).  Prints one JSON line per case and a summary; exit 1 on any
tolerance violation.
  python tests/fuzz_parity.py [cases]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gridcodegenerator_b200 import load_named_robot                       # noqa: E402
from gridcodegenerator_b200.runtime import get_engine                     # noqa: E402
from gridcodegenerator_b200.synthetic import make_states                  # noqa: E402
from oracle import c_oracle as C                                          # noqa: E402

TOL = {"id": 1e-4, "minv": 1e-4, "fd": 1e-4, "id_grad": 1e-3, "fd_grad": 1e-3}


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rng = np.random.default_rng(20261018)
    robots = {n: load_named_robot(n) for n in ("atlas", "hyq", "iiwa14", "mixed5")}
    worst, bad = {}, 0
    for case in range(cases):
        name = ("atlas", "hyq", "iiwa14", "mixed5")[case % 4]
        robot, eng = robots[name], get_engine(robots[name])
        n = robot.n
        alg = ("id", "minv", "fd", "id_grad", "fd_grad")[int(rng.integers(5))]
        fams = [f for f in ("tps", "cps", "pipe", "wps") if f in eng.kernel_kind(alg)]
        fam = fams[int(rng.integers(len(fams)))]
        N = int(rng.choice([1, 2, 31, 32, 33, 63, 255, 256, 257, 1000, int(rng.integers(1, 6000))]))
        pad = int(rng.choice([0, 0, 1, 5]))                       # extra words per input row (stride > 3n)
        q, qd, u, qdd = make_states(n, N, int(rng.integers(1 << 30)))
        rows = np.concatenate([q, qd, u, np.full((N, pad), 7.0, np.float32)], axis=1)
        x = torch.from_numpy(np.ascontiguousarray(rows)).cuda()
        words = {"id": n, "minv": n * n, "fd": n, "id_grad": 2 * n * n, "fd_grad": 2 * n * n}[alg]
        out = torch.full((N + 2, words), 7.0, device="cuda")
        use_qdd = alg in ("id", "id_grad") and bool(rng.integers(2))
        dq = torch.from_numpy(qdd).cuda() if use_qdd else None
        os.environ["GRID_FORCE_KERNEL"] = fam
        if fam == "pipe" and rng.integers(2):
            os.environ["GRID_PIPE_MODE"] = "fused"
        else:
            os.environ.pop("GRID_PIPE_MODE", None)
        stride = 3 * n + pad
        if alg == "id":
            eng.inverse_dynamics_device(out[1:N + 1], x, dq, num_timesteps=N, stride=stride)
        elif alg == "minv":
            eng.direct_minv_device(out[1:N + 1], x, num_timesteps=N, stride=stride)
        elif alg == "fd":
            eng.forward_dynamics_device(out[1:N + 1], x, num_timesteps=N, stride=stride)
        elif alg == "id_grad":
            eng.inverse_dynamics_gradient_device(out[1:N + 1], x, dq, num_timesteps=N, stride=stride)
        else:
            eng.forward_dynamics_gradient_device(out[1:N + 1], x, num_timesteps=N, stride=stride)
        torch.cuda.synchronize()
        o = out.cpu().numpy().astype(np.float64)
        q64, qd64, u64, qdd64 = (a.astype(np.float64) for a in (q, qd, u, qdd))
        third = u64 if alg in ("fd", "fd_grad") else (qdd64 if use_qdd else None)
        ref = C.batch(robot, alg, q64, qd64, third)
        err = float(np.abs(o[1:N + 1] - ref).max() / np.abs(ref).max())
        guards = bool(np.all(o[0] == 7.0) and np.all(o[-1] == 7.0))
        ok = err < TOL[alg] and guards and np.isfinite(o).all()
        bad += not ok
        key = "%s:%s:%s" % (name, alg, fam)
        worst[key] = max(worst.get(key, 0.0), err)
        print(json.dumps({"robot": name, "alg": alg, "kernel": fam, "mode": os.environ.get("GRID_PIPE_MODE", "staged"),
                          "N": N, "stride": stride, "use_qdd": use_qdd, "rel_err": err, "guards_intact": guards,
                          "ok": bool(ok)}), flush=True)
    print(json.dumps({"summary": True, "cases": cases, "violations": int(bad), "worst_rel_err": worst}))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
