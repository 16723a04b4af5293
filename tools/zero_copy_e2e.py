#!/usr/bin/env python3
"""End-to-end FD gradient through HOST buffers: the pipelined H2D / kernel / D2H path of the ABI against one launch
that reads its inputs from and writes its outputs to pinned host memory directly (zero copy over PCIe).
  python tools/zero_copy_e2e.py [robot] [N]           (GPU) -> JSON line"""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                                                       # noqa: E402
import torch                                                             # noqa: E402
from gridcodegenerator_b200 import load_named_robot                     # noqa: E402
from gridcodegenerator_b200.runtime import get_engine                    # noqa: E402
from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u    # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "iiwa14"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
robot = load_named_robot(name)
n = robot.n
eng = get_engine(robot)
q, qd, u, _ = make_states(n, N, 3)
rows = pack_q_qd_u(q, qd, u)
data = eng.make_data(N)
data.h["q_qd_u"][:] = rows
ref = data.forward_dynamics_gradient(N).copy()
h_in = data.h["q_qd_u"].ctypes.data
h_out = data.h["df_du"].ctypes.data
d_in = torch.from_numpy(rows).cuda()
d_out = torch.empty(N, 2 * n * n, device="cuda")
lib = eng.lib
P = ctypes.c_void_p


def launch(out_ptr, in_ptr, cnt, first=0, stream=None):
    rc = lib.grid_forward_dynamics_gradient_device(P(out_ptr + 4 * first * 2 * n * n), P(in_ptr + 4 * first * 3 * n), 3 * n,
                                                   None, None, cnt, 9.81, P(stream) if stream else None)
    assert rc == 0, lib.grid_last_error()


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


res = {"robot": name, "N": N, "h2d_bytes": 4 * N * 3 * n, "d2h_bytes": 4 * N * 2 * n * n}
t = timed(lambda: data.forward_dynamics_gradient(N))
res["pipelined_copies"] = {"ms": t * 1e3, "evals_per_s": N / t}

data.h["df_du"][:] = 0
def zc_both():
    launch(h_out, h_in, N)
t = timed(zc_both)
res["zero_copy_in_out"] = {"ms": t * 1e3, "evals_per_s": N / t, "max_abs_diff_vs_pipelined": float(np.abs(data.h["df_du"][:N] - ref).max())}

data.h["df_du"][:] = 0
pin = torch.from_numpy(rows).pin_memory()
def zc_out():
    d_in.copy_(pin, non_blocking=True)
    launch(h_out, d_in.data_ptr(), N, stream=torch.cuda.current_stream().cuda_stream)
t = timed(zc_out)
res["h2d_copy_zero_copy_out"] = {"ms": t * 1e3, "evals_per_s": N / t, "max_abs_diff_vs_pipelined": float(np.abs(data.h["df_du"][:N] - ref).max())}

# zero-copy launches in chunks on several streams (more CTAs in flight on both PCIe directions)
streams = [torch.cuda.Stream() for _ in range(4)]
for chunks in (2, 4):
    data.h["df_du"][:] = 0
    per = (N + chunks - 1) // chunks
    def zc_chunks():
        for c in range(chunks):
            first = c * per
            cnt = min(per, N - first)
            launch(h_out, h_in, cnt, first, streams[c % 4].cuda_stream)
    t = timed(zc_chunks)
    res["zero_copy_%d_launches" % chunks] = {"ms": t * 1e3, "evals_per_s": N / t,
                                              "max_abs_diff_vs_pipelined": float(np.abs(data.h["df_du"][:N] - ref).max())}
res["pcie_bound_evals_per_s_at_55GBs"] = 55e9 / (4 * 2 * n * n)
print(json.dumps(res), flush=True)
