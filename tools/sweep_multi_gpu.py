#!/usr/bin/env python3
"""Batch sweep of one robot / algorithm on 1..8 GPUs (BASELINE.json config 5: "batch sweep 1 to 1M states at
1/2/4/8 B200, throughput vs FP32/HBM roofline").  Run plain (1 GPU) or under torchrun (one rank per GPU):
  python tools/sweep_multi_gpu.py chain64 id 1,1024,65536,1048576
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      tools/sweep_multi_gpu.py chain64 fd_grad 1024,16384,65536
Weak scaling: every rank runs N states of its own (contiguous shards of an N x G batch, no collective on the data
path); the time of a point is the max over ranks of the CUDA-event time per launch.  One JSON line per point.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gridcodegenerator_b200 import load_named_robot                                      # noqa: E402
from gridcodegenerator_b200.algorithms import algorithmic_bytes, algorithmic_flops      # noqa: E402
from gridcodegenerator_b200.runtime import get_engine                                    # noqa: E402
from gridcodegenerator_b200.sharding import max_over_ranks                               # noqa: E402
from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u, seed_for         # noqa: E402


def main():
    name, alg, sizes = sys.argv[1], sys.argv[2], [int(x) for x in sys.argv[3].split(",")]
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    robot = load_named_robot(name)
    eng = get_engine(robot)
    n = robot.n
    words = {"id": n, "minv": n * n, "fd": n, "id_grad": 2 * n * n, "fd_grad": 2 * n * n}[alg]
    call = {"fd_grad": eng.forward_dynamics_gradient_device, "id_grad": eng.inverse_dynamics_gradient_device,
            "fd": eng.forward_dynamics_device, "minv": eng.direct_minv_device, "id": eng.inverse_dynamics_device}[alg]
    fp32_peak = eng.measure_fp32_tflops(3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    stream = torch.cuda.current_stream()
    for N in sizes:
        q, qd, u, _ = make_states(n, N, seed_for(name) + 100 * rank)
        x = torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda()
        out = torch.empty(N, words, device="cuda")
        for _ in range(3):
            call(out, x, stream=stream)
        torch.cuda.synchronize()
        # enough launches for >= ~20 ms of timed work, at least 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        call(out, x, stream=stream)
        e1.record(stream)
        torch.cuda.synchronize()
        reps = int(min(200, max(5, 20.0 / max(e0.elapsed_time(e1), 1e-3))))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(reps):
            call(out, x, stream=stream)
        e1.record(stream)
        torch.cuda.synchronize()
        us = max_over_ranks(e0.elapsed_time(e1) / reps * 1e3, world, device="cuda")
        if rank == 0:
            ev = world * N / us * 1e6
            print(json.dumps({"robot": name, "alg": alg, "n_gpus": world, "states_per_gpu": N, "us_per_launch": us,
                              "launches_timed": reps, "evals_per_s": ev, "kernel": eng.kernel_kind(alg),
                              "algorithmic_tflops_per_gpu": algorithmic_flops(robot)[alg] * N / us / 1e6,
                              "fp32_frac": algorithmic_flops(robot)[alg] * N / us / 1e6 / fp32_peak,
                              "hbm_gbs_per_gpu": algorithmic_bytes(robot)[alg] * N / us / 1e3,
                              "hbm_frac": algorithmic_bytes(robot)[alg] * N / us / 1e3 / hbm,
                              "fp32_peak_tflops": fp32_peak}), flush=True)
        del x, out
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
