#!/usr/bin/env python3
"""What caps the end-to-end (host-buffer) path when several ranks copy at once?  Every rank copies pinned host memory
to / from its own GPU - H2D only, D2H only, both - with no kernel in between; rank 0 prints the per-rank and the
aggregate rate.  Run plain (1 GPU) or under torchrun:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29577 tools/pcie_aggregate.py
The 8-GPU pool box is a VM with one 32-CPU NUMA node (profiles/r2_8gpu_box_topology.txt): if the aggregate stops
near the value the 8-rank `e2e` of bench.py moves (117 GB/s), the limit is the host side, not the library.
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gridcodegenerator_b200.sharding import max_over_ranks      # noqa: E402


def main():
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes = 256 << 20
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {"n_gpus": world, "bytes_per_copy": nbytes}
    for mode in ("h2d", "d2h", "both"):
        def once():
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_a.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_b, non_blocking=True)
        for _ in range(2):
            once()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 8
        for _ in range(reps):
            once()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0, world, device="cuda")
        moved = reps * nbytes * (2 if mode == "both" else 1)
        res[mode + "_gbs_per_rank"] = moved / dt / 1e9
        res[mode + "_gbs_aggregate"] = world * moved / dt / 1e9
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
