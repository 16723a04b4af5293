#!/bin/bash
# round-2 GPU call E (8 GPUs): bench lines at 8 and 4 ranks (weak value, strong record, e2e), chain-64 sweeps
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
lscpu | grep -E "NUMA|Socket|^CPU\(s\)" > gpurun_out/e_lscpu.txt 2>&1
nvidia-smi topo -m > gpurun_out/e_topo.txt 2>&1
for G in 8 4; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2954$G"
  timeout 300 $TR bench.py --gpus $G --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/e_bench_${G}gpu.json 2> gpurun_out/e_bench_${G}gpu.err; echo "bench $G rc=$?"
  timeout 300 $TR bench.py --gpus $G --robot atlas --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/e_bench_atlas_${G}gpu.json 2> gpurun_out/e_bench_atlas_${G}gpu.err; echo "atlas $G rc=$?"
  timeout 300 $TR tools/sweep_multi_gpu.py chain64 fd_grad 1024,16384,65536 > gpurun_out/e_sweep_chain64_fdgrad_${G}gpu.jsonl 2> gpurun_out/e_sweep_${G}.err; echo "sweep $G rc=$?"
  timeout 300 $TR tools/sweep_multi_gpu.py chain64 id 1024,65536,1048576 > gpurun_out/e_sweep_chain64_id_${G}gpu.jsonl 2>> gpurun_out/e_sweep_${G}.err; echo "sweep id $G rc=$?"
done
