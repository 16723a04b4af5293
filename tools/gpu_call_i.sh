#!/bin/bash
# round-2 GPU call I (8 GPUs): aggregate host<->device copy rate at 2/4/8 ranks; chain-64 sweeps with the final chain kernels
set -x
mkdir -p gpurun_out
for G in 8 4 2; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2956$G"
  timeout 200 $TR tools/pcie_aggregate.py 2> gpurun_out/i_pcie_${G}.err | grep '^{' > gpurun_out/i_pcie_${G}gpu.json; echo "pcie $G rc=$?"
done
for G in 8 4; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2957$G"
  timeout 300 $TR tools/sweep_multi_gpu.py chain64 fd_grad 1024,16384,65536 2> gpurun_out/i_sweep_${G}.err | grep '^{' > gpurun_out/i_sweep_chain64_fdgrad_${G}gpu.jsonl; echo "sweep $G rc=$?"
done
cat gpurun_out/i_pcie_*gpu.json
