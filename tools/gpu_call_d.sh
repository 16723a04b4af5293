#!/bin/bash
# round-2 GPU call D (2 GPUs): several devices in one process, 2-rank bench lines (weak + strong records)
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_consumers.py -m gpu -x -q -k "two_devices or bound_to_its_device" > gpurun_out/d_pytest_2gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/d_pytest_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/d_bench_2gpu.json 2> gpurun_out/d_bench_2gpu.err; echo "bench rc=$?"
timeout 600 $TR bench.py --gpus 2 --robot atlas --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/d_bench_atlas_2gpu.json 2> gpurun_out/d_bench_atlas_2gpu.err; echo "atlas rc=$?"
timeout 600 $TR tools/sweep_multi_gpu.py chain64 fd_grad 1024,16384,65536 > gpurun_out/d_sweep_chain64_fdgrad_2gpu.jsonl 2> gpurun_out/d_sweep.err; echo "sweep rc=$?"
tail -2 gpurun_out/d_bench_2gpu.err gpurun_out/d_sweep.err
