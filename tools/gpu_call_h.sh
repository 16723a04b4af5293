#!/bin/bash
# round-2 GPU call H (1 GPU): parity on the final chain-kernel build, chain-64 batch sweeps 1..1M, full bench lines
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h_pytest.log
tail -6 gpurun_out/h_pytest.log
timeout 600 python tools/sweep_multi_gpu.py chain64 fd_grad 1,8,32,128,256,512,1024,4096,16384,65536,262144 > gpurun_out/h_sweep_chain64_fdgrad_1gpu.jsonl 2> gpurun_out/h_sweep.err; echo "sweep rc=$?"
timeout 600 python tools/sweep_multi_gpu.py chain64 id 1,32,1024,16384,65536,262144,1048576 > gpurun_out/h_sweep_chain64_id_1gpu.jsonl 2>> gpurun_out/h_sweep.err; echo "sweep rc=$?"
timeout 600 python tools/sweep_multi_gpu.py chain64 id_grad 128,1024,16384,65536 > gpurun_out/h_sweep_chain64_idgrad_1gpu.jsonl 2>> gpurun_out/h_sweep.err
timeout 600 python tools/sweep_multi_gpu.py chain64 fd 128,1024,16384,65536,1048576 > gpurun_out/h_sweep_chain64_fd_1gpu.jsonl 2>> gpurun_out/h_sweep.err
timeout 600 python tools/sweep_multi_gpu.py chain64 minv 128,1024,16384,65536 > gpurun_out/h_sweep_chain64_minv_1gpu.jsonl 2>> gpurun_out/h_sweep.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 2 > gpurun_out/h_bench_ref.json 2> gpurun_out/h_bench_ref.err; echo "ref rc=$?"
timeout 300 python tools/pcie_aggregate.py > gpurun_out/h_pcie_1gpu.json 2>&1
python -c "import __graft_entry__ as G; G.smoke()" > gpurun_out/h_smoke.log 2>&1; echo "smoke rc=$?"
