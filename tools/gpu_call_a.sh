#!/bin/bash
# round-2 GPU call A: parity suite, bench (both arms), tensor-core micro-experiment, Atlas bench
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/a_gpus.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
timeout 600 ./tools/micro/tc_minv_gemm 16384 > gpurun_out/a_tc_minv_gemm.jsonl 2>&1; echo "tc rc=$?"
timeout 600 python bench.py --robot atlas --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/a_bench_atlas.json 2> gpurun_out/a_bench_atlas.err; echo "atlas rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err; echo "ref rc=$?"
python -c "import __graft_entry__ as G; G.smoke()" > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?"
