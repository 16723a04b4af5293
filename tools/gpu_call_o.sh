#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/o_pytest.log
tail -6 gpurun_out/o_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --robot atlas --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/o_bench_atlas.json 2> gpurun_out/o_bench_atlas.err; echo "atlas rc=$?"
python -c "import __graft_entry__ as G; G.smoke()"; echo "smoke rc=$?"
