#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n_pytest.log
tail -8 gpurun_out/n_pytest.log
timeout 600 python tools/bench_matrix.py iiwa14:fd_grad:2048:auto iiwa14:fd_grad:4096:auto iiwa14:fd_grad:4096:tps iiwa14:fd_grad:8192:auto iiwa14:fd_grad:8192:tps iiwa14:fd_grad:16384:auto iiwa14:fd_grad:16384:tps iiwa14:fd_grad:18944:auto iiwa14:fd_grad:18944:tps iiwa14:fd_grad:32768:auto iiwa14:fd_grad:65536:auto > gpurun_out/n_matrix.jsonl 2> gpurun_out/n_matrix.err; cat gpurun_out/n_matrix.jsonl
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/n_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['strong'])"
