#!/bin/bash
mkdir -p gpurun_out
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/end_bench.json 2> gpurun_out/end_bench.err; echo "bench rc=$?"
timeout 200 python bench.py --robot atlas --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/end_bench_atlas.json 2> gpurun_out/end_bench_atlas.err; echo "atlas rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/end_bench.json", "gpurun_out/end_bench_atlas.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_launch"], d["e2e"]["value"], d["roofline"]["frac"], d["latency_n128"]["p50_us"], json.dumps(d["strong"].get("single_gpu_projection"))[:300])
PY
timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/end_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/end_pytest_gpu.log
