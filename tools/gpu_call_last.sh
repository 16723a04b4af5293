#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/last_bench.json 2> gpurun_out/last_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --robot atlas --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/last_bench_atlas.json 2> gpurun_out/last_bench_atlas.err; echo "atlas rc=$?"
python -c "import __graft_entry__ as G; G.smoke()"; echo "smoke rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/last_bench.json", "gpurun_out/last_bench_atlas.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["latency_n128"]["p50_us"], d["gpu_launches"], json.dumps(d.get("reference_gpu"))[:160])
PY
