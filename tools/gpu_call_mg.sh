#!/bin/bash
# bench.py under torchrun at HEAD: G ranks of one box
G=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2961$G"
timeout 600 $TR bench.py --gpus $G --steps 20 --warmup 5 2> gpurun_out/mg_bench_${G}.err | grep '^{' > gpurun_out/mg_bench_${G}gpu.json; echo "bench $G rc=$?"
timeout 600 $TR bench.py --gpus $G --robot atlas --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/mg_bench_atlas_${G}.err | grep '^{' > gpurun_out/mg_bench_atlas_${G}gpu.json; echo "atlas $G rc=$?"
python - <<PY
import json
for f in ("gpurun_out/mg_bench_${G}gpu.json", "gpurun_out/mg_bench_atlas_${G}gpu.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], json.dumps(d.get("strong"))[:400])
PY
tail -3 gpurun_out/mg_bench_${G}.err
