#!/bin/bash
mkdir -p gpurun_out
VARIANT_QUICK=1 VARIANT_TASKS=1 timeout 600 python tools/atlas_variants.py run base g4000 > gpurun_out/g7_atlas_merged_runs.jsonl 2> gpurun_out/g7.err; echo rc=$?
python - <<'PY'
import json
for l in open("gpurun_out/g7_atlas_merged_runs.jsonl"):
    d = json.loads(l)
    print(d["variant"], d.get("relerr_vs_c_oracle"), d.get("us_N65536"), d.get("us_N8192"), d.get("us_N128"), {k: round(v, 1) for k, v in d.get("task_us_N65536", {}).items()})
PY
tail -3 gpurun_out/g7.err
