#!/bin/bash
mkdir -p gpurun_out
(timeout 200 python tools/minv_probe.py atlas minv; timeout 200 python tools/minv_probe.py atlas fd; timeout 200 python tools/minv_probe.py hyq minv) > gpurun_out/g5_minv_probe.jsonl 2> gpurun_out/g5_minv_probe.err; echo rc=$?
cat gpurun_out/g5_minv_probe.jsonl; tail -3 gpurun_out/g5_minv_probe.err
ncu --set full --clock-control none --cache-control none -k regex:pipe_kernel -s 4 -c 1 -o /tmp/prof_minv_atlas python tools/minv_probe.py atlas minv > gpurun_out/g5_ncu.log 2>&1
ncu -i /tmp/prof_minv_atlas.ncu-rep --page raw --csv > gpurun_out/g5_prof_minv_atlas_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/g5_prof_minv_atlas_raw.csv "Atlas Minv pipe" | head -50
