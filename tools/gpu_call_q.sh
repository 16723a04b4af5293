#!/bin/bash
# x2 (two states per lane, packed FP32) stage-1 programs on Atlas: parity sample + timing (+ per-task times)
mkdir -p gpurun_out
VARIANT_TASKS=1 VARIANT_QUICK=1 timeout 900 python tools/atlas_variants.py run $VARIANTS > gpurun_out/q_atlas_x2.jsonl 2> gpurun_out/q_atlas_x2.err; echo "rc=$?"
cat gpurun_out/q_atlas_x2.jsonl; tail -5 gpurun_out/q_atlas_x2.err
