#!/bin/bash
# Atlas FD gradient: stage-1 items in chunk-major order (GRID_PIPE_ORDER_CHUNK) vs task-major
mkdir -p gpurun_out
VARIANT_ORDER=1 VARIANT_QUICK=1 timeout 600 python tools/atlas_variants.py run base > gpurun_out/r_atlas_order.jsonl 2> gpurun_out/r_atlas_order.err; echo "rc=$?"
cat gpurun_out/r_atlas_order.jsonl; tail -5 gpurun_out/r_atlas_order.err
