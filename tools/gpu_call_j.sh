#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/atlas_variants.py run base s64 s0 s1024 w4_mb2 w4_mb2_s0 w2_mb4_s0 w1_mb8 lead320 lead80 > gpurun_out/j_atlas_variants.jsonl 2> gpurun_out/j_atlas_variants.err; echo "rc=$?"
tail -3 gpurun_out/j_atlas_variants.err
