#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/atlas_variants.py run base s0 w4_mb2_s0 w2_mb4_s0 w1_mb8 > gpurun_out/l_atlas_stagger.jsonl 2> gpurun_out/l_atlas_stagger.err; echo "rc=$?"
tail -3 gpurun_out/l_atlas_stagger.err
