#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/lps_variants.py run > gpurun_out/g_lps_variants.jsonl 2> gpurun_out/g_lps_variants.err; echo "rc=$?"
cat gpurun_out/g_lps_variants.jsonl; tail -3 gpurun_out/g_lps_variants.err
