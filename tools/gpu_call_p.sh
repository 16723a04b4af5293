#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/lps_variants.py run > gpurun_out/p_lps_bulk.jsonl 2> gpurun_out/p_lps_bulk.err; echo "rc=$?"
cat gpurun_out/p_lps_bulk.jsonl; tail -3 gpurun_out/p_lps_bulk.err
