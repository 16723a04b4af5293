#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/minv_variants.py run > gpurun_out/g4_minv_variants.jsonl 2> gpurun_out/g4_minv_variants.err; echo rc=$?
cat gpurun_out/g4_minv_variants.jsonl; tail -3 gpurun_out/g4_minv_variants.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "chunk_major" 2>&1 | tail -3
