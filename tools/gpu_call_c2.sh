#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chain" > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/c2_pytest.log
timeout 300 python tools/bench_matrix.py chain64:fd_grad:1024:lps chain64:fd_grad:1024:wps > gpurun_out/c2_matrix.jsonl 2> gpurun_out/c2_matrix.err; echo "matrix rc=$?"
tail -15 gpurun_out/c2_matrix.err; cat gpurun_out/c2_matrix.jsonl
