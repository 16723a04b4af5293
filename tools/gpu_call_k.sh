#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/k_pytest.log
tail -30 gpurun_out/k_pytest.log
timeout 600 python tools/bench_matrix.py atlas:fd_grad:65536:auto atlas:fd_grad:8192:auto hyq:fd_grad:65536:auto hyq:fd:16384:auto iiwa14:fd_grad:65536:auto > gpurun_out/k_matrix.jsonl 2> gpurun_out/k_matrix.err; cat gpurun_out/k_matrix.jsonl
