#!/bin/bash
# full GPU suite after the pipe-shell changes (64-state tiles, packed column programs, item order option)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/s_pytest_gpu.log
