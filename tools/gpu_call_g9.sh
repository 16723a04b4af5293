#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/g9_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/g9_pytest_gpu.log
