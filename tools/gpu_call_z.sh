#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/z_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/z_pytest_gpu.log
