#!/bin/bash
# round-2 GPU call F: parity suite after the chain-kernel consumers, ncu of the chain kernels
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
tail -25 gpurun_out/f_pytest.log
CMDC="python tools/bench_matrix.py chain64:fd_grad:4096:lps"
$CMDC > gpurun_out/f_plain_chain.log 2>&1 &&
ncu --set full --clock-control none --cache-control none --import-source on -k regex:"grad_columns|stage_a" -s 6 -c 2 -o gpurun_out/prof_r2_lps_chain64 $CMDC > gpurun_out/f_ncu.log 2>&1
ncu -i gpurun_out/prof_r2_lps_chain64.ncu-rep --page raw --csv > gpurun_out/f_prof_r2_lps_chain64_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8; tail -5 gpurun_out/f_ncu.log
