#!/bin/bash
# round-2 GPU call C: parity suite (chain kernels, wide _inner/_device, Atlas 65 536), chain-64 timings,
# steady-state ncu of the two headline kernels (reports converted to CSV on the box: gpurun_out is capped at 64 MiB)
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -25 gpurun_out/c_pytest.log
timeout 900 python tools/bench_matrix.py chain64:fd_grad:16384:lps chain64:fd_grad:16384:wps chain64:id_grad:16384:lps chain64:id_grad:16384:wps \
    chain64:minv:16384:lps chain64:minv:16384:wps chain64:fd:16384:lps chain64:fd:16384:wps \
    chain64:fd_grad:128:lps chain64:fd_grad:128:wps chain64:fd_grad:256:lps chain64:fd_grad:256:wps chain64:fd_grad:512:lps chain64:fd_grad:512:wps chain64:fd_grad:1024:lps chain64:fd_grad:4096:lps \
    chain64:fd_grad:65536:lps chain64:id_grad:65536:lps chain64:fd:65536:lps chain64:minv:65536:lps \
    atlas:fd_grad:65536:auto atlas:fd_grad:8192:auto atlas:fd_grad:16384:auto atlas:fd_grad:32768:auto \
    > gpurun_out/c_matrix.jsonl 2> gpurun_out/c_matrix.err; echo "matrix rc=$?"
tail -3 gpurun_out/c_matrix.err
# ---- ncu: iiwa14 headline kernel, steady state (mid-run launches, caches not flushed) ----
CMD="python bench.py --profile --steps 20 --warmup 5"
METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__inst_executed.sum
$CMD > gpurun_out/c_plain_iiwa.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 60 --csv --log-file gpurun_out/c_launches_iiwa14.csv $CMD > gpurun_out/c_ncu1.log 2>&1
$CMD > gpurun_out/c_plain_iiwa.log 2>&1 &&
ncu --replay-mode application --clock-control none --cache-control none --metrics $METRICS \
    -k regex:tps_kernel -s 10 -c 10 --csv --log-file gpurun_out/c_traffic_iiwa14.csv $CMD > gpurun_out/c_ncu2.log 2>&1
$CMD > gpurun_out/c_plain_iiwa.log 2>&1 &&
ncu --set full --clock-control none --cache-control none --import-source on -k regex:tps_kernel -s 12 -c 1 -o gpurun_out/prof_r2_tps_iiwa14 $CMD > gpurun_out/c_ncu3.log 2>&1
# ---- ncu: Atlas phase-split kernels ----
CMDA="python bench.py --robot atlas --profile --steps 6 --warmup 3"
$CMDA > gpurun_out/c_plain_atlas.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 40 --csv --log-file gpurun_out/c_launches_atlas.csv $CMDA > gpurun_out/c_ncu6.log 2>&1
$CMDA > gpurun_out/c_plain_atlas.log 2>&1 &&
ncu --replay-mode application --clock-control none --cache-control none --metrics $METRICS \
    -k regex:pipe_kernel -s 6 -c 8 --csv --log-file gpurun_out/c_traffic_atlas.csv $CMDA > gpurun_out/c_ncu4.log 2>&1
$CMDA > gpurun_out/c_plain_atlas.log 2>&1 &&
ncu --set full --clock-control none --cache-control none -k regex:pipe_kernel -s 8 -c 2 -o /tmp/prof_r2_pipe_atlas $CMDA > gpurun_out/c_ncu5.log 2>&1
ncu -i /tmp/prof_r2_pipe_atlas.ncu-rep --page raw --csv > gpurun_out/c_prof_r2_pipe_atlas_raw.csv 2>/dev/null
# ---- ncu: chain-64 chain kernels (one chunk) ----
CMDC="python tools/bench_matrix.py chain64:fd_grad:4096:lps"
$CMDC > gpurun_out/c_plain_chain.log 2>&1 &&
ncu --set full --clock-control none --cache-control none -k regex:lps -s 4 -c 2 -o /tmp/prof_r2_lps_chain64 $CMDC > gpurun_out/c_ncu7.log 2>&1
ncu -i /tmp/prof_r2_lps_chain64.ncu-rep --page raw --csv > gpurun_out/c_prof_r2_lps_chain64_raw.csv 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out/ | tail -25
