#!/bin/bash
mkdir -p gpurun_out
(timeout 150 python tools/st_policy_tps.py iiwa14 twb tcs twb tcs; timeout 150 python tools/st_policy_tps.py hyq twb tcs) > gpurun_out/g12_st_policy_tps.jsonl 2> gpurun_out/g12.err; echo rc=$?
cat gpurun_out/g12_st_policy_tps.jsonl; tail -3 gpurun_out/g12.err
