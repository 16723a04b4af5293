#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/st_policy_run.py st_def st_cs st_wt st_def > gpurun_out/g11_st_policy.jsonl 2> gpurun_out/g11.err; echo rc=$?
cat gpurun_out/g11_st_policy.jsonl; tail -3 gpurun_out/g11.err
