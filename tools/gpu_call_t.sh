#!/bin/bash
mkdir -p gpurun_out
for k in split x2; do timeout 300 python tools/split_tasks.py iiwa14 $k 65536; done > gpurun_out/t_split_tasks.jsonl 2> gpurun_out/t_split_tasks.err; echo "rc=$?"
timeout 300 python tools/split_tasks.py iiwa14 split 524288 >> gpurun_out/t_split_tasks.jsonl 2>> gpurun_out/t_split_tasks.err
cat gpurun_out/t_split_tasks.jsonl; tail -5 gpurun_out/t_split_tasks.err
