#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python tools/zero_copy_e2e.py iiwa14 65536; timeout 300 python tools/zero_copy_e2e.py iiwa14 262144; timeout 300 python tools/zero_copy_e2e.py atlas 16384; timeout 300 python tools/zero_copy_e2e.py hyq 65536) > gpurun_out/v_zero_copy.jsonl 2> gpurun_out/v_zero_copy.err; echo "rc=$?"
cat gpurun_out/v_zero_copy.jsonl; tail -5 gpurun_out/v_zero_copy.err
