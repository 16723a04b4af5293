#!/bin/bash
mkdir -p gpurun_out
VARIANT_ROBOT=iiwa14 VARIANT_FORCE=pipe VARIANT_QUICK=1 VARIANT_TASKS=1 timeout 600 python tools/atlas_variants.py run i_g2800 i_g4200 i_g9000 i_g4200_w16 i_g2800_w16 > gpurun_out/u_iiwa_groups.jsonl 2> gpurun_out/u_iiwa_groups.err; echo "rc=$?"
cut -c1-1500 gpurun_out/u_iiwa_groups.jsonl; tail -5 gpurun_out/u_iiwa_groups.err
