#!/bin/bash
# round-2 GPU call B: full parity suite on the rebuilt libraries, Atlas phase-split variants, chain-64 tensor-core variant
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -5 gpurun_out/b_pytest.log
timeout 900 python tools/atlas_variants.py run base g4000 g3000 g2000 legs_g3000 legs_g4500 split_all split_g4500 > gpurun_out/b_atlas_variants.jsonl 2> gpurun_out/b_atlas_variants.err; echo "variants rc=$?"
VARIANT_ROBOT=chain64 timeout 600 python tools/atlas_variants.py run notc tc > gpurun_out/b_chain64_tc.jsonl 2> gpurun_out/b_chain64_tc.err; echo "tc rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?"
