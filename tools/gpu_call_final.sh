#!/bin/bash
# round-2 final call (1 GPU) at HEAD: parity suite, bench lines (with the reference's emitted kernel beside ours), launch
# list of the bench command, one ncu --set full of the new ABA program on Atlas, family timings of HyQ Minv
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/f_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/f_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --robot atlas --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/f_bench_atlas.json 2> gpurun_out/f_bench_atlas.err; echo "atlas rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 2 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; echo "ref rc=$?"
python -c "import __graft_entry__ as G; G.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"
CMD="python bench.py --profile --steps 20 --warmup 5"
$CMD > gpurun_out/f_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 60 --csv --log-file gpurun_out/f_launches_iiwa14.csv $CMD > gpurun_out/f_ncu1.log 2>&1
cat > /tmp/aba_run.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np, torch, json
from gridcodegenerator_b200 import load_named_robot
from gridcodegenerator_b200.runtime import get_engine
from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u
name, alg, N = sys.argv[1], sys.argv[2], int(sys.argv[3])
eng = get_engine(load_named_robot(name)); n = eng.n
q, qd, u, _ = make_states(n, N, 3)
x = torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda()
out = torch.empty(N, n * n, device="cuda")
res = {"robot": name, "alg": alg, "N": N, "kind": eng.kernel_kind(alg)}
res["us_auto"] = float(np.median(eng.time_launches(alg, out, x, num_timesteps=N, stride=3 * n, reps=20)))
for fam in ("tps", "pipe", "lps"):
    if fam in eng.kernel_kind(alg):
        eng.set_option("GRID_FORCE_KERNEL", fam)
        res["us_" + fam] = float(np.median(eng.time_launches(alg, out, x, num_timesteps=N, stride=3 * n, reps=20)))
eng.set_option("GRID_FORCE_KERNEL", None)
print(json.dumps(res), flush=True)
PY
for N in 16384 32768 65536 262144; do for a in minv id_grad fd_grad; do timeout 120 python /tmp/aba_run.py hyq $a $N; done; done > gpurun_out/f_hyq_families.jsonl 2> gpurun_out/f_hyq_families.err
cat gpurun_out/f_hyq_families.jsonl
timeout 120 python /tmp/aba_run.py atlas aba 65536 > gpurun_out/f_plain_aba.log 2>&1 &&
ncu --set full --clock-control none --cache-control none -k regex:tps_kernel -s 6 -c 1 -o /tmp/prof_aba_atlas python /tmp/aba_run.py atlas aba 65536 > gpurun_out/f_ncu2.log 2>&1
ncu -i /tmp/prof_aba_atlas.ncu-rep --page raw --csv > gpurun_out/f_prof_aba_atlas_raw.csv 2>/dev/null
python - <<'PY'
import json
for f in ("gpurun_out/f_bench.json", "gpurun_out/f_bench_atlas.json", "gpurun_out/f_bench_ref.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("roofline", {}).get("frac"), json.dumps(d.get("reference_gpu"))[:300])
    except Exception as e:
        print(f, "ERR", e)
PY
ls -la gpurun_out | tail -12
