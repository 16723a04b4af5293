#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/fam_run.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np, torch, json
from gridcodegenerator_b200 import load_named_robot
from gridcodegenerator_b200.runtime import get_engine
from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u
name = sys.argv[1]
eng = get_engine(load_named_robot(name)); n = eng.n
NM = 262144
q, qd, u, _ = make_states(n, NM, 3)
x = torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda()
out = torch.empty(NM, 2 * n * n, device="cuda")
for alg in sys.argv[2:]:
    for N in (128, 2048, 16384, 32768, 65536, 262144):
        res = {"robot": name, "alg": alg, "N": N, "kind": eng.kernel_kind(alg)}
        res["us_auto"] = float(np.median(eng.time_launches(alg, out, x, num_timesteps=N, stride=3 * n, reps=20)))
        for fam in ("tps", "pipe"):
            if fam in eng.kernel_kind(alg):
                eng.set_option("GRID_FORCE_KERNEL", fam)
                res["us_" + fam] = float(np.median(eng.time_launches(alg, out, x, num_timesteps=N, stride=3 * n, reps=20)))
        eng.set_option("GRID_FORCE_KERNEL", None)
        print(json.dumps(res), flush=True)
PY
timeout 300 python /tmp/fam_run.py hyq id minv fd id_grad fd_grad > gpurun_out/g2_hyq_families.jsonl 2> gpurun_out/g2_hyq_families.err; echo rc=$?
cat gpurun_out/g2_hyq_families.jsonl; tail -3 gpurun_out/g2_hyq_families.err
