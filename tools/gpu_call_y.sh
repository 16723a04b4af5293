#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/y_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/y_pytest_gpu.log
timeout 600 python - > gpurun_out/y_aba_timings.jsonl 2> gpurun_out/y_aba_timings.err <<'PY'
import json, sys
sys.path.insert(0, ".")
import numpy as np, torch
from gridcodegenerator_b200 import load_named_robot
from gridcodegenerator_b200.runtime import get_engine
from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u
for name, sizes in (("iiwa14", (65536, 32768, 2048, 128)), ("hyq", (65536, 32768, 16384, 128)), ("atlas", (65536, 8192, 256, 128)),
                    ("chain64", (262144, 65536, 16384, 1024, 256, 128))):
    eng = get_engine(load_named_robot(name)); n = eng.n
    NM = max(sizes)
    q, qd, u, _ = make_states(n, NM, 3)
    x = torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda()
    out = torch.empty(NM, n * n, device="cuda")
    for N in sizes:
        r = {"robot": name, "N": N, "kind_fd": eng.kernel_kind("fd"), "fd_at_large": eng.kernel_kind("fd@large"),
             "kind_aba": eng.kernel_kind("aba"), "kind_crba": eng.kernel_kind("crba")}
        for alg in ("aba", "crba", "minv"):
            if eng.kernel_kind(alg) != "none":
                r["us_" + alg] = float(np.median(eng.time_launches(alg, out, x, num_timesteps=N, stride=3 * n, reps=30)))
        r["us_fd_auto"] = float(np.median(eng.time_launches("fd", out, x, num_timesteps=N, stride=3 * n, reps=30)))
        r["fd_evals_per_s"] = N / r["us_fd_auto"] * 1e6
        for fam in ("tps", "pipe", "lps"):
            if fam in eng.kernel_kind("fd"):
                eng.set_option("GRID_FORCE_KERNEL", fam)
                r["us_fd_" + fam] = float(np.median(eng.time_launches("fd", out, x, num_timesteps=N, stride=3 * n, reps=30 if N * n < 2e6 else 5)))
        eng.set_option("GRID_FORCE_KERNEL", None)
        print(json.dumps(r), flush=True)
PY
echo "timings rc=$?"; cut -c1-600 gpurun_out/y_aba_timings.jsonl; tail -3 gpurun_out/y_aba_timings.err
