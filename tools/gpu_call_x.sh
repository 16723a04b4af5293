#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_further_algorithms.py tests/test_abi.py -m gpu -x -q > gpurun_out/x_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/x_pytest.log
timeout 600 python - > gpurun_out/x_aba_timings.jsonl 2> gpurun_out/x_aba_timings.err <<'PY'
import json, sys
sys.path.insert(0, ".")
import numpy as np, torch
from gridcodegenerator_b200 import load_named_robot
from gridcodegenerator_b200.runtime import get_engine
from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u
for name, sizes in (("iiwa14", (65536, 2048, 128)), ("hyq", (65536, 16384, 128)), ("atlas", (65536, 8192, 2048, 512, 128)),
                    ("chain64", (65536, 16384, 1024, 128))):
    eng = get_engine(load_named_robot(name)); n = eng.n
    NM = max(sizes)
    q, qd, u, _ = make_states(n, NM, 3)
    x = torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda()
    out = torch.empty(NM, n * n, device="cuda")
    for N in sizes:
        r = {"robot": name, "N": N, "kind_fd": eng.kernel_kind("fd"), "kind_aba": eng.kernel_kind("aba"), "kind_crba": eng.kernel_kind("crba")}
        for alg in ("aba", "crba", "minv"):
            if eng.kernel_kind(alg) != "none":
                r["us_" + alg] = float(np.median(eng.time_launches(alg, out, x, num_timesteps=N, stride=3 * n, reps=30)))
        r["us_fd_auto"] = float(np.median(eng.time_launches("fd", out, x, num_timesteps=N, stride=3 * n, reps=30)))
        for fam in ("tps", "pipe", "lps", "wps"):
            if fam in eng.kernel_kind("fd"):
                eng.set_option("GRID_FORCE_KERNEL", fam)
                r["us_fd_" + fam] = float(np.median(eng.time_launches("fd", out, x, num_timesteps=N, stride=3 * n, reps=30 if N * n < 2e6 else 5)))
        eng.set_option("GRID_FORCE_KERNEL", None)
        print(json.dumps(r), flush=True)
PY
echo "timings rc=$?"; cat gpurun_out/x_aba_timings.jsonl; tail -3 gpurun_out/x_aba_timings.err
