#!/usr/bin/env python3
"""Reference GPU arm: times the reference's own emitted kernels (baseline/_ref/<robot>/ref_harness*,
built by baseline/make_reference_cuh.py) on this box, next to our kernels, on identical inputs, and
cross-checks the two outputs.  JSON lines on stdout.
  python tools/ref_gpu_bench.py [robot] [N ...]          (REF_ALGS=id,minv,fd,id_grad,fd_grad selects algorithms)
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gridcodegenerator_b200 import load_named_robot                      # noqa: E402
from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u    # noqa: E402


ALGS = os.environ.get("REF_ALGS", "fd_grad").split(",")


def ours(robot, x, N, family, alg="fd_grad"):
    import torch
    from gridcodegenerator_b200.runtime import get_engine
    eng = get_engine(robot)
    n = robot.n
    if family:
        os.environ["GRID_FORCE_KERNEL"] = family
    if family not in eng.kernel_kind(alg):
        os.environ.pop("GRID_FORCE_KERNEL", None)
        raise RuntimeError("no %s kernel for %s" % (family, alg))
    xin = torch.from_numpy(x).cuda()
    words = {"id": n, "minv": n * n, "fd": n, "id_grad": 2 * n * n, "fd_grad": 2 * n * n}[alg]
    out = torch.empty(N, words, device="cuda")
    us = eng.time_launches(alg, out, xin, reps=50)          # event pairs recorded in C
    os.environ.pop("GRID_FORCE_KERNEL", None)
    return float(np.median(us)), out.cpu().numpy()


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "iiwa14"
    sizes = [int(a) for a in sys.argv[2:]] or [128, 65536]
    robot = load_named_robot(name)
    n = robot.n
    ref_dir = os.path.join(ROOT, "baseline", "_ref", name)
    for N, alg in [(N, a) for N in sizes for a in ALGS]:
        q, qd, u, _ = make_states(n, N, 77)
        x = pack_q_qd_u(q, qd, u)
        ours_res = {}
        for fam in ("tps", "wps", "cps"):
            try:
                t, out = ours(robot, x, N, fam, alg)
                ours_res[fam] = out
                print(json.dumps({"impl": "b200_" + fam, "alg": alg, "N": N, "p50_us": t,
                                  "evals_per_s": N / t * 1e6}), flush=True)
            except Exception as e:          # family not built for this robot
                print(json.dumps({"impl": "b200_" + fam, "N": N, "skipped": str(e)[:120]}), flush=True)
        with tempfile.TemporaryDirectory() as td:
            with open(os.path.join(td, "in.bin"), "wb") as f:
                f.write(np.int32(N).tobytes())
                f.write(x.tobytes())
            for exe, threads in (("ref_harness", 256), ("ref_harness", 128), ("ref_harness_r168", 352),
                                 ("ref_harness_r168", 256)):
                path = os.path.join(ref_dir, exe)
                if not os.path.exists(path):
                    print(json.dumps({"impl": "reference_gpu", "unavailable": path}), flush=True)
                    continue
                p = subprocess.run([path, os.path.join(td, "in.bin"), os.path.join(td, "out.bin"), str(threads),
                                    "30", alg], capture_output=True, text=True, timeout=600)
                if p.returncode != 0:
                    print(json.dumps({"impl": "reference_gpu", "exe": exe, "threads": threads, "N": N,
                                      "failed": (p.stdout + p.stderr)[-300:]}), flush=True)
                    continue
                line = json.loads(p.stdout.strip().splitlines()[-1])
                line["exe"] = exe
                ref_out = np.fromfile(os.path.join(td, "out.bin"), dtype=np.float32).reshape(N, -1)
                if ours_res:
                    mine = next(iter(ours_res.values()))
                    line["max_rel_diff_vs_b200"] = float(np.abs(ref_out - mine).max() / np.abs(mine).max())
                print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
