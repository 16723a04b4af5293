#!/usr/bin/env python3
"""Atlas FD-gradient experiments on the phase-split kernels: tagged library variants (fd_grad only) built
here, timed on the B200.
  python tools/atlas_variants.py build [names...]     (CPU, parallel nvcc)
  python tools/atlas_variants.py run   [names...]     (GPU) -> JSON lines
Every variant is checked against the C oracle on a sample before it is timed.
"""
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gridcodegenerator_b200 import load_named_robot                     # noqa: E402
from gridcodegenerator_b200.build import build_robot_library            # noqa: E402
from gridcodegenerator_b200.codegen import KernelPlan                   # noqa: E402

ROBOT = os.environ.get("VARIANT_ROBOT", "atlas")
ALG = "fd_grad"

# name -> KernelPlan keyword arguments (only_algs is added to all)
VARIANTS = {
    "base":      dict(),
    "g4000":     dict(pipe_opts=dict(group_flops=4000)),
    "g3000":     dict(pipe_opts=dict(group_flops=3000)),
    "g2000":     dict(pipe_opts=dict(group_flops=2000)),
    "legs_g3000": dict(pipe_opts=dict(group_flops=3000, single_stage_max_flops=4000)),
    "legs_g4500": dict(pipe_opts=dict(group_flops=4500, single_stage_max_flops=4000)),
    "split_all": dict(pipe_opts=dict(group_flops=3000, single_stage_max_flops=4000, split_sides_above=3500)),
    "split_g4500": dict(pipe_opts=dict(group_flops=4500, single_stage_max_flops=4000, split_sides_above=4500)),
    "w4_g3000":  dict(pipe_opts=dict(group_flops=3000), pipe_warps=4, pipe_min_blocks=(2, 2)),
    # 64-link chain, wide kernel: -Minv dc_du on the tensor cores (3xTF32 mma.sync) vs FP32 FFMA
    "tc":        dict(wps_tc_matmul=True),
    "notc":      dict(),
    "s64":       dict(pipe_sync_every=64),
    "s0":        dict(pipe_sync_every=0),
    "s1024":     dict(pipe_sync_every=1024),
    "w4_mb2":    dict(pipe_warps=4, pipe_min_blocks=(2, 2)),
    "w4_mb2_s0": dict(pipe_warps=4, pipe_min_blocks=(2, 2), pipe_sync_every=0),
    "w2_mb4_s0": dict(pipe_warps=2, pipe_min_blocks=(4, 4), pipe_sync_every=0),
    "w1_mb8":    dict(pipe_warps=1, pipe_min_blocks=(8, 8)),
    "lead320":   dict(pipe_scratch_lead=320),
    "lead80":    dict(pipe_scratch_lead=80),
    # stage-1 programs with two states per lane (FFMA2 / FMUL2 / FADD2)
    "x2":        dict(pipe_x2=True),
    "x2_g4000":  dict(pipe_x2=True, pipe_opts=dict(group_flops=4000)),
    "x2_g3000":  dict(pipe_x2=True, pipe_opts=dict(group_flops=3000)),
    "x2_s0":     dict(pipe_x2=True, pipe_sync_every=0),
    # ... only the programs whose register demand allows it (scratch words re-loaded per cluster of uses)
    "x2m":       dict(pipe_x2=dict(max_live=110, remat_gap=200, lead=80)),
    "x2m_g4000": dict(pipe_x2=dict(max_live=110, remat_gap=200, lead=80), pipe_opts=dict(group_flops=4000)),
    "x2m_ml125": dict(pipe_x2=dict(max_live=125, remat_gap=100, lead=40)),
    "x2m_gap100": dict(pipe_x2=dict(max_live=110, remat_gap=100, lead=40)),
    # iiwa14 (VARIANT_ROBOT=iiwa14 VARIANT_FORCE=pipe): few LARGE column groups - every program under the 64 KB that
    # stay in the instruction cache, scratch words read by 1-3 programs instead of 7
    "i_g2800":   dict(pipe_algs=("fd_grad",), pipe_opts=dict(single_stage_max_flops=0, group_flops=2800)),
    "i_g4200":   dict(pipe_algs=("fd_grad",), pipe_opts=dict(single_stage_max_flops=0, group_flops=4200)),
    "i_g9000":   dict(pipe_algs=("fd_grad",), pipe_opts=dict(single_stage_max_flops=0, group_flops=9000)),
    "i_g4200_w16": dict(pipe_algs=("fd_grad",), pipe_opts=dict(single_stage_max_flops=0, group_flops=4200), pipe_warps=16),
    "i_g2800_w16": dict(pipe_algs=("fd_grad",), pipe_opts=dict(single_stage_max_flops=0, group_flops=2800), pipe_warps=16),
    # iiwa14 thread-per-state FD gradient at 9 / 10 / 12 resident warps per SM (224 / 200 / 168 registers)
    "i_mb8":     dict(tps_min_blocks={"fd_grad": 8}),
    "i_mb9":     dict(tps_min_blocks={"fd_grad": 9}),
    "i_mb10":    dict(tps_min_blocks={"fd_grad": 10}),
    "i_mb12":    dict(tps_min_blocks={"fd_grad": 12}),
    "x2m_w4":    dict(pipe_x2=dict(max_live=110, remat_gap=200, lead=80), pipe_warps=4, pipe_min_blocks=(2, 2)),
}


def plan_for(robot, name):
    return KernelPlan(robot, only_algs=(ALG,), **VARIANTS[name])


def _build(name):
    robot = load_named_robot(ROBOT)
    t = time.time()
    so, info = build_robot_library(robot, plan_for(robot, name), tag="_x" + name)
    spills = [l.strip() for l in info.get("ptxas", "").splitlines() if "spill" in l and " 0 bytes spill stores" not in l]
    tasks = info.get("stats", {}).get("pipe_" + ALG, {})
    return name, time.time() - t, len(spills), tasks.get("scratch_words"), len(tasks.get("tasks", [])), os.path.basename(so)


def build(names):
    with ProcessPoolExecutor(max_workers=int(os.environ.get("BUILD_JOBS", "5"))) as ex:
        for r in ex.map(_build, names):
            print("%-12s %6.1fs  kernels_with_spills=%d scratch_words=%s tasks=%s  %s" % r, flush=True)


def run(names):
    import numpy as np
    import torch
    from gridcodegenerator_b200.runtime import GridEngine
    from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u
    from oracle import c_oracle as C
    robot = load_named_robot(ROBOT)
    n = robot.n
    NMAX = 65536 if ROBOT != "chain64" else 16384
    q, qd, u, _ = make_states(n, NMAX, 3)
    x = torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda()
    out = torch.empty(NMAX, 2 * n * n, device="cuda")
    ref = C.batch(robot, ALG, q[:512], qd[:512], u[:512])
    for name in names:
        try:
            eng = GridEngine(robot, plan=plan_for(robot, name), tag="_x" + name)
            if os.environ.get("VARIANT_FORCE"):
                eng.set_option("GRID_FORCE_KERNEL", os.environ["VARIANT_FORCE"])
            eng.forward_dynamics_gradient_device(out, x)
            torch.cuda.synchronize()
            err = float(np.abs(out[:512].cpu().numpy() - ref).max() / np.abs(ref).max())
            res = {"variant": name, "plan": {k: (dict(v) if isinstance(v, dict) else v) for k, v in VARIANTS[name].items()},
                   "relerr_vs_c_oracle": err}
            if ROBOT == "chain64":
                for N in (16384, 1024, 128):
                    us = eng.time_launches(ALG, out, x, num_timesteps=N, stride=3 * n, reps=5 if N > 1000 else 20)
                    res["us_N%d" % N] = float(np.median(us))
                res["evals_per_s_N16384"] = 16384 / res["us_N16384"] * 1e6
                print(json.dumps(res), flush=True)
                continue
            res["kind"] = eng.kernel_kind(ALG)
            for N in (65536, 8192, 128):
                us = eng.time_launches(ALG, out, x, num_timesteps=N, stride=3 * n, reps=20 if N > 1000 else 100)
                res["us_N%d" % N] = float(np.median(us))
            res["evals_per_s_N65536"] = 65536 / res["us_N65536"] * 1e6
            if os.environ.get("VARIANT_TASKS"):          # per-task times: only that program runs (GRID_PIPE_ONLY_TASK)
                from gridcodegenerator_b200.build import lib_path
                sj = lib_path(robot, "_x" + name, plan_for(robot, name))[:-3] + ".stats.json"
                st = json.load(open(sj)).get("pipe_" + ALG, {}) if os.path.exists(sj) else {}
                res["tasks"] = st.get("tasks")
                res["x2_live"] = st.get("x2_live")
                per = {}
                for stage, cnt in ((0, sum(1 for t in st.get("tasks", []) if t[1] == 0)), (1, sum(1 for t in st.get("tasks", []) if t[1] == 1))):
                    for k in range(cnt):
                        eng.set_option("GRID_PIPE_ONLY_TASK", str(100 * stage + k))
                        us = float(np.median(eng.time_launches(ALG, out, x, num_timesteps=65536, stride=3 * n, reps=5)))
                        if us < 3.0:
                            break
                        per["s%d_t%d" % (stage, k)] = us
                eng.set_option("GRID_PIPE_ONLY_TASK", None)
                res["task_us_N65536"] = per
            if os.environ.get("VARIANT_ORDER"):          # stage-1 item order: chunk-major with chunks of this many states
                for oc in (0, 8192, 16384, 32768):
                    eng.set_option("GRID_PIPE_ORDER_CHUNK", str(oc))
                    for N in (65536, 262144):
                        if N > NMAX:
                            continue
                        us = eng.time_launches(ALG, out, x, num_timesteps=N, stride=3 * n, reps=20)
                        res["us_N%d_order%d" % (N, oc)] = float(np.median(us))
                    eng.forward_dynamics_gradient_device(out, x)
                    torch.cuda.synchronize()
                    res["relerr_order%d" % oc] = float(np.abs(out[:512].cpu().numpy() - ref).max() / np.abs(ref).max())
                    chk = out[-4096:].double().sum().item()
                    res["tail_checksum_order%d" % oc] = chk
                eng.set_option("GRID_PIPE_ORDER_CHUNK", None)
            if os.environ.get("VARIANT_QUICK"):
                print(json.dumps(res), flush=True)
                continue
            for w in (8, 4, 2):                     # CTA width pinned (GRID_PIPE_WARPS) at the strong-scaling shard size
                eng.set_option("GRID_PIPE_WARPS", str(w))
                res["us_N8192_w%d" % w] = float(np.median(eng.time_launches(ALG, out, x, num_timesteps=8192, stride=3 * n, reps=20)))
                res["us_N16384_w%d" % w] = float(np.median(eng.time_launches(ALG, out, x, num_timesteps=16384, stride=3 * n, reps=20)))
            eng.set_option("GRID_PIPE_WARPS", None)
            for ns in (2000, 10000, 40000):          # de-phased instruction streams: CTAs of one SM start ns apart
                eng.set_option("GRID_PIPE_STAGGER_NS", str(ns))
                res["us_N65536_stagger%d" % ns] = float(np.median(eng.time_launches(ALG, out, x, num_timesteps=65536, stride=3 * n, reps=10)))
            eng.set_option("GRID_PIPE_STAGGER_NS", None)
            print(json.dumps(res), flush=True)
        except Exception as e:
            print(json.dumps({"variant": name, "error": str(e)[:300]}), flush=True)


if __name__ == "__main__":
    mode, names = sys.argv[1], sys.argv[2:] or list(VARIANTS)
    (build if mode == "build" else run)(names)
