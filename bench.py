#!/usr/bin/env python3
"""Benchmark: forward-dynamics-gradient evals/s, iiwa14, 65,536 states per GPU (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, one process per GPU)
  python bench.py --impl reference [...]                         reference CPU path (numpy oracle port)

One "step" = one pass of the hot path over one batch of synthetic states = `launches_per_step`
back-to-back launches of the N-state kernel over rotating buffer sets (a single 65 536-state iiwa14
launch lasts 35 us: K = 20 of those would time 0.7 ms, which is jitter under max-over-ranks).
Prints ONE JSON line on rank 0.  See DESIGN.md section "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROBOT = "iiwa14"
ALG = "fd_grad"
BATCH = 65536
GRAVITY = 9.81
METRIC = "fd_grad_evals_per_s_iiwa14_N65536"
UNIT = "evals/s"
FP32_THEORETICAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12       # 74.4, SURVEY.md 8d


def launches_per_step(robot, alg, N):
    """Launches that make one step last ~2.5 ms at 40 algorithmic TFLOP/s: a pure function of the
    configuration, so that both arms print the same `config`."""
    from gridcodegenerator_b200.algorithms import algorithmic_flops
    est_s = algorithmic_flops(robot)[alg] * N / 40e12
    return int(max(1, min(256, np.ceil(2.5e-3 / est_s))))


def make_config(robot_name, robot, alg, N):
    """`config` of the JSON line - identical in our arm and in --impl reference (same workload)."""
    n = robot.n
    out_words = {"id": n, "minv": n * n, "fd": n, "id_grad": 2 * n * n, "fd_grad": 2 * n * n}[alg]
    set_bytes = 4 * N * (3 * n + out_words)
    nsets = max(2, int(np.ceil(300e6 / set_bytes)))
    L = launches_per_step(robot, alg, N)
    return {"workload": "%s %s (forward-dynamics gradient), %d states per GPU per launch, state-major [q|qd|u]" % (
                robot_name, alg, N),
            "states_per_gpu": N, "robot_hash": robot.param_hash(), "launches_per_step": L,
            "states_per_step_per_gpu": N * L,
            "l2": "inputs/outputs rotate over %d buffer sets (%.0f MB > 126 MB L2)" % (nsets, nsets * set_bytes / 1e6)}, nsets, L


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--robot", default=ROBOT)
    ap.add_argument("--alg", default=ALG)
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true",
                    help="minimal run for ncu: W+K launches only, no extended warm-up, e2e, latency or CPU legs")
    ap.add_argument("--cpu-sample", type=int, default=0, help="states in the CPU baseline sample (0 = auto)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# CPU reference path: the oracle port of the reference's numpy implementation (_test.py)
# ---------------------------------------------------------------------------------------------
def _cpu_worker(args):
    robot_name, alg, seed, count = args
    from gridcodegenerator_b200 import load_named_robot
    from gridcodegenerator_b200.synthetic import make_states
    from oracle import rbd_numpy as O
    robot = load_named_robot(robot_name)
    q, qd, u, _ = (x.astype(np.float64) for x in make_states(robot.n, count, seed))
    t0 = time.perf_counter()
    O.batch(robot, alg, q, qd, u if alg in ("fd", "fd_grad") else None, GRAVITY)
    return time.perf_counter() - t0


def cpu_reference_pass(robot_name, alg, total_states, cores, pool):
    """One bounded sample of the workload over `cores` processes; returns (evals/s, single-core ms/eval)."""
    per = max(1, total_states // cores)
    jobs = [(robot_name, alg, 1000 + i, per) for i in range(cores)]
    t0 = time.perf_counter()
    times = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    return per * cores / wall, 1e3 * float(np.mean(times)) / per, per * cores


def run_reference_arm(a):
    """The reference's CPU path (numpy port of _test.py) on all host cores.  A timed step is the FULL
    batch of the configuration when that fits ~8 s (iiwa14, 65 536 states: ~6 s on 16 cores), else the
    largest whole multiple of the core count that does; warm-up steps are small samples (a Python loop
    has nothing to warm beyond imports and the fork pool)."""
    import multiprocessing as mp
    from gridcodegenerator_b200 import load_named_robot
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    robot = load_named_robot(a.robot)
    config, _, _ = make_config(a.robot, robot, a.alg, a.batch)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    with mp.get_context("fork").Pool(cores) as pool:
        rate = 0.0
        for w in range(max(1, min(a.warmup, 2))):          # the first pass also pays imports and the robot load
            rate, _, _ = cpu_reference_pass(a.robot, a.alg, cores * (8 if w == 0 else 64), cores, pool)
        budget_s = 10.0
        sample = a.cpu_sample or int(min(a.batch, max(cores, rate * budget_s)))
        if sample >= a.batch * 0.8:
            sample = a.batch
        vals, t0, done, ms1 = [], time.perf_counter(), 0, 0.0
        for _ in range(a.steps):
            v, ms1, done = cpu_reference_pass(a.robot, a.alg, sample, cores, pool)
            vals.append(v)
        wall = time.perf_counter() - t0
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC if (a.robot, a.alg, a.batch) == (ROBOT, ALG, BATCH) else
        "%s_evals_per_s_%s_N%d" % (a.alg, a.robot, a.batch),
        "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * wall / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%s: %d seeded states per step over %d processes, oracle/rbd_numpy.py (port of "
                                   "reference _test.py:496-520); single-core %.2f ms/eval (SURVEY: 9.74 ms for the "
                                   "reference's own _test.py on the survey box)" % (
                                       "the full batch" if done > a.batch - cores else "bounded sample", done, cores, ms1)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def reference_gpu_arm(robot_name, alg, N, host_in):
    """The reference's OWN emitted CUDA (baseline/_ref/<robot>/ref_harness, built from the unmodified
    reference by baseline/make_reference_cuh.py) timed on this GPU on the same inputs, _compute_only
    mode.  Reported beside our numbers; absent when the harness was not built."""
    import tempfile
    exe = os.path.join(ROOT, "baseline", "_ref", robot_name, "ref_harness")
    if alg != "fd_grad" or not os.path.exists(exe):
        return {"unavailable": "baseline/_ref/%s/ref_harness not built (python baseline/make_reference_cuh.py)" % robot_name}
    out = {}
    with tempfile.TemporaryDirectory() as td:
        for n_states, key in ((N, "n_full"), (128, "n128")):
            with open(os.path.join(td, "in.bin"), "wb") as f:
                f.write(np.int32(n_states).tobytes())
                f.write(np.ascontiguousarray(host_in[:n_states]).tobytes())
            best = None
            for threads in (128, 256):
                p = subprocess.run([exe, os.path.join(td, "in.bin"), os.path.join(td, "out.bin"), str(threads), "20"],
                                   capture_output=True, text=True, timeout=300)
                if p.returncode == 0:
                    r = json.loads(p.stdout.strip().splitlines()[-1])
                    if best is None or r["p50_us"] < best["p50_us"]:
                        best = r
            if best:
                out[key] = {"N": n_states, "threads": best["threads"], "p50_us": best["p50_us"],
                            "evals_per_s": best["evals_per_s"]}
    out["what"] = ("reference GRiDCodeGenerator's emitted forward_dynamics_gradient_compute_only kernel, nvcc sm_100a, "
                   "best of 128/256 threads per block, CUDA events around the call")
    return out


def other_baseline_configs():
    """The other BASELINE.json configurations, timed in the same run (C-side event pairs, 20 launches each,
    buffers reused: L2-resident outputs) - context for the headline line, not part of its timed region."""
    import torch
    from gridcodegenerator_b200 import load_named_robot
    from gridcodegenerator_b200.algorithms import algorithmic_flops
    from gridcodegenerator_b200.runtime import get_engine
    from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u, seed_for
    out = {}
    cases = [("cfg3_hyq_N16384", "hyq", ("id", "minv", "fd"), 16384),
             ("cfg4_atlas_N65536", "atlas", ("fd_grad",), 65536),
             ("cfg4_atlas_N8192_per_gpu_of_8", "atlas", ("fd_grad",), 8192),
             ("cfg5_chain64_N65536", "chain64", ("id",), 65536),
             ("cfg5_chain64_N16384", "chain64", ("minv", "fd", "id_grad", "fd_grad"), 16384),
             ("cfg5_chain64_N128", "chain64", ("fd_grad",), 128)]
    for key, name, algs, N in cases:
        try:
            robot = load_named_robot(name)
            eng = get_engine(robot)
            n = robot.n
            q, qd, u, _ = make_states(n, N, seed_for(name))
            x = torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda()
            res = {}
            for alg in algs:
                words = {"id": n, "minv": n * n, "fd": n, "id_grad": 2 * n * n, "fd_grad": 2 * n * n}[alg]
                o = torch.empty(N, words, device="cuda")
                us = eng.time_launches(alg, o, x, reps=20)
                p50 = float(np.median(us))
                res[alg] = {"p50_us": p50, "evals_per_s": N / p50 * 1e6, "kernel": eng.kernel_kind(alg),
                            "algorithmic_tflops": algorithmic_flops(robot)[alg] * N / p50 / 1e6}
                del o
            out[key] = res
            del x
            torch.cuda.empty_cache()
        except Exception as e:                       # never let the context runs break the headline line
            out[key] = {"unavailable": str(e)[:200]}
    return out


# ---------------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t_lo, t_hi):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t_lo <= t <= t_hi] or [r for (_, r) in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_b200_arm(a):
    import torch
    import torch.distributed as dist
    from gridcodegenerator_b200 import load_named_robot
    from gridcodegenerator_b200.algorithms import algorithmic_bytes, algorithmic_flops
    from gridcodegenerator_b200.hostmem import bind_to_gpu_numa_node
    from gridcodegenerator_b200.runtime import get_engine
    from gridcodegenerator_b200.sharding import max_over_ranks, shard_range
    from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u, seed_for

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    all_cpus = sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else []
    # pinned transfer buffers on the GPU's own NUMA node (gridcodegenerator_b200/hostmem.py)
    numa = bind_to_gpu_numa_node(local) if world > 1 else {"bound": False, "why": "single rank: not bound"}
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    robot = load_named_robot(a.robot)
    eng = get_engine(robot)
    n, N = robot.n, a.batch
    in_words, out_words = 3 * n, {"id": n, "minv": n * n, "fd": n, "id_grad": 2 * n * n, "fd_grad": 2 * n * n}[a.alg]
    config, nsets, L = make_config(a.robot, robot, a.alg, N)

    # every rank owns its own contiguous slice of the global batch (weak scaling: BATCH states per GPU)
    q, qd, u, _ = make_states(n, N, seed_for(a.robot) + 100 * rank)
    host_in = pack_q_qd_u(q, qd, u)
    # rotate over enough buffer sets that the working set exceeds the 126 MB L2
    ins = [torch.from_numpy(host_in).cuda() for _ in range(nsets)]
    outs = [torch.empty(N, out_words, device="cuda") for _ in range(nsets)]
    stream = torch.cuda.current_stream()
    device_call = {"fd_grad": eng.forward_dynamics_gradient_device, "id_grad": eng.inverse_dynamics_gradient_device,
                   "fd": eng.forward_dynamics_device, "minv": eng.direct_minv_device,
                   "id": eng.inverse_dynamics_device}[a.alg]

    def launch(i, T=N):
        k = i % nsets
        device_call(outs[k], ins[k], num_timesteps=T, stride=in_words, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # warm-up: W steps, then keep the GPU busy ~0.4 s so clocks settle and the sampler sees load
    for i in range(a.warmup * (1 if a.profile else L)):
        launch(i)
    torch.cuda.synchronize()
    t_warm = time.perf_counter()
    i = a.warmup * L
    while not a.profile and time.perf_counter() - t_warm < 0.4:
        for _ in range(50):
            launch(i)
            i += 1
        torch.cuda.synchronize()

    # ---- timed region: exactly K steps of L launches, CUDA events on the launching stream ----------
    if a.profile:
        L = 1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches_before = eng.launch_count()
    t_lo = time.perf_counter()
    e0.record(stream)
    for i in range(a.steps * L):
        launch(i)
    e1.record(stream)
    barrier()
    t_hi = time.perf_counter()
    gpu_launches = eng.launch_count() - launches_before
    ms_total = e0.elapsed_time(e1)
    ms_total = max_over_ranks(ms_total, world, device="cuda")
    ms_per_step = ms_total / a.steps
    ms_per_launch = ms_per_step / L
    value = world * N * L / (ms_per_step * 1e-3)
    clocks = sampler.stop(t_lo - 0.3, t_hi + 0.05) if rank == 0 else None

    if a.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": ms_per_step, "gpu_launches": int(gpu_launches)}))
        return

    # ---- strong scaling: the SAME N states sharded over the ranks (north_star: "batch 65536 sharded
    # across 1/2/4/8").  Under torchrun every rank times its own shard (max over ranks); a single rank
    # times the shard sizes of 2/4/8 GPUs on its own GPU - ranks never communicate on the data path, so
    # that is what each GPU of a larger job would run.
    def time_shard(T, reps=40):
        for i in range(5):
            launch(i, T)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        s0.record(stream)
        for i in range(reps):
            launch(i, T)
        s1.record(stream)
        barrier()
        return max_over_ranks(s0.elapsed_time(s1) / reps * 1e3, world, device="cuda")

    strong = {"total_states": N, "t1_us": ms_per_launch * 1e3,
              "what": "one launch over the whole batch on one GPU (t1) vs one launch over each GPU's contiguous shard "
                      "(max over ranks); no collective on the data path"}
    if world > 1:
        first, last = shard_range(N, rank, world)
        tN = time_shard(last - first)
        strong.update(measured_gpus=world, shard_states=N // world, tN_us=tN, speedup=ms_per_launch * 1e3 / tN)
    else:
        proj = {}
        for g in (2, 4, 8):
            tN = time_shard(N // g)
            proj[str(g)] = {"shard_states": N // g, "tN_us": tN, "speedup": ms_per_launch * 1e3 / tN}
        strong["single_gpu_projection"] = proj

    # ---- end to end through the host API (pinned host buffers, H2D + kernel + D2H per step) --
    data = eng.make_data(N)
    data.h["q_qd_u"][:] = host_in
    host_call = {"fd_grad": data.forward_dynamics_gradient, "id_grad": data.inverse_dynamics_gradient,
                 "fd": data.forward_dynamics, "minv": data.direct_minv, "id": data.inverse_dynamics}[a.alg]
    e2e_steps = max(3, min(a.steps, 20))

    def time_host(call):
        for _ in range(3):
            call()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            res = call()
            _ = float(res[0, 0])                 # the result is in host memory when the call returns
        torch.cuda.synchronize()
        return world * N * e2e_steps / max_over_ranks(time.perf_counter() - t0, world, device="cuda")

    e2e_value = time_host(lambda: host_call(N))
    e2e_consumer = None
    if a.alg == "fd_grad" and eng.kernel_kind("fd_vjp") != "none":
        data.consumer_buffers()["lambda"][:] = np.random.default_rng(5).uniform(-1, 1, (N, 2 * n)).astype(np.float32)
        v = time_host(lambda: data.forward_dynamics_gradient_vjp(N, 0.01))
        pcie_floor_s = 4.0 * N * 5 * n / 55e9           # inputs alone, one direction, ~55 GB/s PCIe Gen5 x16
        e2e_consumer = {
            "value": v, "unit": UNIT, "h2d_bytes_per_step": 4 * N * 5 * n, "d2h_bytes_per_step": 4 * N * 5 * n,
            "what": "grid_forward_dynamics_gradient_vjp(grid_data*): the same FD gradient consumed on the device as "
                    "[x+ | A^T lam | B^T lam] (5n words per state back instead of 2n^2); pinned host in -> H2D -> fused "
                    "kernel -> D2H",
            "pcie_bound_evals_per_s": world * N / pcie_floor_s,
            "pcie_bound_note": "H2D of [q|qd|u|lam] alone at ~55 GB/s: no host-buffer API can be faster than this"}
    data.close()

    # ---- N=128 latency (second half of the BASELINE metric), 1 GPU only ----------------------
    lat = None
    if rank == 0:
        small_in, small_out = ins[0][:128], outs[0][:128]
        kw = dict(num_timesteps=128, stride=3 * n, reps=500)
        us = eng.time_launches(a.alg, small_out, small_in, **kw)
        floor = eng.time_launches("noop", small_out, small_in, **kw)
        gus = eng.time_launches(a.alg + "@graph", small_out, small_in, **kw)
        gfloor = eng.time_launches("noop@graph", small_out, small_in, **kw)
        p50 = lambda x: float(np.percentile(x, 50))
        lat = {"p50_us": p50(us), "p90_us": float(np.percentile(us, 90)),
               "min_us": float(us.min()), "kernel": eng.kernel_kind(a.alg),
               "floor_p50_us": p50(floor), "p50_minus_floor_us": p50(us) - p50(floor),
               "graph_p50_us": p50(gus), "graph_floor_p50_us": p50(gfloor),
               "graph_p50_minus_floor_us": p50(gus) - p50(gfloor),
               "floor_what": "same event-pair method around an empty kernel (eager / replayed from a CUDA graph): "
                             "what the measurement itself costs",
               "what": "%s N=128, one CUDA event pair per launch recorded in C (grid_time_launches), 500 launches "
                       "queued back to back; graph_* = the launch captured once (grid_graph_create) and replayed" % a.alg}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline ---------------------------------------------------------------------------
    alg_flops = algorithmic_flops(robot)[a.alg]
    alg_bytes = algorithmic_bytes(robot)[a.alg]
    fp32_peak = eng.measure_fp32_tflops(5)
    kernel_s = ms_per_launch * 1e-3                    # one launch of the dominant kernel (phase-split: its two stages)
    achieved = alg_flops * N / kernel_s / 1e12
    traced = eng.traced_flops(a.alg)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic, traffic_src, executed_ncu = None, None, None
    for rnd in ("r2", "r1_final"):
        fn = {("iiwa14", "fd_grad", 65536): "%s_tps_fdgrad_traffic.json", ("atlas", "fd_grad", 65536):
              "%s_pipe_fdgrad_atlas_traffic.json"}.get((a.robot, a.alg, a.batch))
        if not fn or traffic is not None:
            continue
        try:        # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the ncu capture named in `source`
            tj = json.load(open(os.path.join(ROOT, "profiles", fn % rnd)))
            traffic, traffic_src, executed_ncu = tj["traffic"], tj["source"], tj.get("executed_fp32")
        except (OSError, ValueError, KeyError):
            pass
    roofline = {
        "bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
        "traffic": traffic, "traffic_source": traffic_src,
        "note": "FP32 SIMT (non-tensor) roofline; achieved = algorithmic flops/state (dense reference count, "
                "SURVEY 8d) x states / launch time; peak = FFMA microbenchmark measured in this run "
                "(theoretical %.1f)" % FP32_THEORETICAL_TFLOPS,
        "algorithmic_flops_per_state": alg_flops, "traced_flops_per_state": traced,
        "executed_tflops": traced * N / kernel_s / 1e12, "executed_frac": traced * N / kernel_s / 1e12 / fp32_peak,
        "executed_ncu": executed_ncu,
        "hbm": {"bound": "hbm", "achieved": alg_bytes * N / kernel_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": alg_bytes * N / kernel_s / 1e9 / hbm_peak,
                "peak_source": "measured" if peaks else "fallback", "algorithmic_bytes_per_state": alg_bytes},
    }

    line = {
        "metric": METRIC if (a.robot, a.alg, a.batch) == (ROBOT, ALG, BATCH) else "%s_evals_per_s_%s_N%d" % (a.alg, a.robot, N),
        "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "ms_per_launch": ms_per_launch, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
        "kernel": eng.kernel_kind(a.alg), "clocks": clocks, "gpu_launches": int(gpu_launches),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * N * in_words,
                "d2h_bytes_per_step": 4 * N * out_words, "steps": e2e_steps,
                "what": "grid_forward_dynamics_gradient(grid_data*), the reference's mode-0 host contract: pinned host in "
                        "-> H2D -> kernel -> D2H of df_du (2n^2 words per state) -> pinned host out, one N-state call per "
                        "step",
                "pcie_bound_evals_per_s": world * N / (4.0 * N * out_words / 55e9),
                "pcie_bound_note": "D2H of df_du alone at ~55 GB/s (PCIe Gen5 x16 per GPU)",
                "numa": numa},
        "e2e_consumer": e2e_consumer, "strong": strong,
        "roofline": roofline, "latency_n128": lat,
    }

    line["reference_gpu"] = reference_gpu_arm(a.robot, a.alg, N, host_in)
    if world == 1 and (a.robot, a.alg, a.batch) == (ROBOT, ALG, BATCH):
        line["other_configs"] = other_baseline_configs()

    if not a.no_cpu_baseline:
        import multiprocessing as mp
        if all_cpus:
            os.sched_setaffinity(0, all_cpus)            # the CPU legs use every host core again
        cores = len(all_cpus) or os.cpu_count() or 1
        sample = a.cpu_sample or cores * 96
        with mp.get_context("fork").Pool(cores) as pool:
            cpu_reference_pass(a.robot, a.alg, cores * 4, cores, pool)
            v, ms1, done = cpu_reference_pass(a.robot, a.alg, sample, cores, pool)
        try:
            from oracle import c_oracle as C
            qs, qds, us_, _ = (x.astype(np.float64) for x in make_states(n, N, 4242))
            C.batch(robot, a.alg, qs[:256], qds[:256], us_[:256])
            t0 = time.perf_counter()
            C.batch(robot, a.alg, qs, qds, us_ if a.alg in ("fd", "fd_grad") else None, threads=cores)
            line["cpu_baseline_c"] = {"value": N / (time.perf_counter() - t0), "unit": UNIT, "cores": cores,
                                      "kind": "port", "sample": "full batch of %d states, oracle/rbd_oracle.c "
                                      "(float64 C restatement, pthreads)" % N}
        except Exception as e:                         # gcc missing on the box: the numpy baseline stands
            line["cpu_baseline_c"] = {"unavailable": str(e)[:200]}
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "%d seeded states over %d processes, oracle/rbd_numpy.py (port of reference "
                                          "_test.py:496-520); single-core %.2f ms/eval" % (done, cores, ms1)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_b200_arm(a)


if __name__ == "__main__":
    main()
