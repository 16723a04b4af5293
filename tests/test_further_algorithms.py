"""Further algorithms: mass matrix by CRBA, forward dynamics by ABA.

The reference has neither, so the pins are its OWN algorithms: tests/golden/<robot>_mass.npz holds M(q) assembled
from the reference's test_rnea (tests/golden/make_golden_mass.py), <robot>.npz its Minv and its Minv (u - c).
CPU half: oracle vs those fixtures, traced programs (interpreted in float64) vs the oracle.  GPU half: the kernels
through the C ABI vs the oracle, full-batch properties, ragged batches, forward_dynamics routed to the ABA program."""
import os

import numpy as np
import pytest

from gridcodegenerator_b200 import load_named_robot
from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u, seed_for
from helpers import GOLDEN, load_golden, relerr
from oracle import rbd_numpy as O

MASS_TAGS = ["mixed5", "iiwa14", "hyq", "atlas", "chain64"]


@pytest.mark.parametrize("name", MASS_TAGS)
def test_oracle_crba_matches_mass_matrix_from_reference_rnea(name):
    z = np.load(os.path.join(GOLDEN, name + "_mass.npz"))
    robot = load_named_robot(name)
    assert robot.param_hash() == str(z["robot_hash"]), "rerun tests/golden/make_golden_mass.py"
    for s in range(min(z["q"].shape[0], 4 if robot.n > 32 else 16)):
        M = O.crba(robot, z["q"][s])
        assert relerr(M, z["M"][s]) < 1e-11
        assert np.abs(M - M.T).max() == 0.0


@pytest.mark.parametrize("tag", ["mixed5", "iiwa14", "iiwa14_damped", "hyq", "atlas", "chain64"])
def test_oracle_crba_and_aba_against_reference_minv_and_fd(tag):
    """M Minv = I with the reference's test_minv; ABA = the reference's Minv (u - c) (damping included)."""
    robot, z = load_golden(tag)
    q, qd, u = (z[k].astype(np.float64) for k in ("q", "qd", "u"))
    n = robot.n
    for s in range(3 if n > 32 else 12):
        M = O.crba(robot, q[s])
        assert np.abs(M @ z["minv_dense"][s] - np.eye(n)).max() < 1e-9
        assert relerr(O.aba(robot, q[s], qd[s], u[s]), z["fd_qdd"][s]) < 1e-10


@pytest.mark.parametrize("name", ["mixed5", "iiwa14", "hyq", "pchain4", "atlas"])
def test_traced_programs_match_oracle(name):
    """The straight-line programs the kernels are generated from, interpreted with numpy in float64."""
    from gridcodegenerator_b200.algorithms import TRACERS
    robot = load_named_robot(name)
    n, N = robot.n, 3
    q, qd, u, _ = (x.astype(np.float64) for x in make_states(n, N, seed_for(name) + 11))
    ins = {"gravity": np.float64(9.81)}
    for i in range(n):
        ins["q%d" % i], ins["qd%d" % i], ins["u%d" % i] = q[:, i], qd[:, i], u[:, i]
    qdd = TRACERS["aba"](robot).evaluate(ins)["qdd"]
    assert relerr(qdd, O.batch(robot, "aba", q, qd, u)) < 1e-12
    assert relerr(qdd, O.batch(robot, "fd", q, qd, u)) < 1e-10
    M = TRACERS["crba"](robot).evaluate({k: v for k, v in ins.items() if k[0] == "q" and k[1] != "d"})["M"]
    assert relerr(M, O.batch(robot, "crba", q)) < 1e-12


def test_plan_lists_the_further_algorithms():
    from gridcodegenerator_b200.codegen import KernelPlan
    plan = KernelPlan(load_named_robot("iiwa14"))
    assert plan.extras == {"crba": "tps", "aba": "tps"}
    assert plan.fd_via_aba and plan.fd_aba_min_states == 32768      # FD has its own program: ABA only for big batches
    assert KernelPlan(load_named_robot("iiwa14"), fd_via_aba=False).fd_via_aba is False
    plan = KernelPlan(load_named_robot("pchain4"), only_algs=("fd",))
    assert plan.extras == {"crba": "none", "aba": "none"}


def test_launchers_route_by_batch_size():
    """Generated launchers (HyQ has thread-per-state AND phase-split kernels): Minv / FD leave the phase-split kernels
    above 32 768 states, the gradients when the output outgrows the L2, forward dynamics takes the ABA program for
    large batches unless a kernel family is forced."""
    import re
    from gridcodegenerator_b200.codegen import generate_translation_unit
    from helpers import cached_plan
    robot = load_named_robot("hyq")
    plan = cached_plan("hyq")
    assert plan.pipe_max_states["minv"] == 32768 and plan.pipe_max_states["fd_grad"] == (96 << 20) // (8 * 144)
    src, _ = generate_translation_unit(robot, plan)
    body = {m.group(1): m.group(0) for m in re.finditer(r"cudaError_t launch_(\w+)\(.*?\n}\n", src, re.S)}
    assert "use_pipe(N, 32768)" in body["minv"] and "use_pipe(N, 32768)" in body["fd"]
    assert "use_pipe(N, %d)" % plan.pipe_max_states["fd_grad"] in body["fd_grad"]
    assert "force_kernel == kAuto && N >= 32768) return launch_aba" in body["fd"]
    assert "tps_launch<AlgAba" in body["aba"] and "tps_launch<AlgCrba" in body["crba"]
    assert 'if (!strcmp(alg, "fd@large")) return "tps(aba)";' in src


# ---- GPU half -------------------------------------------------------------------------------------------
def _engine(name):
    import torch
    from gridcodegenerator_b200.runtime import get_engine
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return get_engine(load_named_robot(name)), torch


@pytest.mark.gpu
@pytest.mark.parametrize("name,N", [("iiwa14", 256), ("hyq", 256), ("mixed5", 256), ("pchain4", 256), ("atlas", 64),
                                    ("chain64", 8)])
def test_gpu_crba_and_aba_against_oracle(name, N):
    eng, torch = _engine(name)
    robot, n = eng.robot, eng.n
    q, qd, u, _ = make_states(n, N, seed_for(name) + 21)
    q64, qd64, u64 = (x.astype(np.float64) for x in (q, qd, u))
    x = torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda()
    if eng.kernel_kind("aba") != "none":
        qdd = eng.aba_device(torch.empty(N, n, device="cuda"), x)
        torch.cuda.synchronize()
        assert relerr(qdd.cpu().numpy(), O.batch(robot, "aba", q64, qd64, u64)) < 1e-4
    else:
        assert name == "never"                                  # every named robot has an ABA program
    if eng.kernel_kind("crba") != "none":
        M = eng.crba_device(torch.empty(N, n * n, device="cuda"), x)
        torch.cuda.synchronize()
        assert relerr(M.cpu().numpy(), O.batch(robot, "crba", q64)) < 1e-4
        # compressed layout: q alone, stride n
        M2 = eng.crba_device(torch.empty(N, n * n, device="cuda"), torch.from_numpy(q).cuda())
        torch.cuda.synchronize()
        assert torch.equal(M, M2)
    else:
        from gridcodegenerator_b200.runtime import GridError
        assert name == "chain64"                                # 37 k traced flops: no single-thread program
        with pytest.raises(GridError, match="CRBA"):
            eng.crba_device(torch.empty(N, n * n, device="cuda"), x)


@pytest.mark.gpu
def test_gpu_mass_matrix_goldens_from_reference_rnea():
    for name in ("mixed5", "iiwa14", "hyq", "atlas"):
        eng, torch = _engine(name)
        z = np.load(os.path.join(GOLDEN, name + "_mass.npz"))
        n, N = eng.n, z["q"].shape[0]
        M = eng.crba_device(torch.empty(N, n * n, device="cuda"), torch.from_numpy(z["q"].astype(np.float32)).cuda())
        torch.cuda.synchronize()
        ref = np.transpose(z["M"], (0, 2, 1)).reshape(N, -1)           # column-major per state
        assert relerr(M.cpu().numpy(), ref) < 1e-4, name


@pytest.mark.gpu
@pytest.mark.parametrize("name,N", [("iiwa14", 65536), ("atlas", 16384)])
def test_gpu_full_batch_properties(name, N):
    """M Minv = I and ABA = FD over a full batch, everything computed on the GPU."""
    eng, torch = _engine(name)
    n = eng.n
    q, qd, u, _ = make_states(n, N, 5)
    x = torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda()
    M = eng.crba_device(torch.empty(N, n * n, device="cuda"), x).view(N, n, n)
    Mi = eng.direct_minv_device(torch.empty(N, n * n, device="cuda"), x).view(N, n, n).transpose(1, 2)   # -> [row, col]
    Mi = torch.triu(Mi) + torch.triu(Mi, 1).transpose(1, 2)
    eye = torch.eye(n, device="cuda").expand(N, n, n)
    err = (torch.bmm(M.double(), Mi.double()) - eye).abs().amax(dim=(1, 2))
    assert float(err.max()) < 5e-3 and float(err.median()) < 2e-4
    eng.set_option("GRID_FORCE_KERNEL", None)
    a = eng.aba_device(torch.empty(N, n, device="cuda"), x)
    kinds = [k for k in ("tps", "pipe") if k in eng.kernel_kind("fd")]
    eng.set_option("GRID_FORCE_KERNEL", kinds[0])                       # the Minv-based program: a forced family never takes the ABA route
    try:
        f = eng.forward_dynamics_device(torch.empty(N, n, device="cuda"), x)
        torch.cuda.synchronize()
    finally:
        eng.set_option("GRID_FORCE_KERNEL", None)
    scale = float(f.abs().max())
    assert float((a - f).abs().max()) / scale < 2e-4


@pytest.mark.gpu
@pytest.mark.parametrize("N", [1, 31, 33, 255, 256, 1000, 4099])
def test_gpu_forward_dynamics_routed_to_aba_on_atlas(N):
    """Atlas has no single-thread Minv-based FD: batches of 256 states and more run the ABA program
    (grid_kernel_kind("fd@large")), smaller ones the phase-split kernels; same answers, guard rows intact."""
    eng, torch = _engine("atlas")
    assert eng.kernel_kind("fd@large") == "tps(aba)"
    n = eng.n
    q, qd, u, _ = make_states(n, N, 31)
    guard = torch.full((N + 2, n), 7.0, device="cuda")
    eng.forward_dynamics_device(guard[1:N + 1], torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda(), num_timesteps=N, stride=3 * n)
    torch.cuda.synchronize()
    g = guard.cpu().numpy()
    assert np.all(g[0] == 7.0) and np.all(g[-1] == 7.0)
    M = min(N, 64)
    ref = O.batch(eng.robot, "fd", *(x.astype(np.float64)[-M:] for x in (q, qd, u)))
    assert relerr(g[1:N + 1][-M:], ref) < 1e-4
