"""GPU parity of the consumers fused after the FD gradient and of the round-2 host
features: CUDA-graph entry, several devices in one process, GridData lifetime.

The consumers have no reference counterpart (the reference stops at df_du,
algorithms/_forward_dynamics_gradient.py:159-161); the oracle composes the reference-pinned pieces
(fd, minv, fd_grad) with the explicit-Euler algebra of include/grid_b200.h
(oracle/rbd_numpy.compose_consumer)."""
import gc

import numpy as np
import pytest

from helpers import relerr

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from gridcodegenerator_b200 import load_named_robot                      # noqa: E402
from gridcodegenerator_b200.runtime import GridEngine, GridError, get_engine        # noqa: E402
from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u, seed_for  # noqa: E402
from oracle import c_oracle as C                                        # noqa: E402
from oracle import rbd_numpy as O                                       # noqa: E402

TOL_CONSUMER = 1e-3            # same bar as the gradients they are built from (BASELINE.json north_star)
DT = 0.0125


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def make_lambda(n, N, seed):
    return np.random.default_rng(seed).uniform(-3.0, 3.0, (N, 2 * n)).astype(np.float32)


def run_consumer(eng, alg, q, qd, u, lam, dt=DT):
    n, N = eng.n, q.shape[0]
    x = dev(pack_q_qd_u(q, qd, u))
    if alg == "fd_vjp":
        out = torch.full((N, 5 * n), float("nan"), device="cuda")
        eng.forward_dynamics_gradient_vjp_device(out, x, dev(lam), dt)
    else:
        out = torch.full((N, 2 * n + 3 * n * n), float("nan"), device="cuda")
        eng.forward_dynamics_linearize_device(out, x, dt)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def blockwise_relerr(alg, n, out, ref):
    """max|x - ref| / max|ref| per logical block (x+, A^T lam, B^T lam | x+, A21, A22, B2): a small block
    must not hide behind a large one."""
    cuts = [0, 2 * n, 4 * n, 5 * n] if alg == "fd_vjp" else [0, 2 * n, 2 * n + n * n, 2 * n + 2 * n * n, 2 * n + 3 * n * n]
    return max(relerr(out[:, a:b], ref[:, a:b]) for a, b in zip(cuts, cuts[1:]))


@pytest.mark.parametrize("name,N", [("iiwa14", 256), ("hyq", 256), ("atlas", 96), ("mixed5", 256)])
@pytest.mark.parametrize("alg", ["fd_vjp", "fd_lin"])
def test_consumers_against_numpy_oracle(name, N, alg):
    robot = load_named_robot(name)
    eng = get_engine(robot)
    assert eng.kernel_kind(alg) != "none"
    q, qd, u, _ = make_states(robot.n, N, seed_for(name) + 7)
    lam = make_lambda(robot.n, N, 3)
    out = run_consumer(eng, alg, q, qd, u, lam)
    ref = O.consumer_batch(robot, alg, q.astype(np.float64), qd.astype(np.float64), u.astype(np.float64), DT,
                           lam.astype(np.float64))
    assert np.isfinite(out).all()
    assert blockwise_relerr(alg, robot.n, out, ref) < TOL_CONSUMER, blockwise_relerr(alg, robot.n, out, ref)


@pytest.mark.parametrize("name,N", [("iiwa14", 65536), ("hyq", 16384), ("atlas", 8192)])
@pytest.mark.parametrize("alg", ["fd_vjp", "fd_lin"])
def test_consumers_full_batches_against_c_oracle(name, N, alg):
    """BASELINE batch sizes (Atlas: the 8 192 states one GPU owns of 65 536 over 8), every state checked."""
    robot = load_named_robot(name)
    eng = get_engine(robot)
    q, qd, u, _ = make_states(robot.n, N, seed_for(name) + 8)
    lam = make_lambda(robot.n, N, 4)
    out = run_consumer(eng, alg, q, qd, u, lam)
    ref = C.consumer_batch(robot, alg, q, qd, u, DT, lam.astype(np.float64))
    assert blockwise_relerr(alg, robot.n, out, ref) < TOL_CONSUMER


@pytest.mark.parametrize("family", ["tps", "pipe"])
@pytest.mark.parametrize("name", ["iiwa14", "mixed5"])
def test_consumers_split_and_single_thread_agree(name, family, monkeypatch):
    """The forced phase-split build of single-tree robots (state program + one column program per joint)
    runs the consumers through stage 0 exports (w = Minv lam_v, lam) and stage 1 column groups."""
    import __graft_entry__ as G
    robot = load_named_robot(name)
    eng = GridEngine(robot, plan=G.split_test_plan(robot), tag=G.SPLIT_TEST_TAG)
    monkeypatch.setenv("GRID_FORCE_KERNEL", family)
    N = 1000
    q, qd, u, _ = make_states(robot.n, N, 21)
    lam = make_lambda(robot.n, N, 5)
    q64, qd64, u64 = (x[:128].astype(np.float64) for x in (q, qd, u))
    for alg in ("fd_vjp", "fd_lin"):
        assert family in eng.kernel_kind(alg)
        out = run_consumer(eng, alg, q, qd, u, lam)
        ref = O.consumer_batch(robot, alg, q64, qd64, u64, DT, lam[:128].astype(np.float64))
        assert blockwise_relerr(alg, robot.n, out[:128], ref) < TOL_CONSUMER
        assert np.isfinite(out).all()


@pytest.mark.parametrize("N", [0, 1, 31, 33, 1000])
@pytest.mark.parametrize("name", ["iiwa14", "atlas"])
def test_consumers_ragged_batches_and_guard_rows(name, N):
    robot = load_named_robot(name)
    eng = get_engine(robot)
    n = robot.n
    q, qd, u, _ = make_states(n, 1000, 6)
    lam = make_lambda(n, 1000, 6)
    big = run_consumer(eng, "fd_vjp", q, qd, u, lam)
    guard = torch.full((N + 2, 5 * n), 7.0, device="cuda")
    x = dev(pack_q_qd_u(q[:N], qd[:N], u[:N])) if N else torch.empty(0, 3 * n, device="cuda")
    l = dev(lam[:N]) if N else torch.empty(0, 2 * n, device="cuda")
    eng.forward_dynamics_gradient_vjp_device(guard[1:N + 1], x, l, DT, num_timesteps=N, stride=3 * n)
    torch.cuda.synchronize()
    g = guard.cpu().numpy()
    assert np.all(g[0] == 7.0) and np.all(g[-1] == 7.0)
    assert np.array_equal(g[1:N + 1], big[:N])


def test_vjp_is_the_adjoint_of_the_linearisation():
    """Size-independent property at the BASELINE batch: <lam, A dx + B du> == <A^T lam, dx> + <B^T lam, du>
    with A, B assembled from the fd_lin output and A^T lam, B^T lam taken from the fd_vjp output."""
    robot = load_named_robot("iiwa14")
    eng = get_engine(robot)
    n, N = robot.n, 65536
    q, qd, u, _ = make_states(n, N, 77)
    lam = make_lambda(n, N, 78)
    vjp = run_consumer(eng, "fd_vjp", q, qd, u, lam).astype(np.float64)
    lin = run_consumer(eng, "fd_lin", q, qd, u, lam).astype(np.float64)
    assert np.array_equal(vjp[:, :2 * n], lin[:, :2 * n].astype(np.float64))          # same x+
    A21 = lin[:, 2 * n:2 * n + n * n].reshape(N, n, n).transpose(0, 2, 1)
    A22 = lin[:, 2 * n + n * n:2 * n + 2 * n * n].reshape(N, n, n).transpose(0, 2, 1)
    B2 = lin[:, 2 * n + 2 * n * n:].reshape(N, n, n).transpose(0, 2, 1)
    lq, lv = lam[:, :n].astype(np.float64), lam[:, n:].astype(np.float64)
    ATl_q = lq + np.einsum("sij,si->sj", A21, lv)
    ATl_v = DT * lq + np.einsum("sij,si->sj", A22, lv)
    BTl = np.einsum("sij,si->sj", B2, lv)
    assert relerr(vjp[:, 2 * n:3 * n], ATl_q) < 1e-4
    assert relerr(vjp[:, 3 * n:4 * n], ATl_v) < 1e-4
    assert relerr(vjp[:, 4 * n:], BTl) < 1e-4
    assert np.abs(B2 - B2.transpose(0, 2, 1)).max() == 0.0                              # B2 full symmetric


def test_consumer_host_path_returns_order_n_words():
    """grid_forward_dynamics_gradient_vjp(grid_data*): pinned host in -> H2D -> fused kernel -> D2H of 5n words."""
    robot = load_named_robot("iiwa14")
    eng = get_engine(robot)
    n, N = robot.n, 20000
    q, qd, u, _ = make_states(n, N, 19)
    lam = make_lambda(n, N, 20)
    data = eng.make_data(N)
    data.h["q_qd_u"][:] = pack_q_qd_u(q, qd, u)
    data.consumer_buffers()["lambda"][:] = lam
    host = data.forward_dynamics_gradient_vjp(N, DT).copy()
    assert np.array_equal(host, run_consumer(eng, "fd_vjp", q, qd, u, lam))
    host_lin = data.forward_dynamics_linearize(N, DT).copy()
    assert np.array_equal(host_lin, run_consumer(eng, "fd_lin", q, qd, u, lam))
    data.close()


def test_consumers_unavailable_robot_fails_loudly():
    robot = load_named_robot("chain64")
    eng = get_engine(robot)
    if eng.kernel_kind("fd_vjp") != "none":
        pytest.skip("chain64 has a consumer kernel")
    n = robot.n
    with pytest.raises(GridError, match="consumers"):
        eng.forward_dynamics_gradient_vjp_device(torch.empty(4, 5 * n, device="cuda"), torch.zeros(4, 3 * n, device="cuda"),
                                                 torch.zeros(4, 2 * n, device="cuda"), DT)


# ---- CUDA graph entry -----------------------------------------------------------------------------
@pytest.mark.parametrize("name,alg", [("iiwa14", "fd_grad"), ("atlas", "fd_grad"), ("hyq", "fd"), ("iiwa14", "fd_vjp")])
def test_graph_replay_matches_eager(name, alg):
    robot = load_named_robot(name)
    eng = get_engine(robot)
    n, N = robot.n, 128
    q, qd, u, _ = make_states(n, N, 55)
    x = dev(pack_q_qd_u(q, qd, u))
    lam = dev(make_lambda(n, N, 56)) if alg == "fd_vjp" else None
    words = {"fd_grad": 2 * n * n, "fd": n, "fd_vjp": 5 * n}[alg]
    eager = torch.empty(N, words, device="cuda")
    if alg == "fd_grad":
        eng.forward_dynamics_gradient_device(eager, x)
    elif alg == "fd":
        eng.forward_dynamics_device(eager, x)
    else:
        eng.forward_dynamics_gradient_vjp_device(eager, x, lam, DT)
    out = torch.zeros(N, words, device="cuda")
    graph = eng.make_graph(alg, out, x, in1=lam, dt=DT)
    out.zero_()
    s = torch.cuda.Stream()
    for _ in range(3):
        graph.launch(s)
    torch.cuda.synchronize()
    assert torch.equal(out, eager)
    # new inputs in the same buffers are picked up by the replay
    x.copy_(dev(pack_q_qd_u(q[::-1].copy(), qd[::-1].copy(), u[::-1].copy())))
    graph.launch(s)
    torch.cuda.synchronize()
    assert not torch.equal(out, eager)
    graph.close()
    us = eng.time_launches(alg.replace("fd_vjp", "fd_grad") + "@graph", torch.empty(N, 2 * n * n, device="cuda"), x, reps=20)
    assert us.shape == (20,) and (us > 0).all()


# ---- several devices in one process (ADVICE r1 / VERDICT weak #8) -----------------------------------
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("name,alg", [("atlas", "fd_grad"), ("chain64", "id_grad"), ("iiwa14", "fd_grad")])
def test_two_devices_one_process(name, alg):
    """Kernels above 48 KB of shared memory need their opt-in on EVERY device they run on."""
    robot = load_named_robot(name)
    eng = get_engine(robot)
    n, N = robot.n, 96
    q, qd, u, _ = make_states(n, N, 61)
    outs = []
    for d in (1, 0, 1):
        with torch.cuda.device(d):
            x = torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda(d)
            o = torch.empty(N, 2 * n * n, device="cuda:%d" % d)
            (eng.forward_dynamics_gradient_device if alg == "fd_grad" else eng.inverse_dynamics_gradient_device)(o, x)
            torch.cuda.synchronize(d)
            outs.append(o.cpu().numpy())
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    with torch.cuda.device(0):
        with pytest.raises(GridError, match="current device"):
            eng.forward_dynamics_gradient_device(torch.empty(N, 2 * n * n, device="cuda:1"),
                                                 torch.zeros(N, 3 * n, device="cuda:1"))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_grid_data_is_bound_to_its_device():
    eng = get_engine(load_named_robot("iiwa14"))
    with torch.cuda.device(1):
        data = eng.make_data(64)
        data.forward_dynamics(64)
    with torch.cuda.device(0):
        with pytest.raises(GridError, match="another device"):
            data.forward_dynamics(64)
    data.close()


# ---- GridData lifetime (views must not dangle) --------------------------------------------
def test_results_outlive_their_grid_data():
    robot = load_named_robot("iiwa14")
    eng = get_engine(robot)
    n, N = robot.n, 512
    q, qd, u, _ = make_states(n, N, 23)

    def temp():
        d = eng.make_data(N)
        d.h["q_qd_u"][:] = pack_q_qd_u(q, qd, u)
        return d.forward_dynamics(N)              # the GridData object dies here

    res = temp()
    gc.collect()
    junk = [eng.make_data(N) for _ in range(4)]    # would reuse the freed pinned pages
    for j in junk:
        j.h["qdd"][:] = -1.0
    keep = res.copy()
    d2 = eng.make_data(N)
    d2.h["q_qd_u"][:] = pack_q_qd_u(q, qd, u)
    assert np.array_equal(keep, d2.forward_dynamics(N))
    assert np.array_equal(res, keep)               # still the values, not the junk
    view = d2.h["qdd"]
    d2.close()
    assert np.array_equal(view[:N], keep)          # close() never frees under a live view
    with pytest.raises(GridError, match="closed"):
        d2.forward_dynamics(N)


# ---- consumers on the chain kernels (csrc/grid_lps.cuh) -----------------------------------------------
@pytest.mark.parametrize("name", ["iiwa14", "pchain4"])
@pytest.mark.parametrize("alg", ["fd_vjp", "fd_lin"])
def test_consumers_on_chain_kernels_small_chains(name, alg, monkeypatch):
    """The chain kernels' fused consumers (second articulated-body solve w = Minv lam_v in stage A, the dot with the
    dc_du column in the column kernel; A21 / A22 columns and the mirrored B2 = dt Minv) forced onto small chains with
    damping and prismatic joints: against the composed oracle, ragged last tile, chunked launch."""
    import __graft_entry__ as G
    robot = load_named_robot(name)
    eng = GridEngine(robot, plan=G.lps_test_plan(robot), tag=G.LPS_TEST_TAG)
    assert "lps" in eng.kernel_kind(alg)
    monkeypatch.setenv("GRID_FORCE_KERNEL", "lps")
    N = 4500
    q, qd, u, _ = make_states(robot.n, N, seed_for(name) + 11)
    lam = make_lambda(robot.n, N, 12)
    out = run_consumer(eng, alg, q, qd, u, lam)
    assert np.isfinite(out).all()
    M = 128
    ref = O.consumer_batch(robot, alg, q[:M].astype(np.float64), qd[:M].astype(np.float64), u[:M].astype(np.float64), DT,
                           lam[:M].astype(np.float64))
    assert blockwise_relerr(alg, robot.n, out[:M], ref) < TOL_CONSUMER, blockwise_relerr(alg, robot.n, out[:M], ref)
    tail = O.consumer_batch(robot, alg, q[-3:].astype(np.float64), qd[-3:].astype(np.float64), u[-3:].astype(np.float64), DT,
                            lam[-3:].astype(np.float64))
    assert blockwise_relerr(alg, robot.n, out[-3:], tail) < TOL_CONSUMER
    # the same values as the thread-per-state consumer programs of the default dispatch
    monkeypatch.delenv("GRID_FORCE_KERNEL")
    assert relerr(run_consumer(eng, alg, q[:M], qd[:M], u[:M], lam[:M]), out[:M]) < 1e-4


@pytest.mark.parametrize("alg", ["fd_vjp", "fd_lin"])
def test_consumers_chain64_against_c_oracle(alg):
    """64-link chain: the fused consumers against the C oracle's pieces (fd, minv, fd_grad) composed on the host."""
    robot = load_named_robot("chain64")
    eng = get_engine(robot)
    assert eng.kernel_kind(alg) == "lps"
    N = 1024
    q, qd, u, _ = make_states(robot.n, N, seed_for("chain64") + 13)
    lam = make_lambda(robot.n, N, 14)
    out = run_consumer(eng, alg, q, qd, u, lam)
    ref = C.consumer_batch(robot, alg, q, qd, u, DT, lam.astype(np.float64))
    assert blockwise_relerr(alg, robot.n, out, ref) < TOL_CONSUMER, blockwise_relerr(alg, robot.n, out, ref)
