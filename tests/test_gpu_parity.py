"""GPU parity proper: the CUDA path, called through the C ABI, against the oracle on seeded
states and against the committed reference goldens."""
import os

import numpy as np
import pytest

from helpers import TOL, colmajor_batch, golden_big_count, golden_df_du, load_golden, relerr

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from gridcodegenerator_b200 import load_named_robot                      # noqa: E402
from gridcodegenerator_b200.runtime import GridError, get_engine        # noqa: E402
from gridcodegenerator_b200.synthetic import make_states, pack_q_qd, pack_q_qd_u, seed_for  # noqa: E402
from oracle import rbd_numpy as O                                       # noqa: E402

ALL = ("id", "minv", "fd", "id_grad", "fd_grad")
FAMILIES = ("tps", "wps", "cps", "pipe")


class forced:
    """Pins the kernel family through the library's GRID_FORCE_KERNEL switch."""

    def __init__(self, family):
        self.family = family

    def __enter__(self):
        self.old = os.environ.get("GRID_FORCE_KERNEL")
        if self.family:
            os.environ["GRID_FORCE_KERNEL"] = self.family

    def __exit__(self, *exc):
        if self.old is None:
            os.environ.pop("GRID_FORCE_KERNEL", None)
        else:
            os.environ["GRID_FORCE_KERNEL"] = self.old


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def run_alg(eng, alg, q, qd, u, qdd=None, Minv=None, compressed=False):
    n, N = eng.n, q.shape[0]
    if alg == "id":
        out = torch.empty(N, n, device="cuda")
        eng.inverse_dynamics_device(out, dev(pack_q_qd(q, qd) if compressed else pack_q_qd_u(q, qd, u)),
                                    None if qdd is None else dev(qdd))
    elif alg == "minv":
        out = torch.empty(N, n * n, device="cuda")
        eng.direct_minv_device(out, dev(q if compressed else pack_q_qd_u(q, qd, u)))
    elif alg == "fd":
        out = torch.empty(N, n, device="cuda")
        eng.forward_dynamics_device(out, dev(pack_q_qd_u(q, qd, u)))
    elif alg == "id_grad":
        out = torch.empty(N, 2 * n * n, device="cuda")
        eng.inverse_dynamics_gradient_device(out, dev(pack_q_qd(q, qd) if compressed else pack_q_qd_u(q, qd, u)),
                                             None if qdd is None else dev(qdd))
    elif alg == "fd_grad":
        out = torch.empty(N, 2 * n * n, device="cuda")
        eng.forward_dynamics_gradient_device(out, dev(pack_q_qd_u(q, qd, u)),
                                             None if qdd is None else dev(qdd), None if Minv is None else dev(Minv))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def supported(eng, alg, family=None):
    kind = eng.kernel_kind(alg)
    return kind != "none" if family is None else family in kind


@pytest.mark.parametrize("tag", ["mixed5", "iiwa14", "iiwa14_damped", "hyq", "atlas", "chain64"])
def test_against_reference_goldens(tag):
    robot, z = load_golden(tag)
    eng = get_engine(robot)
    q, qd, u, qdd = z["q"], z["qd"], z["u"], z["qdd"]
    assert q.shape[0] >= 64                       # 64 reference-generated states per robot
    nb = golden_big_count(z)                      # the 64-link chain keeps its big matrices for 8 states
    checks = {
        "id": [(dict(), z["c"]), (dict(qdd=qdd), z["c_qdd"])],
        "minv": [(dict(), colmajor_batch(z["minv_upper"]))],
        "fd": [(dict(), z["fd_qdd"])],
        "id_grad": [(dict(), colmajor_batch(z["dc_du"])), (dict(qdd=qdd[:nb]), colmajor_batch(z["dc_du_qdd"]))],
        "fd_grad": [(dict(), colmajor_batch(golden_df_du(z)))],
    }
    ran = set()
    for family in FAMILIES:
        for alg, cases in checks.items():
            if not supported(eng, alg, family):
                continue
            for kw, ref in cases:
                m = ref.shape[0]
                with forced(family):
                    out = run_alg(eng, alg, q[:m], qd[:m], u[:m], **kw)
                assert relerr(out, ref) < TOL[alg], (tag, family, alg, relerr(out, ref))
                ran.add(alg)
    assert ran == set(ALL), "every algorithm must have a kernel for %s, got %s" % (tag, sorted(ran))


@pytest.mark.parametrize("name,N", [("iiwa14", 256), ("hyq", 256), ("atlas", 64), ("chain64", 8), ("mixed5", 256)])
@pytest.mark.parametrize("family", FAMILIES)
@pytest.mark.parametrize("alg", ALL)
def test_against_oracle_seeded_states(name, N, alg, family):
    robot = load_named_robot(name)
    eng = get_engine(robot)
    if not supported(eng, alg, family):
        pytest.skip("%s has no %s kernel for %s" % (name, family, alg))
    q, qd, u, qdd = make_states(robot.n, N, seed_for(name))
    q64, qd64, u64 = (x.astype(np.float64) for x in (q, qd, u))
    with forced(family):
        out = run_alg(eng, alg, q, qd, u)
    ref = O.batch(robot, alg, q64, qd64, u64 if alg in ("fd", "fd_grad") else None)
    assert relerr(out, ref) < TOL[alg], relerr(out, ref)
    # per-state check too (a single bad state must not hide behind the tensor-wide max)
    per = np.abs(out - ref).max(axis=1) / np.abs(ref).max(axis=1)
    assert per.max() < 20 * TOL[alg], per.max()


@pytest.mark.parametrize("family", FAMILIES)
def test_qdd_minv_overload_and_compressed_layouts(family, monkeypatch):
    monkeypatch.setenv("GRID_FORCE_KERNEL", family)
    robot = load_named_robot("iiwa14")
    eng = get_engine(robot)
    n = robot.n
    q, qd, u, qdd = make_states(n, 64, 99)
    q64, qd64, qdd64 = (x.astype(np.float64) for x in (q, qd, qdd))
    Mu = np.array([O.minv(robot, q64[s], dense=False).flatten(order="F") for s in range(64)], dtype=np.float32)
    Md = np.array([O.minv(robot, q64[s]) for s in range(64)])
    out = run_alg(eng, "fd_grad", q, qd, u, qdd=qdd, Minv=Mu)
    ref = O.batch(robot, "fd_grad_qdd_minv", q64, qd64, qdd64, Minv_in=Md)
    assert relerr(out, ref) < TOL["fd_grad"]
    # USE_COMPRESSED_MEM strides give the same answers as the 3n stride
    for alg in ("id", "minv", "id_grad"):
        a = run_alg(eng, alg, q, qd, u)
        b = run_alg(eng, alg, q, qd, u, compressed=True)
        assert np.array_equal(a, b), alg


@pytest.mark.parametrize("family", FAMILIES)
@pytest.mark.parametrize("N", [0, 1, 31, 32, 33, 1000])
def test_ragged_and_empty_batches(N, family, monkeypatch):
    monkeypatch.setenv("GRID_FORCE_KERNEL", family)
    robot = load_named_robot("iiwa14")
    eng = get_engine(robot)
    n = robot.n
    q, qd, u, _ = make_states(n, 1000, 5)
    big = run_alg(eng, "fd_grad", q, qd, u)
    q, qd, u = q[:N], qd[:N], u[:N]
    guard = torch.full((N + 2, 2 * n * n), 7.0, device="cuda")
    if N:
        eng.forward_dynamics_gradient_device(guard[1:N + 1], dev(pack_q_qd_u(q, qd, u)), num_timesteps=N, stride=3 * n)
    else:
        eng.forward_dynamics_gradient_device(guard[1:1], torch.empty(0, 3 * n, device="cuda"), num_timesteps=0, stride=3 * n)
    torch.cuda.synchronize()
    g = guard.cpu().numpy()
    assert np.all(g[0] == 7.0) and np.all(g[-1] == 7.0)        # no out-of-bounds writes
    assert np.array_equal(g[1:N + 1], big[:N])                  # batch size does not change results


def test_full_size_properties_iiwa_65536():
    """BASELINE size: checks that do not need the oracle at every state - Minv symmetry via
    M*Minv = I on a sample, FD/ID round trip, and fd_grad consistency between overloads."""
    robot = load_named_robot("iiwa14")
    eng = get_engine(robot)
    n, N = robot.n, 65536
    q, qd, u, _ = make_states(n, N, seed_for("iiwa14"))
    qdd = run_alg(eng, "fd", q, qd, u)
    tau = run_alg(eng, "id", q, qd, u, qdd=qdd)                 # ID(FD(u)) == u
    assert relerr(tau, u) < 1e-3
    assert np.isfinite(qdd).all()
    Minv = run_alg(eng, "minv", q, qd, u)
    a = run_alg(eng, "fd_grad", q, qd, u)
    b = run_alg(eng, "fd_grad", q, qd, u, qdd=qdd, Minv=Minv)   # same result from precomputed qdd, Minv
    assert relerr(b, a) < 1e-3
    idx = np.random.default_rng(0).choice(N, 64, replace=False)
    ref = O.batch(robot, "fd_grad", q[idx].astype(np.float64), qd[idx].astype(np.float64), u[idx].astype(np.float64))
    assert relerr(a[idx], ref) < TOL["fd_grad"]


def test_host_path_grid_data():
    robot = load_named_robot("iiwa14")
    eng = get_engine(robot)
    n, N = robot.n, 5000
    q, qd, u, qdd = make_states(n, N, 17)
    data = eng.make_data(N)
    data.h["q_qd_u"][:] = pack_q_qd_u(q, qd, u)
    df = data.forward_dynamics_gradient(N).copy()
    # the host path pipelines the batch in chunks, and small chunks take the latency kernels:
    # same values to rounding, not bit-identical to one big launch
    assert relerr(df, run_alg(eng, "fd_grad", q, qd, u)) < 1e-5
    c = data.inverse_dynamics(N).copy()
    assert np.array_equal(c, run_alg(eng, "id", q, qd, u))
    data.h["qdd"][:] = qdd
    c2 = data.inverse_dynamics(N, use_qdd=True).copy()
    assert np.array_equal(c2, run_alg(eng, "id", q, qd, u, qdd=qdd))
    assert np.array_equal(data.direct_minv(N).copy(), run_alg(eng, "minv", q, qd, u))
    assert np.array_equal(data.forward_dynamics(N).copy(), run_alg(eng, "fd", q, qd, u))
    assert relerr(data.inverse_dynamics_gradient(N).copy(), run_alg(eng, "id_grad", q, qd, u)) < 1e-5
    with pytest.raises(GridError):
        data.forward_dynamics(N + 1)
    data.close()


def test_host_path_grid_data_atlas_phase_split():
    """gridData flow on Atlas: the host path pipelines chunks over several streams, each chunk a
    phase-split launch with its own stream-ordered scratch array."""
    robot = load_named_robot("atlas")
    eng = get_engine(robot)
    n, N = robot.n, 9000
    q, qd, u, _ = make_states(n, N, 18)
    data = eng.make_data(N)
    data.h["q_qd_u"][:] = pack_q_qd_u(q, qd, u)
    for _ in range(2):
        assert np.array_equal(data.forward_dynamics_gradient(N).copy(), run_alg(eng, "fd_grad", q, qd, u))
    assert np.array_equal(data.forward_dynamics(N).copy(), run_alg(eng, "fd", q, qd, u))
    assert np.array_equal(data.direct_minv(N).copy(), run_alg(eng, "minv", q, qd, u))
    assert np.array_equal(data.inverse_dynamics_gradient(N).copy(), run_alg(eng, "id_grad", q, qd, u))
    data.close()


def test_phase_split_chunked_launches_subprocess(tmp_path):
    """GRID_PIPE_CHUNK (read once per process) cuts a batch into chunks that reuse one scratch array;
    ragged last chunk and ragged last tile included.  Same bits as the unchunked launch."""
    import subprocess
    import sys
    robot = load_named_robot("atlas")
    eng = get_engine(robot)
    N = 10007
    q, qd, u, _ = make_states(robot.n, N, 41)
    ref = run_alg(eng, "fd_grad", q, qd, u)
    np.save(tmp_path / "ref.npy", ref)
    code = (
        "import sys, numpy as np, torch\n"
        "sys.path.insert(0, %r)\n"
        "from gridcodegenerator_b200 import load_named_robot\n"
        "from gridcodegenerator_b200.runtime import get_engine\n"
        "from gridcodegenerator_b200.synthetic import make_states, pack_q_qd_u\n"
        "r = load_named_robot('atlas'); e = get_engine(r)\n"
        "q, qd, u, _ = make_states(r.n, %d, 41)\n"
        "x = torch.from_numpy(pack_q_qd_u(q, qd, u)).cuda(); o = torch.empty(%d, 2 * r.n * r.n, device='cuda')\n"
        "e.forward_dynamics_gradient_device(o, x); torch.cuda.synchronize()\n"
        "assert np.array_equal(o.cpu().numpy(), np.load(%r)), 'chunked launch differs'\n"
        "print('chunked ok')\n" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), N, N, str(tmp_path / "ref.npy")))
    env = dict(os.environ, GRID_PIPE_CHUNK="4096", GRID_FORCE_KERNEL="pipe")
    p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "chunked ok" in p.stdout, p.stdout + p.stderr


def test_phase_split_concurrent_streams():
    """Two streams, two batches, launches interleaved: results equal the serial ones (the scratch
    arrays are per call)."""
    robot = load_named_robot("atlas")
    eng = get_engine(robot)
    n = robot.n
    qa, qda, ua, _ = make_states(n, 3000, 31)
    qb, qdb, ub, _ = make_states(n, 2500, 32)
    ra, rb = run_alg(eng, "fd_grad", qa, qda, ua), run_alg(eng, "fd_grad", qb, qdb, ub)
    xa, xb = dev(pack_q_qd_u(qa, qda, ua)), dev(pack_q_qd_u(qb, qdb, ub))
    oa, ob = torch.empty(3000, 2 * n * n, device="cuda"), torch.empty(2500, 2 * n * n, device="cuda")
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for _ in range(4):
        eng.forward_dynamics_gradient_device(oa, xa, stream=sa)
        eng.forward_dynamics_gradient_device(ob, xb, stream=sb)
    torch.cuda.synchronize()
    assert np.array_equal(oa.cpu().numpy(), ra) and np.array_equal(ob.cpu().numpy(), rb)


@pytest.mark.parametrize("name,N,algs", [("iiwa14", 65536, ALL), ("hyq", 16384, ALL), ("atlas", 4096, ALL),
                                         ("chain64", 256, ALL), ("mixed5", 8192, ALL)])
def test_full_batches_against_c_oracle(name, N, algs):
    """BASELINE.json batch sizes, every state checked: the C restatement of the oracle
    (oracle/rbd_oracle.c, pinned to the reference goldens in tests/test_oracle.py) finishes
    these batches in seconds."""
    from oracle import c_oracle as C
    robot = load_named_robot(name)
    eng = get_engine(robot)
    q, qd, u, qdd = make_states(robot.n, N, seed_for(name) + 1)
    q64, qd64, u64 = (x.astype(np.float64) for x in (q, qd, u))
    for alg in algs:
        out = run_alg(eng, alg, q, qd, u)
        ref = C.batch(robot, alg, q64, qd64, u64 if alg in ("fd", "fd_grad") else None)
        assert relerr(out, ref) < TOL[alg], (name, alg, relerr(out, ref))
        per = np.abs(out - ref).max(axis=1) / np.abs(ref).max(axis=1)
        assert per.max() < 50 * TOL[alg], (name, alg, per.max())


@pytest.mark.parametrize("alg", ["minv", "fd", "id_grad", "fd_grad"])
@pytest.mark.parametrize("N", [1, 31, 33, 1000])
def test_pipe_ragged_batches_atlas(N, alg, monkeypatch):
    """Phase-split kernels: ragged last tile, guard rows around the output, same values for any
    batch size, and the scratch words of lanes past the end never reach the output."""
    monkeypatch.setenv("GRID_FORCE_KERNEL", "pipe")
    robot = load_named_robot("atlas")
    eng = get_engine(robot)
    assert "pipe" in eng.kernel_kind(alg)
    n = robot.n
    q, qd, u, _ = make_states(n, 1000, 5)
    big = run_alg(eng, alg, q, qd, u)
    words = big.shape[1]
    guard = torch.full((N + 2, words), 7.0, device="cuda")
    x = dev(pack_q_qd_u(q[:N], qd[:N], u[:N]))
    call = {"minv": eng.direct_minv_device, "fd": eng.forward_dynamics_device,
            "id_grad": eng.inverse_dynamics_gradient_device, "fd_grad": eng.forward_dynamics_gradient_device}[alg]
    call(guard[1:N + 1], x, num_timesteps=N, stride=3 * n)
    torch.cuda.synchronize()
    g = guard.cpu().numpy()
    assert np.all(g[0] == 7.0) and np.all(g[-1] == 7.0)
    assert np.array_equal(g[1:N + 1], big[:N])


@pytest.mark.parametrize("name", ["iiwa14", "mixed5"])
@pytest.mark.parametrize("mode", ["staged", "fused"])
def test_forced_split_of_single_tree_robots(name, mode, monkeypatch):
    """A serial chain (iiwa14) and a tree with prismatic joints (mixed5) forced through the
    state-program + one-column-program-per-joint split (library variant built by
    __graft_entry__.build()): same answers as the oracle, ragged batch, both launch modes."""
    import __graft_entry__ as G
    from gridcodegenerator_b200.runtime import GridEngine
    robot = load_named_robot(name)
    eng = GridEngine(robot, plan=G.split_test_plan(robot), tag=G.SPLIT_TEST_TAG)
    monkeypatch.setenv("GRID_FORCE_KERNEL", "pipe")
    if mode == "fused":
        monkeypatch.setenv("GRID_PIPE_MODE", "fused")
    N = 1000
    q, qd, u, qdd = make_states(robot.n, N, seed_for(name) + 3)
    q64, qd64, u64, qdd64 = (x.astype(np.float64) for x in (q, qd, u, qdd))
    for alg, kw, ref in (("id_grad", {}, O.batch(robot, "id_grad", q64[:128], qd64[:128])),
                         ("id_grad", dict(qdd=qdd), O.batch(robot, "id_grad", q64[:128], qd64[:128], qdd64[:128])),
                         ("fd_grad", {}, O.batch(robot, "fd_grad", q64[:128], qd64[:128], u64[:128]))):
        assert "pipe" in eng.kernel_kind(alg)
        out = run_alg(eng, alg, q, qd, u, **kw)
        assert relerr(out[:128], ref) < TOL[alg], (name, alg, relerr(out[:128], ref))
        assert np.isfinite(out).all()


@pytest.mark.parametrize("N", [1, 31, 33, 63, 64, 65, 1000, 4099])
def test_packed_two_states_per_lane_column_programs(N, monkeypatch):
    """KernelPlan(pipe_x2=...): stage-1 warps take 64-state tiles; column programs under the register-demand
    threshold run two states per lane on FFMA2 / FMUL2 / FADD2 with re-loaded scratch words, the others run the tile
    as two halves (library variant built by __graft_entry__.build()).  Ragged tiles, guard rows, the oracle."""
    import __graft_entry__ as G
    from gridcodegenerator_b200.runtime import GridEngine
    robot = load_named_robot("iiwa14")
    n = robot.n
    eng = GridEngine(robot, plan=G.x2_test_plan(robot), tag=G.X2_TEST_TAG)
    st = eng.build_info.get("stats", {}).get("pipe_fd_grad")
    if st:                                       # freshly built here: both kinds of stage-1 programs exist
        lives = st["x2_live"]
        assert any(v <= 70 for v in lives) and any(v > 70 for v in lives)
    monkeypatch.setenv("GRID_FORCE_KERNEL", "pipe")
    q, qd, u, qdd = make_states(n, N, 77)
    q64, qd64, u64, qdd64 = (x.astype(np.float64) for x in (q, qd, u, qdd))
    M = min(N, 128)
    for alg, kw, ref in (("id_grad", {}, O.batch(robot, "id_grad", q64[-M:], qd64[-M:])),
                         ("id_grad", dict(qdd=qdd), O.batch(robot, "id_grad", q64[-M:], qd64[-M:], qdd64[-M:])),
                         ("fd_grad", {}, O.batch(robot, "fd_grad", q64[-M:], qd64[-M:], u64[-M:]))):
        assert "pipe" in eng.kernel_kind(alg)
        out = run_alg(eng, alg, q, qd, u, **kw)
        assert np.isfinite(out).all()
        assert relerr(out[-M:], ref) < TOL[alg], (alg, N, relerr(out[-M:], ref))
    # guard rows around a ragged batch
    guard = torch.full((N + 2, 2 * n * n), 7.0, device="cuda")
    eng.forward_dynamics_gradient_device(guard[1:N + 1], dev(pack_q_qd_u(q, qd, u)), num_timesteps=N, stride=3 * n)
    torch.cuda.synchronize()
    g = guard.cpu().numpy()
    assert np.all(g[0] == 7.0) and np.all(g[-1] == 7.0)
    assert np.array_equal(g[1:N + 1], out)


@pytest.mark.parametrize("chunk", [2048, 8192])
def test_pipe_chunk_major_item_order_is_bit_identical(chunk, monkeypatch):
    """GRID_PIPE_ORDER_CHUNK (experiment, off by default): the stage-1 items of a two-stage launch walked chunk by
    chunk in descending order, ragged last chunk first - another order of the same independent items."""
    robot = load_named_robot("atlas")
    eng = get_engine(robot)
    N = 20000 + 37
    q, qd, u, _ = make_states(robot.n, N, 41)
    base = run_alg(eng, "fd_grad", q, qd, u)
    monkeypatch.setenv("GRID_PIPE_ORDER_CHUNK", str(chunk))
    assert np.array_equal(run_alg(eng, "fd_grad", q, qd, u), base)
    assert np.array_equal(run_alg(eng, "id_grad", q, qd, u), run_alg(eng, "id_grad", q, qd, u))
    monkeypatch.delenv("GRID_PIPE_ORDER_CHUNK")
    assert np.array_equal(run_alg(eng, "fd_grad", q, qd, u), base)


@pytest.mark.parametrize("name,N", [("atlas", 1000), ("hyq", 4099)])
def test_pipe_fused_variant_matches_staged(name, N, monkeypatch):
    """The SM-partitioned single-kernel variant (GRID_PIPE_MODE=fused: stage-1 warps wait on
    release/acquire flags of stage 0) computes exactly what the two staged kernels compute."""
    monkeypatch.setenv("GRID_FORCE_KERNEL", "pipe")
    robot = load_named_robot(name)
    eng = get_engine(robot)
    q, qd, u, _ = make_states(robot.n, N, 77)
    for alg in ("fd", "id_grad", "fd_grad"):
        monkeypatch.delenv("GRID_PIPE_MODE", raising=False)
        staged = run_alg(eng, alg, q, qd, u)
        monkeypatch.setenv("GRID_PIPE_MODE", "fused")
        for _ in range(3):
            assert np.array_equal(run_alg(eng, alg, q, qd, u), staged), alg


@pytest.mark.parametrize("name,family", [("atlas", "wps"), ("atlas", "pipe"), ("iiwa14", "tps"), ("iiwa14", "cps"),
                                         ("mixed5", "wps")])
def test_repeated_launches_are_bit_identical(name, family, monkeypatch):
    """compute-sanitizer is closed on this pool, so races are hunted the indirect way: the wide
    kernels use named barriers, aliased scratch and shared-memory atomics - any race would show up as
    run-to-run differences."""
    monkeypatch.setenv("GRID_FORCE_KERNEL", family)
    robot = load_named_robot(name)
    eng = get_engine(robot)
    q, qd, u, _ = make_states(robot.n, 300, 123)
    first = run_alg(eng, "fd_grad", q, qd, u)
    for _ in range(4):
        assert np.array_equal(run_alg(eng, "fd_grad", q, qd, u), first)


def test_randomised_sweep_all_families(monkeypatch, capsys):
    """tests/fuzz_parity.py: random batch sizes, strides, seeds, qdd overloads and kernel families (incl.
    both launch modes of the phase-split kernels) against the C oracle, guard rows around every output."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("fuzz_parity", os.path.join(root, "tests", "fuzz_parity.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.setattr("sys.argv", ["fuzz_parity.py", "48"])
    monkeypatch.setenv("GRID_FORCE_KERNEL", "tps")      # restored by monkeypatch after the sweep rewrites it
    monkeypatch.setenv("GRID_PIPE_MODE", "staged")
    assert mod.main() == 0, capsys.readouterr().out[-2000:]


# ---- chain kernels (csrc/grid_lps.cuh): lane per state, rolled joint loops --------------------------------
@pytest.mark.parametrize("name", ["iiwa14", "pchain4"])
@pytest.mark.parametrize("N", [1, 33, 1000, 5000])
def test_chain_kernels_on_small_chains(name, N, monkeypatch):
    """The chain kernels forced onto small serial chains (library variant of __graft_entry__.build()): every
    algorithm against the numpy oracle - damping, prismatic joints (pchain4), ragged tiles, the chunked launch
    (N > 4096) and guard rows included."""
    import __graft_entry__ as G
    from gridcodegenerator_b200.runtime import GridEngine
    robot = load_named_robot(name)                  # (pchain4's URDF carries joint damping)
    eng = GridEngine(robot, plan=G.lps_test_plan(robot), tag=G.LPS_TEST_TAG)
    monkeypatch.setenv("GRID_FORCE_KERNEL", "lps")
    n = robot.n
    q, qd, u, qdd = make_states(n, N, seed_for(name) + 9)
    M = min(N, 96)
    q64, qd64, u64, qdd64 = (x[:M].astype(np.float64) for x in (q, qd, u, qdd))
    cases = [("minv", {}, O.batch(robot, "minv", q64)), ("fd", {}, O.batch(robot, "fd", q64, qd64, u64)),
             ("id_grad", {}, O.batch(robot, "id_grad", q64, qd64)),
             ("id_grad", dict(qdd=qdd), O.batch(robot, "id_grad", q64, qd64, qdd64)),
             ("fd_grad", {}, O.batch(robot, "fd_grad", q64, qd64, u64))]
    for alg, kw, ref in cases:
        assert "lps" in eng.kernel_kind(alg)
        out = run_alg(eng, alg, q, qd, u, **kw)
        assert np.isfinite(out).all()
        assert relerr(out[:M], ref) < TOL[alg], (name, alg, relerr(out[:M], ref))
        per = np.abs(out[:M] - ref).max(axis=1) / np.abs(ref).max(axis=1)
        assert per.max() < 20 * TOL[alg], (name, alg, per.max())
    # USE_QDD_MINV_FLAG overload: the CALLER's Minv must be used (here a deliberately scaled one), through the
    # batched product kernel (tensor cores when n % 16 == 0, FFMA otherwise)
    Md = np.array([1.25 * O.minv(robot, q64[s]) for s in range(M)])
    Mu = np.array([np.triu(Md[s]).flatten(order="F") for s in range(M)], dtype=np.float32)
    out = run_alg(eng, "fd_grad", q[:M], qd[:M], u[:M], qdd=qdd[:M], Minv=Mu)
    ref = O.batch(robot, "fd_grad_qdd_minv", q64, qd64, qdd64, Minv_in=Md)
    assert relerr(out, ref) < TOL["fd_grad"], (name, "fd_grad_qdd_minv", relerr(out, ref))
    # guard rows and batch-size independence
    big = run_alg(eng, "fd_grad", q, qd, u)
    guard = torch.full((N + 2, 2 * n * n), 7.0, device="cuda")
    eng.forward_dynamics_gradient_device(guard[1:N + 1], dev(pack_q_qd_u(q, qd, u)), num_timesteps=N, stride=3 * n)
    torch.cuda.synchronize()
    g = guard.cpu().numpy()
    assert np.all(g[0] == 7.0) and np.all(g[-1] == 7.0) and np.array_equal(g[1:N + 1], big)
    if N >= 33:
        assert np.array_equal(run_alg(eng, "fd_grad", q[:33], qd[:33], u[:33]), big[:33])


def test_chain64_chain_kernels_match_wide_kernels_and_oracle(monkeypatch):
    """64-link chain: the chain kernels (default above 256 states) against the C oracle at 4 096 states for the
    gradients and 65 536 for ID, and against the CTA-per-state kernels they replace."""
    from oracle import c_oracle as C
    robot = load_named_robot("chain64")
    eng = get_engine(robot)
    n, N = robot.n, 4096
    q, qd, u, qdd = make_states(n, N, seed_for("chain64") + 5)
    q64, qd64, u64 = (x.astype(np.float64) for x in (q, qd, u))
    for alg in ("minv", "fd", "id_grad", "fd_grad"):
        assert "lps" in eng.kernel_kind(alg)
        monkeypatch.setenv("GRID_FORCE_KERNEL", "lps")
        out = run_alg(eng, alg, q, qd, u)
        ref = C.batch(robot, alg, q64, qd64, u64 if alg in ("fd", "fd_grad") else None)
        assert relerr(out, ref) < TOL[alg], (alg, relerr(out, ref))
        per = np.abs(out - ref).max(axis=1) / np.abs(ref).max(axis=1)
        assert per.max() < 50 * TOL[alg], (alg, per.max())
        monkeypatch.setenv("GRID_FORCE_KERNEL", "wps")
        wide = run_alg(eng, alg, q[:64], qd[:64], u[:64])
        assert relerr(out[:64], wide) < TOL[alg], alg
        monkeypatch.delenv("GRID_FORCE_KERNEL")
        auto = run_alg(eng, alg, q[:512], qd[:512], u[:512])
        if alg == "fd":                          # default = the single-thread articulated-body program
            assert eng.kernel_kind("fd@large") == "tps(aba)"
            assert relerr(auto, ref[:512]) < TOL[alg]
        else:                                    # default = chain kernels
            assert np.array_equal(auto, out[:512])
    # USE_QDD_MINV_FLAG overload on the chain kernels: dc_du columns at the given qdd, then -Minv_given dc_du as a
    # batched 3xTF32 tensor-core product; feeding FD's own qdd and Minv back must reproduce df_du
    M = 512
    qdd_fd = run_alg(eng, "fd", q[:M], qd[:M], u[:M])
    Minv_fd = run_alg(eng, "minv", q[:M], qd[:M], u[:M])
    ref = C.batch(robot, "fd_grad", q64[:M], qd64[:M], u64[:M])
    pre = run_alg(eng, "fd_grad", q[:M], qd[:M], u[:M], qdd=qdd_fd, Minv=Minv_fd)
    assert relerr(pre, ref) < TOL["fd_grad"], relerr(pre, ref)
    monkeypatch.setenv("GRID_FORCE_KERNEL", "wps")
    wide = run_alg(eng, "fd_grad", q[:32], qd[:32], u[:32], qdd=qdd_fd[:32], Minv=Minv_fd[:32])
    monkeypatch.delenv("GRID_FORCE_KERNEL")
    assert relerr(pre[:32], wide) < 1e-4
    N = 65536
    q, qd, u, _ = make_states(n, N, seed_for("chain64") + 6)
    out = run_alg(eng, "id", q, qd, u)
    ref = C.batch(robot, "id", q.astype(np.float64), qd.astype(np.float64))
    assert relerr(out, ref) < TOL["id"]


def test_atlas_fd_gradient_full_batch_65536_against_c_oracle():
    """BASELINE config 4 at its full size: 65 536 Atlas states through the launch shape the bench uses (8-warp CTAs,
    ticket counters, the large-batch column programs), every state checked against the C oracle;
    and the small-batch column programs (<= 24 576 states) give the same values."""
    from oracle import c_oracle as C
    robot = load_named_robot("atlas")
    eng = get_engine(robot)
    N = 65536
    q, qd, u, _ = make_states(robot.n, N, seed_for("atlas") + 2)
    out = run_alg(eng, "fd_grad", q, qd, u)
    ref = C.batch(robot, "fd_grad", q.astype(np.float64), qd.astype(np.float64), u.astype(np.float64))
    assert relerr(out, ref) < TOL["fd_grad"], relerr(out, ref)
    per = np.abs(out - ref).max(axis=1) / np.abs(ref).max(axis=1)
    assert per.max() < 50 * TOL["fd_grad"], per.max()
    small = run_alg(eng, "fd_grad", q[:8192], qd[:8192], u[:8192])
    assert relerr(small, ref[:8192]) < TOL["fd_grad"]
    assert relerr(small, out[:8192]) < 1e-5


@pytest.mark.parametrize("name,N", [("atlas", 3000), ("hyq", 1000), ("mixed5", 500)])
def test_qdd_minv_overload_on_phase_split_kernels(name, N, monkeypatch):
    """USE_QDD_MINV_FLAG on the phase-split kernels (round 1: 8x slower wide kernel for Atlas): qdd comes through the
    input tile, the caller's Minv is read straight from global memory by the state programs and handed to the column
    programs through the scratch array.  The CALLER's Minv must be used (a scaled one here); the wide kernel agrees."""
    robot = load_named_robot(name)
    eng = get_engine(robot) if name != "mixed5" else None
    if eng is None:
        import __graft_entry__ as G
        from gridcodegenerator_b200.runtime import GridEngine
        eng = GridEngine(robot, plan=G.split_test_plan(robot), tag=G.SPLIT_TEST_TAG)
    assert "pipe" in eng.kernel_kind("fd_grad")
    n = robot.n
    q, qd, u, qdd = make_states(n, N, seed_for(name) + 21)
    M = 48
    q64, qd64, qdd64 = (x[:M].astype(np.float64) for x in (q, qd, qdd))
    Md = np.array([0.8 * O.minv(robot, q64[s]) for s in range(M)])
    Mu_small = np.array([np.triu(Md[s]).flatten(order="F") for s in range(M)], dtype=np.float32)
    Mu = np.tile(Mu_small, (N // M + 1, 1))[:N].copy()
    monkeypatch.setenv("GRID_FORCE_KERNEL", "pipe")
    out = run_alg(eng, "fd_grad", q, qd, u, qdd=qdd, Minv=Mu)
    ref = O.batch(robot, "fd_grad_qdd_minv", q64, qd64, qdd64, Minv_in=Md)
    assert np.isfinite(out).all()
    assert relerr(out[:M], ref) < TOL["fd_grad"], relerr(out[:M], ref)
    monkeypatch.setenv("GRID_FORCE_KERNEL", "wps")
    wide = run_alg(eng, "fd_grad", q[:M], qd[:M], u[:M], qdd=qdd[:M], Minv=Mu[:M])
    assert relerr(out[:M], wide) < 1e-4
    monkeypatch.delenv("GRID_FORCE_KERNEL")
    assert np.array_equal(run_alg(eng, "fd_grad", q[:M], qd[:M], u[:M], qdd=qdd[:M], Minv=Mu[:M]), out[:M]) or name == "mixed5"
