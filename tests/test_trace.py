"""Host logic: the traced per-robot programs (numpy interpretation of the same DAG the CUDA
emitter prints) against the oracle, in float64 and float32."""
import numpy as np
import pytest

from gridcodegenerator_b200 import load_named_robot
from gridcodegenerator_b200.algorithms import TRACERS, algorithmic_flops
from gridcodegenerator_b200.synthetic import make_states
from helpers import TOL, relerr
from oracle import rbd_numpy as O


def _ins(**kw):
    d = {"gravity": 9.81}
    for k, v in kw.items():
        for i in range(v.shape[1]):
            d["%s%d" % (k, i)] = v[:, i]
    return d


def _run(robot, key, q, qd, u, qdd, dtype):
    p = TRACERS[key](robot)
    N = q.shape[0]
    if key == "id":
        return p.evaluate(_ins(q=q, qd=qd), dtype)["c"], O.batch(robot, "id", q, qd), "id"
    if key == "id_qdd":
        return p.evaluate(_ins(q=q, qd=qd, qdd=qdd), dtype)["c"], O.batch(robot, "id", q, qd, qdd), "id"
    if key == "minv":
        return p.evaluate(_ins(q=q), dtype)["Minv"], O.batch(robot, "minv", q), "minv"
    if key == "fd":
        return p.evaluate(_ins(q=q, qd=qd, u=u), dtype)["qdd"], O.batch(robot, "fd", q, qd, u), "fd"
    if key == "id_grad":
        return p.evaluate(_ins(q=q, qd=qd), dtype)["dc_du"], O.batch(robot, "id_grad", q, qd), "id_grad"
    if key == "id_grad_qdd":
        return p.evaluate(_ins(q=q, qd=qd, qdd=qdd), dtype)["dc_du"], O.batch(robot, "id_grad", q, qd, qdd), "id_grad"
    if key == "fd_grad":
        return p.evaluate(_ins(q=q, qd=qd, u=u), dtype)["df_du"], O.batch(robot, "fd_grad", q, qd, u), "fd_grad"
    if key == "fd_grad_qdd_minv":
        Mu = np.array([O.minv(robot, q[s], dense=False).flatten(order="F") for s in range(N)])
        Md = np.array([O.minv(robot, q[s]) for s in range(N)])
        return (p.evaluate(_ins(q=q, qd=qd, qdd=qdd, Minv=Mu), dtype)["df_du"],
                O.batch(robot, "fd_grad_qdd_minv", q, qd, qdd, Minv_in=Md), "fd_grad")
    if key in ("fd_grad_q", "fd_grad_qd"):       # the halves of the mid-size-batch kernel: d/dq block, d/dqd block
        nn = robot.n * robot.n
        ref = O.batch(robot, "fd_grad", q, qd, u)
        return (p.evaluate(_ins(q=q, qd=qd, u=u), dtype)["df_du"],
                ref[:, :nn] if key == "fd_grad_q" else ref[:, nn:], "fd_grad")
    if key == "crba":                            # further algorithms
        return p.evaluate(_ins(q=q), dtype)["M"], O.batch(robot, "crba", q), "minv"
    if key == "aba":
        return p.evaluate(_ins(q=q, qd=qd, u=u), dtype)["qdd"], O.batch(robot, "fd", q, qd, u), "fd"
    if key in ("fd_vjp", "fd_lin"):              # consumers fused after the FD gradient
        lam = np.random.default_rng(2).uniform(-3, 3, (N, 2 * robot.n))
        ins = _ins(q=q, qd=qd, u=u, lam=lam)
        ins["dt"] = 0.0125
        ins = {k: v for k, v in ins.items() if k in p.inputs}
        return p.evaluate(ins, dtype)[key], O.consumer_batch(robot, key, q, qd, u, 0.0125, lam), "fd_grad"
    raise KeyError(key)


@pytest.mark.parametrize("name", ["iiwa14", "hyq", "mixed5"])
@pytest.mark.parametrize("key", list(TRACERS))
def test_traced_program_matches_oracle(name, key):
    robot = load_named_robot(name)
    if name != "mixed5":
        robot = robot.with_damping(0.3)
    q, qd, u, qdd = (x.astype(np.float64) for x in make_states(robot.n, 4, 11))
    out, ref, alg = _run(robot, key, q, qd, u, qdd, np.float64)
    assert relerr(out, ref) < 1e-11
    out32, _, _ = _run(robot, key, q, qd, u, qdd, np.float32)
    assert relerr(out32, ref) < TOL[alg]


@pytest.mark.parametrize("name", ["atlas", "chain64"])
def test_traced_id_large_robots(name):
    robot = load_named_robot(name)
    q, qd, u, qdd = (x.astype(np.float64) for x in make_states(robot.n, 2, 3))
    out, ref, _ = _run(robot, "id_qdd", q, qd, u, qdd, np.float64)
    assert relerr(out, ref) < 1e-11


def test_traced_work_is_below_dense_reference_count():
    robot = load_named_robot("iiwa14")
    alg = algorithmic_flops(robot)
    assert alg["fd_grad"] == 41834 and alg["id"] == 2779
    traced = TRACERS["fd_grad"](robot).op_counts()["flops"]
    assert traced < alg["fd_grad"] / 3


def test_prismatic_joint_traces():
    """A prismatic joint makes r (not E) depend on q."""
    from gridcodegenerator_b200.robot import Robot, spatial_inertia
    I = spatial_inertia(2.0, [0.01, 0.02, 0.03], np.diag([0.1, 0.2, 0.3]))
    E = np.eye(3)
    robot = Robot("pr", [-1, 0, 1], [2, 4, 0], [E, E, E], [[0, 0, 0.1], [0.2, 0, 0], [0, 0.3, 0]], [I, I, I],
                  [0.1, 0.0, 0.2])
    q, qd, u, qdd = (x.astype(np.float64) for x in make_states(3, 3, 5))
    for key in TRACERS:
        out, ref, _ = _run(robot, key, q, qd, u, qdd, np.float64)
        assert relerr(out, ref) < 1e-11, key


@pytest.mark.parametrize("name", ["iiwa14", "hyq"])
@pytest.mark.parametrize("alg", ["id_grad", "fd_grad"])
def test_lane_uniform_column_program(name, alg):
    """The latency kernels run ONE program on every lane; the lane's column enters through 0/1
    masks.  Assembling all 2n columns must reproduce the full gradient."""
    from gridcodegenerator_b200.algorithms import trace_column_program
    robot = load_named_robot(name).with_damping(0.2)
    n = robot.n
    q, qd, u, qdd = (x.astype(np.float64) for x in make_states(n, 2, 9))
    p = trace_column_program(robot, alg)
    base = _ins(q=q, qd=qd, u=u)
    cols = []
    for c in range(2 * n):
        ins = dict(base)
        for i in range(n):
            ins["mq%d" % i] = np.full(2, 1.0 if c == i else 0.0)
            ins["mqd%d" % i] = np.full(2, 1.0 if c == n + i else 0.0)
        cols.append(p.evaluate(ins, np.float64)["col"])
    out = np.concatenate(cols, axis=1)
    ref = O.batch(robot, alg, q, qd, u if alg == "fd_grad" else None)
    assert relerr(out, ref) < 1e-11
    # a lane without a column (all masks zero) produces exact zeros
    ins = dict(base)
    for i in range(n):
        ins["mq%d" % i] = np.zeros(2)
        ins["mqd%d" % i] = np.zeros(2)
    assert np.all(p.evaluate(ins, np.float64)["col"] == 0.0)


@pytest.mark.parametrize("name", ["iiwa14", "hyq"])
@pytest.mark.parametrize("park", [(), ("v", "mXa", "mf"), ("v", "Iv", "mXa", "mf", "Minv")])
def test_paired_and_parked_gradient_programs(name, park):
    """Experimental tps variants (KernelPlan.tps_pairs / tps_v2_park): the (d/dq_j, d/dqd_j) columns
    travel as float2 pairs (FFMA2) and per-joint data can be parked in shared memory and re-loaded
    per column.  Same numbers as the scalar trace."""
    from gridcodegenerator_b200.algorithms import trace_fd_grad_paired, trace_id_grad_paired
    robot = load_named_robot(name).with_damping(0.2)
    q, qd, u, qdd = (x.astype(np.float64) for x in make_states(robot.n, 3, 13))
    p = trace_fd_grad_paired(robot, False, park)
    assert p.op_counts().get("packed", 0) > 0
    assert relerr(p.evaluate(_ins(q=q, qd=qd, u=u))["df_du"], O.batch(robot, "fd_grad", q, qd, u)) < 1e-11
    p = trace_id_grad_paired(robot, True, [x for x in park if x != "Minv"])
    assert relerr(p.evaluate(_ins(q=q, qd=qd, qdd=qdd))["dc_du"], O.batch(robot, "id_grad", q, qd, qdd)) < 1e-11
    if park:
        assert len(p.parks) > 0 and any(k[0] == "ld" for k in p.nodes)


def test_experimental_variants_emit_compilable_looking_code():
    from gridcodegenerator_b200.codegen import emit_alg_struct_looped, emit_alg_struct_v2
    robot = load_named_robot("iiwa14")
    txt, cnt = emit_alg_struct_looped(robot, "AlgFdGrad", "fd_grad")
    assert "for (int col = 0; col < 14; ++col)" in txt and cnt["loop_body_flops"] > 0
    txt, cnt = emit_alg_struct_v2(robot, "fd_grad", ("v", "mXa", "mf"))
    assert "__ffma2_rn(" in txt and "flush_colpair<7>(g_tile, s_warp, 6, cnt, lane);" in txt and "s_park[" in txt
    assert cnt["park_slots"] == 84
