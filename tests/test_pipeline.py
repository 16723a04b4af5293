"""Host-side checks of the phase-split decomposition (gridcodegenerator_b200/pipeline.py): the task
programs, interpreted with numpy in kernel order, must reproduce the oracle - components, column
groups, scratch hand-over and zero blocks included.  The GPU runs of the same programs are in
test_gpu_parity.py."""
import numpy as np
import pytest

from gridcodegenerator_b200 import load_named_robot
from gridcodegenerator_b200.codegen import KernelPlan
from gridcodegenerator_b200.pipeline import PipeVariant, components, emit_pipe_struct
from gridcodegenerator_b200.synthetic import make_states
from helpers import relerr
from oracle import rbd_numpy as O


def rows_for(variant, q, qd, u, qdd):
    return np.concatenate({"id": [q, qd], "id_qdd": [q, qd, qdd], "minv": [q], "fd": [q, qd, u], "id_grad": [q, qd],
                           "id_grad_qdd": [q, qd, qdd], "fd_grad": [q, qd, u]}[variant], axis=1).astype(np.float64)


def oracle_for(robot, variant, q, qd, u, qdd):
    q, qd, u, qdd = (x.astype(np.float64) for x in (q, qd, u, qdd))
    if variant.endswith("_qdd"):
        return O.batch(robot, variant[:-4], q, qd, qdd)
    return O.batch(robot, variant, q, qd, u if variant in ("fd", "fd_grad") else None)


def test_components_are_the_root_subtrees():
    assert [len(c) for c in components(load_named_robot("atlas"))] == [18, 6, 6]
    assert [len(c) for c in components(load_named_robot("hyq"))] == [3, 3, 3, 3]
    assert [len(c) for c in components(load_named_robot("iiwa14"))] == [7]


@pytest.mark.parametrize("name,opts", [("atlas", {}), ("hyq", {}), ("mixed5", dict(single_stage_max_flops=0, group_flops=1)),
                                       ("iiwa14", dict(single_stage_max_flops=0, group_flops=1)),
                                       ("iiwa14", dict(single_stage_max_flops=0, group_flops=4000))])
@pytest.mark.parametrize("variant", ["id", "id_qdd", "minv", "fd", "id_grad", "id_grad_qdd", "fd_grad"])
def test_pipe_programs_match_oracle(name, opts, variant):
    robot = load_named_robot(name)
    q, qd, u, qdd = make_states(robot.n, 6, 11)
    pv = PipeVariant(robot, variant, **opts)
    assert pv.feasible
    out = pv.evaluate(rows_for(variant, q, qd, u, qdd))
    assert not np.isnan(out).any(), "some output word is written by no task"
    ref = oracle_for(robot, variant, q, qd, u, qdd)
    assert relerr(out, ref) < 1e-10
    if opts and variant in ("id_grad", "fd_grad"):
        assert pv.scratch_words > 0 and len(pv.stage_tasks[1]) >= 2


def test_every_output_word_has_exactly_one_writer_and_emission_is_deterministic():
    robot = load_named_robot("atlas")
    pv = PipeVariant(robot, "fd_grad")
    seen = np.zeros(pv.out, dtype=int)
    for t in pv.tasks:
        for name, idx, _ in t.program.outputs:
            if name == "out":
                seen[idx] += 1
    assert (seen == 1).all()
    sc = np.zeros(pv.scratch_words, dtype=int)
    for t in pv.tasks:
        for name, idx, _ in t.program.outputs:
            if name == "sc":
                sc[idx] += 1
    assert (sc == 1).all()
    a, _ = emit_pipe_struct(pv)
    b, _ = emit_pipe_struct(PipeVariant(robot, "fd_grad"))
    assert a == b
    assert "flush2<" in a and "pipe::ldsc(sc_in" in a and "sc_out[" in a and "__syncthreads();" in a
    assert "static constexpr int NT = %d;" % len(pv.tasks) in a


def test_plan_policy():
    """Single-tree robots with thread-per-state programs keep them; forests (Atlas: torso+arms, two
    legs; HyQ: four legs) get one thread per (state, tree) for every algorithm and column groups
    where a tree's gradient does not fit one thread; a 64-link chain's columns are too long."""
    from helpers import cached_plan
    assert cached_plan("iiwa14").pipe == {}
    for name in ("atlas", "hyq"):
        plan = cached_plan(name)
        assert all("pipe" in plan.kind[a] for a in ("minv", "fd", "id_grad", "fd_grad")), plan.kind
        assert plan.kind["id"] == "tps"          # measured: the single-thread RNEA is faster (23 vs 31 us, Atlas 65 536)
    assert not PipeVariant(load_named_robot("chain64"), "id_grad").feasible
    assert cached_plan("chain64").pipe == {}


@pytest.mark.parametrize("variant", ["fd_grad", "id_grad", "fd_vjp", "fd_lin"])
def test_side_split_and_two_stage_legs_are_exact(variant):
    """Round-2 options of the phase-split decomposition - smaller column groups, two-stage legs, and the d/dq and
    d/dqd columns of an expensive joint as separate programs - reproduce the default decomposition bit for bit
    (float64 interpretation of the task programs)."""
    robot = load_named_robot("atlas")
    a = PipeVariant(robot, variant)
    b = PipeVariant(robot, variant, group_flops=3000, single_stage_max_flops=4000, split_sides_above=3500)
    assert len(b.tasks) > len(a.tasks) and any(t.name.endswith("q_j1_1") for t in b.tasks)
    rows = np.random.default_rng(4).uniform(-1.5, 1.5, (3, a.in0 + a.in1))
    oa, ob = a.evaluate(rows, dt=0.02), b.evaluate(rows, dt=0.02)
    assert not np.isnan(ob).any() and np.array_equal(oa, ob)
    assert max(t.flops for t in b.stage_tasks[1]) < 0.6 * max(t.flops for t in a.stage_tasks[1])
