"""Phase-split ("pipe") kernels: traced straight-line programs for robots whose whole
algorithm does not fit one thread.

The thread-per-state kernels (csrc/grid_tps.cuh) run one traced program per state.  That
stops working when the program outgrows 255 registers (Atlas FD-gradient: 72 k traced flops,
47 KB of spill traffic per state).  Two structural facts of the reference algorithms give a
decomposition into programs that DO fit:

  * **Forest components.**  Joints whose root ancestors differ never interact on a fixed
    base: M is block-diagonal, c_i only depends on its own component (the reference's
    ancestor/subtree tests, helpers/_topology_helpers.py:193-215, already encode this as
    structural zeros).  Atlas = {torso+arms+neck (18), left leg (6), right leg (6)}, HyQ =
    4 legs of 3.  One thread runs one (state, component).
  * **Column independence of the gradient.**  Every du-column of dc_du runs its own
    forward/backward recursion (algorithms/_inverse_dynamics_gradient.py:189-541 loops over
    columns inside every wave; oracle _test.py:229-488); the only shared data are the RNEA
    results v, I v, mxS(X a_parent), mxS(f) and, for the FD gradient, Minv
    (algorithms/_forward_dynamics_gradient.py:48-57).  Stage A (one thread per state and
    component) computes those once and writes them to a scratch array; stage B (one thread
    per state and GROUP of columns) reads them back and finishes its columns.

Scratch layout: [tile of 32 states][word][lane] - every access of a warp is one 128-byte
line, word offsets are compile-time immediates.  Output columns are staged per warp in shared
memory and flushed as runs of n contiguous floats per state.

A Program traced here names its inputs "in:<w>" (word w of the state's input row
[in0 | in1]), "sc:<key>" (scratch, key resolved to a word by PipeVariant) and "gravity";
outputs are ("sc", word) and ("out", flat index in the state's output row).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

from .algorithms import (RneaResult, SymRobot, cross_motion_axis, fd_prologue, minv, minv_get, rnea,
                         rnea_grad_columns, vjp_column)
from .ir import Program, V, dot
from .robot import Robot

# variant -> (struct name, IN0 / n, IN1 / n, OUT as f(n))
PIPE_VARIANTS = {
    "id":          ("PipeId",         2, 0, lambda n: n),
    "id_qdd":      ("PipeIdQdd",      2, 1, lambda n: n),
    "minv":        ("PipeMinv",       1, 0, lambda n: n * n),
    "fd":          ("PipeFd",         3, 0, lambda n: n),
    "id_grad":     ("PipeIdGrad",     2, 0, lambda n: 2 * n * n),
    "id_grad_qdd": ("PipeIdGradQdd",  2, 1, lambda n: 2 * n * n),
    "fd_grad":     ("PipeFdGrad",     3, 0, lambda n: 2 * n * n),
    # USE_QDD_MINV_FLAG overload (algorithms/_forward_dynamics_gradient.py:22-25, 202-220): in0 = [q | qd], in1 = qdd;
    # the caller's Minv (n*n per state, column-major, upper triangle) is a third input that the state programs read
    # straight from global memory ("g2:<index>") and hand to the column programs through the scratch array
    "fd_grad_qdd_minv": ("PipeFdGradPre", 2, 1, lambda n: 2 * n * n),
    # consumers fused after the FD gradient (algorithms.trace_fd_consumer documents the outputs);
    # in1 = lam = [lam_q | lam_v]
    "fd_vjp":      ("PipeFdVjp",      3, 2, lambda n: 5 * n),
    "fd_lin":      ("PipeFdLin",      3, 0, lambda n: 2 * n + 3 * n * n),
}
FD_LIKE = ("fd_grad", "fd_vjp", "fd_lin")          # programs that start with RNEA(0), Minv, qdd
GRAD_LIKE = ("id_grad", "fd_grad_pre") + FD_LIKE   # programs with du-columns (two-stage when large)


def _given_minv(p: Program, n: int, ids: Sequence[int]):
    """The caller's Minv restricted to one component, {(r, c): V} for local r <= c, read from the upper triangle of
    the column-major n x n input (entries that couple different components are structural zeros and never read)."""
    return {(r, c): p.inp("g2:%d" % (ids[c] * n + ids[r])) for c in range(len(ids)) for r in range(c + 1)}


def components(robot: Robot) -> List[List[int]]:
    """Joint ids of each tree hanging off the base (contiguous ranges: ids are a DFS pre-order)."""
    out = []
    for i in range(robot.n):
        if robot.parent[i] < 0:
            out.append(robot.get_subtree_by_id(i))
    return out


def subrobot(robot: Robot, ids: Sequence[int], name: Optional[str] = None) -> Robot:
    m = {g: l for l, g in enumerate(ids)}
    return Robot(name or "%s_%d_%d" % (robot.name, ids[0], ids[-1]),
                 [m.get(robot.parent[g], -1) for g in ids], [robot.S_ind[g] for g in ids],
                 [robot.E0[g].copy() for g in ids], [robot.r0[g].copy() for g in ids],
                 [robot.Imats[g].copy() for g in ids], [robot.damping[g] for g in ids])


class PipeTask:
    """One traced program = one thread per (state, task)."""

    def __init__(self, name: str, stage: int, program: Program, runs: List[Tuple[Tuple[int, ...], int]],
                 comp: int = 0):
        self.name, self.stage, self.program, self.runs, self.comp = name, stage, program, runs, comp
        self.counts = program.op_counts()
        self.cost = self.counts["nodes"]              # refined by emit_task (instructions, incl. loads and flushes)

    def program_reads_only_scratch(self) -> bool:
        p = self.program
        live = p.live_nodes()
        return all(k[1] in ("gravity", "dt") or k[1].startswith("sc:")
                   for i, k in enumerate(p.nodes) if live[i] and k[0] == "in")

    @property
    def flops(self) -> int:
        return self.counts["flops"]


# ---- stage A: per-component state programs ---------------------------------------------------
class _Exports:
    """Values stage A can hand to stage B, by name.  Constants (structural zeros of the root
    joints, ...) are folded into the importing program instead of being stored."""

    def __init__(self):
        self.cand: Dict[str, V] = {}

    def add(self, name: str, v: V):
        self.cand[name] = v

    def importer(self, p: Program):
        def imp(name: str) -> V:
            v = self.cand[name]
            if v.is_const:
                return p.const(v.c)
            x = p.inp("sc:%d" % v.i)          # keyed by the producing node: aliases share a word
            return x if v.s > 0 else -x
        return imp


def _state_inputs(p: Program, n: int, ids: Sequence[int], block: int) -> List[V]:
    return [p.inp("in:%d" % (block * n + g)) for g in ids]


def _fd_prologue(p: Program, S: SymRobot, qd, u, g, lam_v=None):
    """RNEA(qdd = 0), Minv, qdd = Minv (u - c) of one component (+ w = Minv lam_v for the fused VJP)."""
    _, Mi, qdd, extra = fd_prologue(S, qd, u, g, [lam_v] if lam_v is not None else [])
    return (Mi, qdd, extra[0]) if lam_v is not None else (Mi, qdd)


def _xnext_outputs(p: Program, n: int, ids: Sequence[int], q, qd, qdd, dt):
    """x+ = [q + dt qd ; qd + dt qdd] rows of one component: two runs of nc words."""
    for l, gid in enumerate(ids):
        p.output("out", gid, q[l] + dt * qd[l])
        p.output("out", n + gid, qd[l] + dt * qdd[l])
    return ((ids[0], n + ids[0]), len(ids))


def _b2_outputs(p: Program, n: int, ids: Sequence[int], Mi, dt):
    """Columns of B2 = dt Minv owned by one component (zeros outside it: Minv is block-diagonal)."""
    nc, base, runs = len(ids), ids[0], []
    for j in range(nc):
        off = 2 * n + 2 * n * n + n * ids[j]
        for ig in range(n):
            l = ig - base
            p.output("out", off + ig, dt * minv_get(Mi, l, j) if 0 <= l < nc else 0.0)
        runs.append(((off,), n))
    return runs


def _trace_stage_a(robot: Robot, ids: Sequence[int], alg: str, use_qdd: bool):
    sub, n = subrobot(robot, ids), robot.n
    nc = sub.n
    p = Program()
    q, qd = _state_inputs(p, n, ids, 0), _state_inputs(p, n, ids, 1)
    g = p.inp("gravity")
    S = SymRobot(p, sub, q)
    Mi = None
    runs: List[Tuple[Tuple[int, ...], int]] = []
    ex = _Exports()
    if alg == "fd_vjp":
        u = _state_inputs(p, n, ids, 2)
        lam_q, lam_v = _state_inputs(p, n, ids, 3), _state_inputs(p, n, ids, 4)
        Mi, qdd, w = _fd_prologue(p, S, qd, u, g, lam_v)
    elif alg in FD_LIKE:
        u = _state_inputs(p, n, ids, 2)
        Mi, qdd = _fd_prologue(p, S, qd, u, g)
    elif alg == "fd_grad_pre":
        qdd = _state_inputs(p, n, ids, 2)                           # in1 follows the 2n words of in0
        Mi = _given_minv(p, n, ids)
    else:
        qdd = _state_inputs(p, n, ids, 2) if use_qdd else None      # in1 follows the 2n words of in0
    if alg in ("fd_vjp", "fd_lin"):
        dt = p.inp("dt")
        runs.append(_xnext_outputs(p, n, ids, q, qd, qdd, dt))
    if alg == "fd_vjp":
        for l, gid in enumerate(ids):
            p.output("out", 4 * n + gid, dt * w[l])
            ex.add("w%d" % l, w[l])
            ex.add("lq%d" % l, lam_q[l])
            ex.add("lv%d" % l, lam_v[l])
        runs.append(((4 * n + ids[0],), nc))
    if alg == "fd_lin":
        runs += _b2_outputs(p, n, ids, Mi, dt)
    R = rnea(S, qd, qdd, g)
    for i in range(nc):
        k = sub.S_ind[i]
        if k < 3:
            ex.add("sin%d" % i, S.sin[i])
            ex.add("cos%d" % i, S.cos[i])
        else:
            ex.add("q%d" % i, q[i])
        ex.add("qd%d" % i, qd[i])
        mXa = cross_motion_axis(p, k, R.Xa[i])
        mf = cross_motion_axis(p, k, R.f[i])
        for r in range(6):
            ex.add("v%d_%d" % (i, r), R.v[i][r])
            ex.add("Iv%d_%d" % (i, r), R.Iv[i][r])
            ex.add("mXa%d_%d" % (i, r), mXa[r])
            ex.add("mf%d_%d" % (i, r), mf[r])
    if Mi is not None and alg != "fd_vjp":          # the costate product needs w = Minv lam_v, not Minv
        for (r, c), v in Mi.items():
            ex.add("M%d_%d" % (r, c), v)
    return p, ex, sub, runs


class _ImportSource:
    def __init__(self, imp):
        self.imp = imp

    def _six(self, nm, i):
        return [self.imp("%s%d_%d" % (nm, i, r)) for r in range(6)]

    def v(self, i):
        return self._six("v", i)

    def Iv(self, i):
        return self._six("Iv", i)

    def mxs_Xa(self, i):
        return self._six("mXa", i)

    def mxs_f(self, i):
        return self._six("mf", i)


def _column_outputs(p: Program, n: int, ids: Sequence[int], j: int, cq, cqd, M):
    """Writes the two full output columns of local joint j (zeros outside the component)."""
    nc, base, jg = len(ids), ids[0], ids[j]
    offs = []
    for s, col in ((0, cq), (1, cqd)):
        if col is None:                              # this program computes only the other side
            continue
        if M is None:
            vals = col
        else:
            rows = sorted(col)
            vals = {i: -dot([M(i, r) for r in rows], [col[r] for r in rows]) for i in range(nc)}
        for ig in range(n):
            l = ig - base
            p.output("out", s * n * n + n * jg + ig, vals.get(l, 0.0) if 0 <= l < nc else 0.0)
        offs.append(s * n * n + n * jg)
    return (tuple(offs), n)


def _lin_column_outputs(p: Program, n: int, ids: Sequence[int], j: int, cq, cqd, M, dt):
    """Columns jg of A21 = dt dqdd/dq and A22 = I + dt dqdd/dqd (zeros / identity outside the component)."""
    nc, base, jg = len(ids), ids[0], ids[j]
    offs = []
    for s, col in ((0, cq), (1, cqd)):
        if col is None:
            continue
        rows = sorted(col)
        scaled = [col[r] * dt for r in rows]
        for ig in range(n):
            l = ig - base
            v = -dot([M(l, r) for r in rows], scaled) if 0 <= l < nc else p.const(0.0)
            if s == 1 and ig == jg:
                v = v + 1.0
            p.output("out", 2 * n + s * n * n + n * jg + ig, v)
        offs.append(2 * n + s * n * n + n * jg)
    return (tuple(offs), n)


def _consume_columns(p: Program, n: int, ids: Sequence[int], alg: str, columns, M, dt, w=None, lam_q=None, lam_v=None):
    """Turns the dc_du column pairs of `columns` into the output runs of `alg`."""
    runs, done, vjp_sides = [], [], (True, True)
    for j, cq, cqd in columns:
        if alg == "fd_vjp":
            aq, av = vjp_column(p, cq, cqd, w, lam_q[j], lam_v[j], dt)
            if aq is not None:
                p.output("out", 2 * n + ids[j], aq)
            if av is not None:
                p.output("out", 3 * n + ids[j], av)
            vjp_sides = (aq is not None, av is not None)
            done.append(j)
        elif alg == "fd_lin":
            runs.append(_lin_column_outputs(p, n, ids, j, cq, cqd, M, dt))
        else:
            runs.append(_column_outputs(p, n, ids, j, cq, cqd, M))
    if done:                                         # a group holds contiguous joints: one run (pair)
        assert done == list(range(done[0], done[0] + len(done)))
        offs = tuple(off for off, on in ((2 * n + ids[done[0]], vjp_sides[0]), (3 * n + ids[done[0]], vjp_sides[1])) if on)
        runs.append((offs, len(done)))
    return runs


def _trace_stage_b(robot: Robot, ids: Sequence[int], sub: Robot, joints: Sequence[int], alg: str, ex: _Exports,
                   sides: Sequence[int] = (0, 1)):
    n, nc = robot.n, sub.n
    p = Program()
    imp = ex.importer(p)
    rev = [sub.S_ind[i] < 3 for i in range(nc)]
    sin = [imp("sin%d" % i) if rev[i] else None for i in range(nc)]
    cos = [imp("cos%d" % i) if rev[i] else None for i in range(nc)]
    q = [None if rev[i] else imp("q%d" % i) for i in range(nc)]
    S = SymRobot(p, sub, q, trig=(sin, cos))
    qd = [imp("qd%d" % i) for i in range(nc)]
    M = (lambda r, c: imp("M%d_%d" % (min(r, c), max(r, c)))) if alg in ("fd_grad", "fd_lin", "fd_grad_pre") else None
    dt = p.inp("dt") if alg in ("fd_vjp", "fd_lin") else None
    kw = {}
    if alg == "fd_vjp":
        kw = dict(w=[imp("w%d" % i) for i in range(nc)], lam_q={j: imp("lq%d" % j) for j in joints},
                  lam_v={j: imp("lv%d" % j) for j in joints})
    runs = _consume_columns(p, n, ids, alg, rnea_grad_columns(S, qd, None, src=_ImportSource(imp), joints=joints,
                                                              sides=sides), M, dt, **kw)
    return p, runs


# ---- single-stage programs: one (state, component) per thread -----------------------------------
def _trace_full(robot: Robot, ids: Sequence[int], alg: str, use_qdd: bool):
    sub, n = subrobot(robot, ids), robot.n
    nc, base = sub.n, ids[0]
    p = Program()
    q = _state_inputs(p, n, ids, 0)
    S = SymRobot(p, sub, q)
    runs: List[Tuple[Tuple[int, ...], int]] = []
    if alg == "minv":
        Mi = minv(S)
        for j in range(nc):
            jg = ids[j]
            for ig in range(n):
                l = ig - base
                p.output("out", jg * n + ig, Mi[(l, j)] if 0 <= l <= j else 0.0)
            runs.append(((jg * n,), n))
        return p, runs
    qd = _state_inputs(p, n, ids, 1)
    g = p.inp("gravity")
    if alg == "id":
        qdd = _state_inputs(p, n, ids, 2) if use_qdd else None
        R = rnea(S, qd, qdd, g)
        for i in range(nc):
            p.output("out", ids[i], R.c[i])
        return p, [((base,), nc)]
    if alg == "fd":
        u = _state_inputs(p, n, ids, 2)
        _, qdd = _fd_prologue(p, S, qd, u, g)
        for i in range(nc):
            p.output("out", ids[i], qdd[i])
        return p, [((base,), nc)]
    M, dt, kw = None, None, {}
    if alg == "fd_vjp":
        u = _state_inputs(p, n, ids, 2)
        lam_q, lam_v = _state_inputs(p, n, ids, 3), _state_inputs(p, n, ids, 4)
        Mi, qdd, w = _fd_prologue(p, S, qd, u, g, lam_v)
    elif alg in FD_LIKE:
        u = _state_inputs(p, n, ids, 2)
        Mi, qdd = _fd_prologue(p, S, qd, u, g)
        M = lambda r, c: minv_get(Mi, r, c)
    elif alg == "fd_grad_pre":
        qdd = _state_inputs(p, n, ids, 2)
        Mi = _given_minv(p, n, ids)
        M = lambda r, c: minv_get(Mi, r, c)
    else:
        qdd = _state_inputs(p, n, ids, 2) if use_qdd else None
    if alg in ("fd_vjp", "fd_lin"):
        dt = p.inp("dt")
        runs.append(_xnext_outputs(p, n, ids, q, qd, qdd, dt))
    if alg == "fd_vjp":
        for l, gid in enumerate(ids):
            p.output("out", 4 * n + gid, dt * w[l])
        runs.append(((4 * n + ids[0],), nc))
        kw = dict(w=w, lam_q=lam_q, lam_v=lam_v)
    if alg == "fd_lin":
        runs += _b2_outputs(p, n, ids, Mi, dt)
    R = rnea(S, qd, qdd, g)
    runs += _consume_columns(p, n, ids, alg, rnea_grad_columns(S, qd, R), M, dt, **kw)
    return p, runs


class PipeVariant:
    """All tasks of one algorithm variant of one robot, plus its scratch layout."""

    def __init__(self, robot: Robot, variant: str, single_stage_max_flops: int = 12000,
                 group_flops: int = 6500, stage_a_max_flops: int = 20000, split_sides_above: int = 0,
                 struct_suffix: str = ""):
        """split_sides_above: a single joint's column pair that costs more than this many flops is traced as two
        programs, d/dq and d/dqd (0 = never)."""
        self.robot, self.variant = robot, variant
        sname, m0, m1, out_fn = PIPE_VARIANTS[variant]
        n = robot.n
        self.struct = sname + struct_suffix
        self.in0, self.in1, self.out = m0 * n, m1 * n, out_fn(n)
        alg = {"id_qdd": "id", "id_grad_qdd": "id_grad", "fd_grad_qdd_minv": "fd_grad_pre"}.get(variant, variant)
        use_qdd = variant.endswith("_qdd")
        self.in2 = n * n if alg == "fd_grad_pre" else 0
        self.tasks: List[PipeTask] = []
        self.scratch_words = 0
        self.feasible = True
        sc_base = 0
        for ci, ids in enumerate(components(robot)):
            # a gradient whose dense operation count is more than 12x the single-stage limit is two-stage for sure
            # (tracing folds 4-5x away): skip the full trace, which for a 64-link chain costs tens of seconds
            hopeless = False
            if alg in GRAD_LIKE:
                from .algorithms import algorithmic_flops
                dense = algorithmic_flops(subrobot(robot, ids))["id_grad" if alg in ("id_grad",) else "fd_grad"]
                hopeless = dense > 12 * max(single_stage_max_flops, 1)
            if not hopeless:
                pf, runs = _trace_full(robot, ids, alg, use_qdd)
                full = PipeTask("c%d_full" % ci, 0, pf, runs, ci)
                if alg not in GRAD_LIKE or full.flops <= single_stage_max_flops:
                    if full.flops > stage_a_max_flops:
                        self.feasible = False
                    self.tasks.append(full)
                    continue
            # two stages: A = state program, B = groups of du-columns
            pa, ex, sub, runs_a = _trace_stage_a(robot, ids, alg, use_qdd)
            nc = sub.n
            # cost of every single column pair, then greedy packing of contiguous joints
            cost = []
            for j in range(nc):
                pb, _ = _trace_stage_b(robot, ids, sub, [j], alg, ex)
                cost.append(pb.op_counts()["flops"])
                if cost[-1] > stage_a_max_flops:     # a single column too long for one thread (64-link chain): the
                    self.feasible = False            # variant is unusable, no point in tracing the other columns
                    break
            if not self.feasible:
                break
            groups, cur, acc = [], [], 0
            for j in range(nc):
                if cur and acc + cost[j] > group_flops:
                    groups.append(cur)
                    cur, acc = [], 0
                cur.append(j)
                acc += cost[j]
            if cur:
                groups.append(cur)
            btasks = []
            used: Dict[int, None] = {}
            for gi, J in enumerate(groups):
                too_long = split_sides_above and len(J) == 1 and cost[J[0]] > split_sides_above
                for sides in (((0,), (1,)) if too_long else ((0, 1),)):
                    pb, runs = _trace_stage_b(robot, ids, sub, J, alg, ex, sides)
                    live = pb.live_nodes()
                    for i, k in enumerate(pb.nodes):
                        if live[i] and k[0] == "in" and k[1].startswith("sc:"):
                            used[int(k[1][3:])] = None
                    btasks.append(("%d%s" % (gi, "" if len(sides) == 2 else "qd"[sides[0]:sides[0] + 1] or "q"), J, pb, runs))
            # scratch words in stage-A production order
            word_of = {node: sc_base + w for w, node in enumerate(sorted(used))}
            sc_base += len(word_of)
            for node, w in word_of.items():
                pa.output("sc", w, V(pa, i=node))
            ta = PipeTask("c%d_A" % ci, 0, pa, runs_a, ci)
            if ta.flops > stage_a_max_flops:
                self.feasible = False
            self.tasks.append(ta)
            for gi, J, pb, runs in btasks:
                t = PipeTask("c%d_B%s_j%d_%d" % (ci, gi, ids[J[0]], ids[J[-1]]), 1, pb, runs, ci)
                if t.flops > stage_a_max_flops:          # a single column too long for one thread (64-link chain)
                    self.feasible = False
                t.sc_word = {("sc:%d" % node): w for node, w in word_of.items()}
                self.tasks.append(t)
        self.scratch_words = sc_base
        # heavy tasks first: CTAs are dispatched in blockIdx order
        self.stage_tasks = [sorted([t for t in self.tasks if t.stage == s], key=lambda t: -t.flops) for s in (0, 1)]
        self.flops = sum(t.flops for t in self.tasks)

    def evaluate(self, in_rows, gravity: float = 9.81, dtype=None, dt: float = 0.0, in2_rows=None):
        """Interprets the task programs with numpy in kernel order (stage 0, then stage 1) - the
        host-side check of the decomposition (tests/test_pipeline.py); never a product path.
        in_rows: (N, IN0 + IN1) array; returns (N, OUT) with NaN in words no task wrote."""
        import numpy as np
        dtype = dtype or np.float64
        rows = np.asarray(in_rows, dtype=dtype)
        N = rows.shape[0]
        out = np.full((N, self.out), np.nan, dtype=dtype)
        scratch = np.full((N, max(1, self.scratch_words)), np.nan, dtype=dtype)
        for stage in (0, 1):
            for t in self.stage_tasks[stage]:
                inputs = {}
                for name in t.program.inputs:
                    if name == "gravity":
                        inputs[name] = dtype(gravity)
                    elif name == "dt":
                        inputs[name] = dtype(dt)
                    elif name.startswith("in:"):
                        inputs[name] = rows[:, int(name[3:])]
                    elif name.startswith("g2:"):
                        inputs[name] = np.asarray(in2_rows, dtype=dtype)[:, int(name[3:])]
                    else:
                        inputs[name] = scratch[:, t.sc_word[name]] if name in getattr(t, "sc_word", {}) else np.nan
                res = t.program.evaluate(inputs, dtype=dtype)
                for name, idx, _ in t.program.outputs:
                    (scratch if name == "sc" else out)[:, idx] = res[name][:, idx]
        return out

    def summary(self) -> Dict[str, object]:
        return {"flops": self.flops, "scratch_words": self.scratch_words,
                "tasks": [(t.name, t.stage, t.flops, getattr(t, "instr", 0)) for st in self.stage_tasks for t in st],
                "x2_live": [getattr(t, "x2_live", None) for t in self.stage_tasks[1]]}


# ---- emission ----------------------------------------------------------------------------------
def _flit(x: float) -> str:
    s = "%.9g" % x
    if "e" not in s and "." not in s and "inf" not in s and "nan" not in s:
        s += ".0"
    return s + "f"


def emit_task(t: PipeTask, fname: str, out_words: int, stage_pad: int, tile_lead: int = 24,
              scratch_lead: int = 160, indent: str = "        ", sync_every: int = 0, x2: bool = False,
              sc_stride: int = 32, remat_gap: int = 0) -> List[str]:
    """One task as a __device__ function.  Input loads are issued `lead` operations before
    their first use (long enough to cover the L2 latency of scratch reads, short enough not to
    pin registers); output runs are flushed as soon as their last value exists.

    x2: the lane runs the program for TWO consecutive states at once, every value a float2 computed with the
    sm_100 packed FP32 instructions (FFMA2 / FMUL2 / FADD2: literals ride as a broadcast immediate, negations as
    operand modifiers) - half the instructions per state for kernels that are bound by instruction supply.  Only
    programs whose inputs are scratch words (stage 1).  sc_stride: states per scratch word row (64 when the
    consumers are x2 programs)."""
    if x2:
        return _emit_task_x2(t, fname, out_words, stage_pad, scratch_lead, indent, sync_every, remat_gap)
    p = t.program
    live = p.live_nodes()
    sc_word = getattr(t, "sc_word", {})
    order = [i for i, k in enumerate(p.nodes) if live[i] and k[0] != "in"]
    pos = {i: c for c, i in enumerate(order)}
    first_use: Dict[int, int] = {}
    for i in order:
        for o in p.operands(i):
            if p.nodes[o][0] == "in" and o not in first_use:
                first_use[o] = pos[i]
    out_pos: Dict[int, List[Tuple[str, int, V]]] = {}
    for (name, idx, v) in p.outputs:
        if not v.is_const and p.nodes[v.i][0] == "in":
            first_use.setdefault(v.i, 0)
    loads_at: Dict[int, List[str]] = {}
    load_pos: Dict[int, int] = {}
    for o, fu in first_use.items():
        name = p.nodes[o][1]
        if name in ("gravity", "dt"):
            stmt, lead = "const float t%d = %s;" % (o, name), 0
        elif name.startswith("in:"):
            stmt, lead = "const float t%d = s_in[%d];" % (o, int(name[3:])), tile_lead
        elif name.startswith("g2:"):
            stmt, lead = "const float t%d = __ldg(g_in2 + %d);" % (o, int(name[3:])), scratch_lead
        else:
            stmt, lead = "const float t%d = pipe::ldsc(sc_in + %d);" % (o, sc_stride * sc_word[name]), scratch_lead
        load_pos[o] = max(0, fu - lead)
        loads_at.setdefault(load_pos[o], []).append(indent + stmt)

    # outputs: scratch words are stored where produced; "out" words wait for their run
    run_of: Dict[int, int] = {}
    for ri, (offs, ln) in enumerate(t.runs):
        for si, off in enumerate(offs):
            for c in range(ln):
                run_of[off + c] = ri
    run_vals: List[Dict[int, V]] = [dict() for _ in t.runs]
    run_last = [-1] * len(t.runs)
    sc_at: Dict[int, List[str]] = {}
    for (name, idx, v) in p.outputs:
        # constants go first; a passed-through input follows its own load
        where = -1 if v.is_const else (load_pos[v.i] if p.nodes[v.i][0] == "in" else pos[v.i])
        if name == "sc":
            expr = _flit(v.c) if v.is_const else "%st%d" % ("-" if v.s < 0 else "", v.i)
            sc_at.setdefault(where, []).append("%ssc_out[%d] = %s;" % (indent, sc_stride * idx, expr))
        else:
            ri = run_of[idx]
            if idx in run_vals[ri]:
                raise ValueError("output word %d written twice" % idx)
            run_vals[ri][idx] = v
            run_last[ri] = max(run_last[ri], where)
    flush_at: Dict[int, List[int]] = {}
    for ri, (offs, ln) in enumerate(t.runs):
        if len(run_vals[ri]) != ln * len(offs):
            raise ValueError("task %s does not cover run %d" % (t.name, ri))
        flush_at.setdefault(run_last[ri], []).append(ri)

    def flush(ri: int) -> List[str]:
        offs, ln = t.runs[ri]
        L = []
        for si, off in enumerate(offs):
            for c in range(ln):
                v = run_vals[ri][off + c]
                expr = _flit(v.c) if v.is_const else "%st%d" % ("-" if v.s < 0 else "", v.i)
                L.append("%ss_stage[%d] = %s;" % (indent, si * ln + c, expr))
        L.append("%s__syncwarp();" % indent)
        if len(offs) == 1:
            L.append("%spipe::flush1<%d, %d, %d>(g_tile, s_warp, %d, cnt, lane);" % (indent, out_words, ln, stage_pad, offs[0]))
        else:
            L.append("%spipe::flush2<%d, %d, %d>(g_tile, s_warp, %d, %d, cnt, lane);" % (
                indent, out_words, ln, stage_pad, offs[0], offs[1]))
        L.append("%s__syncwarp();" % indent)
        return L

    body: List[str] = ["    // %s: %d mul + %d add per state" % (t.name, t.counts["mul"], t.counts["add"]),
                       "    static __device__ __noinline__ void %s(const float *s_in, const float *__restrict__ sc_in,"
                       " float *__restrict__ sc_out, float *s_stage, float *__restrict__ g_tile, const int cnt,"
                       " const int lane, const float *s_warp, const float gravity, const float dt,"
                       " const float *__restrict__ g_in2) {" % fname,
                       # the call boundary hides the address space: without this the staging accesses
                       # compile to generic LD/ST instead of LDS/STS
                       indent + "__builtin_assume(__isShared(s_in)); __builtin_assume(__isShared(s_stage));"
                                " __builtin_assume(__isShared(s_warp));"]
    body += sc_at.get(-1, [])
    for ri in flush_at.get(-1, []):
        body += flush(ri)
    sincos_done = set()
    for c in range(max(1, len(order))):
        if sync_every and c and c % sync_every == 0:
            body.append(indent + "__syncthreads();")
        body += loads_at.get(c, [])
        i = order[c] if c < len(order) else None
        k = p.nodes[i] if i is not None else ("nop",)
        op = k[0]
        if op == "nop":
            pass
        elif op in ("sin", "cos"):
            a = k[1]
            if a not in sincos_done:
                sincos_done.add(a)
                si, ci = p._cse.get(("sin", a)), p._cse.get(("cos", a))
                sn = "t%d" % si if si is not None and live[si] else "unused_s%d" % a
                cn = "t%d" % ci if ci is not None and live[ci] else "unused_c%d" % a
                body.append("%sfloat %s, %s; sincosf(t%d, &%s, &%s);" % (indent, sn, cn, a, sn, cn))
        elif op == "rcp":
            body.append("%sconst float t%d = 1.0f / t%d;" % (indent, i, k[1]))
        elif op == "mul":
            body.append("%sconst float t%d = t%d * t%d;" % (indent, i, k[1], k[2]))
        elif op == "mulc":
            body.append("%sconst float t%d = t%d * %s;" % (indent, i, k[1], _flit(k[2])))
        elif op == "add":
            body.append("%sconst float t%d = t%d %s t%d;" % (indent, i, k[1], "+" if k[3] > 0 else "-", k[2]))
        elif op == "addc":
            body.append("%sconst float t%d = t%d + %s;" % (indent, i, k[1], _flit(k[2])))
        else:
            raise ValueError("pipe emitter: unsupported node %r" % (k,))
        body += sc_at.get(c, [])
        for ri in flush_at.get(c, []):
            body += flush(ri)
    body.append("    }")
    # instruction estimate for the SM partition of the fused kernel: operations, loads, scratch stores,
    # staging stores and the unrolled flush rows; code beyond the 64 KB the instruction cache keeps
    # runs ~3.4x slower per instruction (profiles/r1_ifetch_regions.md)
    n_instr = len(order) + len(first_use) + sum(len(v) for v in sc_at.values())
    n_instr += sum(len(offs) * ln + 48 for (offs, ln) in t.runs)
    n_sass = int(0.8 * n_instr)                  # mul+add pairs fuse into FFMA
    t.cost = int(min(n_sass, 4096) + 3.4 * max(0, n_sass - 4096))
    t.instr = n_instr
    return body


def x2_clusters(t: PipeTask, remat_gap: int):
    """Use positions (in emission order) of every live input of the task, clustered: a new cluster starts where two
    consecutive uses are more than remat_gap operations apart (0 = one cluster).  -> (order, pos, {input: [[pos...]]})"""
    p = t.program
    live = p.live_nodes()
    order = [i for i, k in enumerate(p.nodes) if live[i] and k[0] != "in"]
    pos = {i: c for c, i in enumerate(order)}
    use_pos: Dict[int, List[int]] = {}
    for i in order:
        for o in p.operands(i):
            if p.nodes[o][0] == "in":
                use_pos.setdefault(o, []).append(pos[i])
    for (name, idx, v) in p.outputs:
        if not v.is_const and p.nodes[v.i][0] == "in":
            use_pos.setdefault(v.i, []).insert(0, 0)
    clusters: Dict[int, List[List[int]]] = {}
    for o, us in use_pos.items():
        us = sorted(set(us))
        cl = [[us[0]]]
        for u in us[1:]:
            if remat_gap and u - cl[-1][-1] > remat_gap and p.nodes[o][1].startswith("sc:"):
                cl.append([u])
            else:
                cl[-1].append(u)
        clusters[o] = cl
    return order, pos, clusters


def x2_max_live(t: PipeTask, scratch_lead: int, remat_gap: int) -> int:
    """Values alive at once when the task is emitted in program order with its loads `scratch_lead` operations ahead
    of each cluster of uses: the estimate that decides whether the packed form (two registers per value) fits."""
    p = t.program
    order, pos, clusters = x2_clusters(t, remat_gap)
    last = {}
    for i in order:
        for o in p.operands(i):
            if p.nodes[o][0] != "in":
                last[o] = pos[i]
    delta = [0] * (len(order) + 2)
    for o, b in last.items():
        delta[pos[o]] += 1
        delta[b + 1] -= 1
    for o, cls in clusters.items():
        for cl in cls:
            delta[max(0, cl[0] - scratch_lead)] += 1
            delta[cl[-1] + 1] -= 1
    cur = best = 0
    for d in delta:
        cur += d
        best = max(best, cur)
    return best


def _emit_task_x2(t: PipeTask, fname: str, out_words: int, stage_pad: int, scratch_lead: int, indent: str,
                  sync_every: int, remat_gap: int = 0) -> List[str]:
    """The two-states-per-lane (float2) form of emit_task; see there.  The scratch rows hold 64 states, the lane
    reads words 2*lane, 2*lane+1 with one 64-bit load.  nvcc does not contract the packed intrinsics, so a product
    whose only use is a sum is fused into FFMA2 here.  Output runs go through the warp's 32 staging rows twice:
    first the even states of the 64-state tile (.x), then the odd ones (.y).

    remat_gap > 0: a scratch word whose uses lie more than remat_gap operations apart is loaded again for each
    cluster of uses instead of occupying a register pair in between (M^-1 entries are used by every column of a
    group; packed values cost two registers each and the programs were sized for 255 scalar registers)."""
    p = t.program
    live = p.live_nodes()
    sc_word = getattr(t, "sc_word", {})
    order, pos, clusters = x2_clusters(t, remat_gap)
    uses = [0] * len(p.nodes)
    for i in order:
        for o in p.operands(i):
            uses[o] += 1
    for (name, idx, v) in p.outputs:
        if not v.is_const:
            uses[v.i] += 1
    # product -> the sum that absorbs it
    fused_into: Dict[int, int] = {}
    for i in order:
        k = p.nodes[i]
        if k[0] == "add":
            for cand in (k[2], k[1]):
                if p.nodes[cand][0] in ("mul", "mulc") and uses[cand] == 1 and cand not in fused_into and live[cand]:
                    fused_into[cand] = i
                    break
    loads_at: Dict[int, List[str]] = {}
    load_pos: Dict[int, int] = {}
    n_loads = 0
    for o, cls in clusters.items():
        name = p.nodes[o][1]
        for ci, cl in enumerate(cls):
            var = "t%d" % o if ci == 0 else "t%dr%d" % (o, ci)
            if name in ("gravity", "dt"):
                stmt, lead = "const float2 %s = make_float2(%s, %s);" % (var, name, name), 0
            elif name.startswith("sc:"):
                stmt, lead = "const float2 %s = pipe::ldsc2(sc_in + %d);" % (var, 64 * sc_word[name]), scratch_lead
            else:
                raise ValueError("x2 task %s reads %r: only scratch words and scalars are supported" % (t.name, name))
            at = max(0, cl[0] - lead)
            if ci == 0:
                load_pos[o] = at
            loads_at.setdefault(at, []).append(indent + stmt)
            n_loads += 1
    cur = [0]

    def nm(o: int) -> str:
        """name of node o at the current position (inputs: the cluster that covers it)"""
        cls = clusters.get(o)
        if not cls or len(cls) == 1:
            return "t%d" % o
        ci = 0
        while ci + 1 < len(cls) and cls[ci + 1][0] <= cur[0]:
            ci += 1
        return "t%d" % o if ci == 0 else "t%dr%d" % (o, ci)

    run_of: Dict[int, int] = {}
    for ri, (offs, ln) in enumerate(t.runs):
        for off in offs:
            for c in range(ln):
                run_of[off + c] = ri
    run_vals: List[Dict[int, V]] = [dict() for _ in t.runs]
    run_last = [-1] * len(t.runs)
    for (name, idx, v) in p.outputs:
        if name == "sc":
            raise ValueError("x2 task %s writes scratch words" % t.name)
        where = -1 if v.is_const else (load_pos[v.i] if p.nodes[v.i][0] == "in" else pos[v.i])
        ri = run_of[idx]
        if idx in run_vals[ri]:
            raise ValueError("output word %d written twice" % idx)
        run_vals[ri][idx] = v
        run_last[ri] = max(run_last[ri], where)
    flush_at: Dict[int, List[int]] = {}
    for ri, (offs, ln) in enumerate(t.runs):
        if len(run_vals[ri]) != ln * len(offs):
            raise ValueError("task %s does not cover run %d" % (t.name, ri))
        flush_at.setdefault(run_last[ri], []).append(ri)

    def flush(ri: int) -> List[str]:
        offs, ln = t.runs[ri]
        L = []
        for h, half in enumerate("xy"):
            for si, off in enumerate(offs):
                for c in range(ln):
                    v = run_vals[ri][off + c]
                    expr = _flit(v.c) if v.is_const else "%st%d.%s" % ("-" if v.s < 0 else "", v.i, half)
                    L.append("%ss_stage[%d] = %s;" % (indent, si * ln + c, expr))
            L.append("%s__syncwarp();" % indent)
            cnt_h = "(cnt + 1) >> 1" if h == 0 else "cnt >> 1"
            if len(offs) == 1:
                L.append("%spipe::flush1<%d, %d, %d>(g_tile + %d, s_warp, %d, %s, lane);" % (
                    indent, 2 * out_words, ln, stage_pad, h * out_words, offs[0], cnt_h))
            else:
                L.append("%spipe::flush2<%d, %d, %d>(g_tile + %d, s_warp, %d, %d, %s, lane);" % (
                    indent, 2 * out_words, ln, stage_pad, h * out_words, offs[0], offs[1], cnt_h))
            L.append("%s__syncwarp();" % indent)
        return L

    def bc(c: float) -> str:
        f = _flit(c)
        return "make_float2(%s, %s)" % (f, f)

    def sg(i: int, negate: bool) -> str:
        return "pipe::neg2(%s)" % nm(i) if negate else nm(i)

    def second(kc, negate: bool) -> str:
        """second factor of the product node kc, optionally negated"""
        return sg(kc[2], negate) if kc[0] == "mul" else bc(-kc[2] if negate else kc[2])

    body: List[str] = ["    // %s (two states per lane): %d mul + %d add per state" % (t.name, t.counts["mul"], t.counts["add"]),
                       "    static __device__ __noinline__ void %s(const float *s_in, const float *__restrict__ sc_in,"
                       " float *__restrict__ sc_out, float *s_stage, float *__restrict__ g_tile, const int cnt,"
                       " const int lane, const float *s_warp, const float gravity, const float dt,"
                       " const float *__restrict__ g_in2) {" % fname,
                       indent + "__builtin_assume(__isShared(s_stage)); __builtin_assume(__isShared(s_warp));"]
    for ri in flush_at.get(-1, []):
        body += flush(ri)
    for c in range(max(1, len(order))):
        cur[0] = c
        if sync_every and c and c % sync_every == 0:
            body.append(indent + "__syncthreads();")
        body += loads_at.get(c, [])
        i = order[c] if c < len(order) else None
        k = p.nodes[i] if i is not None else ("nop",)
        op = k[0]
        if op == "nop" or i in fused_into:
            pass
        elif op in ("sin", "cos"):
            body.append("%sconst float2 t%d = make_float2(%sf(%s.x), %sf(%s.y));" % (indent, i, op, nm(k[1]), op, nm(k[1])))
        elif op == "rcp":
            body.append("%sconst float2 t%d = make_float2(1.0f / %s.x, 1.0f / %s.y);" % (indent, i, nm(k[1]), nm(k[1])))
        elif op in ("mul", "mulc"):
            body.append("%sconst float2 t%d = __fmul2_rn(%s, %s);" % (indent, i, nm(k[1]), second(k, False)))
        elif op == "add":
            a, b, rel = k[1], k[2], k[3]
            if fused_into.get(b) == i:            # a + rel * (x * y)
                kc = p.nodes[b]
                body.append("%sconst float2 t%d = __ffma2_rn(%s, %s, %s);" % (indent, i, nm(kc[1]), second(kc, rel < 0), nm(a)))
            elif fused_into.get(a) == i:          # (x * y) + rel * b
                kc = p.nodes[a]
                body.append("%sconst float2 t%d = __ffma2_rn(%s, %s, %s);" % (indent, i, nm(kc[1]), second(kc, False),
                                                                            sg(b, rel < 0)))
            else:
                body.append("%sconst float2 t%d = __fadd2_rn(%s, %s);" % (indent, i, nm(a), sg(b, rel < 0)))
        elif op == "addc":
            body.append("%sconst float2 t%d = __fadd2_rn(%s, %s);" % (indent, i, nm(k[1]), bc(k[2])))
        else:
            raise ValueError("pipe emitter: unsupported node %r" % (k,))
        for ri in flush_at.get(c, []):
            body += flush(ri)
    body.append("    }")
    # per-state instruction estimate (the packed program serves two states)
    n_instr = len(order) - len(fused_into) + n_loads
    n_instr += 2 * sum(len(offs) * ln + 48 for (offs, ln) in t.runs)
    n_instr //= 2
    t.cost = int(min(n_instr, 4096) + 3.4 * max(0, n_instr - 4096))
    t.instr = n_instr
    return body


def emit_pipe_struct(pv: PipeVariant, min_blocks: Tuple[int, int] = (1, 1), warps: int = 8,
                     sync_every: int = 256, scratch_lead: int = 160, chunk_states: int = 0,
                     x2=False, order_chunk_states: int = 0) -> Tuple[str, Dict[str, object]]:
    """min_blocks: resident CTAs per SM the two stage kernels are compiled for (register cap =
    65536 / (32 * warps * min_blocks)); warps: tiles (warps) per CTA.

    x2 (False, True or a dict of options): stage-1 programs of a two-stage variant run two states per lane with
    the packed FP32 instructions (emit_task) when their register demand allows it - at most `max_live` values alive
    at once (two registers each) with scratch words re-loaded for uses more than `remat_gap` operations apart and
    loads issued `lead` operations ahead.  Stage-1 warps then take 64-state tiles; the programs that stay scalar run
    a tile as two halves."""
    xo = dict(max_live=110, remat_gap=200, lead=80)
    if isinstance(x2, dict):
        xo.update(x2)
    x2 = bool(x2) and pv.scratch_words > 0 and all(t.stage == 0 or t.program_reads_only_scratch() for t in pv.tasks)
    x2_tasks = set()
    if x2:
        for ti, t in enumerate(pv.stage_tasks[1]):
            t.x2_live = x2_max_live(t, xo["lead"], xo["remat_gap"])
            if t.x2_live <= xo["max_live"]:
                x2_tasks.add(ti)
        x2 = bool(x2_tasks)
    max_run = max([len(offs) * ln for t in pv.tasks for (offs, ln) in t.runs] + [1])
    # row pitch of the staging tile: even (float2 flushes) when every run length is even, else odd
    # (conflict-free scalar access); never a multiple of 32
    if all(ln % 2 == 0 for t in pv.tasks for (_, ln) in t.runs):
        stage_pad = max_run + (2 if max_run % 32 == 0 else 0)
    else:
        stage_pad = max_run | 1
    txt = ["struct %s {" % pv.struct,
           "    static constexpr int IN0 = %d, IN1 = %d, IN2 = %d, OUT = %d, SCRATCH_WORDS = %d, STAGE_PAD = %d;" % (
               pv.in0, pv.in1, pv.in2, pv.out, pv.scratch_words, stage_pad),
           "    static constexpr int NTASKS0 = %d, NTASKS1 = %d, MINB0 = %d, MINB1 = %d, WARPS = %d;" % (
               len(pv.stage_tasks[0]), len(pv.stage_tasks[1]), min_blocks[0], min_blocks[1], warps),
           "    static constexpr long long TRACED_FLOPS = %d;" % pv.flops,
           "    static constexpr int SYNC_EVERY = %d;" % (sync_every if warps > 1 else 0),
           "    static constexpr int CHUNK_STATES = %d;" % chunk_states,
           "    static constexpr int ORDER_CHUNK_STATES = %d;   // stage-1 item order (grid_pipe.cuh); 0 = task-major" % order_chunk_states,
           "    // X2 = 1: stage-1 warps take 64-state tiles; bit t of X2_MASK: stage-1 task t is a two-states-per-lane",
           "    // (packed FP32) program, the others run the tile as two 32-state halves",
           "    static constexpr int X2 = %d;" % int(x2),
           "    static constexpr unsigned long long X2_MASK = %dull;" % sum(1 << ti for ti in x2_tasks)]
    if len(pv.stage_tasks[1]) > 64:
        raise ValueError("more than 64 stage-1 tasks")
    for s in (0, 1):
        for ti, t in enumerate(pv.stage_tasks[s]):
            is2 = x2 and s == 1 and ti in x2_tasks
            txt += emit_task(t, "s%d_t%d" % (s, ti), pv.out, stage_pad, sync_every=sync_every if warps > 1 else 0,
                             scratch_lead=xo["lead"] if is2 else scratch_lead, x2=is2, sc_stride=64 if x2 else 32,
                             remat_gap=xo["remat_gap"])
    txt.append("    template <int STAGE> static __device__ __forceinline__ void run(const int task, const float *s_in,"
               " const float *__restrict__ sc_in, float *__restrict__ sc_out, float *s_stage,"
               " float *__restrict__ g_tile, const int cnt, const int lane, const float *s_warp, const float gravity,"
               " const float dt, const float *__restrict__ g_in2) {")
    for s in (0, 1):
        if not pv.stage_tasks[s]:
            continue
        txt.append("        if (STAGE == %d) {" % s)
        txt.append("            switch (task) {")
        for ti in range(len(pv.stage_tasks[s])):
            txt.append("            case %d: s%d_t%d(s_in, sc_in, sc_out, s_stage, g_tile, cnt, lane, s_warp, gravity, dt, g_in2); break;"
                       % (ti, s, ti))
        txt.append("            default: break;")
        txt.append("            }")
        txt.append("        }")
    txt += ["    }"]
    # fused kernel (grid_pipe.cuh pipe_fused_kernel): tasks of both stages numbered stage 0 first
    allt = pv.stage_tasks[0] + pv.stage_tasks[1]
    a_index = {t.comp: i for i, t in enumerate(pv.stage_tasks[0])}
    txt.append("    static constexpr int NT = %d;" % len(allt))
    txt.append("    static constexpr int cost(int k) { return %s; }   // estimated instructions per tile" % (
        " ".join("k == %d ? %d :" % (i, t.cost) for i, t in enumerate(allt)) + " 0"))
    txt.append("    // stage-0 task whose scratch words task k reads (-1: none)")
    txt.append("    static __host__ __device__ constexpr int dep(int k) { return %s; }" % (
        " ".join("k == %d ? %d :" % (i, a_index[t.comp]) for i, t in enumerate(allt) if t.stage == 1) + " -1"))
    txt += ["};", ""]
    return "\n".join(txt), pv.summary()
