"""The five rigid-body-dynamics algorithms, traced symbolically per robot.

Each ``trace_*`` function builds a :class:`~gridcodegenerator_b200.ir.Program` whose
inputs are the per-state scalars (``q_i``, ``qd_i``, ``u_i``/``qdd_i``, ``gravity``) and
whose outputs are laid out exactly as the reference kernels write them
(SURVEY.md 8a a9): ``c[n]``, ``Minv[n*n]`` column-major upper-triangular, ``qdd[n]``,
``dc_du[2n*n]`` / ``df_du[2n*n]`` column-major ``n x 2n``.

What each function replaces in the reference:
  SymRobot              helpers/_topology_helpers.py:90-182 (per-q X update) and the
                        mx*/fx* device helpers of helpers/_spatial_algebra_helpers.py:35-256
  rnea / trace_id       algorithms/_inverse_dynamics.py:33-304  (oracle _test.py:5-115)
  minv / trace_minv     algorithms/_direct_minv.py:23-382       (oracle _test.py:117-226)
  trace_fd              algorithms/_forward_dynamics.py:21-112
  rnea_grad_columns     algorithms/_inverse_dynamics_gradient.py:27-650 (oracle _test.py:229-488)
  trace_fd_grad         algorithms/_forward_dynamics_gradient.py:7-57   (oracle _test.py:496-520)

Structure exploited at trace time (the reference multiplies dense 6x6 matrices):
X = [[E,0],[-E r^x,E]] is applied as a constant 3x3 (X_tree, usually a signed
permutation) followed by a planar rotation; spatial inertias and articulated
inertias are kept symmetric; topology (parents, subtrees, ancestor sets) is resolved
in Python so no index arithmetic reaches the GPU.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .ir import Program, V, V2, cross3, dot, matvec, vadd, vsub, vscale, zeros
from .robot import Robot


class SymRobot:
    """Per-state symbolic view of the robot: sin/cos of q and the X_i(q) actions."""

    def __init__(self, p: Program, robot: Robot, q: Sequence[V], trig=None):
        """trig = (sin list, cos list): take sin/cos of the revolute joints from the caller instead
        of computing them from q (stage-B programs of pipeline.py import them from scratch memory);
        q[i] is then only read for prismatic joints."""
        self.p, self.robot, self.n = p, robot, robot.n
        self.q = list(q)
        self.sin: List[Optional[V]] = []
        self.cos: List[Optional[V]] = []
        self.r: List[List[V]] = []
        for i in range(self.n):
            k = robot.S_ind[i]
            E0, r0 = robot.E0[i], robot.r0[i]
            if k < 3:
                self.sin.append(trig[0][i] if trig is not None else p.sin(q[i]))
                self.cos.append(trig[1][i] if trig is not None else p.cos(q[i]))
                self.r.append([p.const(x) for x in r0])
            else:   # prismatic: r = r0 + q * E0[k-3, :]
                self.sin.append(None)
                self.cos.append(None)
                self.r.append([p.const(r0[t]) + q[i] * float(E0[k - 3, t]) for t in range(3)])
        self.I = [[[p.const(x) for x in row] for row in robot.Imats[i]] for i in range(self.n)]

    # -- rotations ---------------------------------------------------------------
    def _EJ(self, i: int, y: Sequence[V], transpose: bool) -> List[V]:
        k = self.robot.S_ind[i]
        if k >= 3:
            return list(y)
        a, b = (k + 1) % 3, (k + 2) % 3
        c, s = self.cos[i], self.sin[i]
        if transpose:
            s = -s
        out = list(y)
        out[a] = c * y[a] + s * y[b]
        out[b] = c * y[b] - s * y[a]
        return out

    def E(self, i: int, x: Sequence[V]) -> List[V]:
        E0 = self.robot.E0[i]
        y = [dot([self.p.const(E0[r, t]) for t in range(3)], x) for r in range(3)]
        return self._EJ(i, y, False)

    def ET(self, i: int, y: Sequence[V]) -> List[V]:
        E0 = self.robot.E0[i]
        z = self._EJ(i, y, True)
        return [dot([self.p.const(E0[t, r]) for t in range(3)], z) for r in range(3)]

    # -- spatial transforms --------------------------------------------------------
    def X_motion(self, i: int, v: Sequence[V]) -> List[V]:
        """X_i v for a motion vector (also how the reference moves F columns forward)."""
        w, l = v[:3], v[3:]
        return self.E(i, w) + self.E(i, vsub(l, cross3(self.r[i], w)))

    def XT_force(self, i: int, f: Sequence[V]) -> List[V]:
        """X_i^T f for a force vector."""
        fl = self.ET(i, f[3:])
        return vadd(self.ET(i, f[:3]), cross3(self.r[i], fl)) + fl

    def X_col(self, i: int, col: int, scale) -> List[V]:
        """scale * X_i[:, col]  (root acceleration: X[:,5]*gravity)."""
        e = zeros(self.p, 6)
        e[col] = self.p.lift(scale)
        return self.X_motion(i, e)

    def I_mul(self, i: int, v: Sequence[V]) -> List[V]:
        return matvec(self.I[i], v)

    def congruence(self, i: int, M: List[List[V]]) -> List[List[V]]:
        """X_i^T M X_i for a symmetric 6x6 M, returned symmetric (upper triangle
        computed, mirrored)."""
        p = self.p
        A = [[M[r][c] for c in range(3)] for r in range(3)]
        B = [[M[r][c + 3] for c in range(3)] for r in range(3)]
        C = [[M[r + 3][c + 3] for c in range(3)] for r in range(3)]

        def rot(Mb, symmetric):
            cols = [self.ET(i, [Mb[r][c] for r in range(3)]) for c in range(3)]   # E^T Mb, by column
            T1 = [[cols[c][r] for c in range(3)] for r in range(3)]
            rows = [self.ET(i, T1[r]) for r in range(3)]                           # (E^T Mb) E, by row
            if symmetric:
                for r in range(3):
                    for c in range(r):
                        rows[r][c] = rows[c][r]
            return rows

        A1, B1, C1 = rot(A, True), rot(B, False), rot(C, True)
        r = self.r[i]
        rx = [[p.const(0.0), -r[2], r[1]], [r[2], p.const(0.0), -r[0]], [-r[1], r[0], p.const(0.0)]]

        def mm(P, Q):
            return [[dot(P[a], [Q[t][b] for t in range(3)]) for b in range(3)] for a in range(3)]

        rxC = mm(rx, C1)
        B2 = [[B1[a][b] + rxC[a][b] for b in range(3)] for a in range(3)]
        B1rx = mm(B1, rx)
        rxB2T = mm(rx, [[B2[b][a] for b in range(3)] for a in range(3)])
        A2 = [[None] * 3 for _ in range(3)]
        for a in range(3):
            for b in range(a, 3):
                A2[a][b] = A1[a][b] - B1rx[a][b] + rxB2T[a][b]
                A2[b][a] = A2[a][b]
        out = [[None] * 6 for _ in range(6)]
        for a in range(3):
            for b in range(3):
                out[a][b] = A2[a][b]
                out[a][b + 3] = B2[a][b]
                out[b + 3][a] = B2[a][b]
                out[a + 3][b + 3] = C1[a][b]
        return out


def cross_motion_axis(p: Program, k: int, v: Sequence[V], alpha=None) -> List[V]:
    """(v x) e_k [* alpha]  - the reference's mx0..mx5 (helpers/_spatial_algebra_helpers.py:62-147)."""
    e = [p.const(1.0 if t == (k % 3) else 0.0) for t in range(3)]
    if k < 3:
        out = cross3(v[:3], e) + cross3(v[3:], e)
    else:
        out = zeros(p, 3) + cross3(v[:3], e)
    return out if alpha is None else vscale(out, alpha)


def cross_force(v: Sequence[V], f: Sequence[V]) -> List[V]:
    """v x* f  - fx_times_v (helpers/_spatial_algebra_helpers.py:181-256)."""
    return vadd(cross3(v[:3], f[:3]), cross3(v[3:], f[3:])) + cross3(v[:3], f[3:])


# ---- RNEA -------------------------------------------------------------------------
class RneaResult:
    __slots__ = ("c", "v", "a", "f", "Xa", "Iv", "f_fpass")


def rnea(sr: SymRobot, qd: Sequence[V], qdd: Optional[Sequence[V]], gravity: V) -> RneaResult:
    p, robot, n = sr.p, sr.robot, sr.n
    v: List[List[V]] = [None] * n
    a: List[List[V]] = [None] * n
    Xa: List[List[V]] = [None] * n
    f: List[List[V]] = [None] * n
    Iv: List[List[V]] = [None] * n
    for i in range(n):
        par, k = robot.parent[i], robot.S_ind[i]
        if par < 0:
            v[i] = zeros(p, 6)
            Xa[i] = sr.X_col(i, 5, gravity)
        else:
            v[i] = sr.X_motion(i, v[par])
            Xa[i] = sr.X_motion(i, a[par])
        v[i][k] = v[i][k] + qd[i]
        a[i] = list(Xa[i])
        if qdd is not None:
            a[i][k] = a[i][k] + qdd[i]
        if par >= 0:
            a[i] = vadd(a[i], cross_motion_axis(p, k, v[i], qd[i]))
        Iv[i] = sr.I_mul(i, v[i])
        f[i] = vadd(sr.I_mul(i, a[i]), cross_force(v[i], Iv[i]))
    f_fpass = [list(x) for x in f]            # forces before the backward accumulation (_test.py:5-76)
    c = [None] * n
    for i in range(n - 1, -1, -1):
        c[i] = f[i][robot.S_ind[i]] + qd[i] * robot.damping[i]
        par = robot.parent[i]
        if par >= 0:
            f[par] = vadd(f[par], sr.XT_force(i, f[i]))
    res = RneaResult()
    res.c, res.v, res.a, res.f, res.Xa, res.Iv, res.f_fpass = c, v, a, f, Xa, Iv, f_fpass
    return res


# ---- Minv ---------------------------------------------------------------------------
def minv(sr: SymRobot, bpass: Optional[dict] = None, on_final=None) -> Dict[Tuple[int, int], V]:
    """Upper-triangular M^-1 as {(row, col): V}, col >= row.  `bpass` (a dict) receives the state
    after the backward pass - Minv entries, F columns, U, Dinv - i.e. what the reference's
    test_minv_bpass returns (_test.py:117-184).  `on_final(i, j, m)` is called the moment entry (i, j)
    has its final value, so that consumers (qdd = Minv (u - c), ...) can use it right there in the trace:
    an entry that is consumed where it is produced does not occupy a register until the end of the pass."""
    p, robot, n = sr.p, sr.robot, sr.n
    zero6 = lambda: zeros(p, 6)
    Mi: Dict[Tuple[int, int], V] = {}
    F: List[Dict[int, List[V]]] = [dict() for _ in range(n)]       # F[i][col] -> 6-vector
    U: List[List[V]] = [None] * n
    Dinv: List[V] = [None] * n
    IA = [[list(row) for row in sr.I[i]] for i in range(n)]
    for i in range(n - 1, -1, -1):
        k, par = robot.S_ind[i], robot.parent[i]
        sub = robot.get_subtree_by_id(i)
        U[i] = [IA[i][r][k] for r in range(6)]
        Dinv[i] = p.rcp(U[i][k])
        for j in sub:
            Fij = F[i].get(j)
            Mi[(i, j)] = (Dinv[i] if j == i else p.const(0.0)) - (Dinv[i] * Fij[k] if Fij is not None else 0.0)
        if par >= 0:
            for j in sub:
                Fij = vadd(F[i].get(j, zero6()), vscale(U[i], Mi[(i, j)]))
                F[i][j] = Fij
                F[par][j] = vadd(F[par].get(j, zero6()), sr.XT_force(i, Fij))
            UD = vscale(U[i], Dinv[i])
            Ia = [[None] * 6 for _ in range(6)]
            for r in range(6):
                for c in range(r, 6):
                    Ia[r][c] = IA[i][r][c] - UD[r] * U[i][c]
                    Ia[c][r] = Ia[r][c]
            Ip = sr.congruence(i, Ia)
            IA[par] = [[IA[par][r][c] + Ip[r][c] for c in range(6)] for r in range(6)]
    if bpass is not None:
        bpass.update(Minv=dict(Mi), F=[dict(Fi) for Fi in F], U=list(U), Dinv=list(Dinv))
    for i in range(n):
        k, par = robot.S_ind[i], robot.parent[i]
        cols = range(i, n)
        if par >= 0:
            w = vscale(sr.XT_force(i, U[i]), Dinv[i])                 # Dinv * X^T U
            for j in cols:
                Fpj = F[par].get(j)
                if Fpj is not None:
                    Mi[(i, j)] = Mi.get((i, j), p.const(0.0)) - dot(w, Fpj)
        for j in cols:
            m = Mi.get((i, j), p.const(0.0))
            Mi[(i, j)] = m
            if on_final is not None:
                on_final(i, j, m)
            Fij = zero6()
            Fij[k] = m
            if par >= 0 and F[par].get(j) is not None:
                Fij = vadd(Fij, sr.X_motion(i, F[par][j]))
            F[i][j] = Fij
    return Mi


def minv_get(Mi: Dict[Tuple[int, int], V], r: int, c: int) -> V:
    return Mi[(r, c)] if r <= c else Mi[(c, r)]


class SymmetricProduct:
    """y = Minv x accumulated entry by entry in the order the Minv pass finalises its upper triangle
    (use as minv(..., on_final=acc)): y_i += M_ij x_j and, off the diagonal, y_j += M_ij x_i."""

    def __init__(self, p: Program, x: Sequence[V]):
        self.p, self.x = p, list(x)
        self.y: List[Optional[V]] = [None] * len(self.x)

    def _add(self, i: int, term: V):
        self.y[i] = term if self.y[i] is None else self.y[i] + term

    def __call__(self, i: int, j: int, m: V):
        self._add(i, m * self.x[j])
        if j != i:
            self._add(j, m * self.x[i])

    def result(self) -> List[V]:
        return [self.p.const(0.0) if v is None else v for v in self.y]


def fd_prologue(sr: SymRobot, qd: Sequence[V], u: Sequence[V], gravity: V, extra: Sequence[Sequence[V]] = ()):
    """The common head of FD and the FD gradient (algorithms/_forward_dynamics_gradient.py:9-20): bias forces
    c = RNEA(q, qd, 0), Minv, qdd = Minv (u - c).  The product rides on the Minv pass (SymmetricProduct).
    `extra`: more vectors to multiply by Minv in the same pass (the costate part lam_v of the fused VJP).
    Returns (R0, Mi, qdd, [Minv x for x in extra])."""
    p, n = sr.p, sr.n
    R0 = rnea(sr, qd, None, gravity)
    acc = [SymmetricProduct(p, [u[i] - R0.c[i] for i in range(n)])] + [SymmetricProduct(p, x) for x in extra]

    def on_final(i, j, m):
        for a in acc:
            a(i, j, m)

    Mi = minv(sr, on_final=on_final)
    return R0, Mi, acc[0].result(), [a.result() for a in acc[1:]]


# ---- RNEA gradient -------------------------------------------------------------------
class _DirectSource:
    """Per-joint state data of the gradient columns taken straight from a traced RNEA."""

    def __init__(self, p: Program, robot: Robot, R: RneaResult):
        self.p, self.robot, self.R = p, robot, R

    def v(self, i):
        return self.R.v[i]

    def Iv(self, i):
        return self.R.Iv[i]

    def mxs_Xa(self, i):
        return cross_motion_axis(self.p, self.robot.S_ind[i], self.R.Xa[i])

    def mxs_f(self, i):
        return cross_motion_axis(self.p, self.robot.S_ind[i], self.R.f[i])


def rnea_grad_columns(sr: SymRobot, qd: Sequence[V], R: Optional[RneaResult], src=None, joints=None, record=None,
                      sides: Sequence[int] = (0, 1)):
    """Yields (j, dc_dq_col, dc_dqd_col) one du-column pair at a time; each col is a
    dict {row i: V} over anc(j) | sub(j) (structural zeros elsewhere).  Columns are
    independent through both passes, which is what lets the emitter finish and store
    one column before starting the next.  `src` (default: the RNEA result itself) is where
    the per-joint state data v, I v, mxS(X a_parent), mxS(f) come from; `joints` restricts
    the columns (pipeline.py traces one group of columns per program).  `record(j, name, i, s, vec)`
    is called with the per-joint intermediates of column j (names "dv", "da", "df_fp", "df": the arrays
    the reference's test_rnea_grad_inner returns, _test.py:229-488).  `sides` = (0,) or (1,) computes only
    the d/dq or only the d/dqd column of each joint (the other one is yielded as None): the two recursions
    share nothing but the state data, which lets pipeline.py halve a column program that is too long."""
    p, robot, n = sr.p, sr.robot, sr.n
    sides = tuple(sides)
    if src is None:
        src = _DirectSource(p, robot, R)
    for j in (range(n) if joints is None else joints):
        kj = robot.S_ind[j]
        sub = robot.get_subtree_by_id(j)
        dv = {0: {}, 1: {}}
        da = {0: {}, 1: {}}
        df = {0: {}, 1: {}}
        for i in sub:
            k = robot.S_ind[i]
            vi, Ivi = src.v(i), src.Iv(i)
            if i == j:
                if 0 in sides:
                    dv[0][i] = cross_motion_axis(p, k, vi)              # == mxS(X v_parent)
                    da[0][i] = vadd(cross_motion_axis(p, k, dv[0][i], qd[i]), src.mxs_Xa(i))
                if 1 in sides:
                    e = zeros(p, 6)
                    e[k] = p.const(1.0)
                    dv[1][i] = e
                    da[1][i] = cross_motion_axis(p, k, vi)
            else:
                par = robot.parent[i]
                for s in sides:
                    dv[s][i] = sr.X_motion(i, dv[s][par])
                    da[s][i] = vadd(sr.X_motion(i, da[s][par]), cross_motion_axis(p, k, dv[s][i], qd[i]))
            for s in sides:
                df[s][i] = vadd(vadd(sr.I_mul(i, da[s][i]), cross_force(dv[s][i], Ivi)),
                                cross_force(vi, sr.I_mul(i, dv[s][i])))
                if record is not None:
                    record(j, "dv", i, s, dv[s][i])
                    record(j, "da", i, s, da[s][i])
                    record(j, "df_fp", i, s, df[s][i])
        cols = ({}, {})
        for i in reversed(sub):
            par = robot.parent[i]
            for s in sides:
                cols[s][i] = df[s][i][robot.S_ind[i]]
                if record is not None:
                    record(j, "df", i, s, df[s][i])
            if i != j:
                for s in sides:
                    df[s][par] = vadd(df[s][par], sr.XT_force(i, df[s][i]))
        # leave the subtree: the dq column also carries -X_j^T (f_j x) S_j
        up = [vsub(df[0][j], src.mxs_f(j)) if 0 in sides else None, df[1][j] if 1 in sides else None]
        i = j
        while robot.parent[i] >= 0:
            up = [sr.XT_force(i, up[s]) if s in sides else None for s in (0, 1)]
            i = robot.parent[i]
            for s in sides:
                cols[s][i] = up[s][robot.S_ind[i]]
                if record is not None:
                    record(j, "df", i, s, up[s])
        if 1 in sides:
            cols[1][j] = cols[1][j] + robot.damping[j]
        yield j, (cols[0] if 0 in sides else None), (cols[1] if 1 in sides else None)


# ---- traced programs -------------------------------------------------------------------
def _inputs(p: Program, n: int, names: Sequence[str]):
    return [[p.inp("%s%d" % (nm, i)) for i in range(n)] for nm in names]


def trace_id(robot: Robot, use_qdd: bool = False) -> Program:
    p = Program()
    n = robot.n
    q, qd = _inputs(p, n, ("q", "qd"))
    qdd = _inputs(p, n, ("qdd",))[0] if use_qdd else None
    g = p.inp("gravity")
    R = rnea(SymRobot(p, robot, q), qd, qdd, g)
    for i in range(n):
        p.output("c", i, R.c[i])
    return p


def trace_minv(robot: Robot) -> Program:
    p = Program()
    n = robot.n
    (q,) = _inputs(p, n, ("q",))
    Mi = minv(SymRobot(p, robot, q))
    for col in range(n):
        for row in range(n):
            p.output("Minv", col * n + row, Mi[(row, col)] if row <= col else 0.0)
    return p


def trace_fd(robot: Robot) -> Program:
    p = Program()
    n = robot.n
    q, qd, u = _inputs(p, n, ("q", "qd", "u"))
    g = p.inp("gravity")
    sr = SymRobot(p, robot, q)
    _, _, qdd, _ = fd_prologue(sr, qd, u, g)
    for i in range(n):
        p.output("qdd", i, qdd[i])
    return p


def trace_id_grad(robot: Robot, use_qdd: bool = False) -> Program:
    p = Program()
    n = robot.n
    q, qd = _inputs(p, n, ("q", "qd"))
    qdd = _inputs(p, n, ("qdd",))[0] if use_qdd else None
    g = p.inp("gravity")
    sr = SymRobot(p, robot, q)
    R = rnea(sr, qd, qdd, g)
    for j, cq, cqd in rnea_grad_columns(sr, qd, R):
        for i in range(n):
            p.output("dc_du", n * j + i, cq.get(i, 0.0))
        for i in range(n):
            p.output("dc_du", n * n + n * j + i, cqd.get(i, 0.0))
    return p


def trace_fd_grad(robot: Robot, use_qdd_minv: bool = False, side: Optional[int] = None) -> Program:
    """df_du = -Minv dc_du.  side = 0 / 1: only the d/dq / d/dqd block (n*n words, output indices 0 .. n*n-1) - the
    two halves of the mid-size-batch kernel (csrc/grid_tps.cuh tps_half_kernel), each with its own copy of the
    column-independent part.  use_qdd_minv=False: inputs (q, qd, u), everything computed
    in one program.  True: inputs (q, qd, qdd, Minv) - the USE_QDD_MINV_FLAG overload
    (algorithms/_forward_dynamics_gradient.py:22-25, 202-220); Minv is read
    symmetrically from its upper triangle (algorithms/_forward_dynamics.py:44)."""
    p = Program()
    n = robot.n
    q, qd = _inputs(p, n, ("q", "qd"))
    g = p.inp("gravity")
    sr = SymRobot(p, robot, q)
    if use_qdd_minv:
        qdd = _inputs(p, n, ("qdd",))[0]
        Mi = {(r, c): p.inp("Minv%d" % (c * n + r)) for r in range(n) for c in range(r, n)}
    else:
        (u,) = _inputs(p, n, ("u",))
        _, Mi, qdd, _ = fd_prologue(sr, qd, u, g)
    R = rnea(sr, qd, qdd, g)
    for j, cq, cqd in rnea_grad_columns(sr, qd, R, sides=(0, 1) if side is None else (side,)):
        for s, col in ((0, cq), (1, cqd)):
            if col is None:
                continue
            rows = sorted(col)
            for i in range(n):
                acc = dot([minv_get(Mi, i, r) for r in rows], [col[r] for r in rows])
                p.output("df_du", (s * n * n if side is None else 0) + n * j + i, -acc)
    return p


# ---- consumers fused after the FD gradient ------------------------------------
# The reference stops at writing df_du (2n^2 floats per state) to global memory and shipping it to
# the host (algorithms/_forward_dynamics_gradient.py:159-161, 235-238).  Trajectory optimisers do
# not want df_du itself but the explicit-Euler integrator Jacobians built from it,
#     x = [q; qd],  x+ = x + dt [qd; qdd(q, qd, u)]
#     A = dx+/dx = [[I, dt I], [dt dqdd/dq, I + dt dqdd/dqd]],   B = dx+/du = [[0], [dt Minv]],
# or only their action on a costate.  Both are traced here INTO the gradient program, so the
# gradient never leaves the registers of the thread that computed it.
#   "fd_vjp": inputs (q, qd, u), lam = [lam_q; lam_v] (2n), dt
#             out[5n] = [x+ (2n) | A^T lam (2n) | B^T lam (n)]            -- O(n) words per state
#             A^T lam = [lam_q + dt (dqdd/dq)^T lam_v ; lam_v + dt lam_q + dt (dqdd/dqd)^T lam_v]
#             with (df_du)^T lam_v = -dc_du^T (Minv lam_v): Minv is applied ONCE (n^2), not per column
#   "fd_lin": inputs (q, qd, u), dt
#             out[2n + 3n^2] = [x+ (2n) | A21 = dt dqdd/dq | A22 = I + dt dqdd/dqd | B2 = dt Minv],
#             the three n x n blocks column-major (B2 full symmetric): the non-constant blocks of A, B
def consumer_out_words(alg: str, n: int) -> int:
    return {"fd_vjp": 5 * n, "fd_lin": 2 * n + 3 * n * n}[alg]


def vjp_column(p: Program, cq: Dict[int, V], cqd: Dict[int, V], w: Sequence[V], lam_q_j: V, lam_v_j: V, dt: V):
    """(A^T lam)[j], (A^T lam)[n + j] from the dc_du columns of joint j and w = Minv lam_v (a column that was not
    computed - None - gives None)."""
    aq = av = None
    if cq is not None:
        rows = sorted(cq)
        aq = lam_q_j - dt * dot([cq[r] for r in rows], [w[r] for r in rows])
    if cqd is not None:
        rows = sorted(cqd)
        av = lam_v_j + dt * (lam_q_j - dot([cqd[r] for r in rows], [w[r] for r in rows]))
    return aq, av


def trace_fd_consumer(robot: Robot, alg: str) -> Program:
    """Thread-per-state program of a fused consumer ("fd_vjp" or "fd_lin")."""
    p = Program()
    n = robot.n
    q, qd, u = _inputs(p, n, ("q", "qd", "u"))
    g, dt = p.inp("gravity"), p.inp("dt")
    sr = SymRobot(p, robot, q)
    lam = [p.inp("lam%d" % i) for i in range(2 * n)] if alg == "fd_vjp" else None
    _, Mi, qdd, extra = fd_prologue(sr, qd, u, g, [lam[n:]] if lam else [])
    R = rnea(sr, qd, qdd, g)
    for i in range(n):
        p.output(alg, i, q[i] + dt * qd[i])
        p.output(alg, n + i, qd[i] + dt * qdd[i])
    if alg == "fd_vjp":
        w = extra[0]
        for j, cq, cqd in rnea_grad_columns(sr, qd, R):
            aq, av = vjp_column(p, cq, cqd, w, lam[j], lam[n + j], dt)
            p.output(alg, 2 * n + j, aq)
            p.output(alg, 3 * n + j, av)
        for i in range(n):
            p.output(alg, 4 * n + i, dt * w[i])
        return p
    if alg != "fd_lin":
        raise ValueError(alg)
    base = 2 * n
    for j, cq, cqd in rnea_grad_columns(sr, qd, R):
        for s, col in ((0, cq), (1, cqd)):
            rows = sorted(col)
            scaled = [col[r] * dt for r in rows]
            for i in range(n):
                v = -dot([minv_get(Mi, i, r) for r in rows], scaled)
                if s == 1 and i == j:
                    v = v + 1.0
                p.output(alg, base + s * n * n + n * j + i, v)
    for j in range(n):
        for i in range(n):
            p.output(alg, base + 2 * n * n + n * j + i, dt * minv_get(Mi, i, j))
    return p


# ---- extra traces used by the emitted header's _inner functions and the facade's test_* methods ----
def trace_id_full(robot: Robot, use_qdd: bool = False) -> Program:
    """c plus the reference's s_vaf block [v(6n) | a(6n) | f(6n)]."""
    p = Program()
    n = robot.n
    q, qd = _inputs(p, n, ("q", "qd"))
    qdd = _inputs(p, n, ("qdd",))[0] if use_qdd else None
    g = p.inp("gravity")
    R = rnea(SymRobot(p, robot, q), qd, qdd, g)
    for i in range(n):
        p.output("c", i, R.c[i])
    for i in range(n):
        for r in range(6):
            p.output("vaf", 6 * i + r, R.v[i][r])
            p.output("vaf", 6 * n + 6 * i + r, R.a[i][r])
            p.output("vaf", 12 * n + 6 * i + r, R.f[i][r])
    return p


def trace_rnea_fpass(robot: Robot, use_qdd: bool = False) -> Program:
    """(v, a, f) after the forward pass only - the reference's test_rnea_fpass (_test.py:5-76):
    output "vaf" = [v(6n) | a(6n) | f before the backward accumulation (6n)], joint-major."""
    p = Program()
    n = robot.n
    q, qd = _inputs(p, n, ("q", "qd"))
    qdd = _inputs(p, n, ("qdd",))[0] if use_qdd else None
    R = rnea(SymRobot(p, robot, q), qd, qdd, p.inp("gravity"))
    for i in range(n):
        for r in range(6):
            p.output("vaf", 6 * i + r, R.v[i][r])
            p.output("vaf", 6 * n + 6 * i + r, R.a[i][r])
            p.output("vaf", 12 * n + 6 * i + r, R.f_fpass[i][r])
    return p


def trace_minv_bpass(robot: Robot) -> Program:
    """State after the backward pass of the Minv algorithm - the reference's test_minv_bpass
    (_test.py:117-184): outputs "Minv" (n*n, row-major [row][col]; only the entries the pass writes:
    col in subtree(row)), "F" (n*6*n, [joint][row][col]), "U" (n*6), "Dinv" (n); everything the
    reference leaves at zero is a constant 0 here."""
    p = Program()
    n = robot.n
    (q,) = _inputs(p, n, ("q",))
    snap: dict = {}
    minv(SymRobot(p, robot, q), bpass=snap)
    for i in range(n):
        for j in range(n):
            p.output("Minv", i * n + j, snap["Minv"].get((i, j), 0.0))
            Fij = snap["F"][i].get(j)
            for r in range(6):
                p.output("F", (i * 6 + r) * n + j, Fij[r] if Fij is not None else 0.0)
        for r in range(6):
            p.output("U", i * 6 + r, snap["U"][i][r])
        p.output("Dinv", i, snap["Dinv"][i])
    return p


GRAD_INNER_ARRAYS = ("dv_dq", "dv_dqd", "da_dq", "da_dqd", "df_fp_dq", "df_fp_dqd", "df_dq", "df_dqd")


def trace_id_grad_inner(robot: Robot) -> Program:
    """Everything the reference's test_rnea_grad_inner returns (_test.py:229-488) from (q, qd, v, a, f):
    "dc_du" plus the eight 6 x n x n arrays of GRAD_INNER_ARRAYS, flat [row][col][joint]
    (d<x>_joint / du_col); entries outside the ancestor/subtree structure are constant zeros."""
    p = Program()
    n = robot.n
    q, qd = _inputs(p, n, ("q", "qd"))
    sr = SymRobot(p, robot, q)
    R = RneaResult()
    R.v = [[p.inp("vaf%d" % (6 * i + r)) for r in range(6)] for i in range(n)]
    R.a = [[p.inp("vaf%d" % (6 * n + 6 * i + r)) for r in range(6)] for i in range(n)]
    R.f = [[p.inp("vaf%d" % (12 * n + 6 * i + r)) for r in range(6)] for i in range(n)]
    R.Iv = [sr.I_mul(i, R.v[i]) for i in range(n)]
    R.Xa = []
    for i in range(n):
        if robot.parent[i] >= 0:
            R.Xa.append(vsub(R.a[i], cross_motion_axis(p, robot.S_ind[i], R.v[i], qd[i])))
        else:
            R.Xa.append(list(R.a[i]))
    R.c = None
    got: Dict[Tuple[str, int, int, int], V] = {}

    def record(j, name, i, s, vec):
        arr = "%s_%s" % (name, "dq" if s == 0 else "dqd")
        for r in range(6):
            got[(arr, r, j, i)] = vec[r]

    for j, cq, cqd in rnea_grad_columns(sr, qd, R, record=record):
        for i in range(n):
            p.output("dc_du", n * j + i, cq.get(i, 0.0))
            p.output("dc_du", n * n + n * j + i, cqd.get(i, 0.0))
    for arr in GRAD_INNER_ARRAYS:
        for r in range(6):
            for j in range(n):
                for i in range(n):
                    p.output(arr, (r * n + j) * n + i, got.get((arr, r, j, i), 0.0))
    return p


def trace_fd_finish(robot: Robot) -> Program:
    """qdd = Minv (u - c) with Minv read symmetrically from its upper triangle
    (algorithms/_forward_dynamics.py:21-49)."""
    p = Program()
    n = robot.n
    (u,) = _inputs(p, n, ("u",))
    c = [p.inp("c%d" % i) for i in range(n)]
    Mi = {(r, cc): p.inp("Minv%d" % (cc * n + r)) for r in range(n) for cc in range(r, n)}
    umc = [u[i] - c[i] for i in range(n)]
    for i in range(n):
        p.output("qdd", i, dot([minv_get(Mi, i, j) for j in range(n)], umc))
    return p


def trace_id_grad_from_vaf(robot: Robot) -> Program:
    """dc_du from (q, qd, v, a, f) - the reference's inverse_dynamics_gradient_inner contract
    (algorithms/_inverse_dynamics_gradient.py:27-41).  X a_parent is recovered from a_i:
    mxS(X a_p) = mxS(a_i - mxS(v_i) qd_i) because mxS_k(e_k) = 0."""
    p = Program()
    n = robot.n
    q, qd = _inputs(p, n, ("q", "qd"))
    sr = SymRobot(p, robot, q)
    R = RneaResult()
    R.v = [[p.inp("vaf%d" % (6 * i + r)) for r in range(6)] for i in range(n)]
    R.a = [[p.inp("vaf%d" % (6 * n + 6 * i + r)) for r in range(6)] for i in range(n)]
    R.f = [[p.inp("vaf%d" % (12 * n + 6 * i + r)) for r in range(6)] for i in range(n)]
    R.Iv = [sr.I_mul(i, R.v[i]) for i in range(n)]
    R.Xa = []
    for i in range(n):
        if robot.parent[i] >= 0:
            R.Xa.append(vsub(R.a[i], cross_motion_axis(p, robot.S_ind[i], R.v[i], qd[i])))
        else:
            R.Xa.append(list(R.a[i]))
    R.c = None
    for j, cq, cqd in rnea_grad_columns(sr, qd, R):
        for i in range(n):
            p.output("dc_du", n * j + i, cq.get(i, 0.0))
        for i in range(n):
            p.output("dc_du", n * n + n * j + i, cqd.get(i, 0.0))
    return p


# ---- lane-uniform column program (latency kernels, csrc/grid_cps.cuh) -------------------------
def rnea_grad_generic_column(sr: SymRobot, qd: Sequence[V], R: RneaResult, mq: Sequence[V], mqd: Sequence[V]):
    """One du-column of dc_du written so that EVERY column runs the same instruction stream:
    mq[i] / mqd[i] are 0/1 values selecting "this column is d/dq_i" / "d/dqd_i".  All joints are
    visited; columns that do not reach a joint carry exact zeros through it.  Returns dc[0..n)."""
    p, robot, n = sr.p, sr.robot, sr.n
    dv: List[List[V]] = [None] * n
    da: List[List[V]] = [None] * n
    df: List[List[V]] = [None] * n
    for i in range(n):
        k, par = robot.S_ind[i], robot.parent[i]
        ek = zeros(p, 6)
        ek[k] = p.const(1.0)
        mv = cross_motion_axis(p, k, R.v[i])
        sdv = vadd(vscale(mv, mq[i]), vscale(ek, mqd[i]))
        sda = vadd(vscale(cross_motion_axis(p, k, R.Xa[i]), mq[i]), vscale(mv, mqd[i]))
        if par >= 0:
            dv[i] = vadd(sr.X_motion(i, dv[par]), sdv)
            da[i] = vadd(vadd(sr.X_motion(i, da[par]), cross_motion_axis(p, k, dv[i], qd[i])), sda)
        else:
            dv[i] = sdv
            da[i] = vadd(cross_motion_axis(p, k, dv[i], qd[i]), sda)
        df[i] = vadd(vadd(sr.I_mul(i, da[i]), cross_force(dv[i], R.Iv[i])),
                     cross_force(R.v[i], sr.I_mul(i, dv[i])))
    dc: List[V] = [None] * n
    for i in range(n - 1, -1, -1):
        k, par = robot.S_ind[i], robot.parent[i]
        dc[i] = df[i][k] + mqd[i] * robot.damping[i]
        if par >= 0:
            leaving = vsub(df[i], vscale(cross_motion_axis(p, k, R.f[i]), mq[i]))
            df[par] = vadd(df[par], sr.XT_force(i, leaving))
    return dc


def trace_column_program(robot: Robot, alg: str, use_qdd: bool = False) -> Program:
    """alg = "id_grad": inputs (q, qd[, qdd], masks) -> col[0..n) = dc_du[:, column];
    alg = "fd_grad": inputs (q, qd, u, masks) -> col = -Minv dc_du[:, column]."""
    p = Program()
    n = robot.n
    q, qd = _inputs(p, n, ("q", "qd"))
    g = p.inp("gravity")
    sr = SymRobot(p, robot, q)
    Mi = None
    if alg == "fd_grad":
        (u,) = _inputs(p, n, ("u",))
        _, Mi, qdd, _ = fd_prologue(sr, qd, u, g)
    else:
        qdd = _inputs(p, n, ("qdd",))[0] if use_qdd else None
    R = rnea(sr, qd, qdd, g)
    mq = [p.inp("mq%d" % i) for i in range(n)]
    mqd = [p.inp("mqd%d" % i) for i in range(n)]
    dc = rnea_grad_generic_column(sr, qd, R, mq, mqd)
    for i in range(n):
        if Mi is None:
            p.output("col", i, dc[i])
        else:
            p.output("col", i, -dot([minv_get(Mi, i, r) for r in range(n)], dc))
    return p


class ColumnSource:
    """Where a gradient column reads the per-joint state data (v, I v, mxS(X a_parent), mxS(f)) from.
    Default: the values traced by rnea() themselves (they stay live in registers from the RNEA to the
    last column that needs them).  With `park` set, the named groups are parked in per-lane shared
    memory after the RNEA and RE-LOADED by every column, which cuts ~100-150 long-lived registers so
    that more warps fit on an SM (the straight-line kernels are bound by per-warp instruction
    supply, profiles/r1b)."""

    def __init__(self, sr: "SymRobot", R: "RneaResult", park: Sequence[str] = ()):
        p, robot, n = sr.p, sr.robot, sr.n
        self.sr, self.R, self.p = sr, R, p
        self.park = set(park)
        self.mXa = [cross_motion_axis(p, robot.S_ind[i], R.Xa[i]) for i in range(n)]
        self.mf = [cross_motion_axis(p, robot.S_ind[i], R.f[i]) for i in range(n)]
        self.h = {}
        for name, vals in (("v", R.v), ("Iv", R.Iv), ("mXa", self.mXa), ("mf", self.mf)):
            if name in self.park:
                self.h[name] = [[p.park(x) for x in vals[i]] for i in range(n)]
        self.cache = {}

    def begin_column(self):
        self.cache = {}

    def _get(self, name, i, direct):
        if name not in self.park:
            return direct
        key = (name, i)
        if key not in self.cache:
            self.cache[key] = [self.p.unpark(h) for h in self.h[name][i]]
        return self.cache[key]

    def v(self, i):
        return self._get("v", i, self.R.v[i])

    def Iv(self, i):
        if "Iv" in self.park or "v" not in self.park:
            return self._get("Iv", i, self.R.Iv[i])
        key = ("Iv", i)                      # recompute I v from the re-loaded v
        if key not in self.cache:
            self.cache[key] = self.sr.I_mul(i, self.v(i))
        return self.cache[key]

    def mxs_Xa(self, i):
        return self._get("mXa", i, self.mXa[i])

    def mxs_f(self, i):
        return self._get("mf", i, self.mf[i])


# ---- paired gradient columns: (d/dq_j, d/dqd_j) travel together as float2 ------------------------
def rnea_grad_columns_paired(sr: SymRobot, qd: Sequence[V], R: RneaResult, src: Optional[ColumnSource] = None):
    """Same recursion as rnea_grad_columns, with the two columns of joint j carried as ONE pair
    value per component (ir.V2): every operation on them multiplies by the same state-level
    scalar, which is exactly what the packed FFMA2/FMUL2/FADD2 instructions of sm_100 offer."""
    p, robot, n = sr.p, sr.robot, sr.n
    src = src or ColumnSource(sr, R)
    for j in range(n):
        src.begin_column()
        sub = robot.get_subtree_by_id(j)
        dv, da, df = {}, {}, {}
        for i in sub:
            k = robot.S_ind[i]
            vi = src.v(i)
            if i == j:
                mv = cross_motion_axis(p, k, vi)
                e = zeros(p, 6)
                e[k] = p.const(1.0)
                da_q = vadd(cross_motion_axis(p, k, mv, qd[i]), src.mxs_Xa(i))
                dv[i] = [p.pack(mv[r], e[r]) for r in range(6)]
                da[i] = [p.pack(da_q[r], mv[r]) for r in range(6)]
            else:
                par = robot.parent[i]
                dv[i] = sr.X_motion(i, dv[par])
                da[i] = vadd(sr.X_motion(i, da[par]), cross_motion_axis(p, k, dv[i], qd[i]))
            df[i] = vadd(vadd(sr.I_mul(i, da[i]), cross_force(dv[i], src.Iv(i))),
                         cross_force(vi, sr.I_mul(i, dv[i])))
        cols = {}
        for i in reversed(sub):
            cols[i] = df[i][robot.S_ind[i]]
            if i != j:
                par = robot.parent[i]
                df[par] = vadd(df[par], sr.XT_force(i, df[i]))
        mf = src.mxs_f(j)
        up = [df[j][r] - p.pack(mf[r], 0.0) for r in range(6)]
        i = j
        while robot.parent[i] >= 0:
            up = sr.XT_force(i, up)
            i = robot.parent[i]
            cols[i] = up[robot.S_ind[i]]
        cols[j] = cols[j] + p.pack(0.0, robot.damping[j])
        yield j, cols


def trace_id_grad_paired(robot: Robot, use_qdd: bool = False, park: Sequence[str] = ()) -> Program:
    p = Program()
    n = robot.n
    q, qd = _inputs(p, n, ("q", "qd"))
    qdd = _inputs(p, n, ("qdd",))[0] if use_qdd else None
    g = p.inp("gravity")
    sr = SymRobot(p, robot, q)
    R = rnea(sr, qd, qdd, g)
    for j, cols in rnea_grad_columns_paired(sr, qd, R, ColumnSource(sr, R, park)):
        for i in range(n):
            c = p.lift2(cols.get(i, 0.0))
            p.output("dc_du", n * j + i, c.x())
            p.output("dc_du", n * n + n * j + i, c.y())
    return p


def trace_fd_grad_paired(robot: Robot, use_qdd_minv: bool = False, park: Sequence[str] = ()) -> Program:
    p = Program()
    n = robot.n
    q, qd = _inputs(p, n, ("q", "qd"))
    g = p.inp("gravity")
    sr = SymRobot(p, robot, q)
    if use_qdd_minv:
        qdd = _inputs(p, n, ("qdd",))[0]
        Mi = {(r, c): p.inp("Minv%d" % (c * n + r)) for r in range(n) for c in range(r, n)}
    else:
        (u,) = _inputs(p, n, ("u",))
        _, Mi, qdd, _ = fd_prologue(sr, qd, u, g)
    R = rnea(sr, qd, qdd, g)
    src = ColumnSource(sr, R, park)
    Mh = {k: p.park(v) for k, v in Mi.items()} if "Minv" in park else None
    for j, cols in rnea_grad_columns_paired(sr, qd, R, src):
        rows = sorted(cols)
        Mj = {k: p.unpark(h) for k, h in Mh.items()} if Mh else Mi
        for i in range(n):
            acc = p.lift2(-dot([minv_get(Mj, i, r) for r in rows], [cols[r] for r in rows]))
            p.output("df_du", n * j + i, acc.x())
            p.output("df_du", n * n + n * j + i, acc.y())
    return p


# ---- further algorithms: mass matrix and O(n) forward dynamics ---------------------
def crba(sr: SymRobot) -> Dict[Tuple[int, int], V]:
    """Composite-rigid-body algorithm: the joint-space mass matrix M(q) as {(row, col): V}, row <= col (entries between
    joints of different branches are structural zeros and absent).  The reference has no CRBA; M is pinned to it
    through its RNEA (column j of M = RNEA(q, 0, e_j) - RNEA(q, 0, 0), tests/golden) and through M Minv = I."""
    p, robot, n = sr.p, sr.robot, sr.n
    Ic = [[list(row) for row in sr.I[i]] for i in range(n)]
    M: Dict[Tuple[int, int], V] = {}
    # one sweep, leaves first: the composite inertia of joint i is final once its children are done - column i of M
    # (F = Ic_i S_i carried up the ancestors) is produced right there and Ic_i dies as soon as it is folded into its
    # parent, so only the open branches hold registers (30 composite inertias kept to a second loop: 5.4 KB of spills
    # per thread on Atlas)
    for i in range(n - 1, -1, -1):
        k, par = robot.S_ind[i], robot.parent[i]
        F = [Ic[i][r][k] for r in range(6)]
        M[(i, i)] = F[k]
        j = i
        while robot.parent[j] >= 0:
            F = sr.XT_force(j, F)
            j = robot.parent[j]
            M[(j, i)] = F[robot.S_ind[j]]
        if par >= 0:
            Ip = sr.congruence(i, Ic[i])
            Ic[par] = [[Ic[par][r][c] + Ip[r][c] for c in range(6)] for r in range(6)]
        Ic[i] = None
    return M


def aba(sr: SymRobot, qd: Sequence[V], u: Sequence[V], gravity: V) -> List[V]:
    """Articulated-body algorithm: qdd = FD(q, qd, u) in O(n), without M^-1.  Same conventions as rnea() above
    (base acceleration X[:,5] * gravity, joint damping); equals the reference's Minv (u - c) up to rounding."""
    p, robot, n = sr.p, sr.robot, sr.n
    v: List[List[V]] = [None] * n
    cb: List[Optional[List[V]]] = [None] * n
    pA: List[List[V]] = [None] * n
    for i in range(n):
        par, k = robot.parent[i], robot.S_ind[i]
        v[i] = zeros(p, 6) if par < 0 else sr.X_motion(i, v[par])
        v[i][k] = v[i][k] + qd[i]
        cb[i] = cross_motion_axis(p, k, v[i], qd[i]) if par >= 0 else None
        pA[i] = cross_force(v[i], sr.I_mul(i, v[i]))
    IA = [[list(row) for row in sr.I[i]] for i in range(n)]
    U: List[List[V]] = [None] * n
    Dinv: List[V] = [None] * n
    uu: List[V] = [None] * n
    for i in range(n - 1, -1, -1):
        par, k = robot.parent[i], robot.S_ind[i]
        U[i] = [IA[i][r][k] for r in range(6)]
        Dinv[i] = p.rcp(U[i][k])
        uu[i] = u[i] - qd[i] * robot.damping[i] - pA[i][k]
        if par >= 0:
            UD = vscale(U[i], Dinv[i])
            Ia = [[None] * 6 for _ in range(6)]
            for r in range(6):
                for c in range(r, 6):
                    Ia[r][c] = IA[i][r][c] - UD[r] * U[i][c]
                    Ia[c][r] = Ia[r][c]
            pa = vadd(vadd(pA[i], matvec(Ia, cb[i])), vscale(UD, uu[i]))
            Ip = sr.congruence(i, Ia)
            IA[par] = [[IA[par][r][c] + Ip[r][c] for c in range(6)] for r in range(6)]
            pA[par] = vadd(pA[par], sr.XT_force(i, pa))
    a: List[List[V]] = [None] * n
    qdd: List[V] = [None] * n
    for i in range(n):
        par, k = robot.parent[i], robot.S_ind[i]
        ap = sr.X_col(i, 5, gravity) if par < 0 else vadd(sr.X_motion(i, a[par]), cb[i])
        qdd[i] = Dinv[i] * (uu[i] - dot(U[i], ap))
        a[i] = list(ap)
        a[i][k] = a[i][k] + qdd[i]
    return qdd


def trace_crba(robot: Robot) -> Program:
    """M(q), n x n column-major, BOTH triangles (unlike Minv's upper-triangular contract)."""
    p = Program()
    n = robot.n
    (q,) = _inputs(p, n, ("q",))
    M = crba(SymRobot(p, robot, q))
    for col in range(n):
        for row in range(n):
            p.output("M", col * n + row, M.get((min(row, col), max(row, col)), 0.0))
    return p


def trace_aba(robot: Robot) -> Program:
    p = Program()
    n = robot.n
    q, qd, u = _inputs(p, n, ("q", "qd", "u"))
    g = p.inp("gravity")
    qdd = aba(SymRobot(p, robot, q), qd, u, g)
    for i in range(n):
        p.output("qdd", i, qdd[i])
    return p


TRACERS = {
    "id": lambda robot: trace_id(robot, False),
    "id_qdd": lambda robot: trace_id(robot, True),
    "minv": trace_minv,
    "fd": trace_fd,
    "id_grad": lambda robot: trace_id_grad(robot, False),
    "id_grad_qdd": lambda robot: trace_id_grad(robot, True),
    "fd_grad": lambda robot: trace_fd_grad(robot, False),
    "fd_grad_qdd_minv": lambda robot: trace_fd_grad(robot, True),
    "fd_grad_q": lambda robot: trace_fd_grad(robot, False, side=0),
    "fd_grad_qd": lambda robot: trace_fd_grad(robot, False, side=1),
    "fd_vjp": lambda robot: trace_fd_consumer(robot, "fd_vjp"),
    "fd_lin": lambda robot: trace_fd_consumer(robot, "fd_lin"),
    "crba": trace_crba,
    "aba": trace_aba,
}

# packed (float2) variants of the gradient programs
PAIRED_TRACERS = {
    "id_grad": lambda robot: trace_id_grad_paired(robot, False),
    "id_grad_qdd": lambda robot: trace_id_grad_paired(robot, True),
    "fd_grad": lambda robot: trace_fd_grad_paired(robot, False),
    "fd_grad_qdd_minv": lambda robot: trace_fd_grad_paired(robot, True),
}


# ---- algorithmic (dense-reference) work, SURVEY.md 8d -------------------------------------
def algorithmic_flops(robot: Robot) -> Dict[str, int]:
    """Dense 6x6 op count of the reference algorithm restricted to topological
    non-zero columns (SURVEY.md 8d closed forms)."""
    n = robot.n
    anc = [len(robot.get_ancestors_by_id(i)) for i in range(n)]
    sub = [len(robot.get_subtree_by_id(i)) for i in range(n)]
    nonroot = [i for i in range(n) if robot.parent[i] >= 0]
    L, n0 = len(nonroot), n - len(nonroot)
    A = sum(anc)
    D = A + n
    B = sum(anc[i] + sub[i] for i in nonroot)
    XI = 36 * n
    ID = 168 * n + 224 * L + 7 * n0
    MINV = n + 2 * sum(sub) + sum(936 + 78 * sub[i] for i in nonroot) + 80 * sum(n - i for i in nonroot)
    IDG = 456 * n + 276 * A + 356 * D + 168 * B
    return {
        "id": XI + ID,
        "minv": XI + MINV,
        "fd": XI + MINV + ID + 3 * n * n,
        "id_grad": XI + ID + IDG,
        "fd_grad": XI + MINV + 2 * ID + 3 * n * n + IDG + 4 * n ** 3,
    }


def algorithmic_bytes(robot: Robot) -> Dict[str, int]:
    n = robot.n
    return {"id": 12 * n, "minv": 4 * (n + n * n), "fd": 16 * n, "id_grad": 4 * (2 * n + 2 * n * n),
            "fd_grad": 4 * (3 * n + 2 * n * n)}
