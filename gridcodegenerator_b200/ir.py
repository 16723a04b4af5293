"""Scalar SSA expression graph used to specialise the dynamics algorithms per robot.

The reference specialises per robot by printing sympy matrices into strings and
looking up topology tables at run time (helpers/_topology_helpers.py:154-172,
260-332).  Here the whole per-state computation is traced once, in Python, into a
hash-consed DAG of float operations in which every robot constant (X_tree entries,
inertias, topology) is a literal.  Multiplications by 0/+-1 and additions of 0
disappear at trace time, identical sub-expressions are shared (so e.g. the two RNEA
passes of the FD gradient share their velocity recursion - the TODO at
algorithms/_forward_dynamics_gradient.py:11-14), negations ride on operand signs,
and dead values are never emitted.  The CUDA emitter (codegen.py) prints the DAG as
straight-line sm_100a code; ``Program.evaluate`` interprets the same DAG with numpy
so the traced algorithms can be unit-tested without a GPU.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence, Tuple, Union

import numpy as np

Number = Union[int, float]


class V:
    """A value: either a compile-time constant or a signed reference to a node."""
    __slots__ = ("p", "c", "i", "s")

    def __init__(self, p: "Program", c=None, i: int = -1, s: int = 1):
        self.p = p
        self.c = c      # python float when constant
        self.i = i      # node index otherwise
        self.s = s      # +1 / -1

    @property
    def is_const(self) -> bool:
        return self.c is not None

    def is_zero(self) -> bool:
        return self.c is not None and self.c == 0.0

    # arithmetic ---------------------------------------------------------------
    def __neg__(self):
        if self.is_const:
            return V(self.p, c=-self.c)
        return V(self.p, i=self.i, s=-self.s)

    def __add__(self, o):
        return self.p.add(self, self.p.lift(o))

    __radd__ = __add__

    def __sub__(self, o):
        return self.p.add(self, -self.p.lift(o))

    def __rsub__(self, o):
        return self.p.add(self.p.lift(o), -self)

    def __mul__(self, o):
        return self.p.mul(self, self.p.lift(o))

    __rmul__ = __mul__

    def __repr__(self):
        return "V(c=%r)" % self.c if self.is_const else "V(%s t%d)" % ("+" if self.s > 0 else "-", self.i)


class Program:
    """Append-only list of nodes.  Node kinds:
         ('in', name)            leaf input
         ('sin', a) ('cos', a)   a = node index (always positive sign)
         ('rcp', a)              1 / a
         ('mul', a, b)           t[a] * t[b]
         ('mulc', a, const)      t[a] * const           (const > 0, != 1)
         ('add', a, b, sb)       t[a] + sb * t[b]
         ('addc', a, const)      t[a] + const
    Operand signs are normalised so that every node is the "positive" form; the
    referencing V carries the sign."""

    def __init__(self):
        self.nodes: List[tuple] = []
        self._cse: Dict[tuple, int] = {}
        self.outputs: List[Tuple[str, int, V]] = []   # (array name, flat index, value)
        self.inputs: Dict[str, int] = {}

    # construction ---------------------------------------------------------------
    def lift(self, x) -> V:
        if isinstance(x, V):
            return x
        return V(self, c=float(x))

    def const(self, x: Number) -> V:
        return V(self, c=float(x))

    def _node(self, key: tuple) -> int:
        i = self._cse.get(key)
        if i is None:
            i = len(self.nodes)
            self.nodes.append(key)
            self._cse[key] = i
        return i

    def inp(self, name: str) -> V:
        i = self._node(("in", name))
        self.inputs[name] = i
        return V(self, i=i)

    def sin(self, a: V) -> V:
        if a.is_const:
            return self.const(np.sin(a.c))
        return V(self, i=self._node(("sin", a.i)), s=a.s)           # odd function

    def cos(self, a: V) -> V:
        if a.is_const:
            return self.const(np.cos(a.c))
        return V(self, i=self._node(("cos", a.i)))                  # even function

    def rcp(self, a: V) -> V:
        if a.is_const:
            return self.const(1.0 / a.c)
        return V(self, i=self._node(("rcp", a.i)), s=a.s)

    def mul(self, a: V, b: V) -> V:
        if a.is_const and b.is_const:
            return self.const(a.c * b.c)
        if a.is_const:
            a, b = b, a
        if b.is_const:
            c = b.c
            if c == 0.0:
                return self.const(0.0)
            if c == 1.0:
                return a
            if c == -1.0:
                return -a
            s = a.s * (1 if c > 0 else -1)
            return V(self, i=self._node(("mulc", a.i, abs(c))), s=s)
        lo, hi = (a, b) if a.i <= b.i else (b, a)
        return V(self, i=self._node(("mul", lo.i, hi.i)), s=a.s * b.s)

    def add(self, a: V, b: V) -> V:
        if a.is_const and b.is_const:
            return self.const(a.c + b.c)
        if a.is_const:
            a, b = b, a
        if b.is_const:
            if b.c == 0.0:
                return a
            # a.s * (t + a.s*c)
            return V(self, i=self._node(("addc", a.i, a.s * b.c)), s=a.s)
        if a.i == b.i:
            if a.s == b.s:
                return self.mul(a, self.const(2.0))
            return self.const(0.0)
        lo, hi = (a, b) if a.i < b.i else (b, a)
        # lo.s * (t_lo + (lo.s*hi.s) * t_hi)
        return V(self, i=self._node(("add", lo.i, hi.i, lo.s * hi.s)), s=lo.s)

    def output(self, name: str, index: int, v) -> None:
        self.outputs.append((name, int(index), self.lift(v)))

    # analysis ---------------------------------------------------------------------
    def live_nodes(self) -> List[bool]:
        live = [False] * len(self.nodes)
        stack = [v.i for (_, _, v) in self.outputs if not v.is_const]
        while stack:
            i = stack.pop()
            if live[i]:
                continue
            live[i] = True
            k = self.nodes[i]
            if k[0] in ("sin", "cos", "rcp", "mulc", "addc"):
                stack.append(k[1])
            elif k[0] in ("mul", "add"):
                stack.append(k[1])
                stack.append(k[2])
        return live

    def op_counts(self) -> Dict[str, int]:
        """Counts of live operations; 'flops' counts mul/add as 1 each (an FMA the
        compiler forms from a mul+add pair is therefore 2)."""
        live = self.live_nodes()
        cnt = {"mul": 0, "add": 0, "sincos": 0, "rcp": 0, "in": 0}
        for i, k in enumerate(self.nodes):
            if not live[i]:
                continue
            if k[0] in ("mul", "mulc"):
                cnt["mul"] += 1
            elif k[0] in ("add", "addc"):
                cnt["add"] += 1
            elif k[0] in ("sin", "cos"):
                cnt["sincos"] += 1
            elif k[0] == "rcp":
                cnt["rcp"] += 1
            else:
                cnt["in"] += 1
        cnt["flops"] = cnt["mul"] + cnt["add"]
        cnt["nodes"] = sum(live)
        return cnt

    # numpy interpreter (host-side tests of the traced algorithms) --------------------
    def evaluate(self, inputs: Dict[str, np.ndarray], dtype=np.float64) -> Dict[str, np.ndarray]:
        """inputs: name -> array of shape (N,) (or scalar).  Returns name -> (N, size)."""
        live = self.live_nodes()
        vals: List = [None] * len(self.nodes)
        N = 1
        for x in inputs.values():
            N = max(N, int(np.size(x)))
        for i, k in enumerate(self.nodes):
            if not live[i]:
                continue
            op = k[0]
            if op == "in":
                vals[i] = np.broadcast_to(np.asarray(inputs[k[1]], dtype=dtype), (N,))
            elif op == "sin":
                vals[i] = np.sin(vals[k[1]]).astype(dtype)
            elif op == "cos":
                vals[i] = np.cos(vals[k[1]]).astype(dtype)
            elif op == "rcp":
                vals[i] = (dtype(1.0) / vals[k[1]]).astype(dtype)
            elif op == "mul":
                vals[i] = vals[k[1]] * vals[k[2]]
            elif op == "mulc":
                vals[i] = vals[k[1]] * dtype(k[2])
            elif op == "add":
                vals[i] = vals[k[1]] + vals[k[2]] if k[3] > 0 else vals[k[1]] - vals[k[2]]
            elif op == "addc":
                vals[i] = vals[k[1]] + dtype(k[2])
        sizes: Dict[str, int] = {}
        for name, idx, _ in self.outputs:
            sizes[name] = max(sizes.get(name, 0), idx + 1)
        out = {name: np.zeros((N, sz), dtype=dtype) for name, sz in sizes.items()}
        for name, idx, v in self.outputs:
            if v.is_const:
                out[name][:, idx] = v.c
            else:
                out[name][:, idx] = vals[v.i] * v.s
        return out


# ---- small symbolic linear algebra on lists of V ---------------------------------
def vec(p: Program, xs: Iterable) -> List[V]:
    return [p.lift(x) for x in xs]


def zeros(p: Program, n: int) -> List[V]:
    return [p.const(0.0) for _ in range(n)]


def vadd(a: Sequence[V], b: Sequence[V]) -> List[V]:
    return [x + y for x, y in zip(a, b)]


def vsub(a: Sequence[V], b: Sequence[V]) -> List[V]:
    return [x - y for x, y in zip(a, b)]


def vscale(a: Sequence[V], s) -> List[V]:
    return [x * s for x in a]


def dot(a: Sequence[V], b: Sequence[V]) -> V:
    acc = None
    for x, y in zip(a, b):
        t = x * y
        if t.is_zero():
            continue
        acc = t if acc is None else acc + t
    return acc if acc is not None else a[0].p.const(0.0)


def cross3(a: Sequence[V], b: Sequence[V]) -> List[V]:
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def matvec(M: Sequence[Sequence], x: Sequence[V]) -> List[V]:
    return [dot(row, x) for row in M]


def matTvec(M: Sequence[Sequence], x: Sequence[V]) -> List[V]:
    rows, cols = len(M), len(M[0])
    return [dot([M[r][c] for r in range(rows)], x) for c in range(cols)]


def matmul(A: Sequence[Sequence], B: Sequence[Sequence]) -> List[List[V]]:
    n, m, k = len(A), len(B[0]), len(B)
    return [[dot(A[i], [B[t][j] for t in range(k)]) for j in range(m)] for i in range(n)]


def transpose(A: Sequence[Sequence]) -> List[List]:
    return [list(r) for r in zip(*A)]
