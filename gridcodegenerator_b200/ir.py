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
        if isinstance(o, V2):
            return o + self
        return self.p.add(self, self.p.lift(o))

    __radd__ = __add__

    def __sub__(self, o):
        if isinstance(o, V2):
            return (-o) + self
        return self.p.add(self, -self.p.lift(o))

    def __rsub__(self, o):
        return self.p.add(self.p.lift(o), -self)

    def __mul__(self, o):
        if isinstance(o, V2):
            return o * self
        return self.p.mul(self, self.p.lift(o))

    __rmul__ = __mul__

    def __repr__(self):
        return "V(c=%r)" % self.c if self.is_const else "V(%s t%d)" % ("+" if self.s > 0 else "-", self.i)


class V2:
    """A PAIR of values that always undergo the same operation with the same scalar - the d/dq_j
    and d/dqd_j columns of the gradients.  Emitted as float2 and computed with the sm_100 packed
    FP32 instructions (FFMA2 / FMUL2 / FADD2: one issue slot, two lanes), whose second operand
    may be a scalar register broadcast to both halves.  Either a constant pair c = (cx, cy) or a
    signed reference to a pair node."""
    __slots__ = ("p", "c", "i", "s")

    def __init__(self, p: "Program", c=None, i: int = -1, s: int = 1):
        self.p, self.c, self.i, self.s = p, c, i, s

    @property
    def is_const(self) -> bool:
        return self.c is not None

    def is_zero(self) -> bool:
        return self.c is not None and self.c[0] == 0.0 and self.c[1] == 0.0

    def __neg__(self):
        if self.is_const:
            return V2(self.p, c=(-self.c[0], -self.c[1]))
        return V2(self.p, i=self.i, s=-self.s)

    def __add__(self, o):
        return self.p.add2(self, self.p.lift2(o))

    __radd__ = __add__

    def __sub__(self, o):
        return self.p.add2(self, -self.p.lift2(o))

    def __rsub__(self, o):
        return self.p.add2(self.p.lift2(o), -self)

    def __mul__(self, o):
        if isinstance(o, V2):
            raise TypeError("pair x pair products do not occur in the column recursions")
        return self.p.mul2(self, self.p.lift(o))

    __rmul__ = __mul__

    def x(self) -> V:
        return self.p.half(self, 0)

    def y(self) -> V:
        return self.p.half(self, 1)

    def __repr__(self):
        return "V2(c=%r)" % (self.c,) if self.is_const else "V2(%s p%d)" % ("+" if self.s > 0 else "-", self.i)


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

    # pairs ----------------------------------------------------------------------------
    def lift2(self, x) -> "V2":
        if isinstance(x, V2):
            return x
        if isinstance(x, V):
            if x.is_const:
                return V2(self, c=(x.c, x.c))
            return self.pack(x, x)
        return V2(self, c=(float(x), float(x)))

    def pack(self, a, b) -> "V2":
        """Pair from two scalars (x = a, y = b)."""
        a, b = self.lift(a), self.lift(b)
        if a.is_const and b.is_const:
            return V2(self, c=(a.c, b.c))
        ha = ("c", a.c) if a.is_const else ("n", a.i, a.s)
        hb = ("c", b.c) if b.is_const else ("n", b.i, b.s)
        return V2(self, i=self._node(("pk", ha, hb)))

    def half(self, P: "V2", which: int) -> V:
        if P.is_const:
            return self.const(P.c[which])
        k = self.nodes[P.i]
        if k[0] == "pk":                       # see through a fresh pack
            h = k[1 + which]
            v = self.const(h[1]) if h[0] == "c" else V(self, i=h[1], s=h[2])
            return v if P.s > 0 else -v
        return V(self, i=self._node(("half", P.i, which)), s=P.s)

    def mul2(self, P: "V2", sc: V) -> "V2":
        if P.is_const:
            if sc.is_const:
                return V2(self, c=(P.c[0] * sc.c, P.c[1] * sc.c))
            return self.pack(sc * P.c[0], sc * P.c[1])
        if sc.is_const:
            c = sc.c
            if c == 0.0:
                return V2(self, c=(0.0, 0.0))
            if c == 1.0:
                return P
            if c == -1.0:
                return -P
            return V2(self, i=self._node(("mul2c", P.i, abs(c))), s=P.s * (1 if c > 0 else -1))
        return V2(self, i=self._node(("mul2", P.i, sc.i)), s=P.s * sc.s)

    def add2(self, P: "V2", Q: "V2") -> "V2":
        if P.is_const and Q.is_const:
            return V2(self, c=(P.c[0] + Q.c[0], P.c[1] + Q.c[1]))
        if P.is_const:
            P, Q = Q, P
        if Q.is_const:
            if Q.is_zero():
                return P
            return V2(self, i=self._node(("add2k", P.i, P.s * Q.c[0], P.s * Q.c[1])), s=P.s)
        if P.i == Q.i:
            if P.s == Q.s:
                return self.mul2(P, self.const(2.0))
            return V2(self, c=(0.0, 0.0))
        lo, hi = (P, Q) if P.i < Q.i else (Q, P)
        return V2(self, i=self._node(("add2", lo.i, hi.i, lo.s * hi.s)), s=lo.s)

    # explicit parking of long-lived values in per-lane shared memory ------------------------
    def park(self, v: V):
        """Stores v in a private shared-memory slot right after it is computed; returns a handle.
        Constants are not parked (the handle is the constant)."""
        if v.is_const:
            return ("c", v.c)
        if not hasattr(self, "parks"):
            self.parks: Dict[int, int] = {}
        slot = self.parks.setdefault(v.i, len(self.parks))
        return ("s", slot, v.i, v.s)

    def unpark(self, handle) -> V:
        """A FRESH load of a parked value (never CSE'd with earlier loads): the register copy
        lives only as long as its uses, instead of from the producer to the last column."""
        if handle[0] == "c":
            return self.const(handle[1])
        self._ld_uid = getattr(self, "_ld_uid", 0) + 1
        return V(self, i=self._node(("ld", handle[1], handle[2], self._ld_uid)), s=handle[3])

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
            stack.extend(self.operands(i))
        return live

    def operands(self, i: int) -> List[int]:
        k = self.nodes[i]
        op = k[0]
        if op in ("sin", "cos", "rcp", "mulc", "addc", "mul2c", "add2k", "half"):
            return [k[1]]
        if op in ("mul", "add", "mul2", "add2"):
            return [k[1], k[2]]
        if op == "pk":
            return [h[1] for h in (k[1], k[2]) if h[0] == "n"]
        if op == "ld":
            return [k[2]]
        return []

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
            elif k[0] in ("mul2", "mul2c"):
                cnt["mul"] += 2
                cnt["packed"] = cnt.get("packed", 0) + 1
            elif k[0] in ("add2", "add2k"):
                cnt["add"] += 2
                cnt["packed"] = cnt.get("packed", 0) + 1
            elif k[0] in ("pk", "half"):
                pass
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
            elif op == "ld":
                vals[i] = vals[k[2]]
            elif op == "pk":
                hv = []
                for h in (k[1], k[2]):
                    hv.append(np.full(N, dtype(h[1])) if h[0] == "c" else vals[h[1]] * dtype(h[2]))
                vals[i] = (hv[0], hv[1])
            elif op == "half":
                vals[i] = vals[k[1]][k[2]]
            elif op == "mul2":
                vals[i] = (vals[k[1]][0] * vals[k[2]], vals[k[1]][1] * vals[k[2]])
            elif op == "mul2c":
                vals[i] = (vals[k[1]][0] * dtype(k[2]), vals[k[1]][1] * dtype(k[2]))
            elif op == "add2":
                a, b = vals[k[1]], vals[k[2]]
                vals[i] = (a[0] + b[0], a[1] + b[1]) if k[3] > 0 else (a[0] - b[0], a[1] - b[1])
            elif op == "add2k":
                vals[i] = (vals[k[1]][0] + dtype(k[2]), vals[k[1]][1] + dtype(k[3]))
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
