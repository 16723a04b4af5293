"""Robot model: the stand-in for the (un-vendored) URDFParser ``robot`` object.

The reference consumes a pre-parsed ``robot`` (reference README.md:8-11); the
parser package is not in the reference tree.  This class provides exactly the
API surface the reference calls (every ``self.robot.*`` call site in
reference ``_test.py``, ``helpers/_topology_helpers.py``,
``helpers/_spatial_algebra_helpers.py`` and ``algorithms/*.py``), so the same
object drives (a) the reference's own numpy oracle, (b) our oracle
restatement and (c) the sm_100a code generator.

Conventions (Featherstone):
  * joint ids are a DFS pre-order (parent < child, subtrees contiguous);
  * ``X_i(q) = X_joint(q) * X_tree`` maps parent-frame motion vectors to child
    frame, ``X = [[E, 0], [-E r^x, E]]``;
  * spatial vectors are ``[angular; linear]``; ``S_i`` is one-hot
    (0..2 revolute about x/y/z, 3..5 prismatic along x/y/z);
  * spatial inertia ``I = [[Ic + m c^x c^xT, m c^x], [m c^xT, m 1]]``.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field
from typing import Callable, List, Sequence

import numpy as np


def skew(v) -> np.ndarray:
    x, y, z = (float(v[0]), float(v[1]), float(v[2]))
    return np.array([[0.0, -z, y], [z, 0.0, -x], [-y, x, 0.0]])


def rot_axis(axis: int, theta: float) -> np.ndarray:
    """Coordinate-transform rotation (Featherstone rx/ry/rz) about a principal axis."""
    c, s = np.cos(theta), np.sin(theta)
    if axis == 0:
        return np.array([[1.0, 0.0, 0.0], [0.0, c, s], [0.0, -s, c]])
    if axis == 1:
        return np.array([[c, 0.0, -s], [0.0, 1.0, 0.0], [s, 0.0, c]])
    return np.array([[c, s, 0.0], [-s, c, 0.0], [0.0, 0.0, 1.0]])


def rpy_to_R(rpy) -> np.ndarray:
    """URDF fixed-axis roll/pitch/yaw -> rotation taking child coords to parent coords."""
    r, p, y = (float(rpy[0]), float(rpy[1]), float(rpy[2]))
    Rx = rot_axis(0, r).T
    Ry = rot_axis(1, p).T
    Rz = rot_axis(2, y).T
    return Rz @ Ry @ Rx


def xform(E: np.ndarray, r) -> np.ndarray:
    """6x6 Plucker motion transform [[E,0],[-E r^x, E]]."""
    X = np.zeros((6, 6))
    X[:3, :3] = E
    X[3:, 3:] = E
    X[3:, :3] = -E @ skew(r)
    return X


def spatial_inertia(mass: float, com, Ic: np.ndarray) -> np.ndarray:
    cx = skew(com)
    I = np.zeros((6, 6))
    I[:3, :3] = Ic + mass * (cx @ cx.T)
    I[:3, 3:] = mass * cx
    I[3:, :3] = mass * cx.T
    I[3:, 3:] = mass * np.eye(3)
    return I


def _snap(a: np.ndarray, tol: float = 1e-12) -> np.ndarray:
    """Snap values within tol of -1/0/1 (rpy = pi/2 rounding noise) so that the
    compile-time sparsity of X_tree is exact."""
    a = np.array(a, dtype=np.float64)
    for t in (-1.0, 0.0, 1.0):
        a[np.abs(a - t) < tol] = t
    return a


class _Named:
    def __init__(self, name: str):
        self._name = name

    def get_name(self) -> str:
        return self._name


@dataclass
class Robot:
    """Fixed-base tree of 1-DoF joints in DFS pre-order."""

    name: str
    parent: List[int]                    # parent joint id, -1 = base
    S_ind: List[int]                     # 0..5
    E0: List[np.ndarray]                 # 3x3 tree rotation per joint
    r0: List[np.ndarray]                 # 3 tree translation per joint (parent coords)
    Imats: List[np.ndarray]              # 6x6 spatial inertia of the link moved by joint i
    damping: List[float]
    joint_names: List[str] = field(default_factory=list)
    link_names: List[str] = field(default_factory=list)
    base_inertia: np.ndarray = field(default_factory=lambda: np.zeros((6, 6)))

    def __post_init__(self):
        n = len(self.parent)
        self.n = n
        self.E0 = [_snap(E) for E in self.E0]
        self.r0 = [np.array(r, dtype=np.float64) for r in self.r0]
        self.Imats = [np.array(I, dtype=np.float64) for I in self.Imats]
        self.damping = [float(d) for d in self.damping]
        if not self.joint_names:
            self.joint_names = ["joint_%d" % i for i in range(n)]
        if not self.link_names:
            self.link_names = ["link_%d" % i for i in range(n)]
        for i, p in enumerate(self.parent):
            if not (-1 <= p < i):
                raise ValueError("joint ids must be a DFS pre-order (parent < child)")
        self._level = [0] * n
        for i, p in enumerate(self.parent):
            self._level[i] = 0 if p == -1 else self._level[p] + 1
        self._anc = []
        for i in range(n):
            a, p = [], self.parent[i]
            while p != -1:
                a.append(p)
                p = self.parent[p]
            self._anc.append(sorted(a))
        self._sub = [[j for j in range(n) if j == i or i in self._anc[j]] for i in range(n)]
        for i in range(n):  # DFS contiguity: the reference relies on it (_direct_minv.py:141)
            if self._sub[i] != list(range(i, i + len(self._sub[i]))):
                raise ValueError("subtree of joint %d is not contiguous: ids are not DFS pre-order" % i)

    # ---- sizes / topology --------------------------------------------------
    def get_num_pos(self) -> int:
        return self.n

    def get_num_vel(self) -> int:
        return self.n

    def get_parent_id(self, i: int) -> int:
        return self.parent[i]

    def get_parent_id_array(self) -> List[int]:
        return list(self.parent)

    def get_unique_parent_ids(self, ids: Sequence[int]) -> List[int]:
        return sorted(set(self.parent[i] for i in ids))

    def has_repeated_parents(self, ids: Sequence[int]) -> bool:
        ps = [self.parent[i] for i in ids]
        return len(ps) != len(set(ps))

    def get_bfs_level_by_id(self, i: int) -> int:
        return self._level[i]

    def get_ids_by_bfs_level(self, level: int) -> List[int]:
        return [i for i in range(self.n) if self._level[i] == level]

    def get_max_bfs_level(self) -> int:
        return max(self._level)

    def get_max_bfs_width(self) -> int:
        return max(len(self.get_ids_by_bfs_level(l)) for l in range(self.get_max_bfs_level() + 1))

    def get_ancestors_by_id(self, i: int) -> List[int]:
        return list(self._anc[i])          # fresh list: the oracle mutates it (_test.py:355-356)

    def get_subtree_by_id(self, i: int) -> List[int]:
        return list(self._sub[i])

    def get_total_ancestor_count(self) -> int:
        return sum(len(a) for a in self._anc)

    def get_total_subtree_count(self) -> int:
        return sum(len(s) for s in self._sub)

    def get_is_ancestor_of(self, j: int, i: int) -> bool:
        """True when joint j is an ancestor of joint i."""
        return j in self._anc[i]

    def get_is_in_subtree_of(self, j: int, i: int) -> bool:
        """True when joint j lies in the subtree rooted at joint i."""
        return j in self._sub[i]

    def is_serial_chain(self) -> bool:
        return all(p == i - 1 for i, p in enumerate(self.parent))

    def are_Ss_identical(self, ids: Sequence[int]) -> bool:
        return len(set(self.S_ind[i] for i in ids)) <= 1

    # ---- joint / link data ---------------------------------------------------
    def get_S_by_id(self, i: int) -> np.ndarray:
        S = np.zeros(6)
        S[self.S_ind[i]] = 1
        return S

    def get_damping_by_id(self, i: int) -> float:
        return self.damping[i]

    def get_joint_by_id(self, i: int) -> _Named:
        return _Named(self.joint_names[i])

    def get_link_by_id(self, i: int) -> _Named:
        return _Named(self.link_names[i])

    def get_Imat_by_id(self, i: int) -> np.ndarray:
        return self.Imats[i]

    def get_Imats_ordered_by_id(self) -> List[np.ndarray]:
        """Base inertia first (the reference drops index 0, _test.py:17)."""
        return [self.base_inertia] + list(self.Imats)

    def get_Imats_dict_by_id(self) -> dict:
        return {i: self.Imats[i] for i in range(self.n)}

    # ---- transforms ----------------------------------------------------------
    def joint_E_r(self, i: int, q: float):
        """(E, r) of X_i(q) = [[E,0],[-E r^x,E]]."""
        k = self.S_ind[i]
        if k < 3:
            return rot_axis(k, q) @ self.E0[i], self.r0[i]
        d = np.zeros(3)
        d[k - 3] = q
        return self.E0[i], self.r0[i] + self.E0[i].T @ d

    def Xmat(self, i: int, q: float) -> np.ndarray:
        E, r = self.joint_E_r(i, q)
        return xform(E, r)

    def get_Xmat_Func_by_id(self, i: int) -> Callable[[float], np.ndarray]:
        return lambda q, _i=i: self.Xmat(_i, float(q))

    def get_Xmat_Funcs_ordered_by_id(self) -> List[Callable[[float], np.ndarray]]:
        return [self.get_Xmat_Func_by_id(i) for i in range(self.n)]

    def get_Xmats_ordered_by_id(self):
        """sympy 6x6 matrices in a symbol literally named ``theta`` (the reference
        string-substitutes it, helpers/_topology_helpers.py:164-167).  Only needed
        to drive the reference's own generator; the B200 generator never uses sympy."""
        import sympy as sp
        th = sp.Symbol("theta")
        out = []
        for i in range(self.n):
            k = self.S_ind[i]
            E0 = sp.Matrix(self.E0[i].tolist())
            r0 = sp.Matrix(self.r0[i].tolist())
            if k < 3:
                c, s = sp.cos(th), sp.sin(th)
                if k == 0:
                    EJ = sp.Matrix([[1, 0, 0], [0, c, s], [0, -s, c]])
                elif k == 1:
                    EJ = sp.Matrix([[c, 0, -s], [0, 1, 0], [s, 0, c]])
                else:
                    EJ = sp.Matrix([[c, s, 0], [-s, c, 0], [0, 0, 1]])
                E, r = EJ * E0, r0
            else:
                d = sp.zeros(3, 1)
                d[k - 3] = th
                E, r = E0, r0 + E0.T * d
            rx = sp.Matrix([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
            X = sp.zeros(6, 6)
            X[:3, :3] = E
            X[3:, 3:] = E
            X[3:, :3] = -E * rx
            out.append(X)
        return out

    # ---- identity ------------------------------------------------------------
    def param_hash(self) -> str:
        """Stable hash of every number the generated code depends on (build cache key,
        golden-fixture guard)."""
        h = hashlib.sha256()
        h.update(self.name.encode())
        h.update(np.array(self.parent, dtype=np.int64).tobytes())
        h.update(np.array(self.S_ind, dtype=np.int64).tobytes())
        for E, r, I in zip(self.E0, self.r0, self.Imats):
            h.update(np.round(E, 12).tobytes())
            h.update(np.round(r, 12).tobytes())
            h.update(np.round(I, 12).tobytes())
        h.update(np.round(np.array(self.damping), 12).tobytes())
        return h.hexdigest()[:16]

    def with_damping(self, damping) -> "Robot":
        d = [float(damping)] * self.n if np.isscalar(damping) else [float(x) for x in damping]
        return Robot(self.name, list(self.parent), list(self.S_ind), [E.copy() for E in self.E0],
                     [r.copy() for r in self.r0], [I.copy() for I in self.Imats], d,
                     list(self.joint_names), list(self.link_names), self.base_inertia.copy())
