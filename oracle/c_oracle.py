"""ctypes front-end of the C oracle (oracle/rbd_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os

import numpy as np

from .build_c import build_c_oracle

_ALG = {"id": 0, "minv": 1, "fd": 2, "id_grad": 3, "fd_grad": 4}
_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_c_oracle())
        _lib.orc_batch.restype = ctypes.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t)) if a is not None else None


def batch(robot, alg, q, qd=None, x=None, gravity=9.81, threads=None):
    """Same flat per-state outputs as rbd_numpy.batch, float64, for a whole batch."""
    lib = _load()
    n = robot.n
    q = np.ascontiguousarray(q, dtype=np.float64)
    N = q.shape[0]
    qd = np.ascontiguousarray(qd if qd is not None else np.zeros_like(q), dtype=np.float64)
    x = None if x is None else np.ascontiguousarray(x, dtype=np.float64)
    words = {"id": n, "minv": n * n, "fd": n, "id_grad": 2 * n * n, "fd_grad": 2 * n * n}[alg]
    out = np.zeros((N, words))
    parent = np.ascontiguousarray(robot.parent, dtype=np.int32)
    S = np.ascontiguousarray(robot.S_ind, dtype=np.int32)
    E0 = np.ascontiguousarray(np.concatenate([E.flatten() for E in robot.E0]))
    r0 = np.ascontiguousarray(np.concatenate(robot.r0))
    I = np.ascontiguousarray(np.concatenate([M.flatten() for M in robot.Imats]))
    damp = np.ascontiguousarray(robot.damping, dtype=np.float64)
    d, i32 = ctypes.c_double, ctypes.c_int
    lib.orc_batch(i32(n), _p(parent, i32), _p(S, i32), _p(E0, d), _p(r0, d), _p(I, d), _p(damp, d), i32(_ALG[alg]),
                  i32(N), _p(q, d), _p(qd, d), _p(x, d), d(gravity), _p(out, d), i32(threads or os.cpu_count() or 1))
    return out
