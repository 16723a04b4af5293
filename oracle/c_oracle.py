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


def consumer_batch(robot, alg, q, qd, u, dt, lam=None, gravity=9.81, threads=None):
    """Fused-consumer outputs ("fd_vjp" / "fd_lin") for a whole batch: the C oracle supplies qdd, Minv
    and df_du per state, oracle/rbd_numpy.compose_consumer applies the integrator algebra."""
    from . import rbd_numpy as O
    n = robot.n
    q = np.asarray(q, dtype=np.float64)
    qd = np.asarray(qd, dtype=np.float64)
    N = q.shape[0]
    qdd = batch(robot, "fd", q, qd, u, gravity, threads)
    Mu = batch(robot, "minv", q, None, None, gravity, threads).reshape(N, n, n).transpose(0, 2, 1)   # upper, dense
    Md = Mu + np.triu(Mu, 1).transpose(0, 2, 1)
    df = batch(robot, "fd_grad", q, qd, u, gravity, threads).reshape(N, 2 * n, n).transpose(0, 2, 1)
    return np.array([O.compose_consumer(alg, n, qdd[s], Md[s], df[s], q[s], qd[s], dt, None if lam is None else lam[s])
                     for s in range(N)])
