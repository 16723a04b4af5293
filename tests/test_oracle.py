"""The numpy oracle against the reference's own outputs (committed goldens, and the live
reference when /root/reference exists in this container)."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

from helpers import golden_big_count, golden_df_du, load_golden, relerr
from oracle import rbd_numpy as O

TAGS = ["mixed5", "iiwa14", "iiwa14_damped", "hyq", "atlas", "chain64"]


# states the (slow, per-state Python) numpy oracle is pinned on; the C oracle below covers all 64
NUMPY_STATES = {"atlas": 12, "chain64": 4}


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_matches_reference_goldens(tag):
    robot, z = load_golden(tag)
    q, qd, u, qdd = (z[k].astype(np.float64) for k in ("q", "qd", "u", "qdd"))
    assert q.shape[0] >= 64                      # fixtures hold 64 reference-generated states per robot
    N = min(NUMPY_STATES.get(tag, q.shape[0]), golden_big_count(z))
    for s in range(N):
        assert relerr(O.rnea(robot, q[s], qd[s])[0], z["c"][s]) < 1e-12
        assert relerr(O.rnea(robot, q[s], qd[s], qdd[s])[0], z["c_qdd"][s]) < 1e-12
        assert relerr(O.minv(robot, q[s]), z["minv_dense"][s]) < 1e-11
        assert relerr(O.minv(robot, q[s], dense=False), z["minv_upper"][s]) < 1e-11
        assert relerr(O.fd(robot, q[s], qd[s], u[s]), z["fd_qdd"][s]) < 1e-10
        assert relerr(O.rnea_grad(robot, q[s], qd[s]), z["dc_du"][s]) < 1e-11
        assert relerr(O.rnea_grad(robot, q[s], qd[s], qdd[s]), z["dc_du_qdd"][s]) < 1e-11
        if s < 2:
            assert relerr(O.fd_grad(robot, q[s], qd[s], u[s]), z["df_du"][s]) < 1e-9


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_pass_level_intermediates(tag):
    """The oracle's RNEA against the reference's pass-level outputs (test_rnea_fpass / test_rnea_bpass,
    _test.py:5-107): v, a, the accumulated f and c."""
    robot, z = load_golden(tag)
    q, qd, qdd = (z[k].astype(np.float64) for k in ("q", "qd", "qdd"))
    for s in range(int(z["n_pass"])):
        c, v, a, f = O.rnea(robot, q[s], qd[s], qdd[s])
        assert relerr(v, z["pl_v"][s]) < 1e-12 and relerr(a, z["pl_a"][s]) < 1e-12
        assert relerr(f, z["pl_f"][s]) < 1e-12 and relerr(c, z["pl_c"][s]) < 1e-12


def test_minv_upper_is_triangular_and_dense_is_symmetric():
    robot, z = load_golden("hyq")
    Mu = O.minv(robot, z["q"][0].astype(np.float64), dense=False)
    assert np.all(np.tril(Mu, -1) == 0)
    Md = O.minv(robot, z["q"][0].astype(np.float64))
    assert np.allclose(Md, Md.T)


@pytest.mark.parametrize("name", ["iiwa14", "hyq", "atlas"])
def test_oracle_invariants(name):
    """M * Minv = I with M from RNEA unit accelerations (gravity 0), and central finite
    differences for both gradients."""
    robot, z = load_golden(name)
    n = robot.n
    q, qd, u = (z[k][0].astype(np.float64) for k in ("q", "qd", "u"))
    M = np.zeros((n, n))
    for j in range(n):
        e = np.zeros(n)
        e[j] = 1.0
        M[:, j] = O.rnea(robot, q, np.zeros(n), e, gravity=0.0)[0]
    assert np.abs(O.minv(robot, q) @ M - np.eye(n)).max() < 1e-9
    eps = 1e-6
    num = np.zeros((n, 2 * n))
    for j in range(n):
        d = np.zeros(n)
        d[j] = eps
        num[:, j] = (O.fd(robot, q + d, qd, u) - O.fd(robot, q - d, qd, u)) / (2 * eps)
        num[:, n + j] = (O.fd(robot, q, qd + d, u) - O.fd(robot, q, qd - d, u)) / (2 * eps)
    assert relerr(O.fd_grad(robot, q, qd, u), num) < 1e-6


@pytest.mark.skipif(not os.path.exists("/root/reference/_test.py"), reason="reference tree not present")
@pytest.mark.parametrize("name", ["iiwa14", "hyq"])
def test_oracle_matches_live_reference(name):
    sys.path.insert(0, "/root")
    from reference import GRiDCodeGenerator as RefGen
    robot, _ = load_golden(name)
    robot = robot.with_damping(0.25)
    g = RefGen(robot)
    rng = np.random.default_rng(7)
    n = robot.n
    q, qd, u = rng.uniform(-3, 3, n), rng.uniform(-2, 2, n), rng.uniform(-20, 20, n)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = g.test_fd_grad(q, qd, u)
        ref_dc = g.test_rnea_grad(q, qd)
    assert relerr(O.fd_grad(robot, q, qd, u), ref) < 1e-10
    assert relerr(O.rnea_grad(robot, q, qd), ref_dc) < 1e-11


@pytest.mark.parametrize("tag", TAGS)
def test_c_oracle_matches_numpy_oracle_and_goldens(tag):
    """Every one of the 64 reference-generated states of every fixture."""
    from helpers import colmajor_batch
    from oracle import c_oracle as C
    robot, z = load_golden(tag)
    q, qd, u, qdd = (z[k].astype(np.float64) for k in ("q", "qd", "u", "qdd"))
    nb = golden_big_count(z)
    assert relerr(C.batch(robot, "id", q, qd), z["c"]) < 1e-12
    assert relerr(C.batch(robot, "id", q, qd, qdd), z["c_qdd"]) < 1e-12
    assert relerr(C.batch(robot, "minv", q[:nb]), colmajor_batch(z["minv_upper"])) < 1e-10
    assert relerr(C.batch(robot, "fd", q, qd, u), z["fd_qdd"]) < 1e-9
    assert relerr(C.batch(robot, "id_grad", q[:nb], qd[:nb]), colmajor_batch(z["dc_du"])) < 1e-10
    assert relerr(C.batch(robot, "id_grad", q[:nb], qd[:nb], qdd[:nb]), colmajor_batch(z["dc_du_qdd"])) < 1e-10
    df = C.batch(robot, "fd_grad", q, qd, u)
    assert relerr(df[:nb], colmajor_batch(z["df_du"])) < 1e-8
    assert relerr(df, colmajor_batch(golden_df_du(z))) < (1e-8 if nb == q.shape[0] else 5e-7)   # float32 tail
    assert relerr(C.batch(robot, "fd_grad", q[:2], qd[:2], u[:2], threads=1),
                  O.batch(robot, "fd_grad", q[:2], qd[:2], u[:2])) < 1e-9
