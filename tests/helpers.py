"""Shared test helpers (CPU side)."""
import os

import numpy as np

from gridcodegenerator_b200 import load_named_robot

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# tolerances of BASELINE.json north_star: max|x-ref| / max|ref| per output tensor
TOL = {"id": 1e-4, "minv": 1e-4, "fd": 1e-4, "id_grad": 1e-3, "fd_grad": 1e-3}


def relerr(x, ref):
    x, ref = np.asarray(x, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    denom = np.abs(ref).max()
    return float(np.abs(x - ref).max() / (denom if denom > 0 else 1.0))


def load_golden(tag):
    z = np.load(os.path.join(GOLDEN, tag + ".npz"))
    name = tag.replace("_damped", "")
    robot = load_named_robot(name)
    if float(z["damping"]) >= 0.0:               # negative = the URDF's own per-joint damping
        robot = robot.with_damping(float(z["damping"]))
    assert robot.param_hash() == str(z["robot_hash"]), (
        "golden fixture %s was generated for different robot parameters; rerun tests/golden/make_golden.py" % tag)
    return robot, z


def colmajor_batch(mats):
    """(N, r, c) -> (N, r*c) column-major per state."""
    mats = np.asarray(mats)
    return np.transpose(mats, (0, 2, 1)).reshape(mats.shape[0], -1)


def golden_df_du(z):
    """All df_du matrices of a fixture as float64 (N, n, 2n): the 64-link chain keeps only its first
    `n_big` states in float64 and the rest as float32 (tests/golden/make_golden.py)."""
    head = np.asarray(z["df_du"], dtype=np.float64)
    if "df_du_f32_tail" in z.files:
        return np.concatenate([head, np.asarray(z["df_du_f32_tail"], dtype=np.float64)])
    return head


def golden_big_count(z):
    """States for which the n x n / n x 2n matrices other than df_du are stored."""
    return int(z["n_big"]) if "n_big" in z.files else z["q"].shape[0]


_PLANS = {}


def cached_plan(name, **kw):
    """KernelPlan of a named robot, built once per test process: the plan of the 64-link chain traces every
    phase-split candidate before rejecting it (two minutes), and several test modules look at the same default plans."""
    from gridcodegenerator_b200.codegen import KernelPlan
    key = (name, repr(sorted(kw.items())))
    if key not in _PLANS:
        _PLANS[key] = KernelPlan(load_named_robot(name), **kw)
    return _PLANS[key]
