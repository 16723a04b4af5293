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
