#!/usr/bin/env python3
"""Generates tests/golden/<robot>_mass.npz: the joint-space mass matrix M(q) assembled from the REFERENCE'S OWN RNEA.

The reference has no CRBA.  Its test_rnea (_test.py:109-115) is linear in qdd with slope M(q):
column j of M = test_rnea(q, 0, e_j) - test_rnea(q, 0, 0).  That pins grid_crba_device / oracle crba() to the
reference code (the other pin is M Minv = I with the reference's test_minv, held in <robot>.npz).

Run (in the build container only):  python tests/golden/make_golden_mass.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root")

from gridcodegenerator_b200 import load_named_robot          # noqa: E402
from gridcodegenerator_b200.synthetic import make_states, seed_for  # noqa: E402
from reference import GRiDCodeGenerator as RefGen             # noqa: E402

CASES = [("mixed5", 16), ("iiwa14", 16), ("hyq", 16), ("atlas", 8), ("chain64", 2)]


def main():
    for name, N in CASES:
        robot = load_named_robot(name)
        g = RefGen(robot)
        n = robot.n
        q = make_states(n, N, seed_for(name))[0].astype(np.float64)
        zero = np.zeros(n)
        M = np.zeros((N, n, n))
        for s in range(N):
            c0 = np.asarray(g.test_rnea(q[s], zero, zero)[0], dtype=np.float64).reshape(n)
            for j in range(n):
                e = np.zeros(n)
                e[j] = 1.0
                M[s, :, j] = np.asarray(g.test_rnea(q[s], zero, e)[0], dtype=np.float64).reshape(n) - c0
        np.savez_compressed(os.path.join(HERE, "%s_mass.npz" % name), robot_hash=robot.param_hash(), q=q, M=M)
        print("%-8s %d states, max |M - M^T| = %.2e" % (name, N, np.abs(M - M.transpose(0, 2, 1)).max()))


if __name__ == "__main__":
    main()
