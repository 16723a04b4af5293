#!/usr/bin/env python3
"""Generates tests/golden/<robot>.npz from the REFERENCE ITSELF.

Imports the unmodified reference (/root/reference, read-only) and runs its numpy
implementation (_test.py: test_rnea 109-115, test_minv 213-226, test_rnea_grad 490-494,
test_fd_grad 496-520) on seeded float32-rounded states, driven by our Robot stand-in
(URDFParser is not vendored).  The reference cannot travel to the GPU box, so the
outputs are committed as fixtures; every fixture carries the robot parameter hash so a
changed synthetic URDF invalidates it loudly.

Run (in the build container only):  python tests/golden/make_golden.py
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root")

from gridcodegenerator_b200 import load_named_robot          # noqa: E402
from gridcodegenerator_b200.synthetic import make_states, seed_for  # noqa: E402
from reference import GRiDCodeGenerator as RefGen             # noqa: E402

CASES = [("mixed5", None, 8), ("iiwa14", 0.0, 8), ("iiwa14", 0.5, 4), ("hyq", 0.0, 8), ("atlas", 0.0, 4), ("chain64", 0.0, 2)]


def main():
    for name, damping, N in CASES:
        robot = load_named_robot(name)
        if damping is None:                      # keep the URDF's own per-joint damping
            damping = -1.0
        else:
            robot = robot.with_damping(damping)
        g = RefGen(robot)
        n = robot.n
        q, qd, u, qdd = make_states(n, N, seed_for(name))
        q64, qd64, u64, qdd64 = (x.astype(np.float64) for x in (q, qd, u, qdd))
        out = dict(c=[], c_qdd=[], minv_dense=[], minv_upper=[], fd_qdd=[], dc_du=[], dc_du_qdd=[], df_du=[])
        with contextlib.redirect_stdout(io.StringIO()):   # _test.py:250-253 prints unconditionally
            for s in range(N):
                out["c"].append(g.test_rnea(q64[s], qd64[s])[0])
                out["c_qdd"].append(g.test_rnea(q64[s], qd64[s], qdd64[s])[0])
                Mi = g.test_minv(q64[s])
                out["minv_dense"].append(Mi)
                out["minv_upper"].append(g.test_minv(q64[s], False))
                out["fd_qdd"].append(Mi @ (u64[s] - out["c"][-1]))      # _test.py:498-501
                out["dc_du"].append(g.test_rnea_grad(q64[s], qd64[s]))
                out["dc_du_qdd"].append(g.test_rnea_grad(q64[s], qd64[s], qdd64[s]))
                out["df_du"].append(g.test_fd_grad(q64[s], qd64[s], u64[s]))
        tag = name if damping <= 0.0 else "%s_damped" % name
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), robot_hash=robot.param_hash(),
                            damping=damping, q=q, qd=qd, u=u, qdd=qdd,
                            **{k: np.array(v) for k, v in out.items()})
        print("wrote", tag, "n=%d N=%d hash=%s" % (n, N, robot.param_hash()))


if __name__ == "__main__":
    main()
