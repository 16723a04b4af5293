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

# (robot, damping, states, states whose big matrices are kept in float64, pass-level states)
# 64 seeded states per robot.  The 64-link chain's matrices are 65 KB per state and
# gradient: all outputs are kept in float64 for the first 8 states (what the oracle is pinned against
# at 1e-11); for the other 56 the n-word outputs stay float64 and df_du is stored as float32 (the
# GPU parity bar is 1e-3 relative; float32 rounding is 6e-8).
CASES = [("mixed5", None, 64, 64, 4), ("iiwa14", 0.0, 64, 64, 4), ("iiwa14", 0.5, 64, 64, 4), ("hyq", 0.0, 64, 64, 4),
         ("atlas", 0.0, 64, 64, 2), ("chain64", 0.0, 64, 8, 1)]
PASS_LEVEL = ("dc_dq", "dc_dqd", "dv_dq", "dv_dqd", "da_dq", "da_dqd", "df_fp_dq", "df_fp_dqd", "df_dq", "df_dqd")


def main():
    import copy
    only = set(sys.argv[1:])
    for name, damping, N, NBIG, NPASS in CASES:
        robot = load_named_robot(name)
        if damping is None:                      # keep the URDF's own per-joint damping
            damping = -1.0
        else:
            robot = robot.with_damping(damping)
        tag = name if damping <= 0.0 else "%s_damped" % name
        if only and tag not in only:
            continue
        g = RefGen(robot)
        n = robot.n
        q, qd, u, qdd = make_states(n, N, seed_for(name))
        q64, qd64, u64, qdd64 = (x.astype(np.float64) for x in (q, qd, u, qdd))
        out = dict(c=[], c_qdd=[], minv_dense=[], minv_upper=[], fd_qdd=[], dc_du=[], dc_du_qdd=[], df_du=[])
        pl = {k: [] for k in ("pl_v", "pl_a", "pl_f_fpass", "pl_c", "pl_f", "pl_Minv_bpass", "pl_F", "pl_U", "pl_Dinv")
              + tuple("pl_" + k for k in PASS_LEVEL)}
        with contextlib.redirect_stdout(io.StringIO()):   # _test.py:250-253 prints unconditionally
            for s in range(N):
                big = s < NBIG
                out["c"].append(g.test_rnea(q64[s], qd64[s])[0])
                out["c_qdd"].append(g.test_rnea(q64[s], qd64[s], qdd64[s])[0])
                Mi = g.test_minv(q64[s])
                out["fd_qdd"].append(Mi @ (u64[s] - out["c"][-1]))      # _test.py:498-501
                out["df_du"].append(g.test_fd_grad(q64[s], qd64[s], u64[s]))
                if big:
                    out["minv_dense"].append(Mi)
                    out["minv_upper"].append(g.test_minv(q64[s], False))
                    out["dc_du"].append(g.test_rnea_grad(q64[s], qd64[s]))
                    out["dc_du_qdd"].append(g.test_rnea_grad(q64[s], qd64[s], qdd64[s]))
                if s < NPASS:                    # pass-level intermediates (_test.py:5-107, 117-202, 229-488)
                    v, a, f = g.test_rnea_fpass(q64[s], qd64[s], qdd64[s])
                    pl["pl_v"].append(v); pl["pl_a"].append(a); pl["pl_f_fpass"].append(copy.deepcopy(f))
                    c, facc = g.test_rnea_bpass(q64[s], qd64[s], copy.deepcopy(f))
                    pl["pl_c"].append(c); pl["pl_f"].append(facc)
                    Mb, F, U, Dinv = g.test_minv_bpass(q64[s])
                    pl["pl_Minv_bpass"].append(copy.deepcopy(Mb)); pl["pl_F"].append(copy.deepcopy(F))
                    pl["pl_U"].append(U); pl["pl_Dinv"].append(Dinv)
                    if n <= 32:                  # 10 arrays of 6 n^2: skipped for the 64-link chain
                        for k, arr in zip(PASS_LEVEL, g.test_rnea_grad_inner(q64[s], qd64[s], v, a, facc)):
                            pl["pl_" + k].append(arr)
        arrays = {k: np.array(v) for k, v in out.items()}
        if NBIG < N:
            arrays["df_du_f32_tail"] = arrays["df_du"][NBIG:].astype(np.float32)
            arrays["df_du"] = arrays["df_du"][:NBIG]
        arrays.update({k: np.array(v) for k, v in pl.items() if v})
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), robot_hash=robot.param_hash(),
                            damping=damping, q=q, qd=qd, u=u, qdd=qdd, n_big=NBIG, n_pass=NPASS, **arrays)
        print("wrote", tag, "n=%d N=%d hash=%s %.1f MB" % (n, N, robot.param_hash(),
                                                          os.path.getsize(os.path.join(HERE, tag + ".npz")) / 1e6))


if __name__ == "__main__":
    main()
