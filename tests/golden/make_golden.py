"""Regenerates tests/golden/rhs_*.npz from the CPU oracle (oracle/, C++ restatement of Trixi.jl's CPU rhs!).

    python tests/golden/make_golden.py

Each fixture holds, for one small case of tests/cases.py: the initial state `u` (Trixi layout, flat), `du` after
ONE rhs!(du, u, t) at t = 0.1, `max_dt`, and the connectivity arrays (interfaces / boundaries / mortars, Int64,
1-based). The GPU parity tests compare libtrixib200 against these files, so the comparison does not depend on
the oracle binary built on the GPU box; the CPU tests check that the oracle still reproduces them bit for bit
(guards the oracle against silent change). The reference itself cannot generate fixtures: it is Julia + CUDA.jl
+ Trixi.jl and Julia is not installed in this image ("Oracle").
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "..", "..", "oracle"), os.path.join(HERE, "..")]
import cases  # noqa: E402

GOLDEN = {  # fixture name -> (case name, level override)
    "c1_advection_1d": ("c1_advection_1d", None),
    "c2_euler_ec_2d": ("c2_euler_ec_2d", 2),
    "c3_euler_sc_3d": ("c3_euler_sc_3d", 2),
    "c4_mhd_alfven_mortar_3d": ("c4_mhd_alfven_mortar_3d", None),
    "c5_euler_ec_3d": ("c5_euler_ec_3d", 2),
    "euler_nonperiodic_2d": ("euler_nonperiodic_2d", 2),
    "euler_ec_mortar_2d": ("euler_ec_mortar_2d", 2),
}
T_RHS = 0.1
CONN = ["interfaces.neighbor_ids", "interfaces.orientations", "boundaries.neighbor_ids", "boundaries.orientations",
        "boundaries.neighbor_sides", "boundaries.n_boundaries_per_direction", "mortars.neighbor_ids",
        "mortars.large_sides", "mortars.orientations"]


def generate(name):
    cname, level = GOLDEN[name]
    c = cases.CASES[cname]
    o = cases.make_oracle(c, level=level)
    u = o.compute_coefficients(0.0)
    du = o.rhs(u, T_RHS)
    out = dict(u=u, du=du, max_dt=np.array([o.max_dt(u)]), t=np.array([T_RHS]))
    if c["vi"] == "shock_capturing_hg":
        out["alpha"] = o.f64("alpha")
    for k in CONN:
        out[k.replace(".", "__")] = o.i64(k)
    return out


if __name__ == "__main__":
    for name in GOLDEN:
        out = generate(name)
        np.savez_compressed(os.path.join(HERE, f"rhs_{name}.npz"), **out)
        print(name, out["u"].size, float(np.abs(out["du"]).max()))
