"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/nextsim_oracle.cpp).

PARITY UNPINNED: the reference holds no golden vectors for this path, and cannot be built
here; these fixtures pin the ORACLE (so that a later edit of the restatement cannot silently change its
arithmetic) and travel to the GPU box, where /root/reference does not exist.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from nextsim_b200 import cases  # noqa: E402
import oracle_bridge as ob  # noqa: E402
from oracle import oracle as orc  # noqa: E402

CASES = {
    # name: (case name, nx, dyn, nranks, open_east, stop_after_substeps)
    "toy_bbm_1sub": ("toy", None, "bbm", 1, False, 1),
    "toy_bbm_step": ("toy", None, "bbm", 1, False, 0),
    "toy_mevp_step": ("toy", None, "mevp", 1, False, 0),
    "toy_evp_2sub": ("toy", None, "evp", 1, False, 2),
    "stable24_bbm_step_open": ("10km_stable", 24, "bbm", 1, True, 0),
    "stable24_bbm_step_3ranks": ("10km_stable", 24, "bbm", 3, True, 0),
}
KEYS = ("M_VT", "M_UM", "M_UT", "M_sigma", "M_damage", "D_tau_a", "D_tau_w", "M_surface", "M_delta_x")
UPD = ("M_conc", "M_thick", "M_snow_thick", "M_ridge_ratio", "M_conc_young", "M_h_young", "M_conc_myi")


def run(spec):
    name, nx, dyn, nranks, open_east, stop = spec
    c = cases.make_case(name, nranks=nranks, dyn=dyn, nx=nx, open_east=open_east)
    if stop:
        c.params.stop_after_substeps = stop
        c.params.skip_ow_smoother = 1
    ranks = ob.make_ranks(c)
    q = ob.orc_params(c.params)
    orc.explicit_solve(ranks, q)
    out = {}
    per = [ob.get_state(R, KEYS) for R in ranks]
    for k in KEYS:
        g = cases.gather_global(c, per, k)
        if k == "M_sigma":
            for i in range(3):
                out["M_sigma%d" % i] = g[i]
        else:
            out[k] = g
    if not stop:
        for R in ranks:
            R.update(q)
        per = [ob.get_state(R, UPD) for R in ranks]
        for k in UPD:
            out["upd_" + k] = cases.gather_global(c, per, k)
    return out


def main():
    for key, spec in CASES.items():
        out = run(spec)
        np.savez_compressed(os.path.join(HERE, key + ".npz"), **out)
        print(key, {k: float(np.abs(v).max()) for k, v in list(out.items())[:3]})


if __name__ == "__main__":
    main()
