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


# SURVEY 8(f) rows 1-2 (regrid check, ice diagnostics, forcing interpolation) on a displaced mesh, one rank
NEXT_CASES = {
    # name: (case name, nx, displacement amplitude in mesh sizes, seed)
    "nextrows_stable24": ("10km_stable", 24, 0.2, 7),
}
FORCING_T = (23741.25, 23741.5, 23741.25 + 0.25 * 0.3771)      # ftime_range[0], [1], current time


def next_inputs(spec):
    """Seeded inputs shared by the oracle run here and the CUDA run of tests/test_gpu_golden.py."""
    name, nx, amp, seed = spec
    c = cases.make_case(name, nranks=1, dyn="bbm", nx=nx, open_east=True)
    rng = np.random.default_rng(seed)
    c.state["M_UM"] = amp * c.gm.resolution * rng.uniform(-1.0, 1.0, 2 * c.gm.nn)
    c.local = [cases.local_fields(c, lm) for lm in c.lms]
    nn = c.gm.nn
    slices = {k: (rng.normal(0.0, 7.0, n), rng.normal(0.0, 7.0, n)) for k, n in
              (("M_wind", 2 * nn), ("M_ocean", 2 * nn), ("M_ssh", nn))}
    return c, slices


def run_next(spec):
    c, slices = next_inputs(spec)
    (R,) = ob.make_ranks(c)
    ang, jmin, jmax, flip, regrid = R.check_regridding(10.0)
    R.update_ice_diagnostics(ob.orc_params(c.params))
    out = {"regrid": np.array([ang, jmin, jmax, float(flip), float(regrid)])}
    for k in ("D_conc", "D_thick", "D_snow_thick", "D_sigma0", "D_sigma1", "D_divergence"):
        out[k] = R.get(k)
    t0, t1, t = FORCING_T
    for k, (d0, d1) in slices.items():
        out["forcing_" + k] = orc.external_data_get_vector(d0, d1, True, t, t0, t1, 0.97, -0.013)
    return out


def main():
    for key, spec in NEXT_CASES.items():
        out = run_next(spec)
        np.savez_compressed(os.path.join(HERE, key + ".npz"), **out)
        print(key, out["regrid"])
    for key, spec in CASES.items():
        out = run(spec)
        np.savez_compressed(os.path.join(HERE, key + ".npz"), **out)
        print(key, {k: float(np.abs(v).max()) for k, v in list(out.items())[:3]})


if __name__ == "__main__":
    main()
