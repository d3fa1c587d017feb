"""GPU parity of FiniteElement::thermo() through the C ABI (nsx_thermo*).

The kernel (nsx_thermo.cu) and the CPU side (oracle.thermo) are the same element function compiled for the two targets,
the kernel with -fmad=false, so the only difference left is the device libm (exp, pow, log, cbrt, atan, hypot: <= 2 ulp).
The CPU side is held BIT FOR BIT to the reference's own bodies by tests/test_thermo_cpu.py, and the golden fixtures used
here were written from those reference bodies.  Bar: 1e-9 relative L2 per field (north_star's tolerance); what is
observed is ~1e-15.  Counters and flags (M_freeze_days, M_freeze_onset) must be identical.
"""
import os

import numpy as np
import pytest

from nextsim_b200 import capi, cases, partition as pt, synthetic as syn
import thermo_common as tc

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-9
EXACT = ("M_freeze_days", "M_freeze_onset")
PRIVATE_IN = syn.THERMO_FORCING + syn.THERMO_STATE


def run_gpu(p, t, dt, S, nx, nranks=1, steps=1, path="tiles"):
    """Scatter the global thermo state to `nranks` handles, run thermo() `steps` times, gather the owned entries."""
    c = cases.make_case("10km_stable", nranks=nranks, nx=nx, young=bool(p.ice_cat_young))
    solvers = cases.make_solvers(c, path=path)
    nn = c.gm.nn
    for s, lm in zip(solvers, c.lms):
        s.upload(**{k: pt.scatter_elem(lm, S[k]) for k in syn.THERMO_ICE},
                 **{k: pt.scatter_nodal2(lm, S[k], nn) for k in ("M_wind", "M_VT", "M_ocean")})
        s.thermo_upload(**{k: pt.scatter_elem(lm, S[k]) for k in PRIVATE_IN})
    q = capi.thermo_default_params()
    for (n, _) in q._fields_:
        setattr(q, n, getattr(p, n))
    for k in range(steps):
        for s in solvers:
            s.thermo(q, dt, t + k * dt / 86400.0)
    per_rank = [s.thermo_download(*tc.OUT_FIELDS) for s in solvers]
    out = {k: pt.gather_elem(c.lms, [d[k] for d in per_rank], c.gm.ne) for k in tc.OUT_FIELDS}
    # the ice state is the dynamics' own: NsxFields sees what thermo() wrote
    d0 = solvers[0].download("M_conc", "M_thick", "M_time_relaxation_damage")
    for k in d0:
        assert np.array_equal(d0[k], per_rank[0][k]), k
    for s in solvers:
        s.close()
    return out


def assert_close(ref, got, what, tol=TOL):
    worst = 0.0
    for k, a in ref.items():
        b = got[k]
        assert np.isfinite(b).all(), (what, k)
        if k in EXACT:
            assert np.array_equal(a, b), (what, k)
            continue
        na = np.linalg.norm(a)
        err = np.linalg.norm(a - b)
        if na == 0.0:
            assert err == 0.0, (what, k, err)
            continue
        worst = max(worst, err / na)
        assert err / na <= tol, "%s: %s rel-L2 %.3e > %.1e" % (what, k, err / na, tol)
    return worst


@pytest.mark.parametrize("name", sorted(tc.OPTION_SETS))
def test_kernel_matches_cpu_every_branch(name):
    p, t, dt, gm, S = tc.make_inputs(name, nx=24)
    ref = tc.run_oracle(p, t, dt, gm, S)
    got = run_gpu(p, t, dt, S, nx=24)
    w = assert_close({k: ref[k] for k in tc.OUT_FIELDS}, got, name)
    print("thermo %s: worst rel-L2 %.2e" % (name, w))


@pytest.mark.parametrize("name", ["defaults", "zero_layer", "alb4_ponds", "nudged_ocean"])
def test_kernel_matches_reference_golden(name):
    """fixtures written from the reference's own thermo() (tests/golden/thermo/make_golden.py), 3 consecutive calls"""
    G = np.load(os.path.join(HERE, "golden", "thermo", "%s.npz" % name))
    p, t, dt, gm, S = tc.make_inputs(name, nx=int(G["nx"]), seed=int(G["seed"]))
    got = run_gpu(p, t, dt, S, nx=int(G["nx"]), steps=int(G["steps"]))
    ref = {k[4:]: G[k] for k in G.files if k.startswith("out_")}
    assert_close(ref, {k: got[k] for k in ref}, "golden " + name)


@pytest.mark.parametrize("path", ["resident", "direct"])
def test_three_ranks_and_other_layouts(path):
    """element-wise, so ranks only differ in numbering: 3 in-process ranks, internal orders of the other two paths"""
    name = "ponds_alb3"
    p, t, dt, gm, S = tc.make_inputs(name, nx=30)
    ref = tc.run_oracle(p, t, dt, gm, S, steps=2)
    got = run_gpu(p, t, dt, S, nx=30, nranks=3, steps=2, path=path)
    assert_close({k: ref[k] for k in tc.OUT_FIELDS}, got, name)


def test_large_mesh_five_steps():
    """160 k elements, five calls: drag coefficients, layer temperatures, ponds and tracers feed back"""
    name = "defaults"
    p, t, dt, gm, S = tc.make_inputs(name, nx=283)
    ref = tc.run_oracle(p, t, dt, gm, S, steps=5)
    got = run_gpu(p, t, dt, S, nx=283, steps=5, path="resident")
    w = assert_close({k: ref[k] for k in tc.OUT_FIELDS}, got, name)
    print("thermo 160k x5: worst rel-L2 %.2e" % w)


def test_inside_the_model_step():
    """thermo() between two explicitSolve()+update() calls on the same handle: the dynamics reads what thermo() wrote"""
    from oracle import oracle as orc  # noqa: F401  (only to make sure the dynamics oracle builds on this box)
    c = cases.make_case("10km_stable", nx=48)
    s = cases.make_solvers(c)[0]
    p = capi.thermo_default_params(dtime_step=c.params.dtime_step)
    S = syn.make_thermo_state(c.gm.ne, c.gm.nn, seed=3, young=True, season="winter")
    lm = c.lms[0]
    s.thermo_upload(**{k: pt.scatter_elem(lm, S[k]) for k in PRIVATE_IN})
    before = s.download("M_conc", "M_thick")
    s.explicit_solve(); s.update()
    mid = s.download("M_conc", "M_thick")
    s.thermo(p, int(c.params.dtime_step), tc.datenum(2018, 2, 3, 0.25))
    after = s.download("M_conc", "M_thick", "M_snow_thick")
    assert not np.array_equal(mid["M_thick"], after["M_thick"])
    assert np.isfinite(after["M_thick"]).all() and (after["M_thick"] >= 0).all() and (after["M_conc"] <= 1.0).all()
    s.explicit_solve(); s.update()
    chk = s.check()
    assert chk.n_nan == 0 and chk.n_range == 0
    assert np.isfinite(s.download("M_VT")["M_VT"]).all()
    del before
    s.close()


def test_rejected_options_and_names():
    c = cases.make_case("toy")
    s = cases.make_solvers(c)[0]
    for over in (dict(newice_type=5, ice_cat_young=0), dict(melt_type=3), dict(alb_scheme=0), dict(newice_type=4, ice_cat_young=0)):
        p = capi.thermo_default_params(**over)
        with pytest.raises(RuntimeError):
            s.thermo(p, 200, 43000.0)
    with pytest.raises(RuntimeError):
        s.thermo(capi.thermo_default_params(), 0, 43000.0)
    with pytest.raises(RuntimeError):
        s.thermo_upload(M_no_such_field=np.zeros(s.ne))
    with pytest.raises(RuntimeError):
        s.thermo_upload(M_conc=np.zeros(s.ne))          # ice state goes through NsxFields
    s.close()
