"""thermo(), CPU side: the element function the kernel is compiled from == the reference, bit for bit.

oracle/ref_fe compiles the reference's OWN thermo(), OWBulkFluxes(), IABulkFluxes(), specificHumidity(), albedo(),
thermoWinton(), thermoIce0(), meltPonds(), iceOceanHeatflux(), freezingPoint(), incomingLongwave(), windSpeedElement() text
(FE.cpp:4965-6962, cut at build time).  oracle.thermo is nsx::thermo::thermo_element() of nextsim_b200/csrc/nsx_thermo.cuh
compiled for the host with the same flags (-O2 -ffp-contract=off).  Same inputs, same options: every state and diagnostic
field must be IDENTICAL, over option sets that reach every branch (both thermodynamic schemes, the four new-ice and two
lateral-melt schemes, the four albedo schemes, melt ponds, nudged / constant ocean, both ice-ocean heat-flux and
freezing-point schemes, every forcing-variable fallback, assimilation flux, temperature-dependent healing, flooding
off, first / last step of the day, the 15 September / 1 August / reset-date midnights).  The golden fixtures of the GPU
tests (tests/golden/thermo/*.npz) are generated from the same reference bodies and are checked here against the host build.
"""
import os

import numpy as np
import pytest

import thermo_common as tc
from oracle import ref_fe

HERE = os.path.dirname(os.path.abspath(__file__))
needs_ref = pytest.mark.skipif(not ref_fe.available(), reason="oracle/_ref/libref_fe.so not built")


def assert_identical(ref, got, what):
    for k, a in ref.items():
        b = got[k]
        if not np.array_equal(a, b, equal_nan=True):
            bad = np.flatnonzero(~((a == b) | (np.isnan(a) & np.isnan(b))))
            raise AssertionError("%s: %s differs from the reference bodies in %d of %d entries, first %d: %r vs %r"
                                 % (what, k, bad.size, a.size, bad[0], a[bad[0]], b[bad[0]]))
        assert np.isfinite(a).all(), (what, k, "non-finite values in the reference result: the inputs left the physical range")


@needs_ref
@pytest.mark.parametrize("name", sorted(tc.OPTION_SETS))
def test_host_build_equals_reference_bodies(name):
    p, t, dt, gm, S = tc.make_inputs(name)
    ref = tc.run_reference(p, t, dt, gm, S)
    got = tc.run_oracle(p, t, dt, gm, S)
    assert_identical(ref, got, name)


@needs_ref
@pytest.mark.parametrize("name", ["defaults", "zero_layer", "alb4_ponds", "nudged_ocean", "last_step_of_day"])
def test_several_steps(name):
    """state carried through 5 calls (drag coefficients, ice temperatures, ponds, tracers feed back)"""
    p, t, dt, gm, S = tc.make_inputs(name, nx=16, seed=4242)
    ref = tc.run_reference(p, t, dt, gm, S, steps=5)
    got = tc.run_oracle(p, t, dt, gm, S, steps=5)
    assert_identical(ref, got, name)


def test_branch_coverage_of_the_inputs():
    """the synthetic state really reaches the branches the parametrisation claims (otherwise parity proves little)"""
    p, t, dt, gm, S = tc.make_inputs("defaults")
    out = tc.run_oracle(p, t, dt, gm, S)
    assert (S["M_conc"] == 0).sum() > 20 and (S["M_conc"] > 0.6).sum() > 200
    assert (out["D_newice"] > 0).sum() > 20                       # supercooled leads form ice
    assert ((out["M_conc"] == 0) & (S["M_conc"] > 0)).sum() > 5   # thin ice melts through / falls under hmin
    assert (out["D_del_hi"] < 0).sum() > 20 and (out["D_del_hi"] > 0).sum() > 20
    assert (out["D_snow2ice"] > 0).sum() > 0                      # flooding
    assert (out["M_conc_young"] < S["M_conc_young"]).sum() > 5    # young ice promoted to old ice
    p, t, dt, gm, S = tc.make_inputs("alb4_ponds")
    out = tc.run_oracle(p, t, dt, gm, S)
    assert (out["D_pond_fraction"] > 0).sum() > 20 and (out["M_lid_volume"] > 0).sum() > 0
    p, t, dt, gm, S = tc.make_inputs("last_step_of_day")
    out = tc.run_oracle(p, t, dt, gm, S)
    assert (out["M_freeze_days"] > S["M_freeze_days"]).sum() > 10
    assert ((out["M_freeze_days"] == 0) & (S["M_freeze_days"] > 0) & (out["M_conc"] > 0)).sum() > 10


def test_dates():
    """datenumToString(M_current_time, "%m%d") (core/include/date.hpp:87-120): the product's decoder (era arithmetic) and the
    stand-in the reference bodies are compiled against (a walk over the calendar) against Python's datetime, day by day
    from 1900 to 2100, plus the carry of a time of day that rounds to 24:00:00.000"""
    import ctypes as C
    import datetime
    from oracle import thermo as oth
    L = oth.lib()
    L.orc_thermo_month_day.argtypes = [C.c_double]
    fns = [L.orc_thermo_month_day]
    if ref_fe.available():
        R = ref_fe.lib()
        R.ref_fe_month_day.argtypes = [C.c_double]
        fns.append(R.ref_fe_month_day)
    epoch = datetime.date(1900, 1, 1)
    for n in list(range(0, 73415, 1)):
        d = epoch + datetime.timedelta(days=n)
        want = 100 * d.month + d.day
        for f in fns:
            assert f(float(n)) == want, (n, d)
    for f in fns:
        assert f(tc.datenum(2018, 9, 15, 0.75)) == 915
        assert f(tc.datenum(2018, 9, 14, 0.999999)) == 914
        assert f(tc.datenum(2018, 9, 14) + (1.0 - 2e-9)) == 915       # 23:59:59.9998 rounds to 24:00:00.000
        assert f(tc.datenum(2020, 2, 28) + (1.0 - 2e-9)) == 229
        assert f(tc.datenum(1900, 2, 28) + 1.0) == 301                 # 1900 is not a leap year


def test_date_branches_fire():
    """the midnights of 15 September / 1 August / the reset date really take their branches (FE.cpp:6003-6007, 6051-6077, 6029-6034)"""
    p, t, dt, gm, S = tc.make_inputs("sept15_midnight")
    out = tc.run_oracle(p, t, dt, gm, S)
    ice = out["M_conc"] > 0
    assert ice.sum() > 200 and (out["M_fyi_fraction"][ice] == 0).all() and (S["M_fyi_fraction"][ice] > 0).any()
    p, t, dt, gm, S = tc.make_inputs("aug01_midnight")
    out = tc.run_oracle(p, t, dt, gm, S)
    ice = out["M_conc"] > 0
    changed = out["M_conc_summer"][ice] != S["M_conc_summer"][ice]
    assert changed.mean() > 0.9                                   # every ice element resets its summer minimum
    assert set(np.unique(out["M_freeze_onset"][ice])) <= {0.0, 1.0} and (out["M_freeze_onset"][ice] == 0).sum() > 100
    p, t, dt, gm, S = tc.make_inputs("reset_by_date_hit")
    out = tc.run_oracle(p, t, dt, gm, S)
    ice = out["M_conc"] > 0
    cmax = np.minimum(1.0, out["M_conc"] + out["M_conc_young"])
    assert np.allclose(out["M_conc_myi"][ice], cmax[ice], rtol=0, atol=0) and (out["D_del_ci_rplnt_myi"][ice] != 0).sum() > 100
    p, t, dt, gm, S = tc.make_inputs("reset_by_date_miss")
    out = tc.run_oracle(p, t, dt, gm, S)
    assert (out["D_del_ci_rplnt_myi"] == 0).all()


def test_rejected_options():
    """the values the reference throws std::logic_error on (FE.cpp:5562, 5653, 6527), and melt_type 3 (OASIS-only)"""
    for over in (dict(newice_type=5, ice_cat_young=0), dict(melt_type=3), dict(melt_type=0), dict(alb_scheme=7),
                 dict(newice_type=4, ice_cat_young=0), dict(newice_type=1, ice_cat_young=1), dict(thermo_type=2)):
        p, t, dt, gm, S = tc.make_inputs("defaults", nx=4)
        for k, v in over.items():
            setattr(p, k, v)
        with pytest.raises(RuntimeError):
            tc.run_oracle(p, t, dt, gm, S)


@pytest.mark.parametrize("name", ["defaults", "zero_layer", "alb4_ponds", "nudged_ocean"])
def test_golden_fixture(name):
    """fixtures written by tests/golden/thermo/make_golden.py from the reference bodies; travel to the GPU box"""
    path = os.path.join(HERE, "golden", "thermo", "%s.npz" % name)
    G = np.load(path)
    p, t, dt, gm, S = tc.make_inputs(name, nx=int(G["nx"]), seed=int(G["seed"]))
    got = tc.run_oracle(p, t, dt, gm, S, steps=int(G["steps"]))
    ref = {k[4:]: G[k] for k in G.files if k.startswith("out_")}
    assert_identical(ref, got, name)
