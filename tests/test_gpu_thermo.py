"""GPU parity of FiniteElement::thermo() through the C ABI (nsx_thermo*).

The kernel (nsx_thermo.cu) and the CPU side (oracle.thermo) are the same element function compiled for the two targets,
the kernel with -fmad=false; the device build differs by CUDA's libm (exp, log, cbrt, atan) and by its marked shortcuts
(integer powers as products, reciprocals of repeated divisors), each <= 1-2 ulp.  The CPU side is held BIT FOR BIT to the
reference's own bodies by tests/test_thermo_cpu.py, and the golden fixtures used here were written from those reference
bodies.  Bar: 1e-9 relative L2 per field (north_star's tolerance); what is observed is 1e-15 ... 3e-13.  Counters and
flags (M_freeze_days, M_freeze_onset) must be identical.
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


CANCELLING = ("D_del_hi", "D_del_hi_young")


def assert_close(ref, got, what, tol=TOL, report=False, rate_scale=None):
    worst = 0.0
    errs = {}
    bad = []
    for k, a in ref.items():
        b = got[k]
        assert np.isfinite(b).all(), (what, k)
        if k in EXACT:
            assert np.array_equal(a, b), (what, k)
            continue
        na = np.linalg.norm(a)
        err = np.linalg.norm(a - b)
        if k in CANCELLING and rate_scale is not None:
            # D_del_hi = (hi - hi_old) * 86400 / dt: the difference of two metre-sized thicknesses that differ by micrometres.
            # One ulp of hi is 1e-16 * hi / |hi - hi_old| of the result, 1e-9 and more for thick ice that barely grows, in the
            # reference's own arithmetic as much as here.  The bar for these rates is 1e-9 of the thickness they difference.
            na = max(na, rate_scale)
        if na == 0.0:
            assert err == 0.0, (what, k, err)
            continue
        worst = max(worst, err / na)
        errs[k] = (err / na, int(np.count_nonzero(np.abs(a - b) > 1e-12 * (np.abs(a) + 1e-300))))
        if err / na > tol:
            bad.append("%s %.3e" % (k, err / na))
    assert not bad, "%s: rel-L2 above %.1e: %s" % (what, tol, ", ".join(bad))
    if report:
        top = sorted(errs.items(), key=lambda kv: -kv[1][0])[:6]
        print("  largest: " + ", ".join("%s %.1e (%d entries off by > 1e-12)" % (k, e, n) for k, (e, n) in top))
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
    # without the micrometre films of the synthetic state (see make_thermo_state): their Winton solve amplifies one ulp to
    # 1e-8, which the single-call tests above tolerate and five consecutive calls do not
    p, t, dt, gm, S = tc.make_inputs(name, nx=283, films=bool(int(os.environ.get("THERMO_TEST_FILMS", "0"))))
    ref = tc.run_oracle(p, t, dt, gm, S, steps=5)
    got = run_gpu(p, t, dt, S, nx=283, steps=5, path="resident")
    slab = np.where(ref["M_conc"] > 0, ref["M_thick"] / np.maximum(ref["M_conc"], 1e-300), 0.0)
    w = assert_close({k: ref[k] for k in tc.OUT_FIELDS}, got, name, report=True, rate_scale=np.linalg.norm(slab) * 86400.0 / dt)
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


@pytest.mark.parametrize("linear", [True, False])
def test_element_forcing_interpolation_bit_exact(linear):
    """ExternalData::get() of the element forcing (externaldata.cpp:366-455) evaluated on the device: bit-exact"""
    from oracle import oracle as orc
    c = cases.make_case("10km_stable", nx=40)
    s = cases.make_solvers(c)[0]
    rng = np.random.default_rng(5)
    t0, t1, t = 43133.25, 43133.5, 43133.25 + 0.25 * 0.6113
    for name, factor, bias in (("M_tair", 1.0, -273.15), ("M_mslp", 1.0, 0.0), ("M_precip", 1.0 / 3600.0, 0.0), ("M_Qsw_in", 0.97, 0.0)):
        d0, d1 = rng.normal(250.0, 30.0, s.ne), rng.normal(250.0, 30.0, s.ne)
        s.thermo_forcing_load(name, 0, d0)
        if linear:
            s.thermo_forcing_load(name, 1, d1)
        s.thermo_forcing_apply(name, linear, t, t0, t1, factor, bias)
        ref = orc.external_data_get_vector(d0, d1, linear, t, t0, t1, factor, bias)
        assert np.array_equal(s.thermo_download(name)[name], ref), name
    with pytest.raises(RuntimeError):
        s.thermo_forcing_apply("M_dair", True, t, t0, t1)           # never loaded
    with pytest.raises(RuntimeError):
        s.thermo_forcing_load("M_conc_upd", 0, np.zeros(s.ne))        # a ModelVariable, not ExternalData
    s.close()


@pytest.mark.parametrize("path", ["resident", "tiles"])
def test_coupled_model_steps(path):
    """FiniteElement::step() between remeshes, device-resident: per step the time interpolation of the air temperature,
    thermo(dt), explicitSolve(), update() -- against the same sequence of the CPU oracles (ExternalData, thermo, dynamics).
    The ice-ocean heat flux uses the exchange scheme, so thermo() reads the velocities the dynamics produced, and the
    dynamics reads the concentration / thickness / healing time thermo() produced."""
    from oracle import oracle as orc
    from oracle import thermo as oth
    import oracle_bridge as ob
    nsteps = 3
    c = cases.make_case("10km_stable", nx=40)
    lm, f = c.lms[0], c.local[0]
    ne, nn = lm.num_elements, lm.num_nodes
    dt = int(c.params.dtime_step)
    over = dict(Qio_type=1, temp_dep_healing=1, dtime_step=float(dt))
    S = syn.make_thermo_state(ne, nn, seed=21, young=True, season="winter")
    # a consistent start: the case's pack ice with the synthetic slab ocean / atmosphere
    for k in syn.THERMO_ICE:
        S[k] = f[k].copy()
    tfr = -0.055 * S["M_sss"]
    S["M_sst"] = tfr + 0.02
    for k in ("M_tice0", "M_tice1", "M_tice2", "M_tsurf_young"):
        S[k] = np.minimum(S[k], -1.0)
    t0, t1, tstart = 43133.0, 43133.25, 43133.0 + 100 * dt / 86400.0
    tair0, tair1 = S["M_tair"].copy(), S["M_tair"] - 3.0

    # ---- CPU ----
    R = ob.make_ranks(c)[0]
    q = ob.orc_params(c.params)
    p = oth.default_params(**over)
    F = {k: S[k].copy() for k in syn.THERMO_FORCING + syn.THERMO_STATE + syn.THERMO_ICE}
    tri0 = np.asarray(lm.indices, np.int64).reshape(-1, 3) - 1
    for k in range(nsteps):
        t = tstart + k * dt / 86400.0
        F["M_tair"] = orc.external_data_get_vector(tair0, tair1, True, t, t0, t1, 1.0, 0.0)
        for name in syn.THERMO_ICE:
            F[name] = R.get(name)
        out = oth.thermo(p, dt, t, tri0, nn, R.get("M_wind"), R.get("M_VT"), R.get("M_ocean"), F)
        F = {name: out[name] for name in F}
        for name in syn.THERMO_ICE:
            R.set(name, out[name])
        orc.explicit_solve([R], q)
        R.update(q)

    # ---- GPU ----
    s = cases.make_solvers(c, path=path)[0]
    pg = capi.thermo_default_params(**over)
    s.thermo_upload(**{k: S[k] for k in syn.THERMO_FORCING + syn.THERMO_STATE})
    s.thermo_forcing_load("M_tair", 0, tair0)
    s.thermo_forcing_load("M_tair", 1, tair1)
    for k in range(nsteps):
        t = tstart + k * dt / 86400.0
        s.thermo_forcing_apply("M_tair", True, t, t0, t1)
        s.thermo(pg, dt, t)
        s.explicit_solve()
        s.update()
    assert s.path == path
    got = s.download("M_VT", "M_conc", "M_thick", "M_damage", "M_sigma", "M_time_relaxation_damage")
    gth = s.thermo_download("M_sst", "M_sss", "M_tice0", "M_tice1", "M_age", "D_Qa", "D_Qo", "M_fyi_fraction")
    worst = 0.0
    for name, a in list(got.items()) + list(gth.items()):
        if name == "M_sigma":
            pairs = [(a[i], R.get("M_sigma%d" % i)) for i in range(3)]
        elif name in gth:
            pairs = [(a, F[name] if name in F else out[name])]
        else:
            pairs = [(a, R.get(name))]
        for x, y in pairs:
            e = ob.rel_l2(x, y)
            worst = max(worst, e)
            assert e <= TOL, "%s after %d coupled steps: rel-L2 %.3e" % (name, nsteps, e)
    print("coupled steps (%s): worst rel-L2 %.2e" % (path, worst))
    s.close()


def test_cpp_shim_thermo(tmp_path):
    """the C++ host shim (FiniteElementGPU::thermo / thermoUpload / thermoDownload) on a one-element mesh: cold air over
    open water at the freezing point forms young ice"""
    import subprocess
    root = os.path.dirname(HERE)
    libdir = os.path.join(root, "nextsim_b200")
    exe = tmp_path / "shim_smoke"
    subprocess.check_call(["g++", "-std=c++17", "-I", root, os.path.join(root, "tests", "cpp", "shim_smoke.cpp"),
                           "-o", str(exe), "-L", libdir, "-lnsx", "-Wl,-rpath," + libdir])
    r = subprocess.run([str(exe)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    assert "handle created" in r.stdout and "thermo: D_newice=" in r.stdout, r.stdout
