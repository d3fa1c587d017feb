"""GPU parity of the SURVEY.md section 8(f) rows built so far (the callers either side of the path):

  row 1  FiniteElement::checkRegridding() (FE.cpp:8298-8309) and updateIceDiagnostics() (FE.cpp:7860-7900)
  row 2  ExternalData time interpolation of wind / ocean / ssh (externaldata.cpp:366-455)

They are element- or node-wise maps of the resident state, so the bar is tighter than for the sub-cycled solve:
on identical inputs jacobians, flip flags, concentrations, divergence and forcing values are BIT-EXACT (the kernels
spell every product / sum with round-to-nearest intrinsics, the oracle is built with -ffp-contract=off); hypot()
and acos() differ from glibc by <= 2 ulp: 1e-14 relative, plus the conditioning of acos for the minimum angle.
"""
import numpy as np
import pytest

from nextsim_b200 import capi, cases
import oracle_bridge as ob
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
DIAG = ("D_conc", "D_thick", "D_snow_thick", "D_sigma", "D_divergence")


@pytest.fixture(autouse=True, params=["tiles", "direct", "resident"])
def solver_path(request, monkeypatch):
    monkeypatch.setenv("NSX_PATH", request.param)
    return request.param


def displaced_case(name, nranks, nx, amp, seed=5, **kw):
    """A case whose M_UM is a random displacement of `amp` x the mesh size (in global numbering, so ranks agree)."""
    c = cases.make_case(name, nranks=nranks, nx=nx, **kw)
    rng = np.random.default_rng(seed)
    nn = c.gm.nn
    um = amp * c.gm.resolution * rng.uniform(-1.0, 1.0, 2 * nn)
    c.state["M_UM"] = um
    c.local = [cases.local_fields(c, lm) for lm in c.lms]
    return c


def regrid_both(c, angle):
    ranks = ob.make_ranks(c)
    solvers = cases.make_solvers(c)
    out = []
    for R, s in zip(ranks, solvers):
        ref = R.check_regridding(angle)
        got = s.check_regridding(angle)
        out.append((ref, got))
        assert got.min_jacobian == ref[1] and got.max_jacobian == ref[2], "jacobian extrema are bit-exact"
        # acos is ill-conditioned for small angles: 1 ulp of its argument (hypot differs from glibc by <= 1 ulp per
        # side) moves the angle by eps/sin(angle); allow 16 ulp of the argument
        cond = 16 * np.finfo(float).eps / max(np.sin(np.deg2rad(ref[0])), 1e-300) * 180.0 / np.pi
        assert abs(got.min_angle - ref[0]) <= 1e-14 * abs(ref[0]) + cond
        assert bool(got.flip) == ref[3]
        assert bool(got.regrid) == ref[4]
    for s in solvers:
        s.close()
    return out


@pytest.mark.parametrize("nranks", [1, 3])
def test_check_regridding_moderate_displacement(nranks):
    out = regrid_both(displaced_case("10km_stable", nranks, 48, 0.15), 10.0)
    assert not any(ref[3] for ref, _ in out)
    # a looser angle threshold triggers the regrid on the same mesh
    out = regrid_both(displaced_case("10km_stable", nranks, 48, 0.15), 40.0)
    assert all(ref[4] for ref, _ in out)


def test_check_regridding_detects_flip():
    out = regrid_both(displaced_case("10km_stable", 2, 32, 0.9), 10.0)
    assert any(ref[3] for ref, _ in out), "a displacement of 0.9 h must flip some triangle"


def test_check_regridding_undisplaced_toy():
    out = regrid_both(cases.make_case("toy"), 10.0)
    (ref, got), = out
    assert got.min_angle > 10.0 and not got.flip and not got.regrid


def diag_from_oracle(R, q):
    R.update_ice_diagnostics(q)
    return {"D_conc": R.get("D_conc"), "D_thick": R.get("D_thick"), "D_snow_thick": R.get("D_snow_thick"),
            "D_sigma": [R.get("D_sigma0"), R.get("D_sigma1")], "D_divergence": R.get("D_divergence")}


def assert_diag(got, ref):
    for k in ("D_conc", "D_thick", "D_snow_thick", "D_divergence"):
        assert np.array_equal(got[k], ref[k]), "%s is bit-exact" % k
    assert np.array_equal(got["D_sigma"][0], ref["D_sigma"][0])
    np.testing.assert_allclose(got["D_sigma"][1], ref["D_sigma"][1], rtol=1e-14, atol=0)


@pytest.mark.parametrize("nranks,young", [(1, True), (1, False), (4, True)])
def test_ice_diagnostics_identical_inputs(nranks, young):
    c = displaced_case("10km_stable", nranks, 48, 0.1, young=young)
    ranks = ob.make_ranks(c)
    q = ob.orc_params(c.params)
    solvers = cases.make_solvers(c)
    for R, s in zip(ranks, solvers):
        s.update_ice_diagnostics()
        assert_diag(s.download(*DIAG), diag_from_oracle(R, q))
    for s in solvers:
        s.close()


def test_ice_diagnostics_after_solve_and_update():
    """After explicitSolve() + update() the diagnostics of the device state equal the host function applied to
    that same state (downloaded), and the regrid check sees the moved mesh."""
    c = cases.make_case("10km_stable", nranks=1, dyn="bbm", nx=64, open_east=True)
    (s,) = cases.make_solvers(c)
    s.explicit_solve()
    s.update()
    s.update_ice_diagnostics()
    got = s.download(*DIAG)
    st = s.download("M_VT", "M_UM", "M_sigma", "M_conc", "M_thick", "M_snow_thick", "M_conc_young", "M_h_young",
                    "M_hs_young")
    (R,) = ob.make_ranks(c)
    for k, v in st.items():
        if k == "M_sigma":
            for i in range(3):
                R.set("M_sigma%d" % i, v[i])
        else:
            R.set(k, v)
    assert_diag(got, diag_from_oracle(R, ob.orc_params(c.params)))
    assert np.abs(got["D_divergence"]).max() > 0
    ref = R.check_regridding(10.0)
    r = s.check_regridding(10.0)
    assert r.min_jacobian == ref[1] and r.max_jacobian == ref[2] and bool(r.regrid) == ref[4]
    s.close()


def test_diagnostics_download_before_compute_is_an_error():
    (s,) = cases.make_solvers(cases.make_case("toy"))
    with pytest.raises(RuntimeError, match="nsx_update_ice_diagnostics"):
        s.download("D_conc")
    s.close()


@pytest.mark.parametrize("nranks", [1, 3])
@pytest.mark.parametrize("linear", [True, False])
def test_forcing_interpolation_bit_exact(nranks, linear):
    c = cases.make_case("10km_stable", nranks=nranks, nx=48)
    solvers = cases.make_solvers(c)
    rng = np.random.default_rng(11)
    t0, t1, t = 23741.25, 23741.5, 23741.25 + 0.25 * 0.3771
    for s in solvers:
        for name, n, factor, bias in (("M_wind", 2 * s.nn, 1.0, 0.0), ("M_ocean", 2 * s.nn, 0.97, 0.0),
                                      ("M_ssh", s.nn, 1.0, -0.013)):
            d0 = rng.normal(0.0, 7.0, n)
            d1 = rng.normal(0.0, 7.0, n)
            s.forcing_load(name, 0, d0)
            if linear:
                s.forcing_load(name, 1, d1)
            s.forcing_apply(name, linear, t, t0, t1, factor, bias)
            ref = orc.external_data_get_vector(d0, d1, linear, t, t0, t1, factor, bias)
            got = s.download(name)[name]
            assert np.array_equal(got, ref), name
    for s in solvers:
        s.close()


def test_forcing_apply_feeds_the_solve():
    """Interpolating the wind on the device and then solving == uploading the host-interpolated wind and solving."""
    c = cases.make_case("10km_stable", nranks=1, dyn="bbm", nx=48)
    c.params.stop_after_substeps = 3
    f = c.local[0]
    w0 = 0.5 * f["M_wind"]
    w1 = 1.7 * f["M_wind"]
    t0, t1, t = 100.0, 100.25, 100.1
    host = orc.external_data_get_vector(w0, w1, True, t, t0, t1, 1.0, 0.0)
    outs = []
    for mode in ("device", "host"):
        (s,) = cases.make_solvers(c)
        if mode == "device":
            s.forcing_load("M_wind", 0, w0)
            s.forcing_load("M_wind", 1, w1)
            s.forcing_apply("M_wind", True, t, t0, t1)
        else:
            s.upload(M_wind=host)
        s.explicit_solve()
        outs.append(s.download("M_VT", "D_tau_a"))
        s.close()
    for k in ("M_VT", "D_tau_a"):
        assert np.array_equal(outs[0][k], outs[1][k]), k


def test_forcing_apply_without_load_is_an_error():
    (s,) = cases.make_solvers(cases.make_case("toy"))
    with pytest.raises(RuntimeError, match="not loaded"):
        s.forcing_apply("M_ssh", True, 1.0, 0.0, 2.0)
    s.close()
