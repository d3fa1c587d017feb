"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerance (BASELINE.json north_star): relative L2 <= 1e-9 on velocity, stress and damage after one model
time step (120 sub-cycles); every other output of explicitSolve()/update() is held to the same bound.

Conditioning.  The reference algorithm itself is not always reproducible to 1e-9 over 120 sub-cycles:
  * EVP with the ice at rest (toy case): delta -> 0, zeta = P/dmin; a 1e-15 relative wind perturbation grows
    to 6e-4 inside the oracle;
  * BBM on the BASELINE.md section 4 state (conc~U(.85,1), damage~U(0,.8)): in very weak elements the viscous
    relaxation time is << the sub-cycle and sigma_n runs into a growing period-2 oscillation around -Pmax
    (FE.cpp:4189-4200); the same perturbation grows to 1e-7..1e-5.
Strict full-step tests therefore use states on which the oracle is reproducible ("toy" for BBM/mEVP,
"10km_stable" otherwise); on the ill-conditioned states parity is checked (a) one sub-cycle ahead from
oracle states sampled along the trajectory and (b) against the oracle's own sensitivity.
"""
import numpy as np
import pytest

from nextsim_b200 import capi, cases
import oracle_bridge as ob
from oracle import oracle as orc

TOL = 1e-9
pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["tiles", "direct", "resident"])
def solver_path(request, monkeypatch):
    """Every parity test runs on both sub-cycle implementations: the TMA tile pipeline (HBM-bound meshes) and the
    direct element/node kernels (L2-resident meshes).  NSX_PATH overrides the size heuristic of nsx_create."""
    monkeypatch.setenv("NSX_PATH", request.param)
    return request.param


def solve_gpu(solvers):
    if len(solvers) == 1:
        solvers[0].explicit_solve()
    else:
        capi.group_explicit_solve(solvers)


def errors(keys, got, ref):
    out = {}
    for r, (g, f) in enumerate(zip(got, ref)):
        for k in keys:
            if k == "M_sigma":
                for i in range(3):
                    out[("M_sigma%d" % i, r)] = ob.rel_l2(g[k][i], f[k][i])
            else:
                out[(k, r)] = ob.rel_l2(g[k], f[k])
    return out


def assert_parity(keys, got, ref, tol=TOL, tag=""):
    for (k, r), e in errors(keys, got, ref).items():
        assert e <= tol, "%s%s rank %d rel-L2 %.3e > %.1e" % (tag, k, r, e, tol)


def run_both(c, do_update=True, tol=TOL, fast=False):
    ranks = ob.make_ranks(c, fast=fast)
    q = ob.orc_params(c.params)
    orc.explicit_solve(ranks, q)
    solvers = cases.make_solvers(c)
    solve_gpu(solvers)
    ref = [ob.get_state(R, cases.STATE_OUT) for R in ranks]
    got = [s.download(*cases.STATE_OUT) for s in solvers]
    assert_parity(cases.STATE_OUT, got, ref, tol)
    if do_update:
        for R in ranks:
            R.update(q)
        for s in solvers:
            s.update()
        ref = [ob.get_state(R, cases.UPDATE_OUT) for R in ranks]
        got = [s.download(*cases.UPDATE_OUT) for s in solvers]
        assert_parity(cases.UPDATE_OUT, got, ref, tol, tag="update:")
    chk = solvers[0].check()
    assert chk.n_nan == 0 and chk.n_range == 0
    for s in solvers:
        s.close()


@pytest.mark.parametrize("dyn", ["bbm", "mevp"])
def test_toy_one_rank(dyn):
    """BASELINE config #1 (stand-in for nextsim.toy.cfg): full model step + update()."""
    run_both(cases.make_case("toy", nranks=1, dyn=dyn))


@pytest.mark.parametrize("dyn", ["bbm", "mevp", "evp"])
@pytest.mark.parametrize("nsub", [1, 2])
def test_toy_first_substeps(dyn, nsub):
    c = cases.make_case("toy", nranks=1, dyn=dyn)
    c.params.stop_after_substeps = nsub
    c.params.skip_ow_smoother = 1
    run_both(c, do_update=False, tol=1e-12)


@pytest.mark.parametrize("dyn,young,open_east", [("bbm", True, True), ("bbm", False, False), ("mevp", True, True),
                                                 ("evp", True, True)])
def test_stable_state_full_step(dyn, young, open_east):
    run_both(cases.make_case("10km_stable", nranks=1, dyn=dyn, nx=64, open_east=open_east, young=young))


@pytest.mark.parametrize("name,nx,nranks,dyn,open_east", [
    ("toy", None, 2, "bbm", True), ("toy", None, 3, "mevp", False), ("toy", None, 4, "bbm", False),
    ("10km_stable", 64, 8, "evp", True), ("10km_stable", 96, 5, "bbm", True)])
def test_partitioned_group(name, nx, nranks, dyn, open_east):
    """Several ranks stepped in lock-step on one GPU: partition indexing, ghost elements, halo push."""
    run_both(cases.make_case(name, nranks=nranks, dyn=dyn, open_east=open_east, nx=nx))


def test_10km_full_size_bbm():
    """BASELINE config #2 mesh at full size (199 712 elements), well-conditioned state."""
    run_both(cases.make_case("10km_stable", nranks=1, dyn="bbm"))


def test_10km_full_size_mevp():
    """BASELINE config #3 (mEVP on the 10 km mesh) with the BASELINE state."""
    run_both(cases.make_case("10km", nranks=1, dyn="mevp"))


def test_multistep_state_stays_resident():
    """Three model steps without re-uploading: device state carries over exactly like the host members."""
    c = cases.make_case("10km_stable", nranks=1, dyn="bbm", nx=48)
    ranks = ob.make_ranks(c)
    q = ob.orc_params(c.params)
    solvers = cases.make_solvers(c)
    for _ in range(3):
        orc.explicit_solve(ranks, q)
        ranks[0].update(q)
        solvers[0].explicit_solve()
        solvers[0].update()
    ref = [ob.get_state(ranks[0], cases.STATE_OUT)]
    got = [solvers[0].download(*cases.STATE_OUT)]
    assert_parity(("M_VT", "M_sigma", "M_damage", "M_UM"), got, ref, tag="3 steps:")
    solvers[0].close()


# ---- the ill-conditioned BASELINE state -----------------------------------------------------------------
CARRY = ("M_VT", "M_UM", "M_UT", "M_damage")


@pytest.mark.parametrize("nx", [128, None])
@pytest.mark.parametrize("k", [0, 40, 80, 119])
def test_baseline_state_one_substep_ahead(k, nx):
    """From the oracle's state after k sub-cycles (weak elements already oscillating at k >= 60), one more
    sub-cycle on both sides agrees to 1e-12.  nx=None is the EXACT configuration bench.py reports (BASELINE config #2:
    nx=316, 199 712 elements, BBM, "large" state, closed coast)."""
    c = cases.make_case("10km", nranks=1, dyn="bbm", nx=nx, open_east=(nx is not None))
    if k:
        c.params.stop_after_substeps = k
        c.params.skip_ow_smoother = 1
        R = ob.make_ranks(c, fast=True)[0]
        orc.explicit_solve([R], ob.orc_params(c.params))
        for key in CARRY:
            c.local[0][key] = R.get(key)
        c.local[0]["M_sigma"] = [R.get("M_sigma%d" % i) for i in range(3)]
    c.params.stop_after_substeps = 1
    c.params.skip_ow_smoother = 1
    run_both(c, do_update=False, tol=1e-12, fast=True)


def test_bench_configuration_first_substeps():
    """The bench configuration itself (nx=316, BBM, BASELINE state) over its first 12 sub-cycles, before the
    ill-conditioned elements have amplified rounding noise: every output within 1e-9 of the oracle."""
    c = cases.make_case("10km", nranks=1, dyn="bbm")
    c.params.stop_after_substeps = 12
    c.params.skip_ow_smoother = 1
    run_both(c, do_update=False, fast=True)


def test_baseline_state_full_step_within_oracle_sensitivity():
    """Full step on the BASELINE state: the CUDA result is as close to the oracle as the oracle is to itself
    under a 1e-15 relative perturbation of the wind (max over three perturbations, factor 100 slack)."""
    def oracle_run(seed):
        c = cases.make_case("10km", nranks=1, dyn="bbm", nx=128, open_east=True)
        if seed is not None:
            rng = np.random.default_rng(seed)
            w = c.local[0]["M_wind"]
            c.local[0]["M_wind"] = w * (1 + 1e-15 * rng.standard_normal(w.size))
        R = ob.make_ranks(c)[0]
        orc.explicit_solve([R], ob.orc_params(c.params))
        return c, ob.get_state(R, ("M_VT", "M_sigma", "M_damage"))

    c, ref = oracle_run(None)
    keys = ("M_VT", "M_sigma", "M_damage")
    sens = {}
    for seed in (1, 2, 3):
        _, p = oracle_run(seed)
        for kk, e in errors(keys, [p], [ref]).items():
            sens[kk] = max(sens.get(kk, 0.0), e)
    solvers = cases.make_solvers(c)
    solvers[0].explicit_solve()
    got = solvers[0].download(*keys)
    err = errors(keys, [got], [ref])
    solvers[0].close()
    print("oracle self-sensitivity:", sens)
    print("cuda vs oracle         :", err)
    for kk, e in err.items():
        assert e <= max(TOL, 100.0 * sens[kk]), (kk, e, sens[kk])


# ---- boundary behaviour of the C ABI ----------------------------------------------------------------------
def test_upload_download_roundtrip_is_bit_exact():
    """The library renumbers internally (Hilbert order, tiles); the host must get back exactly what it sent."""
    c = cases.make_case("10km_stable", nranks=1, nx=40, open_east=True)
    s = cases.make_solvers(c)[0]
    rng = np.random.default_rng(7)
    nn, ne = c.lms[0].num_nodes, c.lms[0].num_elements
    sent = {"M_VT": rng.standard_normal(2 * nn), "M_UM": rng.standard_normal(2 * nn), "M_ssh": rng.standard_normal(nn),
            "M_damage": rng.random(ne), "M_conc": rng.random(ne), "M_sigma": [rng.standard_normal(ne) for _ in range(3)]}
    s.upload(**sent)
    got = s.download(*sent.keys())
    for k, v in sent.items():
        if k == "M_sigma":
            for a, b in zip(got[k], v):
                assert np.array_equal(a, b)
        else:
            assert np.array_equal(got[k], v), k
    s.close()


def test_wave_stress_slot():
    """Optional OASIS wave stress tau_wi enters the nodal solve (FE.cpp:10408-10415, 10510-10517)."""
    c = cases.make_case("10km_stable", nranks=1, dyn="bbm", nx=48)
    rng = np.random.default_rng(3)
    tau = 0.05 * rng.standard_normal(2 * c.lms[0].num_nodes)
    R = ob.make_ranks(c)[0]
    R.set("tau_wi", tau)
    orc.explicit_solve([R], ob.orc_params(c.params))
    s = cases.make_solvers(c)[0]
    s.upload(M_tau_wi=tau)
    s.explicit_solve()
    got = s.download("M_VT", "M_damage")
    assert ob.rel_l2(got["M_VT"], R.get("M_VT")) <= TOL
    assert ob.rel_l2(got["M_damage"], R.get("M_damage")) <= TOL
    # and it matters: without the stress the velocities differ
    R2 = ob.make_ranks(c)[0]
    orc.explicit_solve([R2], ob.orc_params(c.params))
    assert ob.rel_l2(R2.get("M_VT"), R.get("M_VT")) > 1e-6
    s.close()


def test_check_reports_bad_fields():
    """nsx_check is the device-side checkFieldsFast (FE.cpp:14536-14655)."""
    c = cases.make_case("toy", nranks=1)
    s = cases.make_solvers(c)[0]
    assert s.check().n_nan == 0
    vt = c.local[0]["M_VT"].copy()
    vt[5] = np.nan
    vt[7] = 9.0
    s.upload(M_VT=vt)
    chk = s.check()
    assert chk.n_nan == 1 and chk.n_speed == 1 and chk.max_speed == 9.0
    d = c.local[0]["M_damage"].copy()
    d[3] = 1.5
    s.upload(M_damage=d)
    assert s.check().n_range == 1
    s.close()


def test_params_required_and_validated():
    c = cases.make_case("toy", nranks=1)
    s = capi.Solver(c.lms[0])
    with pytest.raises(RuntimeError, match="before nsx_set_params"):
        s.explicit_solve()
    p = capi.default_params()
    p.dynamics_type = 2                       # free_drift: outside the accelerated path
    with pytest.raises(RuntimeError, match="dynamics_type"):
        s.set_params(p)
    s.close()


# ---- full BASELINE sizes: size-independent properties -------------------------------------------------------
@pytest.mark.parametrize("dyn", ["bbm", "mevp"])
def test_3km_full_size_paths_agree_and_invariants(dyn, solver_path, monkeypatch):
    """BASELINE config #4 mesh (2 000 000 elements; the oracle would need minutes): the two independent sub-cycle
    implementations (TMA tile pipeline with Hilbert tiles and redundant halo slots vs. direct element/node
    kernels) must agree to 1e-9 after a full model step + update(), the checkFieldsFast invariants must hold,
    Dirichlet nodes must stay at rest and update() must conserve ice volume."""
    if solver_path != "tiles":
        pytest.skip("runs both paths itself (2e6 elements do not fit the resident path)")
    c = cases.make_case("3km_stable", nranks=1, dyn=dyn, young=False)
    keys = ("M_VT", "M_sigma", "M_damage", "M_UM", "M_thick", "M_conc", "M_surface")
    res = {}
    for path in ("tiles", "direct"):
        monkeypatch.setenv("NSX_PATH", path)
        s = cases.make_solvers(c)[0]
        s.explicit_solve()
        pre = s.download("M_thick", "M_surface")
        s.update()
        res[path] = s.download(*keys)
        chk = s.check()
        assert chk.n_nan == 0 and chk.n_range == 0 and chk.n_speed == 0
        vol0 = pre["M_thick"] * pre["M_surface"]
        vol1 = res[path]["M_thick"] * res[path]["M_surface"]
        ice = pre["M_thick"] > 0
        assert np.allclose(vol1[ice], vol0[ice], rtol=1e-12)
        s.close()
    for k in keys:
        pairs = zip(res["tiles"][k], res["direct"][k]) if k == "M_sigma" else [(res["tiles"][k], res["direct"][k])]
        for a, b in pairs:
            assert ob.rel_l2(a, b) <= TOL, k
    nn = c.gm.nn
    d = c.gm.dirichlet_flags_root - 1
    assert np.all(res["tiles"]["M_VT"][d] == 0.0) and np.all(res["tiles"]["M_VT"][d + nn] == 0.0)
    assert np.abs(res["tiles"]["M_VT"]).max() < 5.0


# ---- full BASELINE sizes against the oracle (threadless -O3 build of the same source, ~25 s per case) -------------
@pytest.mark.parametrize("nranks", [1, 8])
def test_3km_full_size_against_oracle(nranks, solver_path):
    """BASELINE config #4 mesh (nx=1000, 2 000 000 elements), full model step + update(): CUDA vs the oracle to 1e-9 on
    one rank and on 8 ranks (the reference's partition indexing, ghost elements, updateGhosts every sub-cycle)."""
    if solver_path == "resident":
        pytest.skip("2e6 elements do not fit the resident path on one GPU")
    if solver_path == "tiles" and nranks == 8:
        pytest.skip("8 x 2.5e5 elements: the direct path is what nsx_create selects at that size")
    run_both(cases.make_case("3km_stable", nranks=nranks, dyn="bbm", young=False), fast=True)


def test_1km_first_substeps_against_oracle(solver_path):
    """BASELINE config #5 mesh (nx=3162, 19 996 488 elements): 12 sub-cycles, CUDA tile pipeline vs the oracle."""
    if solver_path != "tiles":
        pytest.skip("2e7 elements: the TMA tile pipeline is the path for this size")
    c = cases.make_case("1km_stable", nranks=1, dyn="bbm", young=False)
    c.params.stop_after_substeps = 12
    c.params.skip_ow_smoother = 1
    run_both(c, do_update=False, fast=True)
