"""Pins the oracle's PHYSICS to the reference: oracle == the reference's own function bodies, bit for bit.

oracle/ref_fe cuts FiniteElement::explicitSolve / update / updateSigmaDamage / updateSigmaVP,EVP,MEVP / updateGhosts /
sides / measure / shapeCoeff / jacobian / minAngle / flip / checkRegridding / updateIceDiagnostics / calcCohesion /
initFETensors (and GmshMesh::vertices) out of /root/reference at build time and compiles that text against a stub class
(stub_fe.hpp).  Here the oracle restatement (oracle/nextsim_oracle.cpp) and those bodies run on identical inputs --
mesh, bamg tables, halo lists, fields, options -- and every output must be IDENTICAL (np.array_equal), for BBM, EVP
and mEVP, young ice on/off, Lemieux basal stress, open (Neumann) boundaries, 1 to 4 MPI-style ranks.
Both sides are built -O2 -ffp-contract=off.  No GPU needed; skipped only if oracle/_ref/libref_fe.so is absent
(it is built by __graft_entry__.build() in the container that holds /root/reference).
"""
import numpy as np
import pytest

from nextsim_b200 import cases
import oracle_bridge as ob
from oracle import oracle as orc
from oracle import ref_fe

pytestmark = pytest.mark.skipif(not ref_fe.available(), reason="oracle/_ref/libref_fe.so not built")

SOLVE_OUT = ("M_VT", "M_UM", "M_UT", "M_sigma0", "M_sigma1", "M_sigma2", "M_damage", "D_tau_a", "D_tau_w", "M_surface",
             "M_delta_x")
UPDATE_OUT = ("M_conc", "M_thick", "M_snow_thick", "M_thick_myi", "M_conc_myi", "M_ridge_ratio", "M_conc_young",
              "M_h_young", "M_hs_young", "M_sigma0", "M_sigma1", "M_sigma2", "M_surface", "D_del_ci_ridge_myi")
FIELDS_IN = ("M_damage", "M_conc", "M_thick", "M_snow_thick", "M_conc_young", "M_h_young", "M_hs_young", "M_thick_myi",
             "M_conc_myi", "M_ridge_ratio", "M_element_depth", "M_drag_ui", "M_drag_ui_young", "M_time_relaxation_damage",
             "M_random_number", "M_VT", "M_UM", "M_UT", "M_wind", "M_ocean", "M_ssh", "M_sigma0", "M_sigma1", "M_sigma2")


def make_ref(c, ranks, q, bamg_from_reference=False):
    """Feed the reference bodies from the ORACLE's ranks (its own replay of nodalGrid / bcMarkedNodes / initUpdateGhosts)."""
    F = ref_fe.RefFE(c.nranks)
    for r, R in enumerate(ranks):
        sz = R.sizes()
        tri = R.get("indices").reshape(-1, 3)
        F.set_mesh(r, R.get("coordX"), R.get("coordY"), sz["local_ndof"], tri, R.get("M_mask_dirichlet"), R.get("M_neumann_flags"))
        nec = R.get("NodalElementConnectivity").reshape(sz["num_nodes"], sz["nec_width"])
        nc = R.get("NodalConnectivity").reshape(sz["num_nodes"], sz["nc_width"])
        if bamg_from_reference:
            from oracle import ref_bamg
            nec, nc, _ = ref_bamg.convert(R.get("coordX"), R.get("coordY"), tri)
        F.set_bamg(r, nec, nc)
        for p in range(c.nranks):
            F.set_halo(r, 0, p, R.halo(0, p))
            F.set_halo(r, 1, p, R.halo(1, p))
        F.set(r, "lat", R.get("lat"))
        for k in FIELDS_IN:
            F.set(r, k, R.get(k))
    F.set_params(q)
    for r in range(c.nranks):
        F.calc_cohesion(r, c.C_fix, c.C_alea)            # reference text: C_fix + C_alea*M_random_number
    return F


def assert_identical(F, ranks, names, what):
    for r, R in enumerate(ranks):
        for k in names:
            a, b = F.get(r, k), R.get(k)
            assert a.shape == b.shape, (what, r, k)
            if not np.array_equal(a, b):
                bad = np.flatnonzero(a != b)
                raise AssertionError("%s: rank %d %s differs from the reference bodies in %d of %d entries, first %d: %r vs %r"
                                     % (what, r, k, bad.size, a.size, bad[0], a[bad[0]], b[bad[0]]))


CASES = [
    # name, nx, dyn, nranks, open_east, young, substeps, extra option overrides
    ("toy", None, "bbm", 1, False, True, 120, {}),
    ("toy", None, "bbm", 1, True, False, 120, {}),
    ("toy", None, "mevp", 1, True, True, 120, {}),
    ("toy", None, "evp", 1, False, True, 120, {}),
    ("10km_stable", 64, "bbm", 1, True, True, 120, {}),
    ("10km_stable", 64, "bbm", 3, True, True, 120, {}),
    ("10km_stable", 64, "mevp", 4, False, True, 120, {}),
    ("10km_stable", 48, "evp", 2, True, False, 60, {}),
    ("10km", 64, "bbm", 1, True, True, 120, {}),                         # the bench's (ill-conditioned) state
    ("10km", 64, "bbm", 2, False, True, 120, {}),
    ("10km_stable", 40, "bbm", 2, True, True, 30, {"basal_stress_type": 1, "equal_ridging": 1}),
    ("10km_stable", 40, "bbm", 1, True, True, 30, {"exponent_relaxation_sigma": 4.5, "ocean_turning_angle_rad": 0.4363}),
    # branches gcov showed no other case reaches: the compression cap of the damage criterion (FE.cpp:4218-4221; ~7 % of the
    # elements with this strength), no basal stress (setup.basal_stress-type = none), the classic ice category in update()
    ("10km_stable", 40, "bbm", 2, True, True, 30, {"compr_strength": 2.0e4}),
    ("10km_stable", 40, "bbm", 1, True, True, 30, {"basal_stress_type": 0}),
    ("10km_stable", 40, "bbm", 1, False, False, 30, {"newice_type": 1, "ice_cat_type": 0}),
]


@pytest.mark.parametrize("name,nx,dyn,nranks,open_east,young,substeps,over", CASES)
def test_oracle_equals_reference_bodies(name, nx, dyn, nranks, open_east, young, substeps, over):
    c = cases.make_case(name, nranks=nranks, dyn=dyn, nx=nx, open_east=open_east, young=young, substeps=substeps)
    for k, v in over.items():
        setattr(c.params, k, v)
    if over.get("basal_stress_type"):
        for f in c.local:                                   # shallow water so that the Lemieux term is active
            f["M_element_depth"] = np.full_like(f["M_element_depth"], 3.0)
    ranks = ob.make_ranks(c)
    for R, f in zip(ranks, c.local):
        R.set("M_random_number", f["M_random_number"])
    q = ob.orc_params(c.params)
    F = make_ref(c, ranks, q)
    assert_identical(F, ranks, ("M_Cohesion",), "calcCohesion")
    orc.explicit_solve(ranks, q)
    F.explicit_solve()
    assert_identical(F, ranks, SOLVE_OUT, "explicitSolve")
    for r, R in enumerate(ranks):
        assert np.array_equal(F.shape_coeff(r, R.sizes()["num_elements"]), R.get("M_shape_coeff")), "M_shape_coeff"
    for R in ranks:
        R.update(q)
    F.update()
    assert_identical(F, ranks, UPDATE_OUT, "update")
    # SURVEY 8(f) row 1 on the moved mesh
    for R in ranks:
        R.update_ice_diagnostics(q)
    F.update_ice_diagnostics()
    assert_identical(F, ranks, ("D_conc", "D_thick", "D_snow_thick", "D_sigma0", "D_sigma1", "D_divergence"), "updateIceDiagnostics")
    for r, R in enumerate(ranks):
        a, b = F.check_regridding(r), R.check_regridding(10.0)
        assert a == b, ("checkRegridding", r, a, b)


def test_second_step_stays_identical():
    """Two model steps back to back (state carried over by both sides independently)."""
    c = cases.make_case("10km_stable", nranks=2, dyn="bbm", nx=48, open_east=True)
    ranks = ob.make_ranks(c)
    for R, f in zip(ranks, c.local):
        R.set("M_random_number", f["M_random_number"])
    q = ob.orc_params(c.params)
    F = make_ref(c, ranks, q)
    for _ in range(2):
        orc.explicit_solve(ranks, q)
        F.explicit_solve()
        for R in ranks:
            R.update(q)
        F.update()
    assert_identical(F, ranks, SOLVE_OUT + UPDATE_OUT, "two steps")


def test_reference_bamg_tables_feed_the_reference_bodies():
    """End-to-end reference chain: tables from the reference's BamgConvertMeshx (oracle/ref_bamg) + reference bodies."""
    from oracle import ref_bamg
    if not ref_bamg.available():
        pytest.skip("oracle/_ref/libref_bamg.so not built")
    c = cases.make_case("10km_stable", nranks=2, dyn="bbm", nx=40, open_east=True)
    ranks = ob.make_ranks(c)
    for R, f in zip(ranks, c.local):
        R.set("M_random_number", f["M_random_number"])
    q = ob.orc_params(c.params)
    F = make_ref(c, ranks, q, bamg_from_reference=True)
    orc.explicit_solve(ranks, q)
    F.explicit_solve()
    assert_identical(F, ranks, SOLVE_OUT, "explicitSolve with the reference's bamg tables")


def test_extractor_cut_the_expected_definitions():
    import os
    idx = os.path.join(os.path.dirname(ref_fe._LIB), "ref_fe_bodies.inc.index")
    if not os.path.exists(idx):
        pytest.skip("index only exists where the reference is present")
    names = [l.split()[0] for l in open(idx)]
    for n in ("FiniteElement::explicitSolve", "FiniteElement::update", "FiniteElement::updateSigmaDamage",
              "FiniteElement::updateSigmaVP", "FiniteElement::updateGhosts", "FiniteElement::shapeCoeff",
              "FiniteElement::checkRegridding", "FiniteElement::updateIceDiagnostics", "GmshMesh::vertices"):
        assert n in names, n
