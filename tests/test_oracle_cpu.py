"""CPU tests of the oracle: golden fixtures, known answers and invariants (no GPU needed).

PARITY UNPINNED: the reference has no golden vectors for this path.  What can be pinned
independently is pinned here: the C++11 standard's known answer for minstd_rand (the generator behind the
cohesion noise, FE.cpp:11464-11468), closed forms for single-element states worked out by hand from the
reference formulas, and the checkFieldsFast invariants (FE.cpp:14539-14629).
"""
import glob
import os

import numpy as np
import pytest

from nextsim_b200 import cases, synthetic as syn
import oracle_bridge as ob
from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _golden_cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg


@pytest.mark.parametrize("key", sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz"))))
def test_oracle_reproduces_golden(key):
    mg = _golden_cases()
    out = mg.run_next(mg.NEXT_CASES[key]) if key in mg.NEXT_CASES else mg.run(mg.CASES[key])
    ref = np.load(os.path.join(GOLD, key + ".npz"))
    assert set(out) == set(ref.files)
    for k in ref.files:
        e = ob.rel_l2(out[k], ref[k])
        assert e <= 1e-13, "%s/%s rel-L2 %.3e" % (key, k, e)


def test_partitioned_oracle_equals_single_rank_golden():
    """3 MPI-style ranks with ghost elements + updateGhosts give the single-rank answer (sum order differs)."""
    a = np.load(os.path.join(GOLD, "stable24_bbm_step_open.npz"))
    b = np.load(os.path.join(GOLD, "stable24_bbm_step_3ranks.npz"))
    for k in a.files:
        assert ob.rel_l2(b[k], a[k]) <= 1e-12, k


def test_minstd_known_answer():
    """ISO C++ [rand.predef]: the 10000th value of a default-constructed minstd_rand is 399268537."""
    u = syn.minstd_uniform01(10000)
    assert round(u[-1] * 2147483646.0 + 1.0) == 399268537
    assert round(u[0] * 2147483646.0 + 1.0) == 48271
    assert np.array_equal(syn.minstd_uniform01_fast(10000), u)
    # the oracle's own generator (orc_calc_cohesion) agrees
    R = orc.single_rank_mesh(np.array([0., 1., 0.]), np.array([0., 0., 1.]), np.array([[1, 2, 3]], np.int32))
    R.calc_cohesion(1, 10.0, 2.0)
    assert R.get("M_Cohesion")[0] == 10.0 + 2.0 * u[0]


def one_triangle(side=10000.0):
    x = np.array([0.0, side + 0.7, 0.3])
    y = np.array([0.0, 0.2, side + 0.9])
    tri = np.array([[1, 2, 3]], np.int32)
    return x, y, tri


def one_element_rank(VT, conc=1.0, thick=2.0, damage=0.0, sigma=(0., 0., 0.), dyn="bbm"):
    x, y, tri = one_triangle()
    R = orc.single_rank_mesh(x, y, tri)
    R.bamg_tables()
    R.bc_marked_nodes(np.zeros(0, np.int32), np.zeros(0, np.int32))
    one = np.ones(1)
    for k, v in dict(M_conc=conc, M_thick=thick, M_snow_thick=0.0, M_conc_young=0.0, M_h_young=0.0, M_hs_young=0.0,
                     M_thick_myi=0.0, M_conc_myi=0.0, M_ridge_ratio=0.0, M_element_depth=500.0, M_drag_ui=0.002,
                     M_drag_ui_young=0.002, M_time_relaxation_damage=25 * 86400.0, M_Cohesion=5000.0,
                     M_damage=damage).items():
        R.set(k, one * v)
    for i in range(3):
        R.set("M_sigma%d" % i, one * sigma[i])
    R.set("M_VT", np.asarray(VT, float))
    for k in ("M_UM", "M_UT", "M_wind", "M_ocean"):
        R.set(k, np.zeros(6))
    R.set("M_ssh", np.zeros(3))
    R.set("lat", np.full(3, 80.0))
    return R, (x, y)


def test_delta_x_is_integer_truncated():
    """Quirk Q1 (FE.cpp:10239): std::accumulate with an int initial value, then integer division by 3."""
    from nextsim_b200 import capi
    p = capi.default_params()
    p.substeps = 1
    R, (x, y) = one_element_rank(np.zeros(6))
    orc.explicit_solve([R], ob.orc_params(p))
    s = [np.hypot(x[1] - x[0], y[1] - y[0]), np.hypot(x[2] - x[1], y[2] - y[1]), np.hypot(x[2] - x[0], y[2] - y[0])]
    acc = 0
    for v in s:
        acc = int(acc + v)
    assert R.get("M_delta_x")[0] == float(acc // 3)
    assert R.get("M_delta_x")[0] != pytest.approx(np.mean(s), abs=1e-3)
    jac = (x[1] - x[0]) * (y[2] - y[0]) - (x[2] - x[0]) * (y[1] - y[0])
    assert R.get("M_surface")[0] == 0.5 * abs(jac)
    # shape function gradients reproduce a linear field exactly: sum_i dN_i/dx * x_i = 1, sum dN_i/dx * y_i = 0
    sc = R.get("M_shape_coeff")
    assert np.dot(sc[:3], x) == pytest.approx(1.0, abs=1e-12)
    assert np.dot(sc[:3], y) == pytest.approx(0.0, abs=1e-12)
    assert np.dot(sc[3:], y) == pytest.approx(1.0, abs=1e-12)


def test_bbm_single_element_closed_form():
    """One undamaged element under uniform convergence: sigma follows FE.cpp:4184-4210 by hand."""
    from nextsim_b200 import capi
    p = capi.default_params()
    p.substeps = 1
    p.stop_after_substeps = 1
    p.skip_ow_smoother = 1
    x, y, _ = one_triangle()
    rate = -1e-7                                  # du/dx = dv/dy = rate
    VT = np.concatenate([rate * x, rate * y])
    R, _ = one_element_rank(VT)
    orc.explicit_solve([R], ob.orc_params(p))
    dt = p.dtime_step
    nu = p.nu0
    E = p.young                                    # d=0, conc=1 -> expC = 1
    s_el = dt * E / (1 - nu * nu) * (rate + nu * rate)
    tv = p.undamaged_time_relaxation_sigma
    # sigma_n = 0 at the start -> tildeP = 0
    mult = min(1 - 1e-12, tv / (tv + dt))
    s = s_el * mult
    assert R.get("M_sigma0")[0] == pytest.approx(s, rel=1e-12)
    assert R.get("M_sigma1")[0] == pytest.approx(s, rel=1e-12)
    assert abs(R.get("M_sigma2")[0]) <= 1e-9 * abs(s)
    # pure isotropic compression inside the envelope: no damage, only healing (clamped at 0)
    assert R.get("M_damage")[0] == 0.0


def test_no_ice_element_is_reset():
    """conc <= 0.1 (hard-coded, quirk Q4): sigma and damage are zeroed (FE.cpp:4151-4159)."""
    from nextsim_b200 import capi
    p = capi.default_params()
    p.substeps = 1
    R, _ = one_element_rank(np.zeros(6), conc=0.1, damage=0.4, sigma=(5., 6., 7.))
    orc.explicit_solve([R], ob.orc_params(p))
    assert R.get("M_damage")[0] == 0.0
    assert [R.get("M_sigma%d" % i)[0] for i in range(3)] == [0.0, 0.0, 0.0]


@pytest.mark.parametrize("dyn", ["bbm", "mevp", "evp"])
def test_checkfieldsfast_invariants(dyn):
    """0<=d<=1, 0<=c<=1, |u|<=5 m/s after a step (the reference's only regression net, FE.cpp:14539-14629)."""
    c = cases.make_case("10km_stable", nranks=1, dyn=dyn, nx=32, open_east=True)
    R = ob.make_ranks(c)[0]
    q = ob.orc_params(c.params)
    orc.explicit_solve([R], q)
    R.update(q)
    d, cc, vt = R.get("M_damage"), R.get("M_conc"), R.get("M_VT")
    assert np.isfinite(vt).all() and np.abs(vt).max() < 5.0
    assert d.min() >= 0.0 and d.max() <= 1.0
    assert cc.min() >= 0.0 and cc.max() <= 1.0
    # Dirichlet nodes never move
    dn = R.get("M_dirichlet_flags")
    nn = c.gm.nn
    assert np.all(vt[dn] == 0.0) and np.all(vt[dn + nn] == 0.0)
    # Neumann (open boundary) nodes: mesh displacement restored, total displacement not (FE.cpp:10549-10550)
    ne_ = R.get("M_neumann_flags")
    assert ne_.size > 0
    assert np.all(R.get("M_UM")[ne_] == 0.0)
    assert np.any(R.get("M_UT")[ne_] != 0.0)


def test_update_conserves_ice_volume():
    """Lagrangian update(): thick*area is conserved where no ridging/capping applies (FE.cpp:3970-3975)."""
    c = cases.make_case("10km_stable", nranks=1, dyn="bbm", nx=32, young=False)
    R = ob.make_ranks(c)[0]
    q = ob.orc_params(c.params)
    orc.explicit_solve([R], q)
    h0, a0 = R.get("M_thick"), R.get("M_surface")
    R.update(q)
    h1, a1 = R.get("M_thick"), R.get("M_surface")
    ice = h0 > 0
    assert np.allclose((h1 * a1)[ice], (h0 * a0)[ice], rtol=1e-12)
