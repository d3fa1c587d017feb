"""Oracle functions of the SURVEY.md section 8(f) rows against closed forms (no reference goldens exist for them):
minAngles / flip (FE.cpp:1758-1768, 1824-1839), updateIceDiagnostics (FE.cpp:7860-7900), ExternalData::get
(externaldata.cpp:366-436)."""
import numpy as np

from nextsim_b200 import cases
import oracle_bridge as ob
from oracle import oracle as orc


def rank_for(x, y, tri1):
    return orc.single_rank_mesh(np.asarray(x, float), np.asarray(y, float), np.asarray(tri1, np.int32))


def test_min_angle_closed_forms():
    # right isosceles (45 deg) and a 30-60-90 triangle (30 deg); M_UM = 0
    x = [0.0, 1.0, 0.0, 3.0, 3.0 + np.sqrt(3.0), 3.0]
    y = [0.0, 0.0, 1.0, 0.0, 0.0, 1.0]
    R = rank_for(x, y, [[1, 2, 3], [4, 5, 6]])
    R.set("M_UM", np.zeros(12))
    ang, jmin, jmax, flip, regrid = R.check_regridding(10.0)
    assert abs(ang - 30.0) < 1e-12
    assert jmin == 1.0 and abs(jmax - np.sqrt(3.0)) < 1e-15
    assert not flip and not regrid
    assert R.check_regridding(31.0)[4]                       # angle criterion alone


def test_flip_needs_displacement():
    x = [0.0, 1.0, 0.0, 1.0]
    y = [0.0, 0.0, 1.0, 1.0]
    R = rank_for(x, y, [[1, 2, 3], [2, 4, 3]])
    um = np.zeros(8)
    R.set("M_UM", um)
    assert not R.check_regridding(10.0)[3]
    um[3] = -1.5                                             # node 4 moves left past the diagonal: element 2 inverts
    um[7] = -1.5
    R.set("M_UM", um)
    ang, jmin, jmax, flip, regrid = R.check_regridding(10.0)
    assert jmin < 0 < jmax and flip and regrid


def test_ice_diagnostics_closed_forms():
    c = cases.make_case("toy")
    nn, ne = c.gm.nn, c.gm.ne
    (R,) = ob.make_ranks(c)
    a, b = 3e-6, -1.25e-6
    x, y = R.get("coordX"), R.get("coordY")
    R.set("M_UM", np.zeros(2 * nn))
    R.set("M_VT", np.concatenate([a * x + 0.1, b * y - 0.2]))       # div = a + b everywhere
    s = [np.full(ne, 3.0e3), np.full(ne, -1.0e3), np.full(ne, 1.5e3)]
    for i in range(3):
        R.set("M_sigma%d" % i, s[i])
    q = ob.orc_params(c.params)
    R.update_ice_diagnostics(q)
    np.testing.assert_allclose(R.get("D_divergence"), a + b, rtol=1e-9)
    np.testing.assert_allclose(R.get("D_sigma0"), 1.0e3)
    np.testing.assert_allclose(R.get("D_sigma1"), 2.5e3)            # hypot(2000, 1500)
    f = c.local[0]
    assert np.array_equal(R.get("D_conc"), f["M_conc"] + f["M_conc_young"])
    q.ice_cat_type = 0
    R.update_ice_diagnostics(q)
    assert np.array_equal(R.get("D_thick"), f["M_thick"])


def test_external_data_time_interpolation():
    rng = np.random.default_rng(3)
    d0, d1 = rng.normal(size=50), rng.normal(size=50)
    t0, t1 = 10.0, 10.25
    assert np.array_equal(orc.external_data_get_vector(d0, d1, True, t0, t0, t1, 1.0, 0.0), d0)
    assert np.array_equal(orc.external_data_get_vector(d0, d1, True, t1, t0, t1, 1.0, 0.0), d1)
    mid = orc.external_data_get_vector(d0, d1, True, 10.125, t0, t1, 2.0, 0.5)
    assert np.array_equal(mid, 2.0 * (0.5 * d0 + 0.5 * d1) + 0.5)
    assert np.array_equal(orc.external_data_get_vector(d0, d1, False, 10.2, t0, t1, 3.0, -1.0), 3.0 * d0 - 1.0)
