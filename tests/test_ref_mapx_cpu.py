"""Node latitudes: the host library's nsx_mapx_latlon against the REFERENCE'S OWN mapx (contrib/mapx, built unmodified
into oracle/_ref/libref_mapx.so), called like GmshMesh::lat() calls it (gmshmesh.cpp:1800-1824).  explicitSolve()
uses these latitudes for the Coriolis parameter and the sign of the ocean turning angle (FE.cpp:10351, 10398, 10497).
The projection files are written here with the parameters of the reference's mesh/NpsNextsim.mpp and mesh/NpsASR.mpp
(values only) plus a southern and a spherical variant."""
import numpy as np
import pytest

from nextsim_b200 import capi, synthetic as syn
from oracle import ref_mapx

pytestmark = pytest.mark.skipif(not ref_mapx.available(), reason="oracle/_ref/libref_mapx.so not built and no /root/reference")

MPP = """{name}
{lat0} 0.0 {lat1}	lat0 lon0 lat1
{rot}	    rotation
0.001	   	scale (km/map unit)
{lat0}	135.00	center lat lon
20.00	90.00	lat min max
-180.00	180.00	lon min max
 10.00 15.00	grid
00.00	00.00	label lat lon
1 0 0		cil bdy riv
{tail}"""
ELL = "6378.273	Earth equatorial radius (km)\n0.081816153	eccentricity\n"


@pytest.mark.parametrize("name,lat0,lat1,rot,tail", [
    ("Polar Stereographic Ellipsoid", 90.0, 60.0, -45.0, ELL),          # mesh/NpsNextsim.mpp
    ("Polar Stereographic Ellipsoid", 90.0, 60.0, -175.0, ELL),         # mesh/NpsASR.mpp
    ("Polar Stereographic Ellipsoid", -90.0, -70.0, 30.0, ELL),
    ("Polar Stereographic Ellipsoid", 90.0, 70.0, 0.0, ""),             # default radius / eccentricity
    ("Polar Stereographic", 90.0, 60.0, -45.0, "")])
def test_latlon_matches_reference_mapx(tmp_path, name, lat0, lat1, rot, tail):
    f = tmp_path / "proj.mpp"
    f.write_text(MPP.format(name=name, lat0=lat0, lat1=lat1, rot=rot, tail=tail))
    rng = np.random.default_rng(0)
    x = rng.uniform(-3.5e6, 3.5e6, 4000)             # map units are metres (scale 0.001 km per unit)
    y = rng.uniform(-3.5e6, 3.5e6, 4000)
    x[:3] = [0.0, 1.0, -2.5e6]
    y[:3] = [0.0, -1.0, 0.0]
    lat_r, lon_r = ref_mapx.latlon(f, x, y)
    lat, lon = capi.mapx_latlon(f, x, y)
    assert np.abs(lat - lat_r).max() <= 1e-12, np.abs(lat - lat_r).max()
    dl = np.abs(lon - lon_r)
    dl = np.minimum(dl, 360.0 - dl)
    assert dl[1:].max() <= 1e-11                      # the pole itself has no longitude
    assert np.sign(lat[3:]).min() == np.sign(lat0)


@pytest.mark.parametrize("fname", ["NpsNextsim.mpp", "NpsASR.mpp"])
def test_reference_mpp_files(fname):
    """The projection files the reference ships (mesh/*.mpp), when the reference tree is on this box."""
    import os
    path = os.path.join("/root/reference/mesh", fname)
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    rng = np.random.default_rng(2)
    x, y = rng.uniform(-3e6, 3e6, 5000), rng.uniform(-3e6, 3e6, 5000)
    lat_r, lon_r = ref_mapx.latlon(path, x, y)
    lat, lon = capi.mapx_latlon(path, x, y)
    assert np.abs(lat - lat_r).max() <= 1e-12 and np.abs(lon - lon_r).max() <= 1e-11


def test_partmesh_lat_from_mpp(tmp_path):
    f = tmp_path / "NpsNextsim.mpp"
    f.write_text(MPP.format(name="Polar Stereographic Ellipsoid", lat0=90.0, lat1=60.0, rot=-45.0, tail=ELL))
    m = syn.make_mesh(24, 10e3)
    pm = capi.PartMesh.build(m.x - 1.2e5, m.y + 3.0e5, m.tri)
    pm.lat_from_mpp(f)
    lm = pm.to_local_mesh()
    lat_r, _ = ref_mapx.latlon(f, lm.x, lm.y)
    assert np.abs(lm.lat - lat_r).max() <= 1e-12
    assert lm.lat.min() > 80.0


def test_unsupported_projection_is_an_error(tmp_path):
    f = tmp_path / "aea.mpp"
    f.write_text(MPP.format(name="Azimuthal_Equal_Area", lat0=90.0, lat1=60.0, rot=0.0, tail=""))
    with pytest.raises(RuntimeError, match="not supported"):
        capi.mapx_latlon(f, np.zeros(1), np.zeros(1))
    with pytest.raises(RuntimeError, match="cannot open"):
        capi.mapx_latlon(tmp_path / "missing.mpp", np.zeros(1), np.zeros(1))
