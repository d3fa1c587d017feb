"""A C++ host process (tests/cpp/host_step.cpp: nextsim.cfg -> options, host mesh library, FiniteElementGPU shim,
several model steps with regrid check and diagnostics) gives bit for bit what the ctypes harness gives on the same
case, and stays within 1e-9 of the oracle."""
import subprocess

import numpy as np
import pytest

from nextsim_b200 import cases
import host_step_common as hs
import oracle_bridge as ob
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,nx,dyn,nsteps", [("toy", None, "bbm", 1), ("10km_stable", 48, "mevp", 2)])
def test_cpp_host_process_matches_harness_and_oracle(tmp_path, name, nx, dyn, nsteps):
    exe = hs.build_exe(tmp_path)
    c = cases.make_case(name, nranks=1, dyn=dyn, nx=nx, open_east=True)
    hs.write_case(tmp_path / "case.bin", c)
    hs.write_cfg(tmp_path / "nextsim.cfg", c, dyn)
    r = subprocess.run([str(exe), str(tmp_path / "case.bin"), str(tmp_path / "nextsim.cfg"), str(tmp_path / "out.bin"),
                        str(nsteps)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    got = hs.read_out(tmp_path / "out.bin", c.gm.nn, c.gm.ne)

    (s,) = cases.make_solvers(c)
    for _ in range(nsteps):
        s.explicit_solve()
        s.update()
        rg = s.check_regridding(10.0)
    s.update_ice_diagnostics()
    ref = s.download(*hs.OUT_FIELDS)
    s.close()
    for k in hs.OUT_FIELDS:
        pairs = zip(got[k], ref[k]) if k in ("M_sigma", "D_sigma") else [(got[k], ref[k])]
        for a, b in pairs:
            assert np.array_equal(a, b), "%s differs between the C++ host process and the ctypes harness" % k
    assert got["min_angle"] == rg.min_angle and bool(got["regrid"]) == bool(rg.regrid)

    (R,) = ob.make_ranks(c)
    q = ob.orc_params(c.params)
    for _ in range(nsteps):
        orc.explicit_solve([R], q)
        R.update(q)
    for k in ("M_VT", "M_damage", "M_conc", "M_thick"):
        assert ob.rel_l2(got[k], R.get(k)) <= 1e-9, k
    for i in range(3):
        assert ob.rel_l2(got["M_sigma"][i], R.get("M_sigma%d" % i)) <= 1e-9
