"""Host-side pieces that need no GPU: the C++ FiniteElementGPU shim compiles and links against the C ABI, and the
tile plan (internal numbering + tiles of the sub-cycle kernel) satisfies its invariants on awkward meshes."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from nextsim_b200 import capi, cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_shim_compiles_and_fails_loudly(tmp_path):
    exe = tmp_path / "shim_smoke"
    libdir = os.path.join(ROOT, "nextsim_b200")
    capi.lib()                                         # make sure libnsx.so is built
    subprocess.check_call(["g++", "-std=c++17", "-I", ROOT, os.path.join(ROOT, "tests", "cpp", "shim_smoke.cpp"),
                           "-o", str(exe), "-L", libdir, "-lnsx", "-Wl,-rpath," + libdir])
    cfg = tmp_path / "a.cfg"
    cfg.write_text("[dynamics]\nsubsteps=60\n")
    r = subprocess.run([str(exe), str(cfg)], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout
    assert "threw: nsx_create" in r.stdout or "handle created" in r.stdout


def mesh_struct(lm):
    keep = []
    M = capi.NsxMesh()
    M.num_nodes, M.local_ndof = lm.num_nodes, lm.local_ndof
    M.num_elements, M.local_nelements = lm.num_elements, lm.local_nelements
    for name, arr, conv in (("coord_x", lm.x, capi._f64), ("coord_y", lm.y, capi._f64),
                            ("indices", lm.indices.reshape(-1), capi._i32),
                            ("ghost_nodes", lm.ghostNodes.reshape(-1), capi._u8),
                            ("mask_dirichlet", lm.mask_dirichlet, capi._u8),
                            ("neumann_flags", lm.neumann_flags, capi._i32),
                            ("nodal_element_connectivity", lm.nodal_element_connectivity.reshape(-1), capi._f64),
                            ("nodal_connectivity", lm.nodal_connectivity.reshape(-1), capi._f64),
                            ("lat", lm.lat, capi._f64)):
        a, p = conv(arr)
        keep.append(a)
        setattr(M, name, p)
    M.n_neumann_flags = int(lm.neumann_flags.size)
    M.nec_width = int(lm.nodal_element_connectivity.shape[1])
    M.nc_width = int(lm.nodal_connectivity.shape[1])
    return M, keep


def plan_info(lm, target=208, wave=148):
    M, keep = mesh_struct(lm)
    out = (C.c_int * 10)()
    rc = capi.lib().nsx_plan_info(C.byref(M), target, wave, out, 10)
    assert rc == 0
    return dict(zip(("ntiles", "tile_nodes", "nslots", "max_local_nodes", "max_slots", "max_own_slots",
                     "max_halo_slots", "max_halo_nodes", "stage_bytes", "shrinks"), list(out)))


@pytest.mark.parametrize("name,nx,nranks", [("10km", None, 1), ("10km", 447, 2), ("10km", 200, 5), ("toy", None, 3)])
def test_tile_plan_is_compact(name, nx, nranks):
    """Tiles stay compact on single-rank and partitioned meshes: the redundantly recomputed slots are a modest
    fraction, no tile degenerates into a one-node-wide strip (halo ~ own), and the staged working set of the
    largest tile fits one shared-memory stage without shrinking the tiles over and over."""
    c = cases.make_case(name, nranks=nranks, nx=nx)
    for lm in c.lms:
        info = plan_info(lm)
        assert info["stage_bytes"] <= (227 * 1024 - 64) // 2
        assert info["nslots"] >= lm.num_elements
        assert info["nslots"] <= 1.45 * lm.num_elements + 64 * info["ntiles"]
        assert info["ntiles"] * info["tile_nodes"] >= lm.local_ndof
        assert info["shrinks"] <= 2
        if lm.num_elements > 50000:
            assert info["max_halo_slots"] < 0.6 * info["max_slots"]


def test_cpp_host_step_program_runs_up_to_the_device(tmp_path):
    """tests/cpp/host_step.cpp (a C++ host process: cfg -> options, host mesh library, shim) parses its inputs,
    builds the mesh and tables, and stops at nsx_create when there is no GPU (exit 3): no CPU fallback."""
    import host_step_common as hs
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present: covered by tests/test_gpu_host_step.py")
    except ImportError:
        pass
    exe = hs.build_exe(tmp_path)
    c = cases.make_case("toy")
    hs.write_case(tmp_path / "case.bin", c)
    hs.write_cfg(tmp_path / "nextsim.cfg", c, "bbm")
    r = subprocess.run([str(exe), str(tmp_path / "case.bin"), str(tmp_path / "nextsim.cfg"), str(tmp_path / "out.bin"), "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 3, r.stdout
    assert "case: 1089 nodes, 2048 elements" in r.stdout and "threw: nsx_create" in r.stdout
    assert not (tmp_path / "out.bin").exists()
