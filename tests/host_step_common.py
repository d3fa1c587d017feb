"""Shared by the CPU and GPU tests of tests/cpp/host_step.cpp: builds the program, writes its case file and cfg."""
import os
import subprocess

import numpy as np

from nextsim_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IN_FIELDS = ("M_VT", "M_UM", "M_UT", "M_wind", "M_ocean", "M_ssh", "M_sigma", "M_damage", "M_conc", "M_thick",
             "M_snow_thick", "M_conc_young", "M_h_young", "M_hs_young", "M_thick_myi", "M_conc_myi", "M_ridge_ratio",
             "M_element_depth", "M_drag_ui", "M_drag_ui_young", "M_random_number", "M_time_relaxation_damage")
OUT_FIELDS = ("M_VT", "M_UM", "M_UT", "D_tau_a", "D_tau_w", "M_sigma", "M_damage", "M_conc", "M_thick", "M_snow_thick",
              "M_ridge_ratio", "M_surface", "D_conc", "D_thick", "D_snow_thick", "D_sigma", "D_divergence")


def build_exe(tmp_path):
    exe = tmp_path / "host_step"
    libdir = os.path.join(ROOT, "nextsim_b200")
    capi.lib()
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", ROOT, os.path.join(ROOT, "tests", "cpp", "host_step.cpp"),
                           "-o", str(exe), "-L", libdir, "-lnsx", "-Wl,-rpath," + libdir])
    return exe


def write_case(path, c):
    gm, f = c.gm, c.local[0]
    with open(path, "wb") as fh:
        np.array([gm.nn, gm.ne, gm.dirichlet_flags_root.size, gm.neumann_flags_root.size], np.int32).tofile(fh)
        np.array([gm.resolution], np.float64).tofile(fh)
        for a in (gm.x, gm.y, gm.lat):
            np.ascontiguousarray(a, np.float64).tofile(fh)
        np.ascontiguousarray(gm.tri, np.int32).tofile(fh)
        np.ascontiguousarray(gm.dirichlet_flags_root, np.int32).tofile(fh)
        np.ascontiguousarray(gm.neumann_flags_root, np.int32).tofile(fh)
        for k in IN_FIELDS:
            if k == "M_sigma":
                for i in range(3):
                    np.ascontiguousarray(f[k][i], np.float64).tofile(fh)
            else:
                np.ascontiguousarray(f[k], np.float64).tofile(fh)


def write_cfg(path, c, dyn):
    """The options of the case as a nextsim.cfg (what cases.make_params sets through the struct)."""
    p = c.params
    path.write_text("[setup]\ndynamics-type=%s\n[simul]\ntimestep=%d\n[dynamics]\nsubsteps=%d\nC_lab=%r\nalea_factor=%r\n"
                    "use_coriolis=%s\n" % (dyn, int(p.dtime_step), p.substeps, p.C_lab, p.alea_factor,
                                           "true" if p.use_coriolis else "false"))


def read_out(path, nn, ne):
    raw = np.fromfile(path, np.float64)
    out, o = {}, 0
    for k in OUT_FIELDS:
        if k in ("M_VT", "M_UM", "M_UT", "D_tau_a", "D_tau_w"):
            out[k] = raw[o:o + 2 * nn]; o += 2 * nn
        elif k == "M_sigma":
            out[k] = [raw[o + i * ne:o + (i + 1) * ne] for i in range(3)]; o += 3 * ne
        elif k == "D_sigma":
            out[k] = [raw[o + i * ne:o + (i + 1) * ne] for i in range(2)]; o += 2 * ne
        else:
            out[k] = raw[o:o + ne]; o += ne
    out["min_angle"], out["regrid"] = raw[o], raw[o + 1]
    assert o + 2 == raw.size
    return out
