"""The C-ABI library loads, exports every symbol include/nsx.h declares, and its structs match the binding.
No compute calls here (no GPU in this suite); nsx_create must fail loudly, not fall back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from nextsim_b200 import capi, cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "nsx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(nsx_[a-z_0-9]+)\s*\(", src)
    return sorted(set(names))


def test_header_symbols_exported():
    L = capi.lib()
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "libnsx.so does not export %s" % n
    assert set(capi.EXPORTS) == set(names)
    assert L.nsx_version() == 1


def test_struct_layouts_match_ctypes():
    L = capi.lib()
    out = (C.c_int * 9)()
    assert L.nsx_abi_sizes(out, 9) == 9
    expect = [C.sizeof(capi.NsxDynParams), C.sizeof(capi.NsxMesh), C.sizeof(capi.NsxHalo),
              C.sizeof(capi.NsxFields), C.sizeof(capi.NsxCheck), C.sizeof(capi.NsxTiming), C.sizeof(capi.NsxRegrid),
              C.sizeof(capi.NsxCreateOptions), C.sizeof(capi.NsxThermoParams)]
    assert list(out) == expect


def test_thermo_defaults_follow_options_cpp():
    """model/options.cpp:272-449, 543-548; the oracle's mirror of the struct holds the same values"""
    from oracle import thermo as oth
    p = capi.thermo_default_params()
    q = oth.default_params()
    assert C.sizeof(p) == C.sizeof(q)
    for (n, _) in p._fields_:
        assert getattr(p, n) == getattr(q, n), n
    assert (p.thermo_type, p.newice_type, p.melt_type, p.alb_scheme, p.flooding, p.ice_cat_young) == (1, 4, 2, 3, 1, 1)
    assert (p.alb_ice, p.alb_sn, p.alb_ponds, p.I_0, p.ocean_albedo) == (0.538, 0.8256, 0.30, 0.30, 0.07)
    assert (p.hnull, p.PhiF, p.PhiM, p.h_young_min, p.h_young_max, p.ks) == (0.25, 4.0, 0.5, 0.05, 0.5, 0.3096)
    assert (p.drag_ocean_t, p.drag_ocean_q, p.freezingpoint_mu, p.Csens_io) == (0.83e-3, 1.5e-3, 0.055, 1e-3)
    assert (p.reset_month, p.reset_day, p.freeze_days_threshold, p.reset_by_date, p.equal_melting) == (9, 15, 3.0, 0, 1)
    assert p.time_relaxation_damage == 25 * 86400.0 and p.deltaT_relaxation_damage == 20.0 and p.dtime_step == 200.0


def test_defaults_follow_options_cpp():
    p = capi.default_params()
    assert p.substeps == 120 and p.dtime_step == 200.0
    assert p.dynamics_type == 0 and p.basal_stress_type == 1 and p.ice_cat_type == 1
    assert p.young == 5.9605e8 and p.nu0 == pytest.approx(1 / 3) and p.tan_phi == 0.7
    assert p.compaction_param == -20.0 and p.exponent_relaxation_sigma == 5.0
    assert p.ocean_turning_angle_rad == pytest.approx(np.deg2rad(25.0))
    assert (p.evp_e, p.evp_Pstar, p.evp_C, p.evp_dmin) == (2.0, 27.5e3, 20.0, 1e-9)
    assert (p.mevp_alpha, p.mevp_beta) == (500.0, 500.0)


TOY_CFG = """
[setup]
ice-type=constant_partial
[mesh]
filename=square_with_point.msh
[simul]
timestep=300
duration=1
[debugging]
log-level=debug#info
[thermo]
use_thermo_forcing=false
newice_type=4
[dynamics]
use_coriolis=false
alea_factor=.33
#compression_factor=0
C_lab=1.5e6
"""


def test_cfg_reader(tmp_path):
    f = tmp_path / "toy.cfg"
    f.write_text(TOY_CFG)
    p = capi.params_from_cfg(f)
    assert p.dtime_step == 300.0 and p.C_lab == 1.5e6 and p.alea_factor == 0.33
    assert p.use_coriolis == 0 and p.ocean_turning_angle_rad == 0.0       # FE.cpp:1167-1172
    assert p.compression_factor == 10e3                                    # commented line ignored
    assert p.ice_cat_type == 1
    f.write_text("[setup]\ndynamics-type=mevp\n[dynamics]\nmevp.alpha=300\nsubsteps=200\n[thermo]\nnewice_type=1\n")
    p = capi.params_from_cfg(f)
    assert p.dynamics_type == 4 and p.mevp_alpha == 300.0 and p.substeps == 200 and p.ice_cat_type == 0
    f.write_text("[dynamics]\nnot_an_option=1\n")
    with pytest.raises(RuntimeError, match="unrecognised option"):
        capi.params_from_cfg(f)
    f.write_text("[setup]\ndynamics-type=free_drift\n")
    with pytest.raises(RuntimeError, match="outside the accelerated path"):
        capi.params_from_cfg(f)
    with pytest.raises(RuntimeError, match="cannot open"):
        capi.params_from_cfg(tmp_path / "missing.cfg")
    # po::value<int> options reject non-integers like the reference does (options.cpp:43, 363)
    for bad in ("[simul]\ntimestep=200.5\n", "[dynamics]\nsubsteps=1e2\n"):
        f.write_text(bad)
        with pytest.raises(RuntimeError, match="not an integer"):
            capi.params_from_cfg(f)


def test_create_options_defaults_and_plan_info():
    """NsxCreateOptions replaces the environment knobs of round 1; the resident plan can be inspected without a GPU."""
    o = capi.create_options()
    assert (o.path, o.use_graph, o.overlap, o.ow_skip, o.max_sms, o.tile_nodes) == (0, 1, 1, 1, 0, 0)
    assert capi.create_options(path="resident", max_sms=37).path == 3
    from nextsim_b200 import cases
    c = cases.make_case("10km_stable", nranks=2, nx=96, open_east=True)
    for lm in c.lms:
        d = capi.resident_plan_info(lm, sms=148)
        assert d["fits"] == 1 and d["ntiles"] <= 148 and d["tile_nodes"] <= 768 and d["max_slots"] <= 2 * 768
        assert d["smem_bytes"] <= d["smem_limit"] and d["own_slots"] == lm.num_elements
        assert 0 < d["export_nodes"] < lm.local_ndof and d["early_own_slots"] < d["own_slots"]
    big = cases.make_case("10km_stable", nranks=1, nx=400)             # 320 000 elements: too many slots per SM
    assert capi.resident_plan_info(big.lms[0])["fits"] == 0


REF_CFG = "/root/reference/config-files"


@pytest.mark.skipif(not os.path.isdir(REF_CFG), reason="the reference tree is not present on this box")
def test_reference_config_files_parse():
    """The three nextsim.cfg files the reference ships (config-files/) go through nsx_params_from_cfg unchanged: every
    [dynamics] key they use is known, keys outside the path are ignored, and the values land in NsxDynParams."""
    p = capi.params_from_cfg(os.path.join(REF_CFG, "nextsim.toy.cfg"))
    assert (p.dtime_step, p.C_lab, p.alea_factor, p.use_coriolis, p.ocean_turning_angle_rad) == (300.0, 1.5e6, 0.33, 0, 0.0)
    assert p.dynamics_type == 0 and p.ice_cat_type == 1 and p.substeps == 120
    p = capi.params_from_cfg(os.path.join(REF_CFG, "nextsim.cfg"))
    assert p.dtime_step == 200.0 and p.use_coriolis == 1 and p.substeps == 120
    p = capi.params_from_cfg(os.path.join(REF_CFG, "cpl_run_opa4.cfg"))
    assert (p.compression_factor, p.C_lab, p.substeps, p.time_relaxation_damage_days) == (3e3, 2e6, 75, 15.0)
    assert (p.quad_drag_coef_water, p.basal_k1, p.min_h) == (0.0067, 5.0, 0.1)


def test_every_reference_dynamics_option_is_known(tmp_path):
    """All 37 options of the [dynamics] section (model/options.cpp:309-376) are either consumed or explicitly ignored."""
    keys = ["alea_factor", "young", "C_lab", "nu0", "tan_phi", "compr_strength", "compaction_param", "min_h", "min_c",
            "time_relaxation_damage", "use_temperature_dependent_healing", "deltaT_relaxation_damage",
            "undamaged_time_relaxation_sigma", "exponent_relaxation_sigma", "ERA5_quad_drag_coef_air",
            "ECMWF_quad_drag_coef_air", "ASR_quad_drag_coef_air", "CFSR_quad_drag_coef_air", "lin_drag_coef_air",
            "quad_drag_coef_water", "lin_drag_coef_water", "use_coriolis", "oceanic_turning_angle", "Lemieux_basal_k1",
            "Lemieux_basal_k2", "Lemieux_basal_Cb", "Lemieux_basal_u_0", "Lemieux_basal_u_crit",
            "exponent_compression_factor", "compression_factor", "substeps", "evp.e", "evp.Pstar", "evp.C", "evp.dmin",
            "mevp.alpha", "mevp.beta"]
    assert len(keys) == 37
    f = tmp_path / "all.cfg"
    f.write_text("[dynamics]\n" + "".join("%s=%s\n" % (k, "true" if k.startswith("use_") else "1") for k in keys))
    p = capi.params_from_cfg(f)
    assert p.substeps == 1 and p.evp_dmin == 1.0 and p.basal_u0 == 1.0


def test_create_fails_loudly_without_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    c = cases.make_case("toy")
    with pytest.raises(RuntimeError, match="nsx_create"):
        capi.Solver(c.lms[0], 0)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "nextsim_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, fn)).read()
                for needle in ("import oracle", "from oracle", "liboracle", "orc_", "oracle_bridge"):
                    assert needle not in txt, "%s references the oracle (%s)" % (fn, needle)
