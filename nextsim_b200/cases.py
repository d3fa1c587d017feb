"""Benchmark / parity cases: BASELINE.md section 4 configs turned into host arrays and solver handles.

A Case owns what a neXtSIM host process would own before calling the path: the (partitioned) mesh,
the FiniteElement member vectors in each rank's local numbering, and the options.  Nothing here
touches the oracle; tests and bench.py build the oracle side separately from the same arrays.
"""
import copy
import math

import numpy as np

from . import capi, partition as pt, synthetic as syn

ELEM_FIELDS = ("M_damage", "M_conc", "M_thick", "M_snow_thick", "M_conc_young", "M_h_young", "M_hs_young",
               "M_thick_myi", "M_conc_myi", "M_ridge_ratio", "M_element_depth", "M_drag_ui", "M_drag_ui_young",
               "M_time_relaxation_damage", "M_random_number")
NODAL2_FIELDS = ("M_VT", "M_UM", "M_UT", "M_wind", "M_ocean")
NODAL1_FIELDS = ("M_ssh",)

# name -> (mesh size, state kind, dt, extra option overrides)   BASELINE.md section 4
CONFIGS = {
    "toy": dict(mesh="toy", kind="toy", dt=300.0, C_lab=1.5e6, alea_factor=0.33, use_coriolis=False),
    "10km": dict(mesh="10km", kind="large", dt=200.0, alea_factor=0.33),
    "10km_stable": dict(mesh="10km", kind="stable", dt=200.0, alea_factor=0.33),
    "3km": dict(mesh="3km", kind="large", dt=200.0, alea_factor=0.33),
    "3km_stable": dict(mesh="3km", kind="stable", dt=200.0, alea_factor=0.33),
    # 1 km mesh: BASELINE.md section 4 row 5 quotes dt = 200 s / 120 sub-cycles, but the explicit scheme is unstable there in
    # the reference itself (elastic wave speed sqrt(E/rho_i) ~ 806 m/s x dte 1.67 s = 1.3 km per sub-cycle > the 1 km cells:
    # the oracle's velocities reach 85 m/s and NaN within 12 sub-cycles).  dt = 50 s (dte 0.42 s, CFL ~ 0.34) keeps the same
    # work per step (120 sub-cycles) and a state that stays physical.
    "1km": dict(mesh="1km", kind="large", dt=50.0, alea_factor=0.33),
    "1km_stable": dict(mesh="1km", kind="stable", dt=50.0, alea_factor=0.33),
}


class Case:
    pass


def make_params(cfg, dyn, gm, substeps=120, defaults=None):
    """defaults: an NsxDynParams already holding the option defaults (the CPU arm of bench.py fills one from the oracle so
    that it never loads libnsx.so); None = nsx_params_defaults."""
    p = defaults if defaults is not None else capi.default_params()
    p.dynamics_type = capi.DYN[dyn]
    p.substeps = substeps
    p.dtime_step = cfg["dt"]
    p.C_lab = cfg.get("C_lab", p.C_lab)
    p.alea_factor = cfg.get("alea_factor", p.alea_factor)
    if not cfg.get("use_coriolis", True):
        p.use_coriolis = 0
        p.ocean_turning_angle_rad = 0.0                  # FE.cpp:1167-1172
    scale_coef = math.sqrt(0.1 / gm.resolution)          # FE.cpp:6995-6998
    p.compr_strength = p.compr_strength * scale_coef
    C_fix = p.C_lab * scale_coef
    C_alea = p.alea_factor * C_fix
    return p, C_fix, C_alea


def make_case(name="toy", nranks=1, dyn="bbm", open_east=False, substeps=120, nx=None, young=True, seed=syn.SEED,
              only_rank=None, defaults=None):
    cfg = dict(CONFIGS[name])
    mnx, h = syn.SIZES[cfg["mesh"]]
    if nx is not None:
        mnx = nx
    gm = syn.make_mesh(mnx, h, seed=seed, open_east=open_east)
    st = syn.make_state(gm, kind=cfg["kind"], seed=seed, young=young)
    c = Case()
    c.name, c.dyn, c.nranks = name, dyn, nranks
    c.gm, c.state = gm, st
    c.params, c.C_fix, c.C_alea = make_params(cfg, dyn, gm, substeps, defaults)
    if nranks > 1:
        c.elem_part = pt.partition_elements(gm.x, gm.y, gm.tri, nranks)
        c.ghost_ptr, c.ghost_val = pt.ghost_tags(gm.tri, c.elem_part, nranks)
        if only_rank is None:
            c.lms = pt.nodal_grid(nranks, gm.x, gm.y, gm.tri, c.elem_part, c.ghost_ptr, c.ghost_val)
        else:
            # one process per GPU: this rank's mesh, halo lists, masks and bamg tables from the host library
            # (nsx_partmesh_*); the other ranks' entries stay None
            pm = capi.PartMesh.build(gm.x, gm.y, gm.tri, only_rank, nranks, c.elem_part, c.ghost_ptr, c.ghost_val)
            pm.bc_marked_nodes(gm.dirichlet_flags_root, gm.neumann_flags_root)
            lm = pm.to_local_mesh()
            pm.close()
            lm.lat = pt.scatter_nodal1(lm, gm.lat)
            c.lms = [lm if r == only_rank else None for r in range(nranks)]
            c.local = [local_fields(c, l) if l is not None else None for l in c.lms]
            return c
    else:
        c.elem_part = np.zeros(gm.ne, np.int32)
        c.ghost_ptr = np.zeros(gm.ne + 1, np.int32)
        c.ghost_val = np.zeros(0, np.int32)
        c.lms = pt.nodal_grid(1, gm.x, gm.y, gm.tri)
    for lm in c.lms:
        if only_rank not in (None, lm.rank):
            continue                                  # another process owns that rank: skip its tables
        pt.bc_marked_nodes(lm, gm.dirichlet_flags_root, gm.neumann_flags_root)
        lm.nodal_element_connectivity, lm.nodal_connectivity = pt.bamg_tables(lm.indices, lm.num_nodes)
        lm.lat = pt.scatter_nodal1(lm, gm.lat)
    c.local = [local_fields(c, lm) if only_rank in (None, lm.rank) else None for lm in c.lms]
    return c


def local_fields(c, lm):
    """Scatter the global state into one rank's local numbering (host-side, like the reference's scatterv)."""
    g = c.state
    nn = c.gm.nn
    f = {}
    for k in ELEM_FIELDS:
        f[k] = pt.scatter_elem(lm, g[k])
    f["M_sigma"] = [pt.scatter_elem(lm, g["M_sigma"][i]) for i in range(3)]
    for k in NODAL2_FIELDS:
        f[k] = pt.scatter_nodal2(lm, g[k], nn)
    for k in NODAL1_FIELDS:
        f[k] = pt.scatter_nodal1(lm, g[k])
    f["M_Cohesion"] = c.C_fix + c.C_alea * f["M_random_number"]      # FE.cpp:3909-3914
    return f


UPLOAD_KEYS = ("M_VT", "M_UM", "M_UT", "M_wind", "M_ocean", "M_ssh", "M_sigma", "M_damage", "M_conc", "M_thick",
               "M_snow_thick", "M_conc_young", "M_h_young", "M_hs_young", "M_thick_myi", "M_conc_myi",
               "M_ridge_ratio", "M_element_depth", "M_drag_ui", "M_drag_ui_young", "M_Cohesion",
               "M_time_relaxation_damage")
STATE_OUT = ("M_VT", "M_UM", "M_UT", "M_sigma", "M_damage", "D_tau_a", "D_tau_w", "M_surface", "M_delta_x")
UPDATE_OUT = ("M_conc", "M_thick", "M_snow_thick", "M_thick_myi", "M_conc_myi", "M_ridge_ratio", "M_conc_young",
              "M_h_young", "M_hs_young", "M_sigma", "M_surface", "D_del_ci_ridge_myi")


def make_solvers(c, device=0, **options):
    """One nsx handle per rank, all on `device`, halos wired in-process.  Ranks sharing the GPU split its SMs: the
    resident path runs one persistent launch per rank and they synchronise through flags, so they must be co-resident."""
    if c.nranks > 1 and "max_sms" not in options:
        sms = capi.lib().nsx_device_sm_count(int(device))
        if sms > 0:
            options["max_sms"] = max(1, sms // c.nranks)
    solvers = [capi.Solver(lm, device, **options) for lm in c.lms]
    for s, f in zip(solvers, c.local):
        s.set_params(c.params)
        s.upload(**{k: f[k] for k in UPLOAD_KEYS})
    if c.nranks > 1:
        capi.connect_local_group(solvers)
    return solvers


def gather_global(c, per_rank, key):
    """Owned entries of every rank back into file numbering."""
    nn, ne = c.gm.nn, c.gm.ne
    if key == "M_sigma":
        return [pt.gather_elem(c.lms, [d[key][i] for d in per_rank], ne) for i in range(3)]
    v0 = per_rank[0][key]
    if v0.size == 2 * c.lms[0].num_nodes:
        return pt.gather_nodal2(c.lms, [d[key] for d in per_rank], nn)
    return pt.gather_elem(c.lms, [d[key] for d in per_rank], ne)
