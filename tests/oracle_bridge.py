"""Glue between nextsim_b200.cases (host arrays) and the CPU oracle.  Test infrastructure only."""
import numpy as np

from oracle import oracle as orc

ORC_FIELDS = ("M_damage", "M_conc", "M_thick", "M_snow_thick", "M_conc_young", "M_h_young", "M_hs_young",
              "M_thick_myi", "M_conc_myi", "M_ridge_ratio", "M_element_depth", "M_drag_ui", "M_drag_ui_young",
              "M_time_relaxation_damage", "M_Cohesion", "M_VT", "M_UM", "M_UT", "M_wind", "M_ocean", "M_ssh")


def orc_params(p):
    """NsxDynParams -> OrcParams (same option values, the oracle keeps its own struct)."""
    q = orc.OrcParams()
    for name, _ in orc.OrcParams._fields_:
        if name == "pad_":
            continue
        setattr(q, name, getattr(p, name))
    return q


def make_ranks(c, fast=False):
    """Oracle ranks for a Case: the oracle derives its OWN local numbering, tables and masks from the global
    mesh + tags (independent of nextsim_b200.partition), then receives the scattered fields."""
    gm = c.gm
    if c.nranks == 1:
        ranks = [orc.single_rank_mesh(gm.x, gm.y, gm.tri, fast=fast)]
    else:
        ranks = orc.nodal_grid(c.nranks, gm.x, gm.y, gm.tri, c.elem_part, c.ghost_ptr, c.ghost_val, fast=fast)
    for R, lm, f in zip(ranks, c.lms, c.local):
        R.bamg_tables()
        R.bc_marked_nodes(gm.dirichlet_flags_root, gm.neumann_flags_root)
        R.set("lat", lm.lat)
        for k in ORC_FIELDS:
            R.set(k, f[k])
        for i in range(3):
            R.set("M_sigma%d" % i, f["M_sigma"][i])
    return ranks


def get_state(R, names):
    out = {}
    for k in names:
        if k == "M_sigma":
            out[k] = [R.get("M_sigma%d" % i) for i in range(3)]
        else:
            out[k] = R.get(k)
    return out


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    den = np.linalg.norm(b)
    num = np.linalg.norm(a - b)
    if den == 0.0:
        return 0.0 if num == 0.0 else np.inf
    return num / den
