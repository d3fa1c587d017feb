"""ctypes binding of the CPU oracle (oracle/nextsim_oracle.cpp).

TEST INFRASTRUCTURE ONLY.  Importers allowed: tests/, __graft_entry__.smoke(), and the
cpu_baseline / ``--impl reference`` legs of bench.py.  The product package ``nextsim_b200``
never imports this module.  Physics PINNED bit for bit against the reference's own function bodies
(oracle/ref_fe, tests/test_ref_fe_cpu.py), bamg tables against the reference's own bamg library (oracle/ref_bamg);
nodalGrid() is restated from gmshmesh.cpp and unpinned: see the header of nextsim_oracle.cpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class OrcParams(C.Structure):
    _fields_ = [
        ("dynamics_type", C.c_int), ("basal_stress_type", C.c_int), ("ice_cat_type", C.c_int),
        ("substeps", C.c_int), ("equal_ridging", C.c_int), ("newice_type", C.c_int),
        ("use_young_ice_in_myi_reset", C.c_int), ("stop_after_substeps", C.c_int),
        ("skip_ow_smoother", C.c_int), ("pad_", C.c_int),
        ("dtime_step", C.c_double), ("ocean_turning_angle_rad", C.c_double),
        ("min_h", C.c_double), ("min_c", C.c_double),
        ("young", C.c_double), ("nu0", C.c_double), ("tan_phi", C.c_double),
        ("compr_strength", C.c_double), ("compaction_param", C.c_double),
        ("undamaged_time_relaxation_sigma", C.c_double), ("exponent_relaxation_sigma", C.c_double),
        ("compression_factor", C.c_double), ("exponent_compression_factor", C.c_double),
        ("quad_drag_coef_water", C.c_double),
        ("evp_e", C.c_double), ("evp_Pstar", C.c_double), ("evp_C", C.c_double),
        ("evp_dmin", C.c_double), ("mevp_alpha", C.c_double), ("mevp_beta", C.c_double),
        ("basal_k1", C.c_double), ("basal_k2", C.c_double), ("basal_Cb", C.c_double),
        ("basal_u0", C.c_double),
    ]


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc only, no reference sources copied)."""
    out = os.path.join(_HERE, "_build", "liboracle.so")
    src = os.path.join(_HERE, "nextsim_oracle.cpp")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return out


_libs = {}


def lib(fast=False):
    key = "fast" if fast else "ref"
    if key not in _libs:
        build()
        path = os.path.join(_HERE, "_build", "liboracle_fast.so" if fast else "liboracle.so")
        L = C.CDLL(path)
        L.orc_rank_create.restype = C.c_void_p
        L.orc_last_error.restype = C.c_char_p
        L.orc_rank_size_double.restype = C.c_long
        L.orc_rank_size_int.restype = C.c_long
        L.orc_rank_halo_size.restype = C.c_long
        L.orc_time_subcycles.restype = C.c_double
        L.orc_time_subcycles_mt.restype = C.c_double
        for f in ("orc_rank_destroy", "orc_rank_sizes", "orc_params_defaults"):
            getattr(L, f).restype = None
        _libs[key] = L
    return _libs[key]


def default_params():
    """(OrcParams, C_lab, alea_factor, time_relaxation_damage_days) with the defaults of model/options.cpp, from the
    oracle's own restatement (no product library involved)."""
    q = OrcParams()
    coh = (C.c_double * 3)()
    lib().orc_params_defaults(C.byref(q), coh)
    return q, coh[0], coh[1], coh[2]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Rank:
    """One MPI rank's worth of FiniteElement members, held by the oracle."""

    def __init__(self, fast=False):
        self.L = lib(fast)
        self.h = C.c_void_p(self.L.orc_rank_create())

    def __del__(self):
        try:
            self.L.orc_rank_destroy(self.h)
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError("oracle: " + self.L.orc_last_error().decode())

    def set(self, name, arr):
        arr = np.ascontiguousarray(arr)
        if arr.dtype.kind == "f":
            arr = arr.astype(np.float64, copy=False)
            self._chk(self.L.orc_rank_set_double(self.h, name.encode(), _dp(arr), C.c_long(arr.size)))
        else:
            arr = arr.astype(np.int32, copy=False)
            self._chk(self.L.orc_rank_set_int(self.h, name.encode(), _ip(arr), C.c_long(arr.size)))

    def get(self, name):
        n = self.L.orc_rank_size_double(self.h, name.encode())
        if n >= 0:
            out = np.empty(n, np.float64)
            self._chk(self.L.orc_rank_get_double(self.h, name.encode(), _dp(out), C.c_long(n)))
            return out
        n = self.L.orc_rank_size_int(self.h, name.encode())
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, np.int32)
        self._chk(self.L.orc_rank_get_int(self.h, name.encode(), _ip(out), C.c_long(n)))
        return out

    def sizes(self):
        out = (C.c_int * 6)()
        self.L.orc_rank_sizes(self.h, out)
        return dict(zip(("num_nodes", "local_ndof", "num_elements", "local_nelements",
                         "nec_width", "nc_width"), list(out)))

    def halo(self, which, proc):
        n = self.L.orc_rank_halo_size(self.h, which, proc)
        out = np.empty(n, np.int32)
        if n:
            self._chk(self.L.orc_rank_halo_get(self.h, which, proc, _ip(out)))
        return out

    def set_halo(self, nranks, which, proc, arr):
        arr = np.ascontiguousarray(arr, np.int32)
        self._chk(self.L.orc_rank_halo_set(self.h, nranks, which, proc, _ip(arr), C.c_long(arr.size)))

    def bamg_tables(self):
        self._chk(self.L.orc_bamg_tables(self.h))

    def bc_marked_nodes(self, dirichlet_flags_root, neumann_flags_root):
        flags = np.ascontiguousarray(np.concatenate([dirichlet_flags_root, neumann_flags_root]), np.int32)
        self._chk(self.L.orc_bc_marked_nodes(self.h, _ip(flags), len(dirichlet_flags_root),
                                             len(neumann_flags_root)))

    def calc_cohesion(self, global_num_elements, C_fix, C_alea):
        self._chk(self.L.orc_calc_cohesion(self.h, global_num_elements, C.c_double(C_fix), C.c_double(C_alea)))

    def update(self, params):
        self._chk(self.L.orc_update(self.h, C.byref(params)))

    def check_regridding(self, regrid_angle):
        """Local part of FiniteElement::checkRegridding(): (min_angle, min_jac, max_jac, flip, regrid_local)."""
        out = (C.c_double * 3)()
        flags = (C.c_int * 2)()
        self._chk(self.L.orc_check_regridding(self.h, C.c_double(regrid_angle), out, flags))
        return out[0], out[1], out[2], bool(flags[0]), bool(flags[1])

    def update_ice_diagnostics(self, params):
        self._chk(self.L.orc_update_ice_diagnostics(self.h, C.byref(params)))


def single_rank_mesh(x, y, tri1, fast=False):
    R = Rank(fast)
    x = np.ascontiguousarray(x, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    tri1 = np.ascontiguousarray(tri1, np.int32)
    R._chk(R.L.orc_single_rank_mesh(R.h, x.size, _dp(x), _dp(y), tri1.shape[0], _ip(tri1)))
    return R


def nodal_grid(nranks, x, y, tri1, elem_part, ghost_ptr, ghost_val, fast=False):
    """Replay GmshMesh::nodalGrid + initUpdateGhosts for every rank; returns the list of Ranks."""
    ranks = [Rank(fast) for _ in range(nranks)]
    hs = (C.c_void_p * nranks)(*[r.h for r in ranks])
    x = np.ascontiguousarray(x, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    tri1 = np.ascontiguousarray(tri1, np.int32)
    elem_part = np.ascontiguousarray(elem_part, np.int32)
    ghost_ptr = np.ascontiguousarray(ghost_ptr, np.int32)
    ghost_val = np.ascontiguousarray(ghost_val, np.int32)
    if ghost_val.size == 0:
        ghost_val = np.zeros(1, np.int32)
    L = ranks[0].L
    ranks[0]._chk(L.orc_nodal_grid(nranks, hs, x.size, _dp(x), _dp(y), tri1.shape[0], _ip(tri1),
                                   _ip(elem_part), _ip(ghost_ptr), _ip(ghost_val)))
    return ranks


def _handles(ranks):
    return (C.c_void_p * len(ranks))(*[r.h for r in ranks])


def explicit_solve(ranks, params):
    ranks[0]._chk(ranks[0].L.orc_explicit_solve(len(ranks), _handles(ranks), C.byref(params)))


def update_ghosts(ranks, name="M_VT"):
    ranks[0]._chk(ranks[0].L.orc_update_ghosts(len(ranks), _handles(ranks), name.encode()))


def time_subcycles(ranks, params, nsub, threads=False):
    f = ranks[0].L.orc_time_subcycles_mt if threads else ranks[0].L.orc_time_subcycles
    t = f(len(ranks), _handles(ranks), C.byref(params), nsub)
    if t < 0:
        raise RuntimeError("oracle: " + ranks[0].L.orc_last_error().decode())
    return t


def external_data_get_vector(d0, d1, interp_linear_time, current_time, ftime0, ftime1, factor, bias_correction,
                             fast=False):
    """ExternalData::getVector() for one variable: d0/d1 are its two time slices on the mesh nodes."""
    d0 = np.ascontiguousarray(d0, np.float64)
    d1 = np.ascontiguousarray(d1, np.float64)
    out = np.empty_like(d0)
    L = lib(fast)
    rc = L.orc_external_data_get_vector(C.c_long(d0.size), _dp(d0), _dp(d1), int(interp_linear_time),
                                        C.c_double(current_time), C.c_double(ftime0), C.c_double(ftime1),
                                        C.c_double(factor), C.c_double(bias_correction), _dp(out))
    if rc != 0:
        raise RuntimeError("oracle: " + L.orc_last_error().decode())
    return out
