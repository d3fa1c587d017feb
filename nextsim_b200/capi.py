"""ctypes binding of the C ABI in include/nsx.h (libnsx.so).

This is the binding a Python harness uses; a C++ host (the reference) would include nsx.h directly
(see INTEGRATION.md).  There is no CPU fallback: if libnsx.so is missing it is built with nvcc, and
every compute entry point needs a CUDA device.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_ubyte_p = C.POINTER(C.c_ubyte)

DYN = {"bbm": 0, "evp": 3, "mevp": 4}


class NsxDynParams(C.Structure):
    _fields_ = [
        ("dynamics_type", C.c_int), ("basal_stress_type", C.c_int), ("ice_cat_type", C.c_int),
        ("substeps", C.c_int), ("equal_ridging", C.c_int), ("newice_type", C.c_int),
        ("use_young_ice_in_myi_reset", C.c_int), ("stop_after_substeps", C.c_int),
        ("skip_ow_smoother", C.c_int), ("use_coriolis", C.c_int),
        ("dtime_step", C.c_double), ("ocean_turning_angle_rad", C.c_double),
        ("min_h", C.c_double), ("min_c", C.c_double),
        ("young", C.c_double), ("nu0", C.c_double), ("tan_phi", C.c_double),
        ("compr_strength", C.c_double), ("compaction_param", C.c_double),
        ("undamaged_time_relaxation_sigma", C.c_double), ("exponent_relaxation_sigma", C.c_double),
        ("compression_factor", C.c_double), ("exponent_compression_factor", C.c_double),
        ("quad_drag_coef_water", C.c_double),
        ("evp_e", C.c_double), ("evp_Pstar", C.c_double), ("evp_C", C.c_double), ("evp_dmin", C.c_double),
        ("mevp_alpha", C.c_double), ("mevp_beta", C.c_double),
        ("basal_k1", C.c_double), ("basal_k2", C.c_double), ("basal_Cb", C.c_double), ("basal_u0", C.c_double),
        ("C_lab", C.c_double), ("alea_factor", C.c_double), ("time_relaxation_damage_days", C.c_double),
    ]


class NsxMesh(C.Structure):
    _fields_ = [
        ("num_nodes", C.c_int), ("local_ndof", C.c_int), ("num_elements", C.c_int), ("local_nelements", C.c_int),
        ("coord_x", c_double_p), ("coord_y", c_double_p), ("indices", c_int_p),
        ("ghost_nodes", c_ubyte_p), ("mask_dirichlet", c_ubyte_p),
        ("neumann_flags", c_int_p), ("n_neumann_flags", C.c_int),
        ("nodal_element_connectivity", c_double_p), ("nec_width", C.c_int),
        ("nodal_connectivity", c_double_p), ("nc_width", C.c_int),
        ("lat", c_double_p),
    ]


class NsxHalo(C.Structure):
    _fields_ = [
        ("rank", C.c_int), ("nranks", C.c_int),
        ("n_send_peers", C.c_int), ("send_peer", c_int_p), ("send_ptr", c_int_p), ("send_idx", c_int_p),
        ("n_recv_peers", C.c_int), ("recv_peer", c_int_p), ("recv_ptr", c_int_p), ("recv_idx", c_int_p),
    ]


NODAL2 = ("M_VT", "M_UM", "M_UT", "M_wind", "M_ocean", "M_tau_wi", "D_tau_a", "D_tau_w")
NODAL1 = ("M_ssh",)
ELEM = ("M_damage", "M_conc", "M_thick", "M_snow_thick", "M_conc_young", "M_h_young", "M_hs_young",
        "M_thick_myi", "M_conc_myi", "M_ridge_ratio", "M_element_depth", "M_drag_ui", "M_drag_ui_young",
        "M_Cohesion", "M_time_relaxation_damage", "M_surface", "M_delta_x")


class NsxFields(C.Structure):
    _fields_ = (
        [(n, c_double_p) for n in ("M_VT", "M_UM", "M_UT", "M_wind", "M_ocean", "M_tau_wi", "D_tau_a", "D_tau_w")]
        + [("M_ssh", c_double_p)]
        + [("M_sigma", c_double_p * 3), ("M_damage", c_double_p)]
        + [(n, c_double_p) for n in ("M_conc", "M_thick", "M_snow_thick", "M_conc_young", "M_h_young", "M_hs_young",
                                     "M_thick_myi", "M_conc_myi", "M_ridge_ratio", "M_element_depth", "M_drag_ui",
                                     "M_drag_ui_young", "M_Cohesion", "M_time_relaxation_damage", "M_surface",
                                     "M_delta_x", "M_shape_coeff", "D_del_ci_ridge_myi",
                                     "D_conc", "D_thick", "D_snow_thick")]
        + [("D_sigma", c_double_p * 2), ("D_divergence", c_double_p)]
    )


PATHS = {"auto": 0, "tiles": 1, "direct": 2, "resident": 3}
PATH_NAMES = {v: k for k, v in PATHS.items()}


class NsxCreateOptions(C.Structure):
    _fields_ = [("path", C.c_int), ("tile_nodes", C.c_int), ("max_sms", C.c_int), ("use_graph", C.c_int),
                ("overlap", C.c_int), ("boundary_sms", C.c_int), ("ow_skip", C.c_int), ("pad_", C.c_int)]


def create_options(**over):
    """NsxCreateOptions with the library defaults.  libnsx.so itself reads no environment variable; this HARNESS maps the
    variables the tests and profiling scripts use (NSX_PATH, NSX_TILE_NODES, NSX_NO_GRAPH, NSX_OVERLAP, NSX_BOUNDARY_SMS,
    NSX_OW_SKIP) onto the struct, explicit keyword arguments win."""
    o = NsxCreateOptions()
    lib().nsx_create_options_defaults(C.byref(o))
    env = os.environ
    if env.get("NSX_PATH"):
        o.path = PATHS[env["NSX_PATH"]]
    if env.get("NSX_TILE_NODES"):
        o.tile_nodes = int(env["NSX_TILE_NODES"])
    if env.get("NSX_NO_GRAPH"):
        o.use_graph = 0
    if env.get("NSX_OVERLAP"):
        o.overlap = int(env["NSX_OVERLAP"])
    if env.get("NSX_BOUNDARY_SMS"):
        o.boundary_sms = int(env["NSX_BOUNDARY_SMS"])
    if env.get("NSX_OW_SKIP"):
        o.ow_skip = int(env["NSX_OW_SKIP"])
    for k, v in over.items():
        if v is None:
            continue
        setattr(o, k, PATHS[v] if k == "path" and isinstance(v, str) else int(v))
    return o


class NsxCheck(C.Structure):
    _fields_ = [("n_nan", C.c_int), ("n_speed", C.c_int), ("n_range", C.c_int), ("pad_", C.c_int),
                ("max_speed", C.c_double)]


class NsxRegrid(C.Structure):
    _fields_ = [("min_angle", C.c_double), ("min_jacobian", C.c_double), ("max_jacobian", C.c_double),
                ("flip", C.c_int), ("regrid", C.c_int)]


FORCING = {"M_wind": 0, "M_ocean": 1, "M_ssh": 2}


class NsxTiming(C.Structure):
    _fields_ = [("prep_ms", C.c_float), ("subcycle_ms", C.c_float), ("ow_smoother_ms", C.c_float),
                ("update_ms", C.c_float), ("n_launches", C.c_int), ("n_substeps", C.c_int)]


class NsxThermoParams(C.Structure):
    """include/nsx.h NsxThermoParams: the options of FiniteElement::thermo() (model/options.cpp [thermo], [age], ...)."""
    _fields_ = [(n, C.c_int) for n in (
        "thermo_type", "ocean_constant", "Qio_type", "freezingpoint_type", "newice_type", "melt_type", "alb_scheme", "flooding",
        "use_assim_flux", "temp_dep_healing", "use_meltponds", "force_neutral_atmosphere", "reset_by_date", "equal_melting",
        "use_young_ice_in_myi_reset", "ice_cat_young", "have_sphuma", "have_mixrat", "have_Qlw_in", "have_snowfr",
        "have_snowfall", "have_mld", "reset_month", "reset_day")] + [(n, C.c_double) for n in (
        "dtime_step", "ocean_nudge_timeT_days", "ocean_nudge_timeS_days", "Qdw_const", "Fdw_const", "hnull", "PhiF", "PhiM",
        "assim_flux_exponent", "constant_mld", "I_0", "freeze_days_threshold", "meltpond_runoff_fraction",
        "meltpond_depth_to_fraction", "drag_ocean_t", "drag_ocean_q", "alb_ice", "alb_sn", "alb_ponds", "zref_wind", "zref_temp",
        "limiting_lengthscale", "quad_drag_coef_air", "ocean_albedo", "ks", "freezingpoint_mu", "Csens_io",
        "time_relaxation_damage", "deltaT_relaxation_damage", "h_young_min", "h_young_max")]


def thermo_default_params(**over):
    p = NsxThermoParams()
    lib().nsx_thermo_params_defaults(C.byref(p))
    for k, v in over.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


EXPORTS = (
    "nsx_resident_plan_info", "nsx_create", "nsx_create_ex", "nsx_create_options_defaults", "nsx_device_sm_count", "nsx_destroy", "nsx_last_error", "nsx_version", "nsx_params_defaults", "nsx_params_from_cfg",
    "nsx_set_params", "nsx_upload", "nsx_download", "nsx_explicit_solve", "nsx_update", "nsx_update_ghosts",
    "nsx_check", "nsx_synchronize", "nsx_get_timing", "nsx_get_stream", "nsx_halo_blob_size", "nsx_halo_blob",
    "nsx_halo_connect_blob", "nsx_halo_connect_local", "nsx_halo_finalize", "nsx_group_explicit_solve",
    "nsx_host_register", "nsx_host_unregister", "nsx_abi_sizes", "nsx_tile_info", "nsx_plan_info", "nsx_cfg_last_error",
    "nsx_check_regridding", "nsx_update_ice_diagnostics", "nsx_forcing_load", "nsx_forcing_apply",
    "nsx_thermo_params_defaults", "nsx_thermo_upload", "nsx_thermo_download", "nsx_thermo_upload_many", "nsx_thermo_download_many",
    "nsx_thermo", "nsx_thermo_forcing_load", "nsx_thermo_forcing_apply",
    "nsx_validate_mesh", "nsx_mapx_latlon", "nsx_partmesh_lat_from_mpp", "nsx_mapx_last_error", "nsx_partmesh_read", "nsx_partmesh_build", "nsx_partmesh_bc_marked_nodes", "nsx_partmesh_set_lat",
    "nsx_partmesh_views", "nsx_partmesh_ids", "nsx_partmesh_destroy", "nsx_partmesh_last_error",
)

_lib = None


def lib():
    """Load (building first if needed) libnsx.so.  Raises if it cannot be built: no fallback."""
    global _lib
    if _lib is None:
        # NSX_LIBRARY: a variant built by profiles/*.sh (harness knob; the library itself reads no environment variable)
        path = os.environ.get("NSX_LIBRARY") or _build.build()
        L = C.CDLL(path)
        L.nsx_last_error.restype = C.c_char_p
        L.nsx_last_error.argtypes = [C.c_void_p]
        L.nsx_cfg_last_error.restype = C.c_char_p
        L.nsx_get_stream.restype = C.c_void_p
        L.nsx_get_stream.argtypes = [C.c_void_p]
        L.nsx_params_defaults.restype = None
        L.nsx_host_register.argtypes = [C.c_void_p, C.c_ulong]
        L.nsx_host_unregister.argtypes = [C.c_void_p]
        for f in ("nsx_destroy", "nsx_set_params", "nsx_upload", "nsx_download", "nsx_explicit_solve", "nsx_update",
                  "nsx_update_ghosts", "nsx_check", "nsx_synchronize", "nsx_get_timing", "nsx_halo_finalize"):
            getattr(L, f).argtypes = [C.c_void_p] + ([C.c_void_p] if f in (
                "nsx_set_params", "nsx_upload", "nsx_download", "nsx_check", "nsx_get_timing") else [])
        L.nsx_check_regridding.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
        L.nsx_update_ice_diagnostics.argtypes = [C.c_void_p]
        L.nsx_forcing_load.argtypes = [C.c_void_p, C.c_int, C.c_int, c_double_p]
        L.nsx_forcing_apply.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_double] * 5
        L.nsx_partmesh_last_error.restype = C.c_char_p
        L.nsx_partmesh_read.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p]
        L.nsx_partmesh_build.argtypes = [C.c_int, c_double_p, c_double_p, C.c_int, c_int_p, c_int_p, c_int_p, c_int_p,
                                         C.c_int, C.c_int, C.c_void_p]
        L.nsx_partmesh_bc_marked_nodes.argtypes = [C.c_void_p, c_int_p, C.c_int, c_int_p, C.c_int]
        L.nsx_partmesh_set_lat.argtypes = [C.c_void_p, c_double_p]
        L.nsx_partmesh_views.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.nsx_partmesh_ids.argtypes = [C.c_void_p, C.c_void_p, c_int_p]
        L.nsx_partmesh_destroy.argtypes = [C.c_void_p]
        L.nsx_mapx_last_error.restype = C.c_char_p
        L.nsx_mapx_latlon.argtypes = [C.c_char_p, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]
        L.nsx_partmesh_lat_from_mpp.argtypes = [C.c_void_p, C.c_char_p]
        L.nsx_halo_blob_size.argtypes = [C.c_void_p, C.c_int]
        L.nsx_halo_blob.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.nsx_halo_connect_blob.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.nsx_halo_connect_local.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def default_params():
    p = NsxDynParams()
    lib().nsx_params_defaults(C.byref(p))
    return p


def params_from_cfg(path):
    p = NsxDynParams()
    rc = lib().nsx_params_from_cfg(str(path).encode(), C.byref(p))
    if rc != 0:
        raise RuntimeError("nsx_params_from_cfg: " + lib().nsx_cfg_last_error().decode())
    return p


def _f64(a):
    a = np.ascontiguousarray(a, np.float64)
    return a, a.ctypes.data_as(c_double_p)


def _i32(a):
    a = np.ascontiguousarray(a, np.int32)
    return a, a.ctypes.data_as(c_int_p)


def _u8(a):
    a = np.ascontiguousarray(a, np.uint8)
    return a, a.ctypes.data_as(c_ubyte_p)


def mesh_structs(lm):
    """NsxMesh / NsxHalo views of a partition.LocalMesh (plus the arrays that must stay alive)."""
    keep = []
    M = NsxMesh()
    M.num_nodes, M.local_ndof = lm.num_nodes, lm.local_ndof
    M.num_elements, M.local_nelements = lm.num_elements, lm.local_nelements
    a, M.coord_x = _f64(lm.x); keep.append(a)
    a, M.coord_y = _f64(lm.y); keep.append(a)
    a, M.indices = _i32(lm.indices.reshape(-1)); keep.append(a)
    a, M.ghost_nodes = _u8(lm.ghostNodes.reshape(-1)); keep.append(a)
    a, M.mask_dirichlet = _u8(lm.mask_dirichlet); keep.append(a)
    a, M.neumann_flags = _i32(lm.neumann_flags); keep.append(a)
    M.n_neumann_flags = int(lm.neumann_flags.size)
    a, M.nodal_element_connectivity = _f64(lm.nodal_element_connectivity.reshape(-1)); keep.append(a)
    M.nec_width = int(lm.nodal_element_connectivity.shape[1])
    a, M.nodal_connectivity = _f64(lm.nodal_connectivity.reshape(-1)); keep.append(a)
    M.nc_width = int(lm.nodal_connectivity.shape[1])
    a, M.lat = _f64(lm.lat); keep.append(a)
    H = None
    if lm.nranks > 1:
        H = NsxHalo()
        H.rank, H.nranks = lm.rank, lm.nranks
        sp = sorted(lm.send_to)
        rp = sorted(lm.recv_from)
        H.n_send_peers, H.n_recv_peers = len(sp), len(rp)
        a, H.send_peer = _i32(np.array(sp, np.int32)); keep.append(a)
        a, H.recv_peer = _i32(np.array(rp, np.int32)); keep.append(a)
        sptr = np.cumsum([0] + [lm.send_to[p].size for p in sp])
        rptr = np.cumsum([0] + [lm.recv_from[p].size for p in rp])
        a, H.send_ptr = _i32(sptr); keep.append(a)
        a, H.recv_ptr = _i32(rptr); keep.append(a)
        a, H.send_idx = _i32(np.concatenate([lm.send_to[p] for p in sp]) if sp else np.zeros(0)); keep.append(a)
        a, H.recv_idx = _i32(np.concatenate([lm.recv_from[p] for p in rp]) if rp else np.zeros(0)); keep.append(a)
    return M, H, keep


def resident_plan_info(lm, sms=148):
    """Host-only statistics of the state-resident plan for this rank (see nsx.h)."""
    M, H, keep = mesh_structs(lm)
    out = (C.c_int * 14)()
    L = lib()
    if L.nsx_resident_plan_info(C.byref(M), C.byref(H) if H is not None else None, int(sms), out, 14) != 0:
        raise RuntimeError(L.nsx_last_error(None).decode())
    names = ("fits", "ntiles", "tile_nodes", "slot_space", "max_slots", "max_local_nodes", "smem_bytes", "export_nodes",
             "early_own_slots", "halo_slots", "own_slots", "smem_limit", "gather_half_warp_wavefronts", "gather_half_warp_cells")
    return dict(zip(names, list(out)))


def validate_mesh(lm):
    """Host-only input checks of nsx_create (no GPU): raises RuntimeError with nsx_create's message."""
    M, H, keep = mesh_structs(lm)
    L = lib()
    if L.nsx_validate_mesh(C.byref(M), C.byref(H) if H is not None else None) != 0:
        raise RuntimeError(L.nsx_last_error(None).decode())


class Solver:
    """Thin owner of one nsx_handle.  Methods map 1:1 on the C ABI."""

    def __init__(self, lm, device=0, **options):
        """lm: partition.LocalMesh with bamg tables, BC masks and lat filled in.  options: NsxCreateOptions fields
        (path="auto"|"tiles"|"direct"|"resident", tile_nodes, max_sms, use_graph, ...)."""
        self.L = lib()
        self.lm = lm
        self.nn, self.ne = lm.num_nodes, lm.num_elements
        M, H, keep = mesh_structs(lm)
        h = C.c_void_p()
        self.options = create_options(**options)
        rc = self.L.nsx_create_ex(C.byref(M), C.byref(H) if H is not None else None, int(device), C.byref(self.options),
                                  C.byref(h))
        if rc != 0:
            raise RuntimeError("nsx_create: " + self.L.nsx_last_error(None).decode())
        self.h = h
        self.peers = sorted(set(lm.send_to) | set(lm.recv_from)) if lm.nranks > 1 else []

    def tile_info(self):
        """ntiles, nodes/tile, slots, max local nodes, max slots, boundary tiles, smem bytes, path name."""
        t = (C.c_int * 8)()
        self.L.nsx_tile_info(self.h, t, 8)
        return list(t[:7]) + [PATH_NAMES.get(t[7], "?")]

    @property
    def path(self):
        return self.tile_info()[7]

    def close(self):
        if getattr(self, "h", None):
            self.L.nsx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s: %s" % (what, self.L.nsx_last_error(self.h).decode()))

    def set_params(self, p):
        self._chk(self.L.nsx_set_params(self.h, C.byref(p)), "nsx_set_params")

    def _fields(self, d):
        F = NsxFields()
        keep = []
        for k, v in d.items():
            if k in ("M_sigma", "D_sigma"):
                for i in range(3 if k == "M_sigma" else 2):
                    a = v[i]
                    assert a.dtype == np.float64 and a.flags.c_contiguous and a.size == self.ne
                    getattr(F, k)[i] = a.ctypes.data_as(c_double_p)
                    keep.append(a)
                continue
            assert v.dtype == np.float64 and v.flags.c_contiguous, k
            n = 2 * self.nn if k in NODAL2 else self.nn if k in NODAL1 else 6 * self.ne if k == "M_shape_coeff" else self.ne
            assert v.size == n, (k, v.size, n)
            setattr(F, k, v.ctypes.data_as(c_double_p))
            keep.append(v)
        return F, keep

    def upload(self, **arrays):
        arrays = {k: np.ascontiguousarray(v, np.float64) if k != "M_sigma" else
                  [np.ascontiguousarray(s, np.float64) for s in v] for k, v in arrays.items()}
        F, keep = self._fields(arrays)
        self._chk(self.L.nsx_upload(self.h, C.byref(F)), "nsx_upload")

    def download(self, *names, out=None):
        out = {} if out is None else out
        for k in names:
            if k in out:
                continue
            if k in ("M_sigma", "D_sigma"):
                out[k] = [np.empty(self.ne) for _ in range(3 if k == "M_sigma" else 2)]
            else:
                n = 2 * self.nn if k in NODAL2 else self.nn if k in NODAL1 else 6 * self.ne if k == "M_shape_coeff" else self.ne
                out[k] = np.empty(n)
        F, keep = self._fields({k: out[k] for k in names})
        self._chk(self.L.nsx_download(self.h, C.byref(F)), "nsx_download")
        return out

    def explicit_solve(self):
        self._chk(self.L.nsx_explicit_solve(self.h), "nsx_explicit_solve")

    def update(self):
        self._chk(self.L.nsx_update(self.h), "nsx_update")

    def update_ghosts(self):
        self._chk(self.L.nsx_update_ghosts(self.h), "nsx_update_ghosts")

    def synchronize(self):
        self._chk(self.L.nsx_synchronize(self.h), "nsx_synchronize")

    def check(self):
        c = NsxCheck()
        self._chk(self.L.nsx_check(self.h, C.byref(c)), "nsx_check")
        return c

    def check_regridding(self, regrid_angle):
        """Local part of FiniteElement::checkRegridding() on the resident M_UM (no download)."""
        r = NsxRegrid()
        self._chk(self.L.nsx_check_regridding(self.h, float(regrid_angle), C.byref(r)), "nsx_check_regridding")
        return r

    def update_ice_diagnostics(self):
        self._chk(self.L.nsx_update_ice_diagnostics(self.h), "nsx_update_ice_diagnostics")

    def forcing_load(self, name, slot, data):
        """ExternalData time slice `slot` (0/1) of M_wind / M_ocean ([u | v]) or M_ssh, local numbering."""
        a, p = _f64(data)
        assert a.size == (self.nn if name == "M_ssh" else 2 * self.nn), (name, a.size)
        self._chk(self.L.nsx_forcing_load(self.h, FORCING[name], int(slot), p), "nsx_forcing_load")

    def forcing_apply(self, name, interp_linear_time, current_time, ftime0, ftime1, factor=1., bias_correction=0.):
        self._chk(self.L.nsx_forcing_apply(self.h, FORCING[name], int(bool(interp_linear_time)), float(current_time),
                                           float(ftime0), float(ftime1), float(factor), float(bias_correction)),
                  "nsx_forcing_apply")

    # ---- thermo() ----
    def _thermo_xfer(self, fn, what, arrays):
        names = list(arrays)
        cn = (C.c_char_p * len(names))(*[n.encode() for n in names])
        cp = (c_double_p * len(names))(*[arrays[n].ctypes.data_as(c_double_p) for n in names])
        self._chk(fn(self.h, len(names), cn, cp), what)

    def thermo_upload(self, **arrays):
        """forcing / slab-ocean / tracer fields of thermo() by reference member name, [num_elements] each, one call"""
        a = {k: np.ascontiguousarray(v, np.float64) for k, v in arrays.items()}
        for k, v in a.items():
            assert v.size == self.ne, (k, v.size, self.ne)
        self._thermo_xfer(self.L.nsx_thermo_upload_many, "nsx_thermo_upload_many", a)

    def thermo_download(self, *names):
        out = {k: np.empty(self.ne) for k in names}
        self._thermo_xfer(self.L.nsx_thermo_download_many, "nsx_thermo_download_many", out)
        return out

    def thermo_forcing_load(self, name, slot, data):
        """ExternalData time slice `slot` (0/1) of one element forcing variable of thermo() ("M_tair", ...)"""
        a, p = _f64(data)
        assert a.size == self.ne, (name, a.size)
        self._chk(self.L.nsx_thermo_forcing_load(self.h, name.encode(), int(slot), p), "nsx_thermo_forcing_load")

    def thermo_forcing_apply(self, name, interp_linear_time, current_time, ftime0, ftime1, factor=1., bias_correction=0.):
        self._chk(self.L.nsx_thermo_forcing_apply(self.h, name.encode(), int(bool(interp_linear_time)), C.c_double(current_time),
                                                  C.c_double(ftime0), C.c_double(ftime1), C.c_double(factor),
                                                  C.c_double(bias_correction)), "nsx_thermo_forcing_apply")

    def thermo(self, p, dt, current_time):
        """FiniteElement::thermo(dt) at model time current_time (days since 1900-01-01) on the resident state"""
        self._chk(self.L.nsx_thermo(self.h, C.byref(p), int(dt), C.c_double(current_time)), "nsx_thermo")

    def timing(self):
        t = NsxTiming()
        self._chk(self.L.nsx_get_timing(self.h, C.byref(t)), "nsx_get_timing")
        return t

    # ---- halo wiring ----
    def halo_blob(self, peer):
        n = self.L.nsx_halo_blob_size(self.h, int(peer))
        buf = (C.c_ubyte * n)()
        self._chk(self.L.nsx_halo_blob(self.h, int(peer), buf), "nsx_halo_blob")
        return bytes(buf)

    def halo_connect_blob(self, peer, blob):
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._chk(self.L.nsx_halo_connect_blob(self.h, int(peer), buf), "nsx_halo_connect_blob")

    def halo_connect_local(self, peer, other):
        self._chk(self.L.nsx_halo_connect_local(self.h, int(peer), other.h), "nsx_halo_connect_local")

    def halo_finalize(self):
        self._chk(self.L.nsx_halo_finalize(self.h), "nsx_halo_finalize")


def group_explicit_solve(solvers):
    L = lib()
    hs = (C.c_void_p * len(solvers))(*[s.h for s in solvers])
    rc = L.nsx_group_explicit_solve(len(solvers), hs)
    if rc != 0:
        raise RuntimeError("nsx_group_explicit_solve: " + L.nsx_last_error(solvers[0].h).decode())


def connect_local_group(solvers):
    """Wire the halos of ranks that all live in this process (tests / single-GPU emulation)."""
    by_rank = {s.lm.rank: s for s in solvers}
    for s in solvers:
        for p in s.peers:
            s.halo_connect_local(p, by_rank[p])
    for s in solvers:
        s.halo_finalize()


class PartMesh:
    """One rank's partitioned mesh from the host library: msh-2.2 reader or in-memory tags ->
    nodalGrid numbering, halo lists, boundary masks, bamg tables.  No GPU needed."""

    def __init__(self, handle):
        self.L = lib()
        self.h = handle

    @staticmethod
    def _chk(rc, what):
        if rc != 0:
            raise RuntimeError("%s: %s" % (what, lib().nsx_partmesh_last_error().decode()))

    @classmethod
    def read(cls, path, rank, nranks, fmt="binary", ordering="gmsh"):
        h = C.c_void_p()
        cls._chk(lib().nsx_partmesh_read(str(path).encode(), fmt.encode(), ordering.encode(), int(rank), int(nranks),
                                         C.byref(h)), "nsx_partmesh_read")
        return cls(h)

    @classmethod
    def build(cls, x, y, tri1, rank=0, nranks=1, elem_part=None, ghost_ptr=None, ghost_val=None):
        xa, xp = _f64(x)
        ya, yp = _f64(y)
        ta, tp = _i32(np.asarray(tri1).reshape(-1))
        if nranks > 1:
            pa, pp = _i32(elem_part)
            ga, gp = _i32(ghost_ptr)
            va, vp = _i32(ghost_val if len(ghost_val) else np.zeros(1, np.int32))
        else:
            pp = gp = vp = None
        h = C.c_void_p()
        cls._chk(lib().nsx_partmesh_build(xa.size, xp, yp, ta.size // 3, tp, pp, gp, vp, int(rank), int(nranks),
                                          C.byref(h)), "nsx_partmesh_build")
        return cls(h)

    def close(self):
        if getattr(self, "h", None):
            self.L.nsx_partmesh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bc_marked_nodes(self, dirichlet_flags_root, neumann_flags_root):
        da, dp = _i32(dirichlet_flags_root)
        na, npp = _i32(neumann_flags_root)
        self._chk(self.L.nsx_partmesh_bc_marked_nodes(self.h, dp, da.size, npp, na.size), "nsx_partmesh_bc_marked_nodes")

    def set_lat(self, lat):
        a, p = _f64(lat)
        self._chk(self.L.nsx_partmesh_set_lat(self.h, p), "nsx_partmesh_set_lat")

    def lat_from_mpp(self, mppfile):
        """GmshMesh::lat(): node latitudes from the projection file (mesh.mppfile)."""
        self._chk(self.L.nsx_partmesh_lat_from_mpp(self.h, str(mppfile).encode()), "nsx_partmesh_lat_from_mpp")

    def views(self):
        M, H = NsxMesh(), NsxHalo()
        self._chk(self.L.nsx_partmesh_views(self.h, C.byref(M), C.byref(H)), "nsx_partmesh_views")
        return M, H

    def to_local_mesh(self):
        """Copy into a partition.LocalMesh (what Solver() and the tests consume)."""
        from . import partition as pt
        M, H = self.views()
        ids = (c_int_p * 4)()
        sizes = (C.c_int * 4)()
        self._chk(self.L.nsx_partmesh_ids(self.h, ids, sizes), "nsx_partmesh_ids")
        nn, ne = M.num_nodes, M.num_elements

        def arr(p, n, dt):
            return np.ctypeslib.as_array(p, shape=(n,)).astype(dt, copy=True) if n else np.zeros(0, dt)
        lm = pt.LocalMesh(rank=H.rank, nranks=max(H.nranks, 1), num_nodes=nn, local_ndof=M.local_ndof, num_elements=ne,
                          local_nelements=M.local_nelements, x=arr(M.coord_x, nn, np.float64),
                          y=arr(M.coord_y, nn, np.float64), indices=arr(M.indices, 3 * ne, np.int32).reshape(ne, 3),
                          ghostNodes=arr(M.ghost_nodes, 3 * ne, np.uint8).reshape(ne, 3),
                          node_gid=arr(ids[0], nn, np.int32), node_rid=arr(ids[1], nn, np.int32),
                          elem_gid=arr(ids[2], ne, np.int32), elem_part=arr(ids[3], ne, np.int32),
                          local_ghost=np.zeros(0, np.int32))
        lm.local_ghost = np.sort(lm.node_rid[M.local_ndof:])
        for k in range(H.n_send_peers):
            a, b = H.send_ptr[k], H.send_ptr[k + 1]
            lm.send_to[int(H.send_peer[k])] = arr(C.cast(C.addressof(H.send_idx.contents) + 4 * a, c_int_p), b - a, np.int32)
        for k in range(H.n_recv_peers):
            a, b = H.recv_ptr[k], H.recv_ptr[k + 1]
            lm.recv_from[int(H.recv_peer[k])] = arr(C.cast(C.addressof(H.recv_idx.contents) + 4 * a, c_int_p), b - a, np.int32)
        lm.mask_dirichlet = arr(M.mask_dirichlet, nn, np.uint8)
        lm.neumann_flags = arr(M.neumann_flags, M.n_neumann_flags, np.int32)
        lm.dirichlet_flags = np.nonzero(lm.mask_dirichlet)[0].astype(np.int32)
        lm.lat = arr(M.lat, nn, np.float64)
        lm.nodal_element_connectivity = arr(M.nodal_element_connectivity, nn * M.nec_width, np.float64).reshape(nn, M.nec_width)
        lm.nodal_connectivity = arr(M.nodal_connectivity, nn * M.nc_width, np.float64).reshape(nn, M.nc_width)
        return lm


def mapx_latlon(mppfile, x, y):
    """Inverse polar stereographic map of mesh coordinates (GmshMesh::lat() / lon())."""
    xa, xp = _f64(x)
    ya, yp = _f64(y)
    lat, lon = np.empty(xa.size), np.empty(xa.size)
    L = lib()
    if L.nsx_mapx_latlon(str(mppfile).encode(), xa.size, xp, yp, lat.ctypes.data_as(c_double_p),
                         lon.ctypes.data_as(c_double_p)) != 0:
        raise RuntimeError(L.nsx_mapx_last_error().decode())
    return lat, lon
