"""ctypes binding of oracle/_build/libthermo_oracle.so (thermo_oracle.cpp): FiniteElement::thermo(dt) on the CPU.

TEST INFRASTRUCTURE ONLY.  Importers allowed: tests/ and __graft_entry__.smoke().  See thermo_oracle.cpp for what this
build is (the host instance of the element function the device kernel is compiled from) and what pins it to the
reference (oracle/ref_fe, bit for bit, tests/test_thermo_cpu.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libthermo_oracle.so")

# field order of NsxThermoParams (include/nsx.h)
_INT_FIELDS = ["thermo_type", "ocean_constant", "Qio_type", "freezingpoint_type", "newice_type", "melt_type", "alb_scheme",
               "flooding", "use_assim_flux", "temp_dep_healing", "use_meltponds", "force_neutral_atmosphere", "reset_by_date",
               "equal_melting", "use_young_ice_in_myi_reset", "ice_cat_young", "have_sphuma", "have_mixrat", "have_Qlw_in",
               "have_snowfr", "have_snowfall", "have_mld", "reset_month", "reset_day"]
_DBL_FIELDS = ["dtime_step", "ocean_nudge_timeT_days", "ocean_nudge_timeS_days", "Qdw_const", "Fdw_const", "hnull", "PhiF",
               "PhiM", "assim_flux_exponent", "constant_mld", "I_0", "freeze_days_threshold", "meltpond_runoff_fraction",
               "meltpond_depth_to_fraction", "drag_ocean_t", "drag_ocean_q", "alb_ice", "alb_sn", "alb_ponds", "zref_wind",
               "zref_temp", "limiting_lengthscale", "quad_drag_coef_air", "ocean_albedo", "ks", "freezingpoint_mu", "Csens_io",
               "time_relaxation_damage", "deltaT_relaxation_damage", "h_young_min", "h_young_max"]


class ThermoParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in _INT_FIELDS] + [(n, C.c_double) for n in _DBL_FIELDS]


_lib = None


def build():
    src = [os.path.join(_HERE, "thermo_oracle.cpp"), os.path.join(_HERE, "..", "nextsim_b200", "csrc", "nsx_thermo.cuh"),
           os.path.join(_HERE, "..", "include", "nsx.h")]
    if not os.path.exists(_LIB) or any(os.path.getmtime(_LIB) < os.path.getmtime(s) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "_build/libthermo_oracle.so"])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        # THERMO_ORACLE_LIB: an instrumented build (oracle/thermo_coverage.sh measures the branch coverage of the tests with gcov)
        path = os.environ.get("THERMO_ORACLE_LIB")
        if not path:
            build()
            path = _LIB
        _lib = C.CDLL(path)
        _lib.orc_thermo_field_names.restype = C.c_char_p
        _lib.orc_thermo_last_error.restype = C.c_char_p
        assert _lib.orc_thermo_params_size() == C.sizeof(ThermoParams), "ThermoParams out of step with NsxThermoParams"
    return _lib


def default_params(**over):
    p = ThermoParams()
    lib().orc_thermo_params_defaults(C.byref(p))
    for k, v in over.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def field_names():
    return lib().orc_thermo_field_names().decode().split()


def thermo(p, dt, current_time, tri0, nn, wind, VT, ocean, fields):
    """thermo(dt) over all elements.  `fields`: {reference member name: [ne] float64}; missing names read as zeros.
    Returns a new dict with every field after the call (inputs are not modified)."""
    L = lib()
    tri0 = np.ascontiguousarray(tri0, np.int32).reshape(-1, 3)
    ne = tri0.shape[0]
    names = field_names()
    out = {}
    for n in names:
        a = fields.get(n)
        out[n] = np.zeros(ne) if a is None else np.array(a, np.float64, copy=True).reshape(ne)
    unknown = set(fields) - set(names)
    if unknown:
        raise KeyError("thermo oracle: unknown fields %s" % sorted(unknown))
    nd = [np.ascontiguousarray(a, np.float64).reshape(2 * nn) for a in (wind, VT, ocean)]
    cn = (C.c_char_p * len(names))(*[n.encode() for n in names])
    cp = (C.POINTER(C.c_double) * len(names))(*[out[n].ctypes.data_as(C.POINTER(C.c_double)) for n in names])
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rc = L.orc_thermo(C.byref(p), int(dt), C.c_double(current_time), ne, int(nn), tri0.ctypes.data_as(C.POINTER(C.c_int)),
                      dp(nd[0]), dp(nd[1]), dp(nd[2]), len(names), cn, cp)
    if rc != 0:
        raise RuntimeError("thermo oracle: " + L.orc_thermo_last_error().decode())
    return out
