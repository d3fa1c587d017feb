"""ctypes binding of oracle/_ref/libref_fe.so: the reference's OWN explicitSolve / update / updateSigma* / updateGhosts
function bodies, cut from /root/reference at build time (extract.py) and compiled against a stub header (stub_fe.hpp).

TEST INFRASTRUCTURE ONLY.  build() runs the Makefile when /root/reference is present (this container); on the GPU box
only the prebuilt oracle/_ref/libref_fe.so is used.  tests/test_ref_fe_cpu.py holds the oracle to these bodies BIT FOR
BIT, which is what pins the oracle's physics to the reference.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(os.path.dirname(_HERE), "_ref", "libref_fe.so")
_REF = os.environ.get("NEXTSIM_REFERENCE", "/root/reference")
_lib = None


def build():
    if os.path.isfile(os.path.join(_REF, "model", "finiteelement.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "REF=" + _REF])
    return _LIB if os.path.exists(_LIB) else None


def available():
    return os.path.exists(_LIB) or build() is not None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.ref_fe_create.restype = C.c_void_p
        _lib.ref_fe_last_error.restype = C.c_char_p
        _lib.ref_fe_last_error.argtypes = [C.c_void_p]
        _lib.ref_fe_size_double.restype = C.c_long
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class RefFE:
    """P ranks of the reference's FiniteElement (stub class, reference function bodies)."""

    def __init__(self, nranks):
        self.L = lib()
        self.n = nranks
        self.h = C.c_void_p(self.L.ref_fe_create(nranks))

    def __del__(self):
        try:
            self.L.ref_fe_destroy(self.h)
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError("ref_fe: " + self.L.ref_fe_last_error(self.h).decode())

    def set_mesh(self, r, x, y, ndof, tri1, mask_dirichlet, neumann_flags):
        x = np.ascontiguousarray(x, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        tri1 = np.ascontiguousarray(tri1, np.int32).reshape(-1, 3)
        m = np.ascontiguousarray(mask_dirichlet, np.uint8)
        nf = np.ascontiguousarray(neumann_flags, np.int32)
        self._chk(self.L.ref_fe_set_mesh(self.h, r, int(x.size), int(ndof), _dp(x), _dp(y), int(tri1.shape[0]), _ip(tri1),
                                         m.ctypes.data_as(C.c_void_p), int(nf.size), _ip(nf)))

    def set_bamg(self, r, nec, nc):
        nec = np.ascontiguousarray(nec, np.float64)
        nc = np.ascontiguousarray(nc, np.float64)
        self._chk(self.L.ref_fe_set_bamg(self.h, r, _dp(nec), int(nec.shape[1]), _dp(nc), int(nc.shape[1])))

    def set_halo(self, r, which, proc, idx):
        idx = np.ascontiguousarray(idx, np.int32)
        self._chk(self.L.ref_fe_set_halo(self.h, r, which, proc, _ip(idx), int(idx.size)))

    def set(self, r, name, arr):
        arr = np.ascontiguousarray(arr, np.float64)
        self._chk(self.L.ref_fe_set_double(self.h, r, name.encode(), _dp(arr), C.c_long(arr.size)))

    def get(self, r, name):
        n = self.L.ref_fe_size_double(self.h, r, name.encode())
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, np.float64)
        self._chk(self.L.ref_fe_get_double(self.h, r, name.encode(), _dp(out), C.c_long(n)))
        return out

    def shape_coeff(self, r, ne):
        out = np.empty(6 * ne, np.float64)
        self._chk(self.L.ref_fe_get_shape_coeff(self.h, r, _dp(out)))
        return out

    def set_params(self, orc_params, regrid_angle=10.0):
        self._chk(self.L.ref_fe_set_params(self.h, C.byref(orc_params), C.c_double(regrid_angle)))

    def calc_cohesion(self, r, C_fix, C_alea):
        self._chk(self.L.ref_fe_calc_cohesion(self.h, r, C.c_double(C_fix), C.c_double(C_alea)))

    def explicit_solve(self):
        self._chk(self.L.ref_fe_explicit_solve(self.h))

    def update(self):
        self._chk(self.L.ref_fe_update(self.h))

    def check_regridding(self, r):
        out = (C.c_double * 3)()
        flags = (C.c_int * 2)()
        self._chk(self.L.ref_fe_check_regridding(self.h, r, out, flags))
        return out[0], out[1], out[2], bool(flags[0]), bool(flags[1])

    def update_ice_diagnostics(self):
        self._chk(self.L.ref_fe_update_ice_diagnostics(self.h))

    # ---- thermo(): `params` is a ctypes mirror of NsxThermoParams (oracle.thermo.ThermoParams) ----
    def thermo_setup(self, params, current_time):
        self._chk(self.L.ref_fe_thermo_setup(self.h, C.byref(params), C.c_double(current_time)))

    def thermo(self, dt):
        self._chk(self.L.ref_fe_thermo(self.h, int(dt)))
