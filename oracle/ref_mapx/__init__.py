"""ctypes binding of the UNMODIFIED reference mapx library built into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).
GmshMesh::lat() (core/src/gmshmesh.cpp:1800-1824) = init_mapx(mppfile) + inverse_mapx(map, X, Y, &lat, &lon)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(os.path.dirname(_HERE), "_ref", "libref_mapx.so")
_REF = os.environ.get("NEXTSIM_REFERENCE", "/root/reference")
_lib = None


def build():
    if os.path.isdir(os.path.join(_REF, "contrib", "mapx", "src")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-j8", "REF=" + _REF], stderr=subprocess.DEVNULL)
    return _LIB if os.path.exists(_LIB) else None


def available():
    return os.path.exists(_LIB) or build() is not None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.init_mapx.restype = C.c_void_p
        _lib.init_mapx.argtypes = [C.c_char_p]
        _lib.inverse_mapx.argtypes = [C.c_void_p, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _lib.close_mapx.argtypes = [C.c_void_p]
    return _lib


def latlon(mppfile, x, y):
    """What GmshMesh::lat() / lon() do, node by node."""
    L = lib()
    m = L.init_mapx(str(mppfile).encode())
    if not m:
        raise RuntimeError("init_mapx failed for %s" % mppfile)
    la, lo = C.c_double(), C.c_double()
    lat, lon = np.empty(len(x)), np.empty(len(x))
    for i in range(len(x)):
        L.inverse_mapx(m, float(x[i]), float(y[i]), C.byref(la), C.byref(lo))
        lat[i], lon[i] = la.value, lo.value
    L.close_mapx(m)
    return lat, lon
