"""ctypes binding of the UNMODIFIED reference bamg library built into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

build(): runs oracle/ref_bamg/Makefile when /root/reference is present (this container); on the GPU box only the
prebuilt oracle/_ref/libref_bamg.so is used.  available() tells the tests whether to run or skip.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(os.path.dirname(_HERE), "_ref", "libref_bamg.so")
_REF = os.environ.get("NEXTSIM_REFERENCE", "/root/reference")
_lib = None


def build():
    if os.path.isdir(os.path.join(_REF, "contrib", "bamg", "src")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-j8", "REF=" + _REF])
    return _LIB if os.path.exists(_LIB) else None


def available():
    return os.path.exists(_LIB) or build() is not None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.ref_bamg_convert.restype = C.c_void_p
        _lib.ref_bamg_get.argtypes = [C.c_void_p] * 4
        _lib.ref_bamg_free.argtypes = [C.c_void_p]
    return _lib


def convert(x, y, tri1):
    """BamgConvertMeshx(bamgmesh, bamggeom, index, x, y, nods, nels) as FiniteElement::distributedMeshProcessing calls
    it (FE.cpp:77-80); returns (NodalElementConnectivity, NodalConnectivity, ElementConnectivity)."""
    L = lib()
    idx = np.ascontiguousarray(np.asarray(tri1).reshape(-1), np.int32)
    x = np.ascontiguousarray(x, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    sizes = (C.c_int * 6)()
    h = C.c_void_p(L.ref_bamg_convert(idx.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p),
                                      y.ctypes.data_as(C.c_void_p), int(x.size), int(idx.size // 3), sizes))
    nc = np.empty((sizes[0], sizes[1]))
    nec = np.empty((sizes[2], sizes[3]))
    ec = np.empty((sizes[4], sizes[5]))
    L.ref_bamg_get(h, nc.ctypes.data_as(C.c_void_p), nec.ctypes.data_as(C.c_void_p), ec.ctypes.data_as(C.c_void_p))
    L.ref_bamg_free(h)
    return nec, nc, ec
