"""Builds libnsx.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnsx.so")
SOURCES = ["nsx_api.cu", "nsx_mesh.cpp", "nsx_cfg.cpp", "nsx_partmesh.cpp", "nsx_mapx.cpp"]
HEADERS = ["nsx_kernels.cuh", "nsx_internal.h", "nsx_mesh.h", os.path.join("..", "..", "include", "nsx.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    extra = []
    for k in ("NSX_SUB_TPB", "NSX_SUB_MINB", "NSX_SUB_STAGES", "NSX_SUB_GROUPS", "NSX_DIRECT_TPB", "NSX_DIRECT_MINB", "NSX_SUB_CTAS_PER_SM", "NSX_RES_TPB", "NSX_RES_CTAS"):           # kernel-shape experiments
        if os.environ.get(k):
            extra.append("-D%s=%s" % (k, os.environ[k]))
    if os.environ.get("NSX_DEBUG_CHECKS"):                # device-side bounds checks of the index tables (nsx_kernels.cuh)
        extra.append("-DNSX_DEBUG_CHECKS")
    cmd = [_nvcc()] + NVCC_FLAGS + extra + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc failed building libnsx.so (see %s)" % log)
    if verbose:
        print(r.stdout)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose=True)
