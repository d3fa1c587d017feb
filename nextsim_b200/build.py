"""Builds libnsx.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnsx.so")
SOURCES = ["nsx_api.cu", "nsx_mesh.cpp", "nsx_cfg.cpp", "nsx_partmesh.cpp", "nsx_mapx.cpp"]
# compiled separately with -fmad=false: element-wise physics held to the reference's rounding (see nsx_thermo.cu)
NOFMA_SOURCES = ["nsx_thermo.cu"]
HEADERS = ["nsx_kernels.cuh", "nsx_internal.h", "nsx_mesh.h", "nsx_thermo.cuh", "nsx_thermo_api.cuh", os.path.join("..", "..", "include", "nsx.h")]

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
    deps = [os.path.join(CSRC, s) for s in SOURCES + NOFMA_SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, out=None):
    """out: write the library there instead of nextsim_b200/libnsx.so (kernel-shape experiments: profiles/*.sh build
    variants with the NSX_* macros and the harness loads one through NSX_LIBRARY)."""
    if out is None and not force and not needs_build():
        return LIB
    extra = []
    for k in ("NSX_SUB_TPB", "NSX_SUB_MINB", "NSX_SUB_STAGES", "NSX_SUB_GROUPS", "NSX_DIRECT_TPB", "NSX_DIRECT_MINB", "NSX_SUB_CTAS_PER_SM", "NSX_RES_TPB", "NSX_RES_CTAS"):           # kernel-shape experiments
        if os.environ.get(k):
            extra.append("-D%s=%s" % (k, os.environ[k]))
    if os.environ.get("NSX_DEBUG_CHECKS"):                # device-side bounds checks of the index tables (nsx_kernels.cuh)
        extra.append("-DNSX_DEBUG_CHECKS")
    log = os.path.join(HERE, "build.log")
    objdir = os.path.join(HERE, "_obj")
    os.makedirs(objdir, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]
    thermo = []
    for k in ("NSX_THERMO_TPB", "NSX_THERMO_MINB"):      # launch shape of k_thermo (nsx_thermo.cu)
        if os.environ.get(k):
            thermo.append("-D%s=%s" % (k, os.environ[k]))
    for k in ("NSX_THERMO_FAST_MINMAX", "NSX_THERMO_EXACT_DIV"):     # arithmetic experiments of k_thermo
        if os.environ.get(k):
            thermo.append("-D%s" % k)
    objs, cmds = [], []
    for src in NOFMA_SOURCES:
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ("" if out is None else "_" + os.path.basename(out)) + ".o")
        fmad = [] if os.environ.get("NSX_THERMO_FMAD") else ["-fmad=false"]
        cmds.append([_nvcc()] + compile_flags + thermo + fmad + ["-c", os.path.join(CSRC, src), "-o", obj])
        objs.append(obj)
    cmds.append([_nvcc()] + NVCC_FLAGS + extra + [os.path.join(CSRC, s) for s in SOURCES] + objs + ["-o", out or LIB])
    text = ""
    for cmd in cmds:
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        text += " ".join(cmd) + "\n" + r.stdout
        with open(log, "w") as f:
            f.write(text)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError("nvcc failed building libnsx.so (see %s)" % log)
    if verbose:
        print(text)
    return out or LIB


if __name__ == "__main__":
    build(force=True, verbose=True, out=os.environ.get("NSX_BUILD_OUT"))
