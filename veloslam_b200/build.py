"""In-tree build of the CUDA C-ABI library (sm_100a only).

`nvcc` cross-compiles without a GPU; the resulting `veloslam_b200/libveloslam_b200.so` is
git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libveloslam_b200.so")
SOURCES = [os.path.join(CSRC, "vs_capi.cu")]
DEPS = SOURCES + [os.path.join(CSRC, f) for f in ("vs_kernels.cuh", "vs_single_pass.cuh", "vs_device.cuh", "vs_layout.cuh")] + \
    [os.path.join(os.path.dirname(HERE), "include", "veloslam_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    # device code keeps the reference's separate multiplies and adds everywhere (the decode path
    # uses explicit __dmul_rn/__dadd_rn already; this covers the N3 geodesy kernel)
    "-fmad=false",
    # host side restates reference arithmetic: no FMA contraction there either
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden",
    "-shared", "-cudart", "static",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=...)")


def build_library(force=False, verbose=False, extra_flags=(), out=None):
    """Compile libveloslam_b200.so if missing or older than its sources.  extra_flags / out:
    A/B builds of kernel geometry variants (-DVS_DEC_STAGES=..., see vs_kernels.cuh) into
    another file."""
    out = out or LIB
    if (not force and out == LIB and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in DEPS)):
        return LIB
    tmp = _tmp_name(out)
    cmd = [find_nvcc()] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", tmp] + SOURCES
    subprocess.check_call(cmd)
    os.replace(tmp, out)
    return out


def _tmp_name(out):
    """Outputs are written next to their final place and renamed into it: several ranks of one job
    may decide to (re)build the same file at once, and none of them may ever load a half-written
    one."""
    d, b = os.path.split(out)
    return os.path.join(d, f".{os.getpid()}.tmp.{b}")


FACADE_DIR = os.path.join(HERE, "cpp")
FACADE_LIB = os.path.join(HERE, "libveloslam_facade.so")
FACADE_SOURCES = ["FrameArena.cpp", "type_defs.cpp", "TransformManager.cpp", "vtkPacketFile.cpp", "HDLManager.cpp", "CoordiTran.cpp", "TimeSolver.cpp", "HDLSource.cpp", "INSSource.cpp",
                  "CalibrationFile.cpp", "HDLParser.cpp"]
DRIVER_SRC = os.path.join(os.path.dirname(HERE), "tests", "cpp", "facade_driver.cpp")
DRIVER_EXE = os.path.join(os.path.dirname(HERE), "tests", "cpp", "facade_driver")
CXXFLAGS = ["-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-Wall"]


def build_facade(force=False):
    """C++ facade (HDLParser / HDLFrame / TransformManager / TimeLine with the reference's
    names) over the C ABI, plus the test driver executable."""
    build_library()
    srcs = [os.path.join(FACADE_DIR, s) for s in FACADE_SOURCES]
    deps = srcs + [os.path.join(FACADE_DIR, f) for f in os.listdir(FACADE_DIR)] + [LIB]
    if (force or not os.path.exists(FACADE_LIB)
            or os.path.getmtime(FACADE_LIB) < max(os.path.getmtime(d) for d in deps)):
        tmp = _tmp_name(FACADE_LIB)
        subprocess.check_call(["g++"] + CXXFLAGS + ["-shared", "-o", tmp] + srcs +
                              ["-L", HERE, "-lveloslam_b200", "-lpthread", "-Wl,-rpath,$ORIGIN"])
        os.replace(tmp, FACADE_LIB)
    if (force or not os.path.exists(DRIVER_EXE)
            or os.path.getmtime(DRIVER_EXE) < max(os.path.getmtime(DRIVER_SRC),
                                                  os.path.getmtime(FACADE_LIB))):
        tmp = _tmp_name(DRIVER_EXE)
        subprocess.check_call(["g++"] + CXXFLAGS + ["-o", tmp, DRIVER_SRC, "-I", FACADE_DIR,
                               "-L", HERE, "-lveloslam_facade", "-lveloslam_b200", "-lpthread",
                               "-Wl,-rpath,$ORIGIN/../../veloslam_b200"])
        os.replace(tmp, DRIVER_EXE)
    return FACADE_LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
    print(build_facade(force=True))
