"""Build recipe for oracle/_ref: the reference's own hot-path sources, compiled where they lie
under /root/reference against the stand-in headers in oracle/ref_shim (see its README).

TEST INFRASTRUCTURE.  Outputs go only to oracle/_ref/ (git-ignored, travels to the GPU box).
/root/reference exists only in the build container: on the GPU box the prebuilt library is
used as is.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("VELOSLAM_REFERENCE", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT_DIR, "libvelo_ref.so")
SHIM = os.path.join(HERE, "ref_shim")
REF_SOURCES = ["TransformManager.cxx", "type_defs.cxx", "HDLFrame.cxx", "vtkPacketFileWriter.cxx",
               "CoordiTran.cpp", "TimeSolver.cxx"]   # HDLParser.cxx is included by ref_capi.cpp
CXXFLAGS = ["-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-w", "-DLINUX"]


def reference_present():
    return os.path.isfile(os.path.join(REF, "HDLParser.cxx"))


def build_ref(force=False):
    """Compile oracle/_ref/libvelo_ref.so; returns its path, or None when the reference
    sources are not available and no prebuilt library exists."""
    if not reference_present():
        return LIB if os.path.exists(LIB) else None
    deps = [os.path.join(SHIM, "ref_capi.cpp")] + [os.path.join(REF, s) for s in REF_SOURCES] + \
        [os.path.join(REF, "HDLParser.cxx"), os.path.join(REF, "TimeLine.h"),
         os.path.join(REF, "type_defs.h")]
    for dirpath, _, files in os.walk(SHIM):
        deps += [os.path.join(dirpath, f) for f in files]
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in deps)):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["g++"] + CXXFLAGS + ["-shared", "-I", SHIM, "-I", REF, "-o", LIB,
                                os.path.join(SHIM, "ref_capi.cpp")] + \
        [os.path.join(REF, s) for s in REF_SOURCES]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build_ref(force=True))
