"""Build recipe for the CPU oracle (test infrastructure; see velo_oracle.h)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libvelo_oracle.so")


def build_oracle(force=False):
    """Compile oracle/libvelo_oracle.so with the Makefile next to this file."""
    src = [os.path.join(HERE, f) for f in ("velo_oracle.cpp", "velo_oracle.h", "Makefile")]
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= max(os.path.getmtime(s) for s in src)):
        return LIB
    subprocess.check_call(["make", "-C", HERE, "-s"] + (["-B"] if force else []))
    return LIB


if __name__ == "__main__":
    print(build_oracle(force=True))
