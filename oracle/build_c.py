"""Build the C/OpenMP twin of the oracle (oracle/fct_c.c -> oracle/_fct_c.so) with gcc.  TEST INFRASTRUCTURE.

    python oracle/build_c.py [--force]

The shared object is git-ignored (built artefact) but travels to the GPU box with the repo snapshot."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "fct_c.c")
LIB = os.path.join(HERE, "_fct_c.so")
LIB_NATIVE = os.path.join(HERE, "_fct_c_native.so")


def needs_build():
    return not os.path.exists(LIB) or os.path.getmtime(SRC) > os.path.getmtime(LIB)


def build(force=False, native=False):
    """gcc -O3 -fopenmp; compilers are tried in turn ($CC, /usr/bin/gcc, gcc, cc) because some toolchain wrappers ship
    without libgomp; the last resort is a build without OpenMP (one thread: bench.py then reports cores = 1).
    native=True: -march=native build for THIS host into _fct_c_native.so (always rebuilt: the file may have travelled
    from another machine); raises if OpenMP is not available so that the caller falls back to the portable object."""
    if native:
        ccs = [c for c in (os.environ.get("CC"), "/usr/bin/gcc", "gcc", "cc") if c]
        for cc in ccs:
            try:
                subprocess.check_call([cc, "-fopenmp", "-march=native", "-O3", "-shared", "-fPIC", "-std=c11", SRC, "-o",
                                       LIB_NATIVE, "-lm"], stderr=subprocess.DEVNULL)
                return LIB_NATIVE
            except (subprocess.CalledProcessError, OSError):
                continue
        raise RuntimeError("no native OpenMP build")
    if not force and not needs_build():
        return LIB
    ccs = [c for c in (os.environ.get("CC"), "/usr/bin/gcc", "gcc", "cc") if c]
    base = ["-O3", "-shared", "-fPIC", "-std=c11", SRC, "-o", LIB, "-lm"]
    # no -march=native here: this shared object is built in the build container and travels to a different host
    for flags in (["-fopenmp"], []):
        for cc in ccs:
            try:
                subprocess.check_call([cc] + flags + base, stderr=subprocess.DEVNULL)
                return LIB
            except (subprocess.CalledProcessError, OSError):
                continue
    raise RuntimeError("oracle/fct_c.c could not be compiled (no working C compiler)")


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
