#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY (oracle/): build the reference CPU oracle from the sources where
they lie under /root/reference/src (never copied), outputs only into oracle/_ref/.

Recipe = the reference's own `ckernel=yes` object list (Makefile:229) and flags
(Makefile:152-164,199: -O3 -fopenmp -DSW4_CROUTINES -DSW4_OPENMP -Isrc -Isrc/double), with
the two missing third-party pieces replaced by stand-ins written for this repo:
host/stubs/mpi.h (single rank) and host/stubs/dspev_stub.C (3x3 symmetric eigenvalues).

Products:
  oracle/_ref/sw4lite_ref   the unmodified reference program (CPU, OpenMP)
  oracle/_ref/libsw4ref.so  the same objects + oracle/ref_shim.C (extern "C" access to the
                            reference kernels and to a set-up EW object); in this library the
                            reference's EW::timesteploop symbol is weakened so that the shim's
                            empty definition is used by the reference constructor.
"""
import os, subprocess, sys, shutil
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SW4_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "src")
OUT = os.path.join(HERE, "_ref")
OBJ = os.path.join(OUT, "obj")
STUBS = os.path.join(os.path.dirname(HERE), "host", "stubs")   # single-rank mpi.h, 3x3 dspev_ (written for this repository)

OBJS = ("main EW Sarray Source SuperGrid GridPointSource time_functions EW_cuda ew-cfromfort "
        "rhs4sg rhs4sg_rev EWCuda CheckPoint Parallel_IO EW-dg MaterialData MaterialBlock "
        "Polynomial SecondOrderSection Filter TimeSeries sacsubc curvilinear-c rhs4sgcurv "
        "rhs4sgcurv_rev").split()

CXX = os.environ.get("SW4B200_CXX", "/usr/bin/g++")  # $CXX in this image points at a wrapper that cannot link -fopenmp
FLAGS = ["-O3", "-fopenmp", "-fPIC", "-w", "-DSW4_CROUTINES", "-DSW4_OPENMP",
         "-I", STUBS, "-I", SRC, "-I", os.path.join(SRC, "double")]
TSL_SYMBOL = "_ZN2EW12timesteploopERSt6vectorI6SarraySaIS1_EES4_"


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise SystemExit("oracle/_ref build failed")


def newer(target, *deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print("oracle/_ref: %s not present, keeping prebuilt files" % SRC)
        return False
    os.makedirs(OBJ, exist_ok=True)

    def cc(name):
        src = os.path.join(SRC, name + ".C")
        obj = os.path.join(OBJ, name + ".o")
        if not newer(obj, src, os.path.join(STUBS, "mpi.h")):
            run([CXX] + FLAGS + ["-c", src, "-o", obj])
        return obj

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(cc, OBJS))
    stub = os.path.join(OBJ, "dspev_stub.o")
    if not newer(stub, os.path.join(STUBS, "dspev_stub.C")):
        run([CXX, "-O2", "-fPIC", "-c", os.path.join(STUBS, "dspev_stub.C"), "-o", stub])
    shim = os.path.join(OBJ, "ref_shim.o")
    if not newer(shim, os.path.join(HERE, "ref_shim.C"), os.path.join(STUBS, "mpi.h")):
        run([CXX] + FLAGS + ["-c", os.path.join(HERE, "ref_shim.C"), "-o", shim])

    exe = os.path.join(OUT, "sw4lite_ref")
    if not newer(exe, *objs, stub):
        run([CXX, "-fopenmp", "-o", exe] + objs + [stub])

    ew_weak = os.path.join(OBJ, "EW_weak.o")
    ew_o = os.path.join(OBJ, "EW.o")
    if not newer(ew_weak, ew_o):
        run(["objcopy", "--weaken-symbol=" + TSL_SYMBOL, ew_o, ew_weak])
    lib = os.path.join(OUT, "libsw4ref.so")
    libobjs = [o for o in objs if not o.endswith("/main.o") and not o.endswith("/EW.o")] + [ew_weak]
    if not newer(lib, shim, stub, *libobjs):
        run([CXX, "-shared", "-fopenmp", "-o", lib, shim] + libobjs + [stub])
    if verbose:
        print("oracle/_ref: built", exe, "and", lib)
    return True


if __name__ == "__main__":
    build()
