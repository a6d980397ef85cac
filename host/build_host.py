#!/usr/bin/env python3
"""Build `host/_build/sw4lite_b200`: the reference program (its own main, .in parser, set-up, sources, receivers,
error norms -- compiled from the sources WHERE THEY LIE under /root/reference/src, unmodified and never copied)
with its GPU operator layer (src/EW_cuda.C, src/device-routines.C, src/EWCuda.C) replaced by
host/EW_cuda_b200.C + libsw4b200.so.

Recipe = the reference's own CUDA configuration (Makefile.cuda:45-49,99: nvcc -x cu -dc -DSW4_CROUTINES -DSW4_CUDA
-DSW4_NONBLOCKING, object list minus the three replaced files) for sm_100a.  The two third-party pieces this image
lacks are stood in for by host/stubs/mpi.h (single rank) and host/stubs/dspev_stub.C (3x3 symmetric
eigenvalues), as for the CPU oracle.  Needs /root/reference at BUILD time only; the binary travels to the GPU box.
"""
import os, subprocess, sys, shutil
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("SW4_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "src")
OUT = os.path.join(HERE, "_build")
OBJ = os.path.join(OUT, "obj")
STUBS = os.path.join(HERE, "stubs")
NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
LIBDIR = os.path.join(ROOT, "sw4lite_b200")

# Makefile.cuda:99 minus EW_cuda, device-routines, EWCuda
OBJS = ("main EW Source rhs4sg rhs4sg_rev SuperGrid GridPointSource time_functions_cu ew-cfromfort Sarray "
        "CheckPoint Parallel_IO EW-dg MaterialData MaterialBlock Polynomial SecondOrderSection TimeSeries sacsubc "
        "curvilinear-c rhs4sgcurv rhs4sgcurv_rev").split()
FLAGS = ["-O3", "-x", "cu", "-dc", "-gencode", "arch=compute_100a,code=sm_100a", "-w", "-DSW4_CROUTINES", "-DSW4_CUDA",
         "-DSW4_NONBLOCKING", "-I", STUBS, "-I", SRC, "-I", os.path.join(SRC, "double")]
EXE = os.path.join(OUT, "sw4lite_b200")


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise SystemExit("host build failed")


def newer(target, *deps):
    return os.path.exists(target) and all(os.path.getmtime(d) <= os.path.getmtime(target) for d in deps)


def build_slab_driver(verbose=True):
    """host/_build/slab_driver: the C++ multi-GPU z-slab driver on the C-ABI (plain g++, needs neither nvcc nor the reference)"""
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(HERE, "slab_driver.C")
    exe = os.path.join(OUT, "slab_driver")
    lib = os.path.join(LIBDIR, "libsw4b200.so")
    if not os.path.exists(lib):
        raise SystemExit("host build: libsw4b200.so is not built")
    if not newer(exe, src, os.path.join(ROOT, "include", "sw4b200.h"), lib):
        run(["/usr/bin/g++", "-O2", "-std=c++17", "-o", exe, src, "-L", LIBDIR, "-lsw4b200", "-Wl,-rpath,$ORIGIN/../../sw4lite_b200", "-lpthread"])
    if verbose:
        print("host: built", exe)
    return True


def build(verbose=True):
    build_slab_driver(verbose)
    if not os.path.isdir(SRC):
        if verbose:
            print("host: %s not present, keeping the prebuilt %s" % (SRC, EXE))
        return os.path.exists(EXE)
    os.makedirs(OBJ, exist_ok=True)
    stub_h = os.path.join(STUBS, "mpi.h")

    def cc(name):
        src = os.path.join(SRC, name + ".C")
        obj = os.path.join(OBJ, name + ".o")
        if not newer(obj, src, stub_h):
            run([NVCC] + FLAGS + ["-c", src, "-o", obj])
        return obj

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(cc, OBJS))
    bind_src = os.path.join(HERE, "EW_cuda_b200.C")
    bind = os.path.join(OBJ, "EW_cuda_b200.o")
    if not newer(bind, bind_src, os.path.join(ROOT, "include", "sw4b200.h"), stub_h):
        run([NVCC] + FLAGS + ["-c", bind_src, "-o", bind])
    stub = os.path.join(OBJ, "dspev_stub.o")
    if not newer(stub, os.path.join(STUBS, "dspev_stub.C")):
        run(["/usr/bin/g++", "-O2", "-fPIC", "-c", os.path.join(STUBS, "dspev_stub.C"), "-o", stub])
    lib = os.path.join(LIBDIR, "libsw4b200.so")
    if not os.path.exists(lib):
        raise SystemExit("host build: libsw4b200.so is not built")
    if not newer(EXE, bind, stub, *objs):
        run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-o", EXE] + objs + [bind, stub, "-L", LIBDIR, "-lsw4b200",
            "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../sw4lite_b200", "-lcudart"])
    if verbose:
        print("host: built", EXE)
    return True


if __name__ == "__main__":
    build()
