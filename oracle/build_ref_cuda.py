#!/usr/bin/env python3
"""MEASUREMENT INFRASTRUCTURE ONLY (oracle/): build the reference's OWN CUDA program for the B200, from the sources where
they lie under /root/reference/src (never copied), output only into oracle/_ref/cuda/.

This is the comparator BASELINE.md 4.5 / SURVEY 2a name: the kernels to beat on the same box (rhs4_v2 & co.,
device-routines.C:9992-10555, launched by RHSPredCU_center / RHSCorrCU_center, EW_cuda.C:1228-1410).  Recipe = the
reference's Makefile.cuda (:45-49 flags, :99 object list): `nvcc -O3 -x cu -dc -DSW4_CROUTINES -DSW4_CUDA
-DSW4_NONBLOCKING`, with -arch=sm_100 in place of sm_60 and the same two stand-ins as the CPU oracle for the third-party
pieces this image lacks (host/stubs/mpi.h: single rank; host/stubs/dspev_stub.C).  The binary is used for TIMING ONLY
(bench.py --impl reference-cuda): its bcfortsg<> kernel covers one k-plane per side window (SURVEY 8a trap 11), so it is not
a parity oracle.

Product: oracle/_ref/cuda/sw4lite_ref_cuda (git-ignored, travels to the GPU box with the snapshot).
"""
import os, subprocess, sys, shutil
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SW4_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "src")
OUT = os.path.join(HERE, "_ref", "cuda")
OBJ = os.path.join(OUT, "obj")
STUBS = os.path.join(os.path.dirname(HERE), "host", "stubs")
NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
EXE = os.path.join(OUT, "sw4lite_ref_cuda")

# Makefile.cuda:99
OBJS = ("main EW Source rhs4sg rhs4sg_rev SuperGrid GridPointSource time_functions_cu ew-cfromfort EW_cuda Sarray "
        "device-routines EWCuda CheckPoint Parallel_IO EW-dg MaterialData MaterialBlock Polynomial SecondOrderSection "
        "TimeSeries sacsubc curvilinear-c rhs4sgcurv rhs4sgcurv_rev").split()
ARCH = ["-arch=sm_100"]
FLAGS = ["-O3", "-x", "cu", "-dc", "-w", "-DSW4_CROUTINES", "-DSW4_CUDA", "-DSW4_NONBLOCKING", "-I", STUBS, "-I", SRC,
         "-I", os.path.join(SRC, "double")] + ARCH


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout[-4000:] + r.stderr[-4000:])
        raise SystemExit("oracle/_ref/cuda build failed")


def newer(target, *deps):
    return os.path.exists(target) and all(os.path.getmtime(d) <= os.path.getmtime(target) for d in deps)


def build(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print("oracle/_ref/cuda: %s not present, keeping the prebuilt %s" % (SRC, EXE))
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
    stub = os.path.join(OBJ, "dspev_stub.o")
    if not newer(stub, os.path.join(STUBS, "dspev_stub.C")):
        run(["/usr/bin/g++", "-O2", "-fPIC", "-c", os.path.join(STUBS, "dspev_stub.C"), "-o", stub])
    if not newer(EXE, stub, *objs):
        run([NVCC] + ARCH + ["-o", EXE] + objs + [stub, "-lcudart"])
    if verbose:
        print("oracle/_ref/cuda: built", EXE)
    return True


if __name__ == "__main__":
    build()
