#!/usr/bin/env python3
"""Build libsw4b200.so (hand-written CUDA for sm_100a + the extern "C" layer) with nvcc, in-tree.
The .so is git-ignored but travels to the GPU box with the repository snapshot."""
import os, subprocess, sys, shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsw4b200.so")
NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared", "-Xptxas=-v"]


def sources():
    out = []
    for root, _, files in os.walk(SRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".h"))]
    out.append(os.path.join(HERE, "..", "include", "sw4b200.h"))
    return out


def build(verbose=True, force=False, out=None, extra=()):
    """out / extra: a variant of the library under another name with extra nvcc flags (A/B measurements: scripts/ab_variants.sh)"""
    deps = sources()
    lib = out or LIB
    if not force and os.path.exists(lib) and all(os.path.getmtime(d) <= os.path.getmtime(lib) for d in deps):
        return lib
    cmd = [NVCC] + FLAGS + list(extra) + ["-o", lib, os.path.join(SRC, "sw4b200.cu"), "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(HERE, "build.log") if out is None else lib + ".log"
    open(log, "w").write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise SystemExit("libsw4b200.so build failed")
    if verbose:
        print("built", lib, "(ptxas log in %s)" % log)
    return lib


if __name__ == "__main__":
    # python -m sw4lite_b200.build [-f] [-o variant.so -D...]
    args = [a for a in sys.argv[1:] if a != "-f"]
    out = None
    if "-o" in args:
        n = args.index("-o")
        out = os.path.abspath(args[n + 1])
        del args[n:n + 2]
    build(force="-f" in sys.argv, out=out, extra=args)
