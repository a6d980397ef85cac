#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY (oracle/): compile the CPU restatement oracle/sw4_oracle.c into
oracle/libsw4oracle.so (git-ignored; travels to the GPU box with the snapshot)."""
import os, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))
CC = os.environ.get("SW4B200_CC", "/usr/bin/gcc")


def build(verbose=True):
    src = os.path.join(HERE, "sw4_oracle.c")
    lib = os.path.join(HERE, "libsw4oracle.so")
    deps = [src, os.path.join(HERE, "..", "sw4lite_b200", "csrc", "sbp4_tables.h")]
    if os.path.exists(lib) and all(os.path.getmtime(d) <= os.path.getmtime(lib) for d in deps):
        return lib
    cmd = [CC, "-O2", "-fopenmp", "-fPIC", "-shared", "-std=gnu99", "-ffp-contract=off", "-o", lib, src, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise SystemExit("oracle port build failed")
    if verbose:
        print("oracle: built", lib)
    return lib


if __name__ == "__main__":
    build()
