#!/usr/bin/env python3
"""Generate the committed golden vectors from the REFERENCE (oracle/_ref/libsw4ref.so built from
/root/reference by oracle/build_ref.py).  Run in the build container only:
    python tests/golden/make_golden.py
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim
from tests.fields import Box, random_fields

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    acof, ghcof, bope, sbop = refshim.get_stencil_coefficients()
    np.savez(os.path.join(HERE, "sbp_coefficients.npz"), acof=acof, ghcof=ghcof, bope=bope, sbop=sbop)
    # small rhs4sg case with both closures, both layouts
    dims = (14, 13, 19); seed = 21; h = 0.25; onesided = (0, 0, 0, 0, 1, 1)
    box = Box(*dims)
    out = dict(dims=np.array(dims), seed=seed, h=h, onesided=np.array(onesided))
    for corder in (1, 0):
        f = random_fields(box, seed=seed, corder=corder)
        lu = np.zeros(3 * box.npts)
        refshim.rhs4sg(corder, box.bounds, box.nk - 4, onesided, acof, bope, ghcof, lu, f["u"], f["mu"], f["la"], h,
                       f["strx"], f["stry"], f["strz"])
        out["lu_c%d" % corder] = lu
    np.savez_compressed(os.path.join(HERE, "rhs4sg_small.npz"), **out)
    # small curvilinear case: rhs4sgcurv with the free-surface closure rows, addsgd4c, freesurfcurvisg (SoA layout)
    from tests.test_gpu_curvilinear import curv_fields
    dims = (12, 11, 14); seed = 61
    box = Box(*dims)
    f = curv_fields(box, seed, 1)
    onesided = (0, 0, 0, 0, 1, 0)
    lu = np.zeros(3 * box.npts)
    refshim.rhs4sgcurv(1, box.bounds, f["u"], f["mu"], f["la"], f["met"], f["jac"], lu, onesided, acof, bope, ghcof,
                       f["strx"], f["stry"])
    up = f["up"].copy()
    refshim.addsgdc(1, 4, box.bounds, up, f["u"], f["um"], f["rho"], f["dcx"], f["dcy"], f["strx"], f["stry"], f["jac"],
                    f["cox"], f["coy"], 0.02)
    forcing = np.random.default_rng(seed + 5).uniform(-1, 1, 3 * box.ni * box.nj)
    ug = f["u"].copy()
    refshim.freesurfcurvisg(1, box.bounds, box.nk - 4, 5, ug, f["mu"], f["la"], f["met"], sbop, forcing, f["strx"], f["stry"])
    ghost = ug.reshape(3, box.nk, box.nj, box.ni)[:, 1].copy()      # plane k=0
    np.savez_compressed(os.path.join(HERE, "curvilinear_small.npz"), dims=np.array(dims), seed=seed,
                        lu=lu, sgd_update=up - f["up"], forcing=forcing, ghost=ghost)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
