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
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
