"""CPU tests (-m "not gpu"): the oracle restatement (oracle/sw4_oracle.c) against the reference
itself (oracle/_ref/libsw4ref.so, when it has been built in this container) and against the
committed golden vectors generated from the reference (tests/golden/)."""
import os
import numpy as np
import pytest

from oracle import port, refshim
from tests.fields import Box, random_fields, harness_fields, relerr

TOL = 1e-13  # port vs reference: same algebra, different association
needs_ref = pytest.mark.skipif(not refshim.available(), reason="oracle/_ref not built (needs /root/reference)")
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_coefficients_match_golden():
    acof, ghcof, bope, sbop = port.get_stencil_coefficients()
    g = np.load(os.path.join(GOLD, "sbp_coefficients.npz"))
    assert np.array_equal(acof, g["acof"]) and np.array_equal(bope, g["bope"])
    assert np.array_equal(ghcof, g["ghcof"]) and np.array_equal(sbop, g["sbop"])


@needs_ref
def test_coefficients_match_reference():
    a = port.get_stencil_coefficients()
    b = refshim.get_stencil_coefficients()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)  # bit exact


def _rhs(mod, corder, box, nk, onesided, f, h):
    acof, ghcof, bope, sbop = port.get_stencil_coefficients()
    lu = np.zeros(3 * box.npts)
    mod.rhs4sg(corder, box.bounds, nk, onesided, acof, bope, ghcof, lu, f["u"], f["mu"], f["la"], h,
               f["strx"], f["stry"], f["strz"])
    return lu


@needs_ref
@pytest.mark.parametrize("corder", [1, 0])
@pytest.mark.parametrize("onesided", [(0, 0, 0, 0, 0, 0), (0, 0, 0, 0, 1, 0), (0, 0, 0, 0, 1, 1), (0, 0, 0, 0, 0, 1)])
def test_rhs4sg_vs_reference(corder, onesided):
    box = Box(21, 18, 25)            # interior 17 x 14 x 21
    nk = box.nk - 4
    f = random_fields(box, seed=7, corder=corder)
    a = _rhs(port, corder, box, nk, onesided, f, 0.37)
    b = _rhs(refshim, corder, box, nk, onesided, f, 0.37)
    assert relerr(a, b) < TOL


@needs_ref
def test_rhs4sg_harness_fields_vs_reference():
    n = 40
    box = Box(n, n, n)
    h = 1.0 / (n - 1)
    f = harness_fields(box, h)
    f.update(strx=np.ones(n), stry=np.ones(n), strz=np.ones(n))
    a = _rhs(port, 1, box, n - 4, (0, 0, 0, 0, 1, 1), f, h)
    b = _rhs(refshim, 1, box, n - 4, (0, 0, 0, 0, 1, 1), f, h)
    # smooth fields: lu = (1/h^2) * (differences that cancel), so the association order of the
    # two implementations shows up amplified by ~1/h^2; 1e-12 still holds at this size
    assert relerr(a, b) < 1e-12


@pytest.mark.parametrize("corder", [1, 0])
def test_rhs4sg_vs_golden(corder):
    g = np.load(os.path.join(GOLD, "rhs4sg_small.npz"))
    box = Box(*g["dims"])
    f = random_fields(box, seed=int(g["seed"]), corder=corder)
    a = _rhs(port, corder, box, box.nk - 4, tuple(g["onesided"]), f, float(g["h"]))
    assert relerr(a, g["lu_c%d" % corder]) < TOL


@needs_ref
@pytest.mark.parametrize("corder", [1, 0])
def test_pred_corr_dpdmt_vs_reference(corder):
    box = Box(13, 11, 9)
    f = random_fields(box, seed=3, corder=corder)
    res = {}
    for name, mod in (("port", port), ("ref", refshim)):
        up = f["up"].copy()
        mod.predfort(corder, box.bounds, up, f["u"], f["um"], f["lu"], f["fo"], f["rho"], 0.013)
        u2 = np.zeros_like(up)
        mod.dpdmtfort(corder, box.bounds, up, f["u"], f["um"], u2, 1 / 0.013)
        up2 = up.copy()
        mod.corrfort(corder, box.bounds, up2, f["lu"], f["fo"], f["rho"], 0.013 ** 2)
        res[name] = (up, u2, up2)
    for a, b in zip(res["port"], res["ref"]):
        assert relerr(a, b) < 1e-15


@needs_ref
@pytest.mark.parametrize("corder", [1, 0])
@pytest.mark.parametrize("order", [4, 6])
def test_addsgd_vs_reference(corder, order):
    box = Box(17, 15, 14)
    f = random_fields(box, seed=11, corder=corder)
    out = {}
    for name, mod in (("port", port), ("ref", refshim)):
        up = f["up"].copy()
        mod.addsgd(corder, order, box.bounds, up, f["u"], f["um"], f["rho"], f["dcx"], f["dcy"], f["dcz"],
                   f["strx"], f["stry"], f["strz"], f["cox"], f["coy"], f["coz"], 0.02)
        out[name] = up
    assert relerr(out["port"], out["ref"]) < TOL
    assert not np.array_equal(out["ref"], f["up"])


def _bc_case(box, bctype):
    """windows as EW::setup_boundary_arrays builds them (EW.C:3347-3420)"""
    b = box.bounds
    wind = np.zeros(36, dtype=np.int32)
    nb = []
    for s in range(6):
        w = list(b)
        if bctype[s] == 0:      # stress free: the boundary plane itself
            if s == 4: w[4] = w[5] = b[4] + 2
            if s == 5: w[4] = w[5] = b[5] - 2
        else:                   # two ghost layers
            lo = 2 * (s // 2)
            if s % 2 == 0: w[lo + 1] = w[lo] + 1
            else: w[lo] = w[lo + 1] - 1
        wind[6 * s:6 * s + 6] = w
        nb.append((w[1] - w[0] + 1) * (w[3] - w[2] + 1) * (w[5] - w[4] + 1))
    return wind, nb


@needs_ref
@pytest.mark.parametrize("corder", [1, 0])
@pytest.mark.parametrize("bctype", [(2, 2, 2, 2, 0, 2), (1, 1, 1, 1, 0, 0), (3, 3, 3, 3, 2, 2), (2, 2, 2, 2, 2, 2)])
def test_bcfortsg_vs_reference(corder, bctype):
    box = Box(15, 13, 12)
    f = random_fields(box, seed=5, corder=corder)
    wind, nb = _bc_case(box, bctype)
    r = np.random.default_rng(9)
    bforce = [r.uniform(-1, 1, 3 * n) for n in nb]
    _, _, _, sbop = port.get_stencil_coefficients()
    out = {}
    for name, mod in (("port", port), ("ref", refshim)):
        u = f["u"].copy()
        mod.bcfortsg(corder, box.bounds, wind, box.ni - 4, box.nj - 4, box.nk - 4, u, 0.1, bctype, sbop,
                     f["mu"], f["la"], 0.0, bforce, f["strx"], f["stry"])
        out[name] = u
    assert relerr(out["port"], out["ref"]) < 1e-14
    assert not np.array_equal(out["ref"], f["u"])
