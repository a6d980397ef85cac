"""CPU tests of the curvilinear part of the oracle (oracle/curv_port.py, numpy): pinned against the reference
build (oracle/_ref, when present) and against golden vectors generated from the reference
(tests/golden/curvilinear_small.npz, tests/golden/make_golden.py)."""
import os
import numpy as np
import pytest

from oracle import refshim, curv_port, port
from tests.fields import Box, relerr, pack3
from tests.test_gpu_curvilinear import curv_fields

GOLD = os.path.join(os.path.dirname(__file__), "golden", "curvilinear_small.npz")
needs_ref = pytest.mark.skipif(not refshim.available(), reason="oracle/_ref/libsw4ref.so not present")


@needs_ref
@pytest.mark.parametrize("corder", [1, 0])
@pytest.mark.parametrize("top", [0, 1])
def test_rhs4sgcurv_vs_reference(corder, top):
    box = Box(15, 13, 17)
    f = curv_fields(box, 71, corder)
    acof, ghcof, bope, _ = refshim.get_stencil_coefficients()
    onesided = (0, 0, 0, 0, top, 0)
    ref = np.full(3 * box.npts, 3.0); out = np.full(3 * box.npts, 3.0)
    args = (f["u"], f["mu"], f["la"], f["met"], f["jac"])
    refshim.rhs4sgcurv(corder, box.bounds, *args, ref, onesided, acof, bope, ghcof, f["strx"], f["stry"])
    curv_port.rhs4sgcurv(corder, box.bounds, *args, out, onesided, acof, bope, ghcof, f["strx"], f["stry"])
    assert relerr(out, ref) < 1e-13


@needs_ref
@pytest.mark.parametrize("corder", [1, 0])
@pytest.mark.parametrize("order", [4, 6])
def test_addsgdc_vs_reference(corder, order):
    box = Box(13, 16, 11)
    f = curv_fields(box, 72, corder)
    ref = f["up"].copy(); out = f["up"].copy()
    args = (f["u"], f["um"], f["rho"], f["dcx"], f["dcy"], f["strx"], f["stry"], f["jac"], f["cox"], f["coy"], 0.017)
    refshim.addsgdc(corder, order, box.bounds, ref, *args)
    curv_port.addsgdc(corder, order, box.bounds, out, *args)
    assert relerr(out - f["up"], ref - f["up"]) < 1e-12


@needs_ref
@pytest.mark.parametrize("corder", [1, 0])
def test_freesurfcurvisg_vs_reference(corder):
    box = Box(14, 12, 10)
    f = curv_fields(box, 73, corder)
    forcing = np.random.default_rng(2).uniform(-1, 1, 3 * box.ni * box.nj)
    _, _, _, sbop = refshim.get_stencil_coefficients()
    ref = f["u"].copy(); out = f["u"].copy()
    refshim.freesurfcurvisg(corder, box.bounds, box.nk - 4, 5, ref, f["mu"], f["la"], f["met"], sbop, forcing, f["strx"], f["stry"])
    curv_port.freesurfcurvisg(corder, box.bounds, box.nk - 4, 5, out, f["mu"], f["la"], f["met"], sbop, forcing, f["strx"], f["stry"])
    assert (ref != f["u"]).sum() == 3 * (box.ni - 4) * (box.nj - 4)
    assert relerr(out, ref) < 1e-13


def test_curvilinear_oracle_vs_golden():
    g = np.load(GOLD)
    box = Box(*[int(x) for x in g["dims"]])
    f = curv_fields(box, int(g["seed"]), 1)
    acof, ghcof, bope, sbop = port.get_stencil_coefficients()
    lu = np.zeros(3 * box.npts)
    curv_port.rhs4sgcurv(1, box.bounds, f["u"], f["mu"], f["la"], f["met"], f["jac"], lu, (0, 0, 0, 0, 1, 0), acof, bope, ghcof,
                         f["strx"], f["stry"])
    assert relerr(lu, g["lu"]) < 1e-13
    up = f["up"].copy()
    curv_port.addsgdc(1, 4, box.bounds, up, f["u"], f["um"], f["rho"], f["dcx"], f["dcy"], f["strx"], f["stry"], f["jac"],
                      f["cox"], f["coy"], 0.02)
    assert relerr(up - f["up"], g["sgd_update"]) < 1e-12
    ug = f["u"].copy()
    curv_port.freesurfcurvisg(1, box.bounds, box.nk - 4, 5, ug, f["mu"], f["la"], f["met"], sbop, g["forcing"], f["strx"], f["stry"])
    assert relerr(ug.reshape(3, box.nk, box.nj, box.ni)[:, 1], g["ghost"]) < 1e-13


def test_enforce_cart_topo_roundtrip():
    """interface injection: afterwards the two grids agree on the 5 shared planes"""
    ni, nj = 9, 8
    cart = Box(ni, nj, 12); curv = Box(ni, nj, 9)
    r = np.random.default_rng(1)
    uc = pack3([r.uniform(-1, 1, (cart.nk, nj, ni)) for _ in range(3)], 1)
    ut = pack3([r.uniform(-1, 1, (curv.nk, nj, ni)) for _ in range(3)], 1)
    curv_port.enforce_cart_topo(1, uc, cart.bounds, ut, curv.bounds)
    C = uc.reshape(3, cart.nk, nj, ni); T = ut.reshape(3, curv.nk, nj, ni)
    for q in range(5):
        assert np.array_equal(C[:, q], T[:, curv.nk - 5 + q])
