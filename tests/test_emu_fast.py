"""CPU test of the fast interior kernels' algebra and index logic: the CUDA source of
sw4lite_b200/csrc/rhs4sg_fast2.cu and rhs4sg_fast4.cu compiled by g++ through tests/emu/cuda_emu.h (one OS thread per
CUDA thread) against the oracle.  This is test infrastructure for the kernel source, not a product
path: the library itself has no CPU implementation."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest

from tests.fields import Box, random_fields, relerr
from tests.cpu_step import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
LIB = os.path.join(EMU, "libemu_fast.so")
SRC = [os.path.join(EMU, "emu_fast.cpp"), os.path.join(EMU, "cuda_emu.h"),
       os.path.join(HERE, "..", "sw4lite_b200", "csrc", "rhs4sg_fast2.cu"),
       os.path.join(HERE, "..", "sw4lite_b200", "csrc", "rhs4sg_fast4.cu"),
       os.path.join(HERE, "..", "sw4lite_b200", "csrc", "fast_common.cuh"),
       os.path.join(HERE, "..", "sw4lite_b200", "csrc", "tmem.cuh"),
       os.path.join(HERE, "..", "sw4lite_b200", "csrc", "tma.cuh"),
       os.path.join(HERE, "..", "sw4lite_b200", "csrc", "common.cuh")]
_dp = C.POINTER(C.c_double)
GEN = 2  # generation of the fast kernel under test, set per test by the fixture below


@pytest.fixture(autouse=True, params=[4, 40, 2], ids=["fast4-pairs", "fast4-padded-rows", "fast2"])
def generation(request):
    """4: rhs4sg_fast4.cu (x-pair register blocking, z state in tensor memory -- emulated here as a per-thread
    array) on a grid with even ni; 40: the same kernel on a grid with ODD ni whose rows are padded to an even pitch
    (what sw4b200_grid_create does; the pad column holds NaN here: it must never reach a result); 2: rhs4sg_fast2.cu"""
    global GEN
    GEN = request.param
    yield


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in SRC):
        subprocess.check_call(["/usr/bin/g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-pthread", "-DSW4B200_EMULATE",
                               "-ffp-contract=off", "-o", LIB, SRC[0]])
    lib = C.CDLL(LIB)
    lib.emu_rhs_fast.argtypes = [C.c_int] * 12 + [_dp] * 6 + [C.c_double] + [_dp] * 5 + [C.c_double]
    lib.emu_rhs_fast.restype = C.c_int
    lib.emu_force_general.argtypes = [C.c_int]
    return lib


def d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def even(dims):
    """the fourth generation stages rows with 16-byte bulk copies: ni (= nx + 4 ghost points) must be even; odd
    grids are served by the second generation (launch_fast4 dispatches)"""
    if GEN == 4:
        return ((dims[0] + 1) // 2 * 2,) + tuple(dims[1:])
    if GEN == 40:
        return (dims[0] // 2 * 2 + 1,) + tuple(dims[1:])
    return dims


def _pad(a, box, fill=np.nan):
    """rows of ni values -> rows of ni+1 values (row pitch of a padded block)"""
    if a is None:
        return None
    r = a.reshape(-1, box.ni)
    return np.ascontiguousarray(np.concatenate([r, np.full((r.shape[0], 1), fill)], axis=1)).ravel()


def run(emu, epi, box, klo, khi, kchunk, f, cof, out, out2=None, um=None, rho=None, fo=None, fac=0.0):
    if GEN != 40:
        rc = emu.emu_rhs_fast(GEN, epi, *box.bounds, box.ni, klo, khi, kchunk, d(f["u"]), d(f["mu"]), d(f["la"]), d(f["strx"]),
                              d(f["stry"]), d(f["strz"]), cof, d(out), d(out2), d(um), d(rho), d(fo), fac)
        assert rc == 0
        return
    assert box.ni % 2 == 1
    SENT = 7.25
    inplace = um is out
    pout = _pad(out, box, SENT); pout2 = _pad(out2, box, SENT)
    pum = pout if inplace else _pad(um, box)
    rc = emu.emu_rhs_fast(4, epi, *box.bounds, box.ni + 1, klo, khi, kchunk, d(_pad(f["u"], box)), d(_pad(f["mu"], box)),
                          d(_pad(f["la"], box)), d(np.append(f["strx"], np.nan)), d(f["stry"]), d(f["strz"]), cof, d(pout), d(pout2),
                          d(pum), d(_pad(rho, box)), d(_pad(fo, box)), fac)
    assert rc == 0
    for dst, src in ((out, pout), (out2, pout2)):
        if dst is not None:
            r = src.reshape(-1, box.ni + 1)
            assert np.all(r[:, -1] == SENT)           # the pad column is never written
            dst[:] = r[:, :-1].ravel()


def cpu_lu(box, f, h, onesided=(0,) * 6, nk=None):
    O = oracle()
    acof, ghcof, bope, _ = O.get_stencil_coefficients()
    lu = np.zeros(3 * box.npts)
    O.rhs4sg(1, box.bounds, nk if nk is not None else box.nk - 4, onesided, acof, bope, ghcof, lu, f["u"], f["mu"],
             f["la"], h, f["strx"], f["stry"], f["strz"])
    return lu


@pytest.mark.parametrize("dims,kchunk", [((45, 22, 20), 16), ((37, 13, 23), 7), ((70, 9, 9), 5)])
def test_emu_lu_matches_oracle(emu, dims, kchunk):
    box = Box(*even(dims))
    f = random_fields(box, seed=21)
    h = 0.7
    ref = cpu_lu(box, f, h)
    out = np.full(3 * box.npts, 55.0)
    run(emu, 0, box, box.kfirst + 2, box.klast - 2, kchunk, f, 1 / h ** 2, out)
    a = out.reshape(3, box.nk, box.nj, box.ni); b = ref.reshape(3, box.nk, box.nj, box.ni)
    assert relerr(a[:, 2:-2, 2:-2, 2:-2], b[:, 2:-2, 2:-2, 2:-2]) < 1e-13
    shell = np.ones((box.nk, box.nj, box.ni), dtype=bool); shell[2:-2, 2:-2, 2:-2] = False
    assert np.all(a[:, shell] == 55.0)          # nothing outside the interior is written


def test_emu_row_range_between_closures(emu):
    """rows 7..nk-6 only (the closure rows belong to another kernel)"""
    box = Box(*even((40, 20, 26)))
    nk = box.nk - 4
    f = random_fields(box, seed=22)
    ref = cpu_lu(box, f, 1.0, onesided=(0, 0, 0, 0, 1, 1)).reshape(3, box.nk, box.nj, box.ni)
    out = np.zeros(3 * box.npts)
    run(emu, 0, box, 7, nk - 6, 100, f, 1.0, out)
    a = out.reshape(3, box.nk, box.nj, box.ni)
    k0 = 7 - box.kfirst; k1 = nk - 6 - box.kfirst
    assert relerr(a[:, k0:k1 + 1, 2:-2, 2:-2], ref[:, k0:k1 + 1, 2:-2, 2:-2]) < 1e-13
    assert np.all(a[:, :k0] == 0) and np.all(a[:, k1 + 1:] == 0)


@pytest.mark.parametrize("k0,nrows", [(9, 2), (4, 2), (12, 5), (3, 1)])
def test_emu_short_launches(emu, k0, nrows):
    """a few planes only (the face rows of a z-slab run: planes next to a halo face)"""
    box = Box(*even((40, 20, 26)))
    f = random_fields(box, seed=24)
    ref = cpu_lu(box, f, 1.0).reshape(3, box.nk, box.nj, box.ni)
    out = np.zeros(3 * box.npts)
    klo = box.kfirst + k0; khi = klo + nrows - 1
    run(emu, 0, box, klo, khi, 100, f, 1.0, out)
    a = out.reshape(3, box.nk, box.nj, box.ni)
    assert relerr(a[:, k0:k0 + nrows, 2:-2, 2:-2], ref[:, k0:k0 + nrows, 2:-2, 2:-2]) < 1e-13
    assert np.all(a[:, :k0] == 0) and np.all(a[:, k0 + nrows:] == 0)


def test_emu_predictor_and_corrector_epilogues(emu):
    box = Box(*even((41, 19, 15)))
    f = random_fields(box, seed=23)
    h, dt = 0.4, 0.05
    if GEN in (4, 40):      # dense forcing is not on the fourth generation's path (launch_fast4 sends it to the second)
        f["fo"] = np.zeros_like(f["fo"])
    O = oracle()
    lu = cpu_lu(box, f, h)
    up = np.zeros(3 * box.npts)
    O.predfort(1, box.bounds, up, f["u"], f["um"], lu, f["fo"], f["rho"], dt * dt)
    out = np.zeros(3 * box.npts); out2 = np.zeros(3 * box.npts)
    run(emu, 1, box, box.kfirst + 2, box.klast - 2, 6, f, 1 / h ** 2, out, out2=out2, um=f["um"], rho=f["rho"], fo=None if GEN in (4, 40) else f["fo"], fac=dt * dt)
    inner = (slice(None), slice(2, -2), slice(2, -2), slice(2, -2))
    r4 = lambda x: x.reshape(3, box.nk, box.nj, box.ni)
    assert relerr(r4(out)[inner], r4(up)[inner]) < 1e-13
    acc = (r4(lu) + r4(f["fo"])) / f["rho"].reshape(box.nk, box.nj, box.ni)
    assert relerr(r4(out2)[inner], acc[inner]) < 1e-13
    # corrector: up + dt^4/(12 rho) (L(u) + fo), in place
    ref = f["up"].copy()
    O.corrfort(1, box.bounds, ref, lu, f["fo"], f["rho"], dt ** 4)
    out = f["up"].copy()
    run(emu, 2, box, box.kfirst + 2, box.klast - 2, 6, f, 1 / h ** 2, out, um=out, rho=f["rho"], fo=None if GEN in (4, 40) else f["fo"], fac=dt ** 4 / 12)
    assert relerr(r4(out)[inner], r4(ref)[inner]) < 1e-13


@pytest.mark.parametrize("epi", [0, 1, 2])
def test_emu_plain_tiles_take_the_body_without_stretching(emu, epi):
    """tiles on which strx = stry = 1 run the march compiled without the stretching factors (rhs4sg_fast4.cu, NOSTR):
    equal to the oracle and bit-identical to the general body on the same tiles; here 2 x 2 tiles of which the
    first column and the first row are stretched (a supergrid layer) and one tile is plain"""
    if GEN == 2:
        pytest.skip("rhs4sg_fast4.cu only")
    box = Box(*even((60, 30, 14)))
    f = random_fields(box, seed=31)
    f["strx"][10:] = 1.0
    f["stry"][9:] = 1.0
    f["fo"] = np.zeros_like(f["fo"])
    h, dt = 0.6, 0.04
    lu = cpu_lu(box, f, h)
    O = oracle()
    if epi == 0:
        ref = lu
    elif epi == 1:
        ref = np.zeros(3 * box.npts)
        O.predfort(1, box.bounds, ref, f["u"], f["um"], lu, f["fo"], f["rho"], dt * dt)
    else:
        ref = f["up"].copy()
        O.corrfort(1, box.bounds, ref, lu, f["fo"], f["rho"], dt ** 4)
    res = []
    try:
        for force in (0, 1):
            emu.emu_force_general(force)
            emu.emu_nostr_ctas()
            out = f["up"].copy() if epi == 2 else np.zeros(3 * box.npts)
            out2 = np.zeros(3 * box.npts) if epi == 1 else None
            run(emu, epi, box, box.kfirst + 2, box.klast - 2, 7, f, 1 / h ** 2, out, out2=out2,
                um=out if epi == 2 else (f["um"] if epi == 1 else None), rho=f["rho"] if epi else None,
                fac=(dt * dt if epi == 1 else dt ** 4 / 12))
            res.append((out, out2))
            # one of the 2 x 2 tiles is plain; 10 interior planes in chunks of 7 -> 2 CTAs per tile
            # (the Lu epilogue is one launch of all tiles with the general march)
            assert emu.emu_nostr_ctas() == (0 if force or epi == 0 else 2)
    finally:
        emu.emu_force_general(0)
    inner = (slice(None), slice(2, -2), slice(2, -2), slice(2, -2))
    r4 = lambda x: x.reshape(3, box.nk, box.nj, box.ni)
    assert relerr(r4(res[0][0])[inner], r4(ref)[inner]) < 1e-13
    assert np.array_equal(res[0][0], res[1][0])
    if epi == 1:
        assert np.array_equal(res[0][1], res[1][1])
