"""-m gpu parity tests: every CUDA operator of libsw4b200.so, called through the C-ABI, against
the oracle (the reference CPU kernels when oracle/_ref is present, else the pinned restatement)
on the same seeded inputs.  Tolerance: 1e-12 relative (max|a-b|/max|b|), fp64, as north_star states;
elementwise updates agree to 1e-15."""
import ctypes as C
import numpy as np
import pytest

from tests.fields import Box, random_fields, harness_fields, relerr
from tests.gpuutil import Dev, oracle, ints
from tests.test_oracle import _bc_case

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def dev():
    return Dev()


def _coef():
    return oracle().get_stencil_coefficients()


def gpu_rhs4sg(dev, corder, box, nk, onesided, f, h, lu0=None):
    d = {k: dev.put(f[k]) for k in ("u", "mu", "la", "strx", "stry", "strz")}
    lu = dev.put(lu0 if lu0 is not None else np.zeros(3 * box.npts))
    dev.check(dev.lib.sw4b200_rhs4sg(corder, *box.bounds, nk, ints(onesided), dev.p(lu), dev.p(d["u"]), dev.p(d["mu"]),
                                     dev.p(d["la"]), h, dev.p(d["strx"]), dev.p(d["stry"]), dev.p(d["strz"]), None))
    return dev.get(lu)


def cpu_rhs4sg(corder, box, nk, onesided, f, h):
    acof, ghcof, bope, sbop = _coef()
    lu = np.zeros(3 * box.npts)
    oracle().rhs4sg(corder, box.bounds, nk, onesided, acof, bope, ghcof, lu, f["u"], f["mu"], f["la"], h,
                    f["strx"], f["stry"], f["strz"])
    return lu


@pytest.mark.parametrize("corder", [1, 0])
@pytest.mark.parametrize("onesided", [(0, 0, 0, 0, 0, 0), (0, 0, 0, 0, 1, 0), (0, 0, 0, 0, 1, 1), (0, 0, 0, 0, 0, 1)])
def test_rhs4sg_random(dev, corder, onesided):
    box = Box(45, 38, 29)   # ragged: not a multiple of any tile size
    nk = box.nk - 4
    f = random_fields(box, seed=7, corder=corder)
    a = gpu_rhs4sg(dev, corder, box, nk, onesided, f, 0.37)
    b = cpu_rhs4sg(corder, box, nk, onesided, f, 0.37)
    assert relerr(a, b) < TOL


@pytest.mark.parametrize("dims", [(5, 5, 5), (6, 9, 5), (33, 5, 17), (37, 70, 19), (68, 13, 40)])
def test_rhs4sg_edge_sizes(dev, dims):
    """minimum block (one interior point), sizes around tile boundaries; no closures possible below 12 planes"""
    box = Box(*dims)
    f = random_fields(box, seed=3, corder=1)
    onesided = (0, 0, 0, 0, 1, 1) if box.nk - 4 >= 12 else (0, 0, 0, 0, 0, 0)
    a = gpu_rhs4sg(dev, 1, box, box.nk - 4, onesided, f, 1.0)
    b = cpu_rhs4sg(1, box, box.nk - 4, onesided, f, 1.0)
    assert relerr(a, b) < TOL


@pytest.mark.parametrize("dims", [(32, 16, 14), (34, 18, 20), (62, 30, 15), (64, 32, 16), (96, 47, 21), (130, 33, 12), (30, 16, 12)])
@pytest.mark.parametrize("onesided", [(0, 0, 0, 0, 0, 0), (0, 0, 0, 0, 1, 1)])
def test_rhs4sg_tma_path_sizes(dev, dims, onesided):
    """even ni (= nx + 4 ghost points): the TMA-staged fourth-generation interior kernel (x-pairs, 16-byte accesses,
    tensor-memory z state); sizes at and around its 32x16 tile and its 36x20 TMA box ((30,16,12): narrower than a
    box, served by the cp.async kernel)"""
    box = Box(*dims)
    f = random_fields(box, seed=11, corder=1)
    os_ = onesided if box.nk - 4 >= 12 else (0, 0, 0, 0, 0, 0)
    a = gpu_rhs4sg(dev, 1, box, box.nk - 4, os_, f, 0.7)
    b = cpu_rhs4sg(1, box, box.nk - 4, os_, f, 0.7)
    assert relerr(a, b) < TOL


def test_rhs4sg_leaves_ghost_points_untouched_even_ni(dev):
    box = Box(40, 21, 15)
    f = random_fields(box, seed=4, corder=1)
    lu0 = np.full(3 * box.npts, 123.0)
    a = gpu_rhs4sg(dev, 1, box, box.nk - 4, (0,) * 6, f, 1.0, lu0=lu0).reshape(3, box.nk, box.nj, box.ni)
    inner = np.zeros((box.nk, box.nj, box.ni), dtype=bool)
    inner[2:-2, 2:-2, 2:-2] = True
    assert np.all(a[:, ~inner] == 123.0) and np.all(a[:, inner] != 123.0)


def test_rhs4sg_leaves_ghost_points_untouched(dev):
    box = Box(20, 17, 15)
    f = random_fields(box, seed=4, corder=1)
    lu0 = np.full(3 * box.npts, 123.0)
    a = gpu_rhs4sg(dev, 1, box, box.nk - 4, (0,) * 6, f, 1.0, lu0=lu0).reshape(3, box.nk, box.nj, box.ni)
    inner = np.zeros((box.nk, box.nj, box.ni), dtype=bool)
    inner[2:-2, 2:-2, 2:-2] = True
    assert np.all(a[:, ~inner] == 123.0) and np.all(a[:, inner] != 123.0)


def test_rhs4sg_harness_fields_128(dev):
    """config 2 fields (tests/testil) at 128^3: smooth data, cancellation amplified by 1/h^2"""
    n = 128
    box = Box(n, n, n)
    h = 1.0 / (n - 1)
    f = harness_fields(box, h)
    f.update(strx=np.ones(n), stry=np.ones(n), strz=np.ones(n))
    a = gpu_rhs4sg(dev, 1, box, n - 4, (0, 0, 0, 0, 1, 1), f, h)
    b = cpu_rhs4sg(1, box, n - 4, (0, 0, 0, 0, 1, 1), f, h)
    # scale: the operator's natural magnitude (1/h^2)*|coef|*|u| rather than the cancelled result
    scale = np.max(np.abs(f["mu"]) * 2 + np.abs(f["la"])) * np.max(np.abs(f["u"])) / h ** 2
    assert np.max(np.abs(a - b)) / scale < TOL
    assert relerr(a, b) < 1e-10


def test_rhs4sg_linearity(dev):
    """size-independent property: L(a u1 + b u2) = a L(u1) + b L(u2)"""
    box = Box(70, 66, 40)
    f1 = random_fields(box, seed=1)
    f2 = random_fields(box, seed=2)
    f2.update({k: f1[k] for k in ("mu", "la", "strx", "stry", "strz")})
    os_ = (0, 0, 0, 0, 1, 1)
    l1 = gpu_rhs4sg(dev, 1, box, box.nk - 4, os_, f1, 0.5)
    l2 = gpu_rhs4sg(dev, 1, box, box.nk - 4, os_, f2, 0.5)
    f3 = dict(f1); f3["u"] = 0.3 * f1["u"] - 1.7 * f2["u"]
    l3 = gpu_rhs4sg(dev, 1, box, box.nk - 4, os_, f3, 0.5)
    assert relerr(l3, 0.3 * l1 - 1.7 * l2) < 1e-12


@pytest.mark.parametrize("corder", [1, 0])
def test_pred_corr_dpdmt(dev, corder):
    box = Box(23, 19, 17)
    f = random_fields(box, seed=3, corder=corder)
    O = oracle()
    up = f["up"].copy()
    O.predfort(corder, box.bounds, up, f["u"], f["um"], f["lu"], f["fo"], f["rho"], 0.013)
    u2 = np.zeros_like(up)
    O.dpdmtfort(corder, box.bounds, up, f["u"], f["um"], u2, 1 / 0.013)
    up2 = up.copy()
    O.corrfort(corder, box.bounds, up2, f["lu"], f["fo"], f["rho"], 0.013 ** 2)
    d = {k: dev.put(f[k]) for k in ("u", "um", "lu", "fo", "rho")}
    gup = dev.put(f["up"])
    dev.check(dev.lib.sw4b200_predfort(corder, *box.bounds, dev.p(gup), dev.p(d["u"]), dev.p(d["um"]), dev.p(d["lu"]),
                                       dev.p(d["fo"]), dev.p(d["rho"]), 0.013, None))
    assert relerr(dev.get(gup), up) < 1e-15
    gup = dev.put(up)                 # identical inputs for the next two (dpdmt cancels, amplifying input ulps)
    gu2 = dev.zeros(3 * box.npts)
    dev.check(dev.lib.sw4b200_dpdmtfort(*box.bounds, dev.p(gup), dev.p(d["u"]), dev.p(d["um"]), dev.p(gu2), 1 / 0.013, None))
    assert relerr(dev.get(gu2), u2) < 1e-15
    dev.check(dev.lib.sw4b200_corrfort(corder, *box.bounds, dev.p(gup), dev.p(d["lu"]), dev.p(d["fo"]), dev.p(d["rho"]),
                                       0.013 ** 2, None))
    assert relerr(dev.get(gup), up2) < 1e-15


@pytest.mark.parametrize("corder", [1, 0])
@pytest.mark.parametrize("order", [4, 6])
def test_addsgd(dev, corder, order):
    box = Box(37, 21, 18)
    f = random_fields(box, seed=11, corder=corder)
    up = f["up"].copy()
    oracle().addsgd(corder, order, box.bounds, up, f["u"], f["um"], f["rho"], f["dcx"], f["dcy"], f["dcz"],
                    f["strx"], f["stry"], f["strz"], f["cox"], f["coy"], f["coz"], 0.02)
    names = ("u", "um", "rho", "dcx", "dcy", "dcz", "strx", "stry", "strz", "cox", "coy", "coz")
    d = {k: dev.put(f[k]) for k in names}
    gup = dev.put(f["up"])
    dev.check(dev.lib.sw4b200_addsgd(corder, order, *box.bounds, dev.p(gup), *[dev.p(d[k]) for k in names], 0.02, None))
    assert relerr(dev.get(gup), up) < TOL


@pytest.mark.parametrize("corder", [1, 0])
@pytest.mark.parametrize("bctype", [(2, 2, 2, 2, 0, 2), (1, 1, 1, 1, 0, 0), (3, 3, 3, 3, 2, 2), (2, 2, 2, 2, 2, 2)])
def test_bcfortsg(dev, corder, bctype):
    box = Box(35, 13, 12)
    f = random_fields(box, seed=5, corder=corder)
    wind, nb = _bc_case(box, bctype)
    r = np.random.default_rng(9)
    bforce = [r.uniform(-1, 1, 3 * n) for n in nb]
    _, _, _, sbop = _coef()
    u = f["u"].copy()
    oracle().bcfortsg(corder, box.bounds, wind, box.ni - 4, box.nj - 4, box.nk - 4, u, 0.1, bctype, sbop,
                      f["mu"], f["la"], 0.0, bforce, f["strx"], f["stry"])
    gu = dev.put(f["u"])
    d = {k: dev.put(f[k]) for k in ("mu", "la", "strx", "stry")}
    gbf = [dev.put(b) for b in bforce]
    ptrs = (C.c_void_p * 6)(*[g.data_ptr() for g in gbf])
    dev.check(dev.lib.sw4b200_bcfortsg(corder, *box.bounds, ints(wind), box.ni - 4, box.nj - 4, box.nk - 4, dev.p(gu),
                                       0.1, ints(bctype), dev.p(d["mu"]), dev.p(d["la"]), ptrs, dev.p(d["strx"]),
                                       dev.p(d["stry"]), None))
    assert relerr(dev.get(gu), u) < 1e-14


def cpu_pred(corder, box, nk, onesided, f, h, dt, dense_fo):
    lu = cpu_rhs4sg(corder, box, nk, onesided, f, h)
    up = np.zeros_like(lu)
    fo = f["fo"] if dense_fo else np.zeros_like(lu)
    oracle().predfort(corder, box.bounds, up, f["u"], f["um"], lu, fo, f["rho"], dt * dt)
    return up


@pytest.mark.parametrize("corder", [1, 0])
@pytest.mark.parametrize("dense_fo", [False, True])
def test_fused_predictor(dev, corder, dense_fo):
    box = Box(41, 37, 30)
    nk = box.nk - 4
    os_ = (0, 0, 0, 0, 1, 1)
    f = random_fields(box, seed=17, corder=corder)
    h, dt = 0.4, 0.05
    ref = cpu_pred(corder, box, nk, os_, f, h, dt, dense_fo)
    names = ("u", "um", "mu", "la", "rho")
    d = {k: dev.put(f[k]) for k in names + ("fo", "strx", "stry", "strz")}
    gup = dev.put(np.full(3 * box.npts, 7.0))
    dev.check(dev.lib.sw4b200_rhs4_pred(corder, *box.bounds, nk, ints(os_), dev.p(gup), *[dev.p(d[k]) for k in names],
                                        dev.p(d["fo"]) if dense_fo else None, dev.p(d["strx"]), dev.p(d["stry"]),
                                        dev.p(d["strz"]), h, dt, None))
    assert relerr(dev.get(gup), ref) < TOL


@pytest.mark.parametrize("corder", [1, 0])
@pytest.mark.parametrize("sg_order", [0, 4, 6])
def test_fused_corrector(dev, corder, sg_order):
    """dpdmt + rhs4sg + corrfort + addsgd in one call, against the oracle's unfused sequence"""
    box = Box(41, 37, 30)
    nk = box.nk - 4
    os_ = (0, 0, 0, 0, 1, 1)
    f = random_fields(box, seed=19, corder=corder)
    h, dt, beta = 0.4, 0.05, 0.02
    O = oracle()
    uacc = np.zeros(3 * box.npts)
    O.dpdmtfort(corder, box.bounds, f["up"], f["u"], f["um"], uacc, 1 / (dt * dt))
    fa = dict(f); fa["u"] = uacc
    lu = cpu_rhs4sg(corder, box, nk, os_, fa, h)
    ref = f["up"].copy()
    O.corrfort(corder, box.bounds, ref, lu, f["fo"], f["rho"], dt ** 4)
    if sg_order:
        O.addsgd(corder, sg_order, box.bounds, ref, f["u"], f["um"], f["rho"], f["dcx"], f["dcy"], f["dcz"],
                 f["strx"], f["stry"], f["strz"], f["cox"], f["coy"], f["coz"], beta)
    names = ("up", "u", "um", "mu", "la", "rho", "fo", "strx", "stry", "strz", "dcx", "dcy", "dcz", "cox", "coy", "coz")
    d = {k: dev.put(f[k]) for k in names}
    gout = dev.zeros(3 * box.npts)
    dev.check(dev.lib.sw4b200_rhs4_corr(corder, *box.bounds, nk, ints(os_), dev.p(gout), *[dev.p(d[k]) for k in names],
                                        beta, sg_order, h, dt, None))
    assert relerr(dev.get(gout), ref) < TOL


def test_host_entry_point(dev):
    box = Box(40, 30, 20)
    f = random_fields(box, seed=23)
    os_ = (0, 0, 0, 0, 1, 0)
    ref = cpu_rhs4sg(1, box, box.nk - 4, os_, f, 0.3)
    lu = np.zeros(3 * box.npts)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    dev.check(dev.lib.sw4b200_rhs4sg_host(1, *box.bounds, box.nk - 4, ints(os_), dp(lu), dp(f["u"]), dp(f["mu"]),
                                          dp(f["la"]), 0.3, dp(f["strx"]), dp(f["stry"]), dp(f["strz"])))
    assert relerr(lu, ref) < TOL


def test_point_forces_and_gather(dev):
    box = Box(12, 11, 10)
    f = random_fields(box, seed=29)
    idx = np.array([5, 77, 300, 1000], dtype=np.int64)
    fv = np.random.default_rng(1).uniform(-1, 1, 12)
    ref = f["up"].copy().reshape(3, -1)
    for n, p in enumerate(idx):
        ref[:, p] += 0.25 / f["rho"][p] * fv[3 * n:3 * n + 3]
    gup = dev.put(f["up"]); grho = dev.put(f["rho"]); gidx = dev.put(idx); gf = dev.put(fv)
    dev.check(dev.lib.sw4b200_add_point_forces(1, box.npts, dev.p(gup), dev.p(grho), 4, dev.p(gidx), dev.p(gf), 0.25, None))
    assert relerr(dev.get(gup), ref.ravel()) < 1e-15
    out = dev.zeros(12)
    dev.check(dev.lib.sw4b200_gather_points(1, box.npts, dev.p(gup), 4, dev.p(gidx), dev.p(out), None))
    assert np.array_equal(dev.get(out).reshape(4, 3), ref[:, idx].T)
