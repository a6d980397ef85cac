"""-m gpu parity tests of the curvilinear-grid operators (rhs4sgcurv, addsgd4c/6c, freesurfcurvisg,
enforceCartTopo) through the C-ABI against the reference's own CPU kernels (oracle/_ref), and whole time steps
of the reference's topography regression (pytest/reference/topo/curvilinear.in: Gaussian hill, curvilinear grid
on top of a Cartesian grid, supergrid, moment source, 3 stations) against the reference's EW object.
Tolerance: 1e-12 relative (max|a-b|/max|b|) per operator application and per step, fp64."""
import os
import numpy as np
import pytest

from oracle import refshim
from tests.fields import Box, random_fields, pack3, relerr
from tests.gpuutil import Dev, ints

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refshim.available(), reason="oracle/_ref/libsw4ref.so not present")]
TOL = 1e-12
INPUTS = os.path.join(os.path.dirname(__file__), "golden", "inputs")


@pytest.fixture(scope="module")
def dev():
    return Dev()


def curv_fields(box, seed, corder):
    """random fields + a metric with the magnitudes EW::metric produces (curvilinear-c.C:159-164):
    met1 ~ sqrt(z_r), met2,met3 ~ slopes/sqrt(z_r), met4 ~ h/sqrt(z_r), jac ~ h^2 z_r"""
    f = random_fields(box, seed=seed, corder=corder)
    r = np.random.default_rng(seed + 1000)
    shp = (box.nk, box.nj, box.ni)
    m = [r.uniform(0.8, 1.2, shp), r.uniform(-0.3, 0.3, shp), r.uniform(-0.3, 0.3, shp), r.uniform(0.8, 1.2, shp)]
    if corder:
        f["met"] = np.ascontiguousarray(np.stack([c.ravel() for c in m]).ravel())
    else:
        f["met"] = np.ascontiguousarray(np.stack([c.ravel() for c in m], axis=1).ravel())
    f["jac"] = r.uniform(0.7, 1.4, shp).ravel()
    return f


@pytest.mark.parametrize("corder", [1, 0])
@pytest.mark.parametrize("top", [0, 1])
def test_rhs4sgcurv_random(dev, corder, top):
    box = Box(37, 30, 21)
    f = curv_fields(box, 31, corder)
    onesided = (0, 0, 0, 0, top, 0)
    acof, ghcof, bope, _ = refshim.get_stencil_coefficients()
    ref = np.zeros(3 * box.npts)
    refshim.rhs4sgcurv(corder, box.bounds, f["u"], f["mu"], f["la"], f["met"], f["jac"], ref, onesided, acof, bope, ghcof,
                       f["strx"], f["stry"])
    d = {k: dev.put(f[k]) for k in ("u", "mu", "la", "met", "jac", "strx", "stry")}
    lu = dev.zeros(3 * box.npts)
    dev.check(dev.lib.sw4b200_rhs4sgcurv(corder, *box.bounds, dev.p(d["u"]), dev.p(d["mu"]), dev.p(d["la"]), dev.p(d["met"]),
                                         dev.p(d["jac"]), dev.p(lu), ints(onesided), dev.p(d["strx"]), dev.p(d["stry"]), None))
    out = dev.get(lu)
    assert np.abs(ref).max() > 0
    assert relerr(out, ref) < TOL


@pytest.mark.parametrize("dims", [(5, 5, 5), (33, 6, 13), (9, 37, 15)])
def test_rhs4sgcurv_edge_sizes(dev, dims):
    box = Box(*dims)
    f = curv_fields(box, 32, 1)
    top = 1 if box.klast >= 8 else 0
    onesided = (0, 0, 0, 0, top, 0)
    acof, ghcof, bope, _ = refshim.get_stencil_coefficients()
    ref = np.full(3 * box.npts, 7.0)
    refshim.rhs4sgcurv(1, box.bounds, f["u"], f["mu"], f["la"], f["met"], f["jac"], ref, onesided, acof, bope, ghcof,
                       f["strx"], f["stry"])
    d = {k: dev.put(f[k]) for k in ("u", "mu", "la", "met", "jac", "strx", "stry")}
    lu = dev.put(np.full(3 * box.npts, 7.0))
    dev.check(dev.lib.sw4b200_rhs4sgcurv(1, *box.bounds, dev.p(d["u"]), dev.p(d["mu"]), dev.p(d["la"]), dev.p(d["met"]),
                                         dev.p(d["jac"]), dev.p(lu), ints(onesided), dev.p(d["strx"]), dev.p(d["stry"]), None))
    out = dev.get(lu)
    assert relerr(out, ref) < TOL
    # ghost points are left alone, as in the reference
    a = out.reshape(3, box.nk, box.nj, box.ni)
    shell = np.ones((box.nk, box.nj, box.ni), dtype=bool); shell[2:-2, 2:-2, 2:-2] = False
    assert np.all(a[:, shell] == 7.0)


@pytest.mark.parametrize("corder", [1, 0])
@pytest.mark.parametrize("order", [4, 6])
def test_addsgdc(dev, corder, order):
    box = Box(29, 33, 17)
    f = curv_fields(box, 41, corder)
    beta = 0.013
    ref = f["up"].copy()
    refshim.addsgdc(corder, order, box.bounds, ref, f["u"], f["um"], f["rho"], f["dcx"], f["dcy"], f["strx"], f["stry"],
                    f["jac"], f["cox"], f["coy"], beta)
    d = {k: dev.put(f[k]) for k in ("up", "u", "um", "rho", "dcx", "dcy", "strx", "stry", "jac", "cox", "coy")}
    dev.check(dev.lib.sw4b200_addsgdc(corder, order, *box.bounds, dev.p(d["up"]), dev.p(d["u"]), dev.p(d["um"]), dev.p(d["rho"]),
                                      dev.p(d["dcx"]), dev.p(d["dcy"]), dev.p(d["strx"]), dev.p(d["stry"]), dev.p(d["jac"]),
                                      dev.p(d["cox"]), dev.p(d["coy"]), beta, None))
    out = dev.get(d["up"])
    assert np.abs(ref - f["up"]).max() > 0
    assert relerr(out - f["up"], ref - f["up"]) < TOL      # the update itself, not up + update
    assert relerr(out, ref) < 1e-15 * 50


@pytest.mark.parametrize("corder", [1, 0])
def test_freesurfcurvisg(dev, corder):
    box = Box(31, 26, 12)
    nz = box.nk - 4
    f = curv_fields(box, 51, corder)
    r = np.random.default_rng(9)
    forcing = r.uniform(-1, 1, 3 * box.ni * box.nj)
    _, _, _, sbop = refshim.get_stencil_coefficients()
    ref = f["u"].copy()
    refshim.freesurfcurvisg(corder, box.bounds, nz, 5, ref, f["mu"], f["la"], f["met"], sbop, forcing, f["strx"], f["stry"])
    d = {k: dev.put(f[k]) for k in ("u", "mu", "la", "met", "strx", "stry")}
    dfo = dev.put(forcing)
    dev.check(dev.lib.sw4b200_freesurfcurvisg(corder, *box.bounds, nz, 5, dev.p(d["u"]), dev.p(d["mu"]), dev.p(d["la"]),
                                              dev.p(d["met"]), dev.p(dfo), dev.p(d["strx"]), dev.p(d["stry"]), None))
    out = dev.get(d["u"])
    changed = ref != f["u"]
    assert changed.sum() == 3 * (box.ni - 4) * (box.nj - 4)
    assert np.array_equal(out[~changed], f["u"][~changed])
    assert relerr(out[changed], ref[changed]) < TOL


def test_enforce_cart_topo(dev):
    """EW::enforceCartTopo (EW.C:3504-3531) restated in numpy as the checker"""
    ni, nj = 14, 11
    cart = Box(ni, nj, 13, kfirst=-1)
    curv = Box(ni, nj, 10, kfirst=-1)
    r = np.random.default_rng(3)
    for corder in (1, 0):
        uc = [r.uniform(-1, 1, (cart.nk, nj, ni)) for _ in range(3)]
        ut = [r.uniform(-1, 1, (curv.nk, nj, ni)) for _ in range(3)]
        rc = [a.copy() for a in uc]; rt = [a.copy() for a in ut]
        for c in range(3):
            for q in range(2):
                rc[c][q] = rt[c][curv.nk - 1 - 4 + q]
            for q in range(3):
                rt[c][curv.nk - 1 - q] = rc[c][4 - q]
        duc = dev.put(pack3(uc, corder)); dut = dev.put(pack3(ut, corder))
        dev.check(dev.lib.sw4b200_enforce_cart_topo(corder, dev.p(duc), *cart.bounds, dev.p(dut), curv.kfirst, curv.klast, None))
        assert np.array_equal(dev.get(duc), pack3(rc, corder))
        assert np.array_equal(dev.get(dut), pack3(rt, corder))


# ---------------------------------------------------------------------------------------------- whole steps
def blocks_from_reference(ew):
    from sw4lite_b200.solver import GridBlock, GridStack
    blocks = []
    for g, G in enumerate(ew.grids):
        curv = bool(ew.topo) and g == ew.ngrids - 1
        blk = GridBlock(ew.corder, G.bounds, (G.nx, G.ny, G.nz), G.h, ew.dt, G.onesided, G.bctype, G.wind,
                        sg_order=ew.sgorder if ew.usesg else 0, beta=ew.beta if ew.usesg else 0.0, curvilinear=curv)
        names = ["mu", "lambda", "rho", "strx", "stry", "dcx", "dcy", "cox", "coy"]
        names += ["metric", "jac"] if curv else ["strz", "dcz", "coz"]
        for name in names:
            blk.upload(name, ew.array(name, g))
        blocks.append(blk)
    return GridStack(blocks, ncart=ew.ncart)


def _golden_station(name):
    """samples of a USGS-format station file of the reference (pytest/reference/topo/curvilinear-output/)"""
    rows = [l.split() for l in open(os.path.join(os.path.dirname(__file__), "golden", "curvilinear-output", name)) if not l.startswith("#")]
    return np.array(rows, dtype=np.float64)


def test_topography_run_matches_reference(tmp_path):
    """config 4 (small): pytest/reference/topo/curvilinear.in stepped side by side with the reference for the
    whole run (63 steps); the three stations must reproduce the reference's golden station files"""
    from tests.test_gpu_step import SourceMap
    ew = refshim.RefEW(os.path.join(INPUTS, "curvilinear.in"), str(tmp_path))
    assert ew.topo == 1 and ew.ngrids == 2 and ew.ncart == 1 and ew.corder == 1
    stack = blocks_from_reference(ew)
    srcs = [SourceMap(ew, g) for g in range(ew.ngrids)]
    for g, blk in enumerate(stack.blocks):
        if len(srcs[g].points):
            blk.set_source_points(srcs[g].points)
    assert sum(len(s.points) for s in srcs) > 0
    recs, _ = ew.receivers()          # (grid, i, j, k) of sta01..sta03
    assert len(recs) == 3
    for g, blk in enumerate(stack.blocks):
        pts = [r[1:4] for r in recs if r[0] == g]
        if pts:
            blk.set_receiver_points(np.array(pts, dtype=np.int32))
    t = ew.tstart
    worst = 0.0
    traces = [[np.zeros(3)] for _ in recs]
    for step in range(ew.nsteps):
        fa = ew.eval_forces(t, False); fta = ew.eval_forces(t, True)
        f = [s.reduce(fa) for s in srcs]; ftt = [s.reduce(fta) for s in srcs]
        ew.step()
        rec = stack.step(f, ftt, record=True)
        t += ew.dt
        nxt = [0] * ew.ngrids
        for n, r in enumerate(recs):
            traces[n].append(rec[r[0]][nxt[r[0]]].copy())
            nxt[r[0]] += 1
        refs = [ew.array("U", g) for g in range(ew.ngrids)]
        ours = [blk.download("U") for blk in stack.blocks]
        scale = max(np.abs(r).max() for r in refs)
        assert scale > 0
        for g in range(ew.ngrids):
            e = np.abs(ours[g] - refs[g]).max() / scale
            worst = max(worst, e)
            assert e < TOL, "step %d grid %d: %g" % (step + 1, g, e)
    print("topography run: %d steps, worst per-step rel. diff %.3g" % (ew.nsteps, worst))
    for n, name in enumerate(("sta01.txt", "sta02.txt", "sta03.txt")):
        gold = _golden_station(name)
        mine = np.array(traces[n])
        assert gold.shape[0] == mine.shape[0] == ew.nsteps + 1
        # the reference's own goldens differ between builds at the 1e-14 level; gate at 1e-10
        assert np.abs(mine - gold[:, 1:4]).max() <= 1e-10 * np.abs(gold[:, 1:4]).max(), name
