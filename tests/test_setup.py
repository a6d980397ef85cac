"""Host-side set-up of the synthetic problems (sw4lite_b200/setup.py, mirrored in C++ by host/slab_driver.C) against the
reference's own parser + setupRun (oracle/_ref): supergrid damping / stretching / corner-taper arrays
(SuperGrid.C:108-198, EW.C:4671-4801), the time step (EW::computeDT, EW.C:5041-5066), boundary types, one-sided flags and
windows (EW.C:4805-4867, 3347-3420), layered materials, and the z-slab decomposition rule (EW::decomp1d, EW.C:2963-2985)."""
import os
import numpy as np
import pytest

from oracle import refshim
from sw4lite_b200.setup import CartesianProblem, supergrid_1d
from sw4lite_b200.slabs import slab_range

needs_ref = pytest.mark.skipif(not refshim.available(), reason="oracle/_ref/libsw4ref.so not present")


def ref_problem(tmp_path, nx, ny, nz, h, gp, ztop):
    txt = "\n".join([
        "fileio path=%s verbose=0" % str(tmp_path / "out"),
        "grid nx=%d ny=%d nz=%d h=%g" % (nx, ny, nz, h),
        "time steps=5",
        "developer checkfornan=0 reporttiming=0 corder=1 cfl=1.3",
        "supergrid gp=%d" % gp,
        "block vp=4000 vs=2000 r=2600",
        "block vp=6000 vs=3464 r=2700 z1=%g" % ztop,
        "source x=%g y=%g z=%g mxy=1e18 t0=0 freq=10 type=C6SmoothBump" % (0.5 * nx * h, 0.5 * ny * h, 0.3 * nz * h), ""])
    f = tmp_path / "setup.in"
    f.write_text(txt)
    return refshim.RefEW(str(f), str(tmp_path))


@needs_ref
@pytest.mark.parametrize("nx,ny,nz,gp", [(61, 50, 45, 12), (100, 72, 70, 30)])
def test_cartesian_problem_equals_the_references_setup(tmp_path, nx, ny, nz, gp):
    h = 25.0
    ztop = 0.6 * nz * h
    ew = ref_problem(tmp_path, nx, ny, nz, h, gp, ztop)
    prob = CartesianProblem(nx, ny, nz, h=h, vp=4000.0, vs=2000.0, rho=2600.0, gp=gp, beta=ew.beta, corder=1,
                            layers=[(ztop, 6000.0, 3464.0, 2700.0)])
    G = ew.grids[0]
    assert G.bounds == prob.bounds and (G.nx, G.ny, G.nz) == (nx, ny, nz)
    assert list(G.onesided) == list(prob.onesided) and list(G.bctype) == list(prob.bctype)
    assert np.array_equal(G.wind, prob.wind)
    assert ew.sgorder == 4 and ew.usesg == 1
    for name in ("strx", "stry", "strz", "dcx", "dcy", "dcz", "cox", "coy", "coz"):
        a, b = np.array(ew.array(name, 0)), getattr(prob, name)
        assert a.shape == b.shape
        assert np.max(np.abs(a - b)) <= 4e-16 * max(1.0, np.max(np.abs(a))), name
    # the reference rounds dt so that an integer number of steps reaches t; with `time steps=` it keeps the CFL value
    assert abs(ew.dt - prob.dt) <= 1e-15 * prob.dt
    for name, mine in (("mu", prob.mu), ("lambda", prob.la), ("rho", prob.rho)):
        ref = np.array(ew.array(name, 0))
        assert np.max(np.abs(ref - mine)) <= 1e-15 * np.max(np.abs(ref)), name


def test_supergrid_profiles_are_tapers():
    x = np.arange(-2, 103) * 10.0
    dc, stretch, corner = supergrid_1d(x, True, True, 0.0, 1000.0, 300.0)
    assert dc[50] == 0 and stretch[50] == 1 and corner[50] == 1            # untouched interior
    assert np.all(stretch > 0) and stretch.min() >= 1e-4 * (1 - 1e-10) and corner[2:-2].min() >= 0.33 - 1e-12
    assert np.all(dc >= 0) and dc[0] > 0 and dc[-1] > 0 and np.all(dc[32:73] == 0)


@pytest.mark.parametrize("nz,n", [(128, 8), (341, 8), (1900, 8), (41, 3), (24, 2)])
def test_slabs_tile_the_grid(nz, n):
    owned = [slab_range(nz, r, n) for r in range(n)]
    assert owned[0][0] == 1 and owned[-1][1] == nz
    for a, b in zip(owned[:-1], owned[1:]):
        assert b[0] == a[1] + 1
    sizes = [b - a + 1 for a, b in owned]
    assert max(sizes) - min(sizes) <= 5      # decomp1d balances the padded blocks; the end slabs own up to 2+2 more planes


@needs_ref
def test_loh1_problem_equals_the_references_setup(tmp_path):
    """BASELINE.json config 3 at h=100: CartesianProblem.loh1 + tests/golden/loh1-h100-setup.npz against the reference's parser,
    set-up and source discretisation for tests/loh1/LOH.1-h100.in"""
    from tests.test_gpu_step import SourceMap
    here = os.path.dirname(os.path.abspath(__file__))
    ew = refshim.RefEW(os.path.join(here, "golden", "inputs", "LOH.1-h100.in"), str(tmp_path))
    prob = CartesianProblem.loh1(100.0)
    fx = np.load(os.path.join(here, "golden", "loh1-h100-setup.npz"))
    G = ew.grids[0]
    assert G.bounds == prob.bounds and list(G.bctype) == list(prob.bctype) and np.array_equal(G.wind, prob.wind)
    assert prob.nsteps == ew.nsteps == int(fx["nsteps"]) and abs(prob.dt - ew.dt) <= 1e-16 and float(fx["dt"]) == ew.dt
    for name, mine in (("mu", prob.mu), ("lambda", prob.la), ("rho", prob.rho)):
        ref = np.array(ew.array(name, 0))
        assert np.max(np.abs(ref - mine)) <= 1e-15 * np.max(np.abs(ref)), name
    for name in ("strx", "stry", "strz", "dcx", "dcy", "dcz", "cox", "coy", "coz"):
        assert np.max(np.abs(np.array(ew.array(name, 0)) - getattr(prob, name))) <= 4e-16, name
    assert abs(ew.beta - prob.beta) < 1e-16
    src = SourceMap(ew)
    assert np.array_equal(src.points, fx["ijk"])
    for s in (0, 17, 21, 40, 300):
        t = ew.tstart + s * ew.dt
        f = src.reduce(ew.eval_forces(t, False)); ftt = src.reduce(ew.eval_forces(t, True))
        assert np.abs(f - fx["F0"] * fx["g"][s]).max() <= 1e-14 * max(np.abs(f).max(), 1e-300)
        assert np.abs(ftt - fx["F0"] * fx["gtt"][s]).max() <= 1e-14 * max(np.abs(ftt).max(), 1e-300)
    recs, _ = ew.receivers()
    assert np.array_equal(np.array(recs)[:, 1:4], fx["rec"])
