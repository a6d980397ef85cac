"""Check point files (sw4lite_b200/checkpoint.py) against the reference's own CheckPoint class, on the CPU: the reference
program (oracle/_ref/sw4lite_ref, unmodified) must be able to restart from a file written by our writer and end with the
error norms of its uninterrupted run, and our reader must recover from a file written by the reference the same wavefield the
reference holds at that cycle.  -m gpu: a device-resident slab run saved and restored mid-run equals the uninterrupted run."""
import os
import subprocess
import numpy as np
import pytest

from oracle import refshim
from sw4lite_b200 import checkpoint as cp

needs_ref = pytest.mark.skipif(not (refshim.available() and os.path.exists(refshim.EXE)), reason="oracle/_ref not built")

BASE = """fileio verbose=1 path=%s
grid x=2.0 y=1.6 z=1.2 h=0.04
time t=0.6
testpointsource rho=1 cp=1.6 cs=0.8 halfspace=1
supergrid gp=8
source x=1.0 y=0.8 z=0.44 Mxx=1 Myy=1 Mzz=1 Mxy=0 Mxz=0 Myz=0 t0=0 freq=1 type=C6SmoothBump
developer checkfornan=0 cfl=1.3 reporttiming=0 corder=0
"""
NCYCLE = 9


def program(d, extra):
    d.mkdir()
    inp = d / "run.in"
    inp.write_text(BASE % (str(d / "out")) + extra)
    r = subprocess.run([refshim.EXE, str(inp)], cwd=str(d), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    err = [p for p in d.rglob("PointSourceErr.txt")]
    assert err, r.stdout[-2000:]
    return [float(x) for x in open(err[0]).read().split()], r.stdout


def stepped(tmp_path):
    """the reference's EW stepped to cycle NCYCLE by the shim: (ew, Um, U) with arrays shaped (3, nk, nj, ni)"""
    d = tmp_path / "shim"
    d.mkdir()
    inp = d / "run.in"
    inp.write_text(BASE % (str(d / "out")))
    ew = refshim.RefEW(str(inp), str(d))
    assert ew.corder == 0 and ew.ngrids == 1
    for _ in range(NCYCLE):
        ew.step()
    G = ew.grids[0]
    shp = lambda a: np.moveaxis(np.array(a).reshape(G.nk, G.nj, G.ni, 3), 3, 0)
    return ew, G, shp(ew.array("Um", 0)), shp(ew.array("U", 0))


@needs_ref
def test_reference_program_restarts_from_our_file(tmp_path):
    full, _ = program(tmp_path / "full", "")
    ew, G, um, u = stepped(tmp_path)
    path = str(tmp_path / "ours.sw4checkpoint")
    sizes = [(G.nx, G.ny, G.nz)]
    cp.write_header(path, ew.t, NCYCLE, sizes)
    cp.write_planes(path, sizes, 0, 0, 1, um[:, 2:-2, 2:-2, 2:-2])
    cp.write_planes(path, sizes, 0, 1, 1, u[:, 2:-2, 2:-2, 2:-2])
    again, out = program(tmp_path / "restart", "restart file=%s\n" % path)
    assert full[1] > 0 and full[0] == again[0]
    for a, b in zip(again, full):
        assert abs(a - b) <= 1e-13 * abs(b), (again, full)


@needs_ref
def test_our_reader_recovers_the_references_file(tmp_path):
    _, out = program(tmp_path / "write", "checkpoint cycle=%d file=chk\n" % NCYCLE)
    files = [p for p in (tmp_path / "write").rglob("*.sw4checkpoint")]
    assert len(files) == 1, out[-1500:]
    ew, G, um, u = stepped(tmp_path)
    t, cycle, sizes = cp.read_header(str(files[0]))
    assert cycle == NCYCLE and sizes == [(G.nx, G.ny, G.nz)] and abs(t - ew.t) < 1e-12
    a = cp.read_planes(str(files[0]), sizes, 0, 0, 1, G.nz)
    b = cp.read_planes(str(files[0]), sizes, 0, 1, 1, G.nz)
    assert np.abs(u).max() > 0
    assert np.array_equal(a, um[:, 2:-2, 2:-2, 2:-2]) and np.array_equal(b, u[:, 2:-2, 2:-2, 2:-2])
    # planes read one slab at a time (what a restarted z-slab does) are the same values
    mid = cp.read_planes(str(files[0]), sizes, 0, 1, 7, 5)
    assert np.array_equal(mid, b[:, 6:11])


@pytest.mark.gpu
@pytest.mark.parametrize("nslabs,nx", [(1, 40), (2, 40), (3, 41)])
def test_device_slabs_saved_and_restored_mid_run(tmp_path, nslabs, nx):
    """f4 on the device path: z-slab blocks write their own planes of Um and U into one reference-format file; fresh blocks
    restored from it (halo planes from the file, ghost points from the boundary kernels) continue bit-identically"""
    import torch
    from sw4lite_b200.setup import CartesianProblem
    prob = CartesianProblem(nx, 34, 46, h=100.0, gp=7, corder=1, layers=[(1500.0, 6000.0, 3464.0, 2700.0)])
    prob.add_point_force(18, 14, 8, (1e12, 2e12, -1e12), freq=2.0)
    prob.add_point_force(20, 16, 30, (-2e12, 1e12, 1e12), freq=3.0)
    r = np.random.default_rng(3)
    u0 = r.uniform(-1e-3, 1e-3, 3 * prob.npts); um0 = u0 + r.uniform(-1e-5, 1e-5, 3 * prob.npts)
    nij = prob.ni * prob.nj
    full = lambda a: a.reshape(3, prob.nk, nij)

    def make():
        slabs = [prob.make_block(rank=q, nranks=nslabs) for q in range(nslabs)]
        buf = [[torch.zeros(s.halo_doubles(True), dtype=torch.float64, device="cuda") for _ in range(2)] for s in slabs]
        return slabs, buf

    def exchange(slabs, buf, with_acc=False):
        for q, s in enumerate(slabs):
            for side in (0, 1):
                if (side == 0 and q > 0) or (side == 1 and q < nslabs - 1):
                    s.pack(side, buf[q][side], with_acc=with_acc)
        for q, s in enumerate(slabs):
            if q > 0:
                s.unpack(0, buf[q - 1][1], with_acc=with_acc)
            if q < nslabs - 1:
                s.unpack(1, buf[q + 1][0], with_acc=with_acc)

    def steps(slabs, buf, first, n):
        for m in range(first, first + n):
            t = m * prob.dt
            f, ftt = prob.forces(t), prob.forces(t, tt=True)
            for s in slabs:
                s.predictor_part(1, f[s.src_sel])
            exchange(slabs, buf, True)
            for s in slabs:
                s.predictor_part(2, f[s.src_sel]); s.enforce_bc(); s.corrector_part(1, ftt[s.src_sel])
            exchange(slabs, buf)
            for s in slabs:
                s.corrector_part(2, ftt[s.src_sel]); s.enforce_bc(); s.cycle()

    slabs, buf = make()
    for s in slabs:
        k0 = s.bounds[4] - prob.bounds[4]
        s.upload("U", np.ascontiguousarray(full(u0)[:, k0:k0 + s.nk]).ravel())
        s.upload("Um", np.ascontiguousarray(full(um0)[:, k0:k0 + s.nk]).ravel())
    steps(slabs, buf, 0, 3)
    path = str(tmp_path / "slabs.sw4checkpoint")
    for q, s in enumerate(slabs):
        cp.save_slab(s, prob, path, 3 * prob.dt, 3, rank=q)
    steps(slabs, buf, 3, 3)
    fresh, buf2 = make()
    for s in fresh:
        t, cycle = cp.load_slab(s, prob, path)
        assert cycle == 3 and abs(t - 3 * prob.dt) < 1e-15
    steps(fresh, buf2, 3, 3)
    for a, b in zip(slabs, fresh):
        ua, ub = a.download("U").reshape(3, a.nk, nij)[:, 2:-2], b.download("U").reshape(3, b.nk, nij)[:, 2:-2]
        assert np.abs(ua).max() > 0 and np.array_equal(ua, ub)
        assert np.array_equal(a.download("Um").reshape(3, a.nk, nij)[:, 2:-2], b.download("Um").reshape(3, b.nk, nij)[:, 2:-2])
