"""-m gpu: whole time steps of the Cartesian path on the GPU against the reference CPU kernels
stepped by the reference's own EW object (oracle/_ref), on the reference's pointsource test
(config 1: tests/pointsource/pointsource.in, 201x201x101, free surface + supergrid, 23 steps).
Gates: per-step max|gpu-cpu|/max|cpu| <= 1e-12 on the new solution, and the final
`Errors at time ... Linf ... L2 ...` line equal to the golden one to printed precision."""
import os
import numpy as np
import pytest

from oracle import refshim
from tests.fields import relerr

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not refshim.available(), reason="oracle/_ref/libsw4ref.so not present")
INPUTS = os.path.join(os.path.dirname(__file__), "golden", "inputs")

GOLDEN_LINE = "Errors at time 0.6 Linf = 0.569416 L2 = 0.0245361 norm of solution = 3.7439"   # reference README.md:66
GOLDEN_FILE = (0.6, 0.569416364119, 0.0245360919934, 3.74390307494)  # pytest/reference/pointsource/pointsource-h0p04/PointSourceErr.txt


def block_from_reference(ew, g=0, **kw):
    """build our device block from the arrays the reference's setup produced"""
    from sw4lite_b200.solver import GridBlock
    G = ew.grids[g]
    blk = GridBlock(ew.corder, G.bounds, (G.nx, G.ny, G.nz), G.h, ew.dt, G.onesided, G.bctype, G.wind,
                    sg_order=ew.sgorder if ew.usesg else 0, beta=ew.beta if ew.usesg else 0.0, **kw)
    for name in ("mu", "lambda", "rho", "strx", "stry", "strz", "dcx", "dcy", "dcz", "cox", "coy", "coz"):
        blk.upload(name, ew.array(name, g))
    return blk


class SourceMap:
    """unique source points of grid g and the reduction of per-source forces onto them
    (EW::Force sums the sources sharing a grid point, EW.C:3092-3121)"""

    def __init__(self, ew, g=0):
        idx, _, ident = ew.point_sources()
        self.sel = [r for r in range(len(ident) - 1) if idx[ident[r], 0] == g]
        self.ranges = [(ident[r], ident[r + 1]) for r in self.sel]
        self.points = np.array([idx[a, 1:4] for a, _ in self.ranges], dtype=np.int32).reshape(-1, 3)

    def reduce(self, f):
        out = np.zeros((len(self.ranges), 3))
        for n, (a, b) in enumerate(self.ranges):
            for s in range(a, b):
                out[n] += f[s]
        return out


@needs_ref
def test_pointsource_per_step_parity_and_error_line(tmp_path):
    ew = refshim.RefEW(os.path.join(INPUTS, "pointsource.in"), str(tmp_path))
    assert ew.ngrids == 1 and ew.nsteps == 23 and ew.corder == 1
    blk = block_from_reference(ew)
    src = SourceMap(ew)
    blk.set_source_points(src.points)
    worst = 0.0
    t = ew.tstart
    for step in range(ew.nsteps):
        f = src.reduce(ew.eval_forces(t, False))
        ftt = src.reduce(ew.eval_forces(t, True))
        ew.step()                       # reference CPU step (rotates: new solution is now "U")
        blk.step(f, ftt)
        t += ew.dt
        ours = blk.download("U")
        ref = ew.array("U", 0)
        e = relerr(ours, ref)
        worst = max(worst, e)
        assert e < 1e-12, "step %d: %g" % (step + 1, e)
    errs = ew.pointsource_error(ew.t, [blk.download("U")])
    line = "Errors at time %g Linf = %g L2 = %g norm of solution = %g" % (ew.t, errs[0], errs[1], errs[2])
    print(line, " worst per-step rel. diff %.3g" % worst)
    assert line == GOLDEN_LINE
    for a, b in zip((ew.t,) + tuple(errs), GOLDEN_FILE):
        assert abs(a - b) <= 1e-10 * abs(b)


# ----------------------------------------------------------------------------------------------
# synthetic blocks against the oracle's unfused kernel sequence (tests/cpu_step.py)
def _problem(corder=1, nz=34):
    from sw4lite_b200.setup import CartesianProblem
    prob = CartesianProblem(37, 30, nz, h=100.0, gp=7, corder=corder, layers=[(1500.0, 6000.0, 3464.0, 2700.0)])
    prob.add_point_force(18, 14, 8, (1e12, 2e12, -1e12), freq=2.0)
    prob.add_point_force(20, 16, nz // 2 + 1, (-2e12, 1e12, 1e12), freq=3.0)
    return prob


def _initial(prob, seed=5):
    r = np.random.default_rng(seed)
    u0 = r.uniform(-1e-3, 1e-3, 3 * prob.npts)
    return u0, u0 + r.uniform(-1e-5, 1e-5, 3 * prob.npts)


@pytest.mark.parametrize("corder", [1, 0])
def test_synthetic_block_steps_match_oracle(corder):
    """free surface (SBP closure rows) + supergrid layers + layered medium + two point forces, 6 steps"""
    from tests.cpu_step import OracleStepper
    prob = _problem(corder)
    blk = prob.make_block()
    cpu = OracleStepper(prob)
    u0, um0 = _initial(prob)
    blk.upload("U", u0); blk.upload("Um", um0)
    cpu.U[:] = u0; cpu.Um[:] = um0
    t = 0.0
    for step in range(6):
        f, ftt = prob.forces(t), prob.forces(t, tt=True)
        blk.step(f, ftt); cpu.step(f, ftt)
        t += prob.dt
        assert relerr(blk.download("U"), cpu.U) < 1e-12, "step %d" % (step + 1)


def test_damping_zonly_kernel_is_bit_identical_to_the_general_kernel():
    """boxes in which only dcz is non-zero take the streaming z-only damping kernel: same bits as the general kernel"""
    import sw4lite_b200 as S
    from sw4lite_b200.setup import CartesianProblem
    prob = CartesianProblem(70, 66, 40, h=100.0, gp=8, corder=1, layers=[(1500.0, 6000.0, 3464.0, 2700.0)])
    prob.add_point_force(30, 28, 9, (1e12, 2e12, -1e12), freq=2.0, t0=0.0)
    u0, um0 = _initial(prob)
    out = []
    lib = S.load()
    try:
        for zonly in (1, 0):
            S.lib.check(lib.sw4b200_set_option(b"sgd_zonly", zonly))
            blk = prob.make_block()
            blk.upload("U", u0); blk.upload("Um", um0)
            n0 = lib.sw4b200_kernel_launch_count()
            for s in range(3):
                blk.step(prob.forces(s * prob.dt), prob.forces(s * prob.dt, tt=True))
            out.append((blk.download("U"), lib.sw4b200_kernel_launch_count() - n0))
    finally:
        S.lib.check(lib.sw4b200_set_option(b"sgd_zonly", 1))
    assert np.array_equal(out[0][0], out[1][0])
    assert out[0][1] == out[1][1] > 0   # same boxes either way, only the kernel of the z-only boxes differs


def test_resident_run_equals_stepwise_and_records():
    """sw4b200_grid_run (sources/receivers device resident, no host sync) == the per-step API, bit for bit"""
    prob = _problem()
    u0, um0 = _initial(prob)
    rec = np.array([[5, 6, 1], [18, 14, 1], [30, 20, 3]], dtype=np.int32)
    n = 5
    f_all = np.array([prob.forces(s * prob.dt) for s in range(n)])
    ftt_all = np.array([prob.forces(s * prob.dt, tt=True) for s in range(n)])
    a = prob.make_block(); b = prob.make_block()
    for blk in (a, b):
        blk.upload("U", u0); blk.upload("Um", um0); blk.set_receiver_points(rec)
    recs = np.array([a.step(f_all[s], ftt_all[s], record=True) for s in range(n)])
    b.set_source_series(f_all, ftt_all)
    b.run(0, 2); b.run(2, 3)
    assert np.array_equal(a.download("U"), b.download("U")) and np.array_equal(a.download("Um"), b.download("Um"))
    assert np.array_equal(recs, b.fetch_records(0, n))
    assert np.abs(recs).max() > 0


@pytest.mark.parametrize("nslabs", [2, 3])
def test_slabs_on_one_gpu_reproduce_single_block(nslabs):
    """z-slab blocks (face rows first, halo planes moved with pack/unpack, bulk rows after) driven from
    one process on one GPU == the undivided block, bit for bit"""
    import torch
    prob = _problem(nz=41)
    u0, um0 = _initial(prob)
    whole = prob.make_block()
    whole.upload("U", u0); whole.upload("Um", um0)
    slabs = [prob.make_block(rank=r, nranks=nslabs) for r in range(nslabs)]
    nij = whole.ni * whole.nj
    full = lambda a: a.reshape(3, whole.nk, nij)
    for s in slabs:
        k0 = s.bounds[4] - whole.bounds[4]
        s.upload("U", np.ascontiguousarray(full(u0)[:, k0:k0 + s.nk]).ravel())
        s.upload("Um", np.ascontiguousarray(full(um0)[:, k0:k0 + s.nk]).ravel())
    buf = [[torch.zeros(s.halo_doubles(True), dtype=torch.float64, device="cuda") for _ in range(2)] for s in slabs]

    def exchange(with_acc=False):
        for r, s in enumerate(slabs):
            for side in (0, 1):
                if (side == 0 and r > 0) or (side == 1 and r < nslabs - 1):
                    s.pack(side, buf[r][side], with_acc=with_acc)
        for r, s in enumerate(slabs):
            if r > 0:
                s.unpack(0, buf[r - 1][1], with_acc=with_acc)
            if r < nslabs - 1:
                s.unpack(1, buf[r + 1][0], with_acc=with_acc)

    t = 0.0
    for step in range(5):
        f, ftt = prob.forces(t), prob.forces(t, tt=True)
        whole.step(f, ftt)
        for s in slabs:
            s.predictor_part(1, f[s.src_sel])
        exchange(with_acc=True)
        for s in slabs:
            s.predictor_part(2, f[s.src_sel]); s.enforce_bc(); s.corrector_part(1, ftt[s.src_sel])
        exchange()
        for s in slabs:
            s.corrector_part(2, ftt[s.src_sel]); s.enforce_bc(); s.cycle()
        t += prob.dt
    ref = full(whole.download("U"))
    assert np.abs(ref).max() > 0
    for s in slabs:
        k0 = s.bounds[4] - whole.bounds[4]
        own = s.download("U").reshape(3, s.nk, nij)[:, 2:-2]
        assert np.array_equal(own, ref[:, k0 + 2:k0 + s.nk - 2]), "slab at k0=%d" % s.bounds[4]


@needs_ref
def test_loh1_h100_station_matches_golden(tmp_path):
    """config 3: LOH.1 layer over half-space (tests/loh1/LOH.1-h100.in: 301x301x171, free surface, supergrid
    gp=30 on five sides, three material blocks, Gaussian moment source, 536 steps).  The reference's set-up
    (parser, materials, supergrid arrays, source discretisation, time functions) feeds the device block, which
    runs all steps device-resident; the station trace must reproduce the reference's golden sta10.txt
    (tests/loh1/loh1-h100-sta10/sta10.txt; the reference's own builds differ from it at the 1e-14 level)."""
    ew = refshim.RefEW(os.path.join(INPUTS, "LOH.1-h100.in"), str(tmp_path))
    assert ew.ngrids == 1 and ew.corder == 1 and ew.nsteps == 536
    blk = block_from_reference(ew)
    src = SourceMap(ew)
    blk.set_source_points(src.points)
    recs, _ = ew.receivers()
    assert len(recs) == 1
    blk.set_receiver_points(np.array([recs[0][1:4]], dtype=np.int32))
    n = ew.nsteps
    times = ew.tstart + ew.dt * np.arange(n)
    f_all = np.array([src.reduce(ew.eval_forces(t, False)) for t in times])
    ftt_all = np.array([src.reduce(ew.eval_forces(t, True)) for t in times])
    blk.set_source_series(f_all, ftt_all)
    blk.run(0, n)
    trace = blk.fetch_records(0, n)[:, 0, :]
    gold = np.array([l.split() for l in open(os.path.join(os.path.dirname(__file__), "golden", "loh1-h100-sta10", "sta10.txt"))
                     if not l.startswith("#")], dtype=np.float64)
    assert gold.shape[0] == n + 1
    scale = np.abs(gold[:, 1:4]).max()
    err = np.abs(trace - gold[1:, 1:4]).max() / scale
    print("LOH.1-h100 sta10: max rel. diff to the golden trace %.3g (amplitude %.3g)" % (err, scale))
    assert scale > 0 and err < 1e-9
