"""CPU (gloo, world_size 2 and 3) test of the z-slab host logic: decomposition, ownership of sources,
halo exchange sequencing and boundary-condition handling on halo faces (sw4lite_b200/slabs.py).
The compute of each slab is done by the ORACLE's kernels (tests/cpu_step.py) standing in for the
device block, so the test isolates the multi-rank logic: the slab run must reproduce the single-block
run bit for bit."""
import os
import socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sw4lite_b200.slabs import decomp1d, slab_range, HaloExchange, SlabStepper


def test_decomp1d_matches_reference_rule():
    # EW::decomp1d examples worked by hand from EW.C:2963-2985 (olap = 4)
    assert decomp1d(100, 0, 2) == (1, 52) and decomp1d(100, 1, 2) == (49, 100)
    assert decomp1d(101, 0, 2) == (1, 53) and decomp1d(101, 1, 2) == (50, 101)
    for nz, n in ((100, 2), (101, 3), (128, 8), (37, 4)):
        owned = [slab_range(nz, r, n) for r in range(n)]
        assert owned[0][0] == 1 and owned[-1][1] == nz
        for a, b in zip(owned[:-1], owned[1:]):
            assert b[0] == a[1] + 1          # owned ranges tile 1..nz without gaps or overlap


class CpuSlab:
    """numpy stand-in for sw4lite_b200.solver.GridBlock on one slab"""

    def __init__(self, prob, rank, nranks):
        from tests.cpu_step import OracleStepper
        bounds, onesided, bctype, self.halo_lo, self.halo_hi = prob.slab(rank, nranks)
        self.st = OracleStepper(prob, bounds=bounds, onesided=onesided, bctype=bctype)
        self.ni, self.nj, self.nk = self.st.ni, self.st.nj, self.st.nk
        self.kown = (bounds[4] + 2, bounds[5] - 2)
        self.prob = prob

    def _planes(self, k0):
        nij = self.ni * self.nj
        return self.st.Up.reshape(3, self.nk, nij)[:, k0:k0 + 2, :]

    def halo_doubles(self, with_acc):
        return 6 * self.ni * self.nj

    def pack(self, side, t, with_acc=False):
        n = self.halo_doubles(False)
        t[:n].copy_(torch.from_numpy(np.ascontiguousarray(self._planes(2 if side == 0 else self.nk - 4)).ravel()))

    def unpack(self, side, t, with_acc=False):
        n = self.halo_doubles(False)
        self._planes(0 if side == 0 else self.nk - 2)[...] = t[:n].numpy().reshape(3, 2, -1)

    def predictor_part(self, part, f):
        if part == 1:
            self.st.predictor(f)

    def corrector_part(self, part, ftt):
        if part == 1:
            self.st.corrector(ftt)

    def begin_exchange(self, ex, with_acc=False):
        ex.exchange(with_acc)

    def end_exchange(self, ex, with_acc=False):
        pass

    def enforce_bc(self):
        self.st.enforce_bc()

    def cycle(self):
        self.st.cycle()


def _problem():
    from sw4lite_b200.setup import CartesianProblem
    prob = CartesianProblem(20, 18, 41, h=50.0, gp=6, layers=[(600.0, 5000.0, 2800.0, 2700.0)])
    prob.add_point_force(9, 8, 5, (1e10, -2e10, 3e10), freq=3.0)
    prob.add_point_force(11, 10, 22, (2e10, 1e10, -1e10), freq=4.0)     # lands on/next to a slab face for 2 and 3 ranks
    prob.add_point_force(8, 9, 30, (-1e10, 1e10, 1e10), freq=5.0)
    return prob


def _worker(rank, nranks, port, nsteps, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=nranks)
    try:
        prob = _problem()
        blk = CpuSlab(prob, rank, nranks)
        stepper = SlabStepper(blk, HaloExchange(blk, rank, nranks))
        t = 0.0
        for _ in range(nsteps):
            stepper.step(prob.forces(t), prob.forces(t, tt=True))
            t += prob.dt
        k0, k1 = blk.kown
        own = blk.st.U.reshape(3, blk.nk, -1)[:, 2:-2, :]
        out[rank] = (k0, k1, own.copy())
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.parametrize("nranks", [2, 3])
def test_slab_run_reproduces_single_block(nranks):
    from tests.cpu_step import OracleStepper
    nsteps = 6
    prob = _problem()
    ref = OracleStepper(prob)
    t = 0.0
    for _ in range(nsteps):
        ref.step(prob.forces(t), prob.forces(t, tt=True))
        t += prob.dt
    full = ref.U.reshape(3, ref.nk, -1)
    assert np.abs(full).max() > 0
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(nranks, _free_port(), nsteps, out), nprocs=nranks, join=True)
    assert len(out) == nranks
    for r in range(nranks):
        k0, k1, own = out[r]
        a = full[:, k0 - prob.bounds[4]:k1 - prob.bounds[4] + 1, :]
        assert np.array_equal(own, a), "slab %d differs from the single-block run" % r
