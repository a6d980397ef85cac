"""z-slab decomposition of one Cartesian grid over the GPUs of one node and the halo exchange that
replaces the reference's MPI x-y halo swap (EW::communicate_array, EW.C:3247-3317;
communicate_arrayCU_X/Y, EW_cuda.C:1699-1997) on this path: one process per GPU, 2 planes of the new
solution per face, moved after the predictor and after the corrector (EW.C:2616, 2743).

With the (i,j,k,c) layout a k-plane of one component is one contiguous run of ni*nj doubles, so a halo is 3 contiguous
runs per face and needs no pack kernel: on the GPU the exchange lives in the library (csrc/exchange.cu, sw4b200_grid_exchange_begin /
_end: grouped ncclSend / ncclRecv straight from the field arrays into the neighbour's halo planes, NCCL over NVLink).  The
exchange of the face planes overlaps the computation of the slab's remaining rows: the face rows are computed first
(sw4b200_grid_*_part(1)), their transfer is started on the communication stream, then the bulk (part 2) runs; the boundary
conditions wait for both.  `HaloExchange` below is the host-logic stand-in used by the CPU (gloo) tests, where a numpy block
takes the place of the device block.
"""
import numpy as np


def decomp1d(nglobal, myid, nproc, olap=4):
    """EW::decomp1d (EW.C:2963-2985): block [s,e] of 1..nglobal for rank myid, blocks overlap by olap"""
    nlocal = (nglobal + (nproc - 1) * olap) // nproc
    deficit = (nglobal + (nproc - 1) * olap) % nproc
    if myid < deficit:
        s = myid * (nlocal - olap) + myid + 1
        nlocal += 1
    else:
        s = myid * (nlocal - olap) + deficit + 1
    return s, s + nlocal - 1


def slab_range(nz, rank, nranks):
    """interior planes [k0,k1] OWNED by slab `rank`: the decomp1d block minus the 2 padding planes
    towards each neighbour"""
    s, e = decomp1d(nz, rank, nranks)
    return s + (2 if rank > 0 else 0), e - (2 if rank < nranks - 1 else 0)


class HaloExchange:
    """CPU stand-in of the library's exchange (tests/test_slabs_cpu.py): moves the face planes of Up between neighbouring slabs
    with torch.distributed send/recv on packed buffers.  `blk` needs: ni, nj, pack(side) ->
    tensor(3*2*ni*nj), unpack(side, tensor); torch.distributed must be initialised when nranks > 1."""

    def __init__(self, blk, rank, nranks, device=None):
        import torch
        self.torch = torch
        self.blk, self.rank, self.nranks = blk, rank, nranks
        n = 2 * blk.halo_doubles(False)   # Up and (after the predictor) the stored acceleration; device rows may be padded
        kw = dict(dtype=torch.float64, device=device if device is not None else "cpu")
        self.send = [torch.zeros(n, **kw), torch.zeros(n, **kw)]
        self.recv = [torch.zeros(n, **kw), torch.zeros(n, **kw)]
        self.lo = rank - 1 if rank > 0 else None
        self.hi = rank + 1 if rank < nranks - 1 else None
        self.bytes_per_exchange = 8 * (n // 2) * ((self.lo is not None) + (self.hi is not None))   # Up only

    def exchange(self, with_acc=False):
        """blocking form (the overlapped form is GridBlock.begin_exchange / end_exchange)"""
        import torch.distributed as dist
        n = self.blk.halo_doubles(with_acc)
        ops = []
        for side, peer in ((0, self.lo), (1, self.hi)):
            if peer is None:
                continue
            self.blk.pack(side, self.send[side], with_acc=with_acc)
            ops.append(dist.P2POp(dist.isend, self.send[side][:n], peer))
            ops.append(dist.P2POp(dist.irecv, self.recv[side][:n], peer))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for side, peer in ((0, self.lo), (1, self.hi)):
            if peer is not None:
                self.blk.unpack(side, self.recv[side], with_acc=with_acc)


class SlabStepper:
    """one time step of a z-slab (EW.C:2527-2842 with the halo swap on k): face rows -> exchange
    (overlapped with the bulk rows) -> boundary conditions, twice per step.

    `curv`: the curvilinear block under the topography (config 4), held whole by the rank that owns the top
    Cartesian slab (rank 0): it steps between the face rows and the bulk rows of the Cartesian slab, so its
    work also overlaps the halo transfer, and is coupled to the Cartesian slab by EW::enforceCartTopo
    (EW.C:3504-3531) after the boundary conditions, as in EW::enforceBC (EW.C:3500)."""

    def __init__(self, blk, exchange, curv=None):
        self.blk, self.ex, self.curv = blk, exchange, curv

    def _bc(self):
        b = self.blk
        b.enforce_bc()
        if self.curv is not None:
            from . import lib as L
            self.curv.enforce_bc()
            L.check(b.lib.sw4b200_grid_enforce_cart_topo(b.h, self.curv.h))

    def step(self, f=None, ftt=None, fcurv=None, fttcurv=None):
        b = self.blk
        b.predictor_part(1, f)
        b.begin_exchange(self.ex, with_acc=True)    # Up and the stored acceleration of the face planes
        if self.curv is not None:
            self.curv.predictor(fcurv)
        b.predictor_part(2, f)
        b.end_exchange(self.ex, with_acc=True)
        self._bc()
        b.corrector_part(1, ftt)
        b.begin_exchange(self.ex)
        if self.curv is not None:
            self.curv.corrector(fttcurv)
        b.corrector_part(2, ftt)
        b.end_exchange(self.ex)
        self._bc()
        b.cycle()
        if self.curv is not None:
            self.curv.cycle()
