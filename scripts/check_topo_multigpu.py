#!/usr/bin/env python3
"""Multi-GPU check of a topography run (config 4 pattern; run under torchrun, one rank per GPU, NCCL):
the Cartesian grid of pytest/reference/topo/curvilinear.in is z-slab decomposed over the ranks, the curvilinear
grid under the topography lives on rank 0 and is coupled through EW::enforceCartTopo; the result must equal
the single-GPU GridStack run bit for bit.  The set-up arrays come from the reference's own set-up (oracle/_ref),
which is test infrastructure.
   torchrun --nproc-per-node N scripts/check_topo_multigpu.py [nsteps] [input.in] [balance]
input.in: a file of tests/golden/inputs (default curvilinear.in; gaussianHill-rev.in = BASELINE.json config 4, 128 x 128 x 1900
Cartesian + 128 x 128 x 106 curvilinear points).  balance = 1: the rank holding the curvilinear grid owns fewer Cartesian planes
(a curvilinear point costs about three Cartesian ones).  Prints the device time per step of the slab run (max over
ranks) and of the single-GPU run next to the bit-identity verdict."""
import os
import sys
import tempfile
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim
from sw4lite_b200.solver import GridBlock, GridStack, boundary_windows, bProcessor
from sw4lite_b200.slabs import SlabStepper, slab_range
from sw4lite_b200 import lib as L
from tests.test_gpu_step import SourceMap


def block(ew, g, device, curv=False, krange=None, halo=(False, False)):
    G = ew.grids[g]
    bounds = list(G.bounds); onesided = list(G.onesided); bctype = list(G.bctype)
    k0 = 0
    if krange is not None:
        k0 = krange[0] - 2 - G.bounds[4]
        bounds[4], bounds[5] = krange[0] - 2, krange[1] + 2
        if halo[0]:
            onesided[4] = 0; bctype[4] = bProcessor
        if halo[1]:
            onesided[5] = 0; bctype[5] = bProcessor
    blk = GridBlock(ew.corder, bounds, (G.nx, G.ny, G.nz), G.h, ew.dt, onesided, bctype, boundary_windows(bounds, bctype),
                    sg_order=ew.sgorder if ew.usesg else 0, beta=ew.beta if ew.usesg else 0.0, curvilinear=curv,
                    halo_lo=halo[0], halo_hi=halo[1], device=device)
    nk = blk.nk
    nij = G.ni * G.nj
    for name in ("mu", "lambda", "rho") + (("jac",) if curv else ()):
        blk.upload(name, np.ascontiguousarray(ew.array(name, g).reshape(G.nk, nij)[k0:k0 + nk]).ravel())
    if curv:
        blk.upload("metric", ew.array("metric", g))
    for name in ("strx", "stry", "dcx", "dcy", "cox", "coy"):
        blk.upload(name, ew.array(name, g))
    if not curv:
        for name in ("strz", "dcz", "coz"):
            blk.upload(name, ew.array(name, g)[k0:k0 + nk])
    return blk


def weighted_ranges(nz, world, ncurv_planes, weight=3.0, kmin=8):
    """owned Cartesian planes per rank when rank 0 also steps `ncurv_planes` curvilinear planes costing `weight` each"""
    if world == 1:
        return [(1, nz)]
    share = (nz + weight * ncurv_planes) / world
    n0 = int(max(kmin, round(share - weight * ncurv_planes)))
    rest = nz - n0
    out = [(1, n0)]
    k = n0 + 1
    for r in range(1, world):
        n = rest // (world - 1) + (1 if r - 1 < rest % (world - 1) else 0)
        out.append((k, k + n - 1))
        k += n
    assert out[-1][1] == nz
    return out


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    infile = sys.argv[2] if len(sys.argv) > 2 else "curvilinear.in"
    balance = len(sys.argv) > 3 and sys.argv[3] == "1"
    L.init(local); L.comm_init(rank, world)
    lib = L.load()
    tmp = tempfile.mkdtemp()
    saved = os.dup(1); os.dup2(2, 1)          # the reference prints its set-up log
    ew = refshim.RefEW(os.path.join(ROOT, "tests", "golden", "inputs", infile), tmp)
    sys.stdout.flush(); os.dup2(saved, 1)
    assert ew.topo == 1 and ew.ngrids == 2
    srcs = [SourceMap(ew, g) for g in range(2)]
    # reference run on this rank's GPU
    stack = GridStack([block(ew, 0, local), block(ew, 1, local, curv=True)], ncart=1)
    for g, b in enumerate(stack.blocks):
        if len(srcs[g].points):
            b.set_source_points(srcs[g].points)
    # slab run
    G0 = ew.grids[0]
    if balance:
        k0, k1 = weighted_ranges(G0.nz, world, ew.grids[1].nz)[rank]
    else:
        k0, k1 = slab_range(G0.nz, rank, world)
    cart = block(ew, 0, local, krange=(k0, k1), halo=(rank > 0, rank < world - 1))
    sel = [n for n, pt in enumerate(srcs[0].points) if k0 <= pt[2] <= k1]
    if sel:
        cart.set_source_points(srcs[0].points[sel])
    curv = None
    if rank == 0:
        curv = block(ew, 1, local, curv=True)
        if len(srcs[1].points):
            curv.set_source_points(srcs[1].points)
    cart.set_neighbours(rank - 1 if rank > 0 else None, rank + 1 if rank < world - 1 else None)
    stepper = SlabStepper(cart, None, curv=curv)
    # a random initial wavefield (the same in both runs) so that every plane carries signal from step 1
    r = np.random.default_rng(11)
    nij = G0.ni * G0.nj
    for g, G in enumerate(ew.grids):
        u0 = r.uniform(-1e-3, 1e-3, 3 * G.npts); um0 = u0 + r.uniform(-1e-5, 1e-5, 3 * G.npts)
        stack.blocks[g].upload("U", u0); stack.blocks[g].upload("Um", um0)
        if g == 0:
            c0 = cart.bounds[4] - G.bounds[4]
            cart.upload("U", np.ascontiguousarray(u0.reshape(3, G.nk, nij)[:, c0:c0 + cart.nk]).ravel())
            cart.upload("Um", np.ascontiguousarray(um0.reshape(3, G.nk, nij)[:, c0:c0 + cart.nk]).ravel())
        elif curv is not None:
            curv.upload("U", u0); curv.upload("Um", um0)
    nsteps += 2
    times = [ew.tstart + s * ew.dt for s in range(nsteps)]
    forces = []
    for t in times:
        fa = ew.eval_forces(t, False); fta = ew.eval_forces(t, True)
        forces.append(([m.reduce(fa) for m in srcs], [m.reduce(fta) for m in srcs]))
    import ctypes as C

    def timed(fn):
        L.check(lib.sw4b200_sync_device()); dist.barrier(); L.check(lib.sw4b200_sync_device())
        L.check(lib.sw4b200_timer_start())
        fn()
        ms = C.c_double(0)
        L.check(lib.sw4b200_timer_stop_ms(C.byref(ms)))
        t = torch.tensor([ms.value], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def run_stack():
        for f, ftt in forces:
            stack.step(f, ftt)

    def run_slabs():
        for f, ftt in forces:
            stepper.step(f[0][sel] if sel else None, ftt[0][sel] if sel else None, f[1] if curv else None, ftt[1] if curv else None)

    # two untimed steps first (kernel attributes, NCCL's lazily built peer connections), on both runs alike
    warm, forces = forces[:2], forces[2:]
    for f, ftt in warm:
        stack.step(f, ftt)
        stepper.step(f[0][sel] if sel else None, ftt[0][sel] if sel else None, f[1] if curv else None, ftt[1] if curv else None)
    nsteps -= 2
    ms_stack = timed(run_stack)
    ms_slabs = timed(run_slabs)
    cart.sync()
    ref = stack.blocks[0].download("U").reshape(3, G0.nk, nij)[:, k0 - G0.bounds[4]:k1 - G0.bounds[4] + 1]
    mine = cart.download("U").reshape(3, cart.nk, nij)[:, 2:-2]
    same = np.array_equal(mine, ref) and np.abs(ref).max() > 0
    msg = "rank %d/%d Cartesian planes %d..%d: %s (scale %.3g)" % (rank, world, k0, k1, "bit-identical" if same else "DIFFERENT max|diff| %.3g" % np.abs(mine - ref).max(), np.abs(ref).max())
    if curv is not None:
        a = curv.download("U"); b = stack.blocks[1].download("U")
        sc = np.array_equal(a, b) and np.abs(b).max() > 0
        same = same and sc
        msg += "; curvilinear grid: %s (scale %.3g)" % ("bit-identical" if sc else "DIFFERENT max|diff| %.3g" % np.abs(a - b).max(), np.abs(b).max())
    print(msg, flush=True)
    if rank == 0:
        pts = sum(G.nx * G.ny * G.nz for G in ew.grids)
        print("TIMING %s: %d grid points (Cartesian %dx%dx%d + curvilinear %dx%dx%d), %d steps; z-slabs over %d GPU(s)%s: %.3f ms/step = %.3f Gpts/s; "
              "single GPU (all ranks run it side by side): %.3f ms/step = %.3f Gpts/s" % (
                  infile, pts, G0.nx, G0.ny, G0.nz, ew.grids[1].nx, ew.grids[1].ny, ew.grids[1].nz, nsteps, world,
                  " (balanced)" if balance else "", ms_slabs / nsteps, pts * nsteps / ms_slabs / 1e6, ms_stack / nsteps, pts * nsteps / ms_stack / 1e6), flush=True)
    ok = torch.tensor([1 if same else 0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(ok.item()) == 1 else 1)


if __name__ == "__main__":
    main()
