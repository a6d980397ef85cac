#!/usr/bin/env python3
"""Multi-GPU check (run under torchrun, one rank per GPU, NCCL): a z-slab run with the overlapped halo
exchange (sw4lite_b200/slabs.py: face rows -> exchange on the comm stream || bulk rows -> BC) must reproduce
the undivided single-GPU run of the same problem bit for bit.  Prints one line per rank and exits non-zero
on any difference.   torchrun --nproc-per-node N scripts/check_slabs_multigpu.py [nz]"""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sw4lite_b200.setup import CartesianProblem
from sw4lite_b200.slabs import SlabStepper
from sw4lite_b200 import lib as L


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nz = int(sys.argv[1]) if len(sys.argv) > 1 else 24 * world
    nx = int(sys.argv[2]) if len(sys.argv) > 2 else 70
    prob = CartesianProblem(nx, 45, nz, h=100.0, gp=7, corder=1, layers=[(1500.0, 6000.0, 3464.0, 2700.0)])
    prob.add_point_force(30, 20, 8, (1e12, 2e12, -1e12), freq=2.0)
    prob.add_point_force(40, 25, nz // 2 + 1, (-2e12, 1e12, 1e12), freq=3.0)
    prob.add_point_force(35, 22, nz - 9, (1e12, 1e12, 1e12), freq=2.5)
    r = np.random.default_rng(5)
    u0 = r.uniform(-1e-3, 1e-3, 3 * prob.npts); um0 = u0 + r.uniform(-1e-5, 1e-5, 3 * prob.npts)
    nij = prob.ni * prob.nj
    full = lambda a: a.reshape(3, prob.nk, nij)
    L.init(local); L.comm_init(rank, world)
    blk = prob.make_block(device=local, rank=rank, nranks=world, comm=True)
    k0 = blk.bounds[4] - prob.bounds[4]
    blk.upload("U", np.ascontiguousarray(full(u0)[:, k0:k0 + blk.nk]).ravel())
    blk.upload("Um", np.ascontiguousarray(full(um0)[:, k0:k0 + blk.nk]).ravel())
    stepper = SlabStepper(blk, None)
    nsteps = 6
    t = 0.0
    for s in range(nsteps):
        stepper.step(prob.forces(t)[blk.src_sel], prob.forces(t, tt=True)[blk.src_sel])
        t += prob.dt
    blk.sync()
    mine = blk.download("U").reshape(3, blk.nk, nij)[:, 2:-2]
    # the undivided run, on every rank's own GPU (small problem)
    whole = prob.make_block(device=local)
    whole.upload("U", u0); whole.upload("Um", um0)
    t = 0.0
    for s in range(nsteps):
        whole.step(prob.forces(t), prob.forces(t, tt=True))
        t += prob.dt
    ref = full(whole.download("U"))[:, k0 + 2:k0 + blk.nk - 2]
    same = np.array_equal(mine, ref)
    print("rank %d/%d planes %d..%d: %s (max|diff| %.3g, scale %.3g)" % (rank, world, blk.bounds[4] + 2, blk.bounds[5] - 2,
          "bit-identical" if same else "DIFFERENT", np.abs(mine - ref).max(), np.abs(ref).max()), flush=True)
    ok = torch.tensor([1 if same and np.abs(ref).max() > 0 else 0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(ok.item()) == 1 else 1)


if __name__ == "__main__":
    main()
