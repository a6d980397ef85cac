"""Check point / restart of device-resident z-slabs in the reference's file format (CheckPoint::write_checkpoint /
read_checkpoint, reference CheckPoint.C:251-440; called at EW.C:2778-2791 and :2403-2415).

File layout (native endianness, written by the reference with one or many writers through Parallel_IO):
    int precision (8)   int ngrids   double time   int cycle   ngrids x 6 int {1, nx, 1, ny, 1, nz}
    then per grid:  Um, U   each nx*ny*nz*3 doubles, interior points only, index order (c, i, j, k) with c fastest
(the reference extracts with Sarray::extract_subarray, Sarray.C:616-633, which addresses the array as (c,i,j,k); its check
points are therefore only meaningful with `developer corder=0` -- the file itself has that order whatever the run uses).
Um is the solution one step before `time`, U the solution at `time`; `cycle` the number of completed steps.

A k-plane of the file is one contiguous run, so every z-slab writes and reads its own planes at their offsets without
any gather: rank 0 writes the header and sizes the file, every rank then writes planes k0..k1 of Um and U straight from its
device block; on restart a rank reads its planes plus the two halo planes per neighbour (which the file holds too, so no
exchange is needed), and the physical ghost points are rebuilt with the boundary-condition kernels exactly as the
reference does after reading (enforceBC on U and Um, EW.C:2410-2414).  Host side only; numpy.
"""
import os
import struct
import numpy as np


def _header_bytes(ng):
    return 3 * 4 + 8 + ng * 6 * 4


def write_header(path, time, cycle, sizes):
    """sizes: [(nx, ny, nz)] per grid.  Creates the file at its final size."""
    ng = len(sizes)
    with open(path, "wb") as f:
        f.write(struct.pack("=iidi", 8, ng, float(time), int(cycle)))
        for (nx, ny, nz) in sizes:
            f.write(struct.pack("=6i", 1, nx, 1, ny, 1, nz))
        total = _header_bytes(ng) + sum(2 * 3 * 8 * nx * ny * nz for (nx, ny, nz) in sizes)
        f.truncate(total)


def read_header(path):
    with open(path, "rb") as f:
        prec, ng, time, cycle = struct.unpack("=iidi", f.read(20))
        if prec != 8:
            raise ValueError("check point %s: precision %d (only double is written by the reference)" % (path, prec))
        sizes = []
        for _ in range(ng):
            g = struct.unpack("=6i", f.read(24))
            sizes.append((g[1] - g[0] + 1, g[3] - g[2] + 1, g[5] - g[4] + 1))
    return time, cycle, sizes


def _offset(sizes, g, which, k):
    """byte offset of interior plane k (1-based) of array `which` (0: Um, 1: U) of grid g"""
    off = _header_bytes(len(sizes))
    for (nx, ny, nz) in sizes[:g]:
        off += 2 * 3 * 8 * nx * ny * nz
    nx, ny, nz = sizes[g]
    return off + which * 3 * 8 * nx * ny * nz + (k - 1) * 3 * 8 * nx * ny


def write_planes(path, sizes, g, which, k0, planes):
    """planes: (3, n, ny, nx) interior values of planes k0..k0+n-1 -> the file"""
    a = np.ascontiguousarray(np.moveaxis(np.asarray(planes, dtype=np.float64), 0, -1))   # (n, ny, nx, 3)
    with open(path, "r+b") as f:
        f.seek(_offset(sizes, g, which, k0))
        f.write(a.tobytes())


def read_planes(path, sizes, g, which, k0, n):
    nx, ny, nz = sizes[g]
    with open(path, "rb") as f:
        f.seek(_offset(sizes, g, which, k0))
        a = np.frombuffer(f.read(3 * 8 * nx * ny * n), dtype=np.float64).reshape(n, ny, nx, 3)
    return np.ascontiguousarray(np.moveaxis(a, -1, 0))


def _field(blk, name, corder):
    """download a 3-component field of a block as (3, nk, nj, ni)"""
    a = blk.download(name)
    if corder:
        return a.reshape(3, blk.nk, blk.nj, blk.ni)
    return np.moveaxis(a.reshape(blk.nk, blk.nj, blk.ni, 3), 3, 0)


def _pack(full, corder):
    if corder:
        return np.ascontiguousarray(full).ravel()
    return np.ascontiguousarray(np.moveaxis(full, 0, 3)).ravel()


def save_slab(blk, prob, path, time, cycle, rank=0, barrier=None):
    """write the owned planes of Um and U of this slab's block.  Call on every rank after a completed step (the solution at
    `time` is the block's U).  barrier: callable that synchronises the ranks (rank 0 creates the file first)."""
    sizes = [(prob.nx, prob.ny, prob.nz)]
    if rank == 0:
        write_header(path, time, cycle, sizes)
    if barrier is not None:
        barrier()
    k0, k1 = blk.bounds[4] + 2, blk.bounds[5] - 2             # owned interior planes
    for which, name in ((0, "Um"), (1, "U")):
        f = _field(blk, name, prob.corder)
        write_planes(path, sizes, 0, which, k0, f[:, 2:-2, 2:-2, 2:-2])
    if barrier is not None:
        barrier()


def load_slab(blk, prob, path):
    """restore U and Um of this slab's block from a check point: own planes and the halo planes come from the file, the
    physical ghost points from the boundary-condition kernels.  Returns (time, cycle)."""
    time, cycle, sizes = read_header(path)
    if sizes[0] != (prob.nx, prob.ny, prob.nz):
        raise ValueError("check point %s holds a %s grid, the problem is %s" % (path, sizes[0], (prob.nx, prob.ny, prob.nz)))
    kb, ke = blk.bounds[4], blk.bounds[5]
    ka, kz = max(kb, 1), min(ke, prob.nz)                     # planes of the block that exist in the file
    for which in (0, 1):                                      # Um first, then U: each goes through Up -> BC -> cycle
        full = np.zeros((3, blk.nk, blk.nj, blk.ni))
        full[:, ka - kb:kz - kb + 1, 2:-2, 2:-2] = read_planes(path, sizes, 0, which, ka, kz - ka + 1)
        blk.upload("Up", _pack(full, prob.corder))
        blk.enforce_bc()
        blk.cycle()
    return time, cycle
