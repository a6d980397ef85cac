"""Host-side mirror of the reference's time-step driver for this path: EW::timesteploop
(reference EW.C:2339-2928) restricted to what happens between "fields are on the device" and
"cycle the solution arrays".  All compute is in libsw4b200.so (C-ABI, include/sw4b200.h); this
module only sequences calls and moves small host arrays (source amplitudes, receiver samples).
"""
import ctypes as C
import numpy as np

from . import lib as L

# boundaryConditionType of the reference (src/double/sw4.h:37)
bStressFree, bDirichlet, bSuperGrid, bPeriodic, bCCInterface, bRefInterface, bAEInterface, bProcessor, bNone = range(9)

_dp = L.c_dp


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def boundary_windows(bounds, bctype):
    """m_BndryWindow as EW::setup_boundary_arrays builds it (EW.C:3347-3420): stress-free sides
    hold the single boundary plane, Dirichlet/supergrid/periodic sides the two ghost layers."""
    wind = np.zeros(36, dtype=np.int32)
    for s in range(6):
        w = [999, -999, 999, -999, 999, -999]
        if bctype[s] in (bStressFree, bDirichlet, bSuperGrid, bPeriodic):
            w = list(bounds)
            lo = 2 * (s // 2)
            if bctype[s] == bStressFree:
                w[lo] = w[lo + 1] = (bounds[lo] + 2) if s % 2 == 0 else (bounds[lo + 1] - 2)
            elif s % 2 == 0:
                w[lo + 1] = w[lo] + 1
            else:
                w[lo] = w[lo + 1] - 1
        wind[6 * s:6 * s + 6] = w
    return wind


class GridBlock:
    """device-resident state of one grid block (sw4b200_grid)"""

    def __init__(self, corder, bounds, nglobal, h, dt, onesided, bctype, wind=None, sg_order=4, beta=0.0,
                 curvilinear=False, halo_lo=False, halo_hi=False, device=0):
        self.lib = L.init(device)
        d = L.GridDesc()
        d.corder = int(corder)
        (d.ifirst, d.ilast, d.jfirst, d.jlast, d.kfirst, d.klast) = [int(x) for x in bounds]
        d.nx, d.ny, d.nz = [int(x) for x in nglobal]
        d.h, d.dt = float(h), float(dt)
        if wind is None:
            wind = boundary_windows(bounds, bctype)
        for s in range(6):
            d.onesided[s] = int(onesided[s]); d.bctype[s] = int(bctype[s])
        for s in range(36):
            d.wind[s] = int(wind[s])
        d.sg_order, d.beta = int(sg_order), float(beta)
        d.curvilinear = int(bool(curvilinear)); d.halo_lo = int(bool(halo_lo)); d.halo_hi = int(bool(halo_hi))
        self.desc = d
        self.bounds = tuple(int(x) for x in bounds)
        self.ni = d.ilast - d.ifirst + 1; self.nj = d.jlast - d.jfirst + 1; self.nk = d.klast - d.kfirst + 1
        self.npts = self.ni * self.nj * self.nk
        self.h = self.lib.sw4b200_grid_create(C.byref(d))
        if not self.h:
            raise L.Sw4b200Error(self.lib.sw4b200_last_error().decode())
        self.nsrc = 0
        self.nrec = 0

    def close(self):
        if self.h:
            self.lib.sw4b200_grid_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def size(self, name):
        return int(self.lib.sw4b200_grid_array_size(self.h, name.encode()))

    def upload(self, name, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        n = self.size(name)
        if a.size != n:
            raise ValueError("array '%s' has %d values, the block needs %d" % (name, a.size, n))
        L.check(self.lib.sw4b200_grid_upload(self.h, name.encode(), _d(a)))

    def download(self, name):
        a = np.empty(self.size(name))
        L.check(self.lib.sw4b200_grid_download(self.h, name.encode(), _d(a)))
        return a

    def device_ptr(self, name):
        return self.lib.sw4b200_grid_device_ptr(self.h, name.encode())

    def set_source_points(self, ijk):
        a = np.ascontiguousarray(ijk, dtype=np.int32).reshape(-1, 3)
        self.nsrc = len(a)
        L.check(self.lib.sw4b200_grid_set_source_points(self.h, self.nsrc, a.ctypes.data_as(L.c_ip)))

    def set_receiver_points(self, ijk):
        a = np.ascontiguousarray(ijk, dtype=np.int32).reshape(-1, 3)
        self.nrec = len(a)
        L.check(self.lib.sw4b200_grid_set_receiver_points(self.h, self.nrec, a.ctypes.data_as(L.c_ip)))

    def predictor(self, f=None):
        f = np.ascontiguousarray(f, dtype=np.float64) if (f is not None and self.nsrc) else None
        L.check(self.lib.sw4b200_grid_predictor(self.h, _d(f)))

    def enforce_bc(self):
        L.check(self.lib.sw4b200_grid_enforce_bc(self.h))

    def corrector(self, ftt=None):
        ftt = np.ascontiguousarray(ftt, dtype=np.float64) if (ftt is not None and self.nsrc) else None
        L.check(self.lib.sw4b200_grid_corrector(self.h, _d(ftt)))

    def cycle(self):
        L.check(self.lib.sw4b200_grid_cycle(self.h))

    def predictor_part(self, part, f=None):
        f = np.ascontiguousarray(f, dtype=np.float64) if (f is not None and self.nsrc) else None
        L.check(self.lib.sw4b200_grid_predictor_part(self.h, int(part), _d(f)))

    def corrector_part(self, part, ftt=None):
        ftt = np.ascontiguousarray(ftt, dtype=np.float64) if (ftt is not None and self.nsrc) else None
        L.check(self.lib.sw4b200_grid_corrector_part(self.h, int(part), _d(ftt)))

    def fill_profile(self, name, kvalues):
        a = np.ascontiguousarray(kvalues, dtype=np.float64)
        if a.size != self.nk:
            raise ValueError("profile of '%s' has %d values, the block has %d planes" % (name, a.size, self.nk))
        L.check(self.lib.sw4b200_grid_fill_profile(self.h, name.encode(), _d(a)))

    def set_source_series(self, f_all, ftt_all):
        """f_all, ftt_all: (nsteps, nsrc, 3) source amplitudes of every step, kept on the device"""
        f = np.ascontiguousarray(f_all, dtype=np.float64); ftt = np.ascontiguousarray(ftt_all, dtype=np.float64)
        nsteps = f.shape[0] if self.nsrc else 0
        L.check(self.lib.sw4b200_grid_set_source_series(self.h, nsteps, _d(f), _d(ftt)))

    def run(self, first_step, nsteps):
        """nsteps whole time steps with no host synchronisation (sources/receivers device resident)"""
        L.check(self.lib.sw4b200_grid_run(self.h, int(first_step), int(nsteps)))

    def fetch_records(self, first_step, nsteps):
        out = np.zeros((nsteps, max(self.nrec, 1), 3))
        if self.nrec:
            L.check(self.lib.sw4b200_grid_fetch_records(self.h, int(first_step), int(nsteps), _d(out)))
        return out[:, :self.nrec]

    def record_resident(self, step):
        """sample the receivers from the new solution of `step` into the device-resident record (no host sync)"""
        L.check(self.lib.sw4b200_grid_record_resident(self.h, int(step)))

    def set_stream(self, st):
        L.check(self.lib.sw4b200_grid_set_stream(self.h, int(st)))

    # ---- z-slab halo planes (see slabs.py).  torch is used for device buffers/streams only.
    def pack(self, side, tensor, stream=None, with_acc=False):
        """the two interior planes next to face `side` (0 low-k, 1 high-k) of Up (+ the stored
        acceleration if with_acc) -> tensor"""
        L.check(self.lib.sw4b200_grid_pack_halo(self.h, int(side), int(with_acc), C.c_void_p(tensor.data_ptr()), stream))

    def unpack(self, side, tensor, stream=None, with_acc=False):
        L.check(self.lib.sw4b200_grid_unpack_halo(self.h, int(side), int(with_acc), C.c_void_p(tensor.data_ptr()), stream))

    def halo_doubles(self, with_acc):
        return int(self.lib.sw4b200_grid_halo_doubles(self.h, int(with_acc)))

    # ---- halo exchange inside the library (csrc/exchange.cu): grouped ncclSend/ncclRecv straight from / into the field
    # arrays on the library's communication stream.  `ex` is unused here (the CPU stand-in of the gloo tests needs it).
    def set_neighbours(self, rank_lo, rank_hi):
        L.check(self.lib.sw4b200_grid_set_neighbours(self.h, -1 if rank_lo is None else int(rank_lo), -1 if rank_hi is None else int(rank_hi)))

    def begin_exchange(self, ex=None, with_acc=False):
        """start moving the face planes of Up (and of the stored acceleration) once the face rows queued so far are done;
        the caller goes on launching the bulk rows on the compute stream"""
        L.check(self.lib.sw4b200_grid_exchange_begin(self.h, int(with_acc)))

    def end_exchange(self, ex=None, with_acc=False):
        L.check(self.lib.sw4b200_grid_exchange_end(self.h))

    def record(self):
        out = np.zeros(3 * max(self.nrec, 1))
        if self.nrec:
            L.check(self.lib.sw4b200_grid_record(self.h, _d(out)))
        return out[:3 * self.nrec].reshape(-1, 3)

    def step(self, f=None, ftt=None, record=False):
        """one full time step of a block without neighbours (EW.C:2527-2842)"""
        self.predictor(f)
        self.enforce_bc()
        self.corrector(ftt)
        self.enforce_bc()
        rec = self.record() if record else None
        self.cycle()
        return rec

    def sync(self):
        L.check(self.lib.sw4b200_grid_sync(self.h))


class GridStack:
    """All grid blocks of a run that live on one device, ordered like the reference's per-grid vectors
    (mU[g], EW.h): Cartesian grids first, the curvilinear grid under the topography last.  Sequences one
    time step over them the way EW::timesteploop does (EW.C:2527-2842): predictor on every grid, boundary
    conditions + the Cartesian/curvilinear interface injection (EW::enforceBC ends with enforceCartTopo,
    EW.C:3500), corrector on every grid, boundary conditions again, cycle."""

    def __init__(self, blocks, ncart=None):
        self.blocks = list(blocks)
        self.ncart = len(self.blocks) if ncart is None else int(ncart)
        self.topo = self.ncart < len(self.blocks)
        if self.topo and (self.ncart < 1 or len(self.blocks) != self.ncart + 1):
            raise ValueError("a topography run has its Cartesian grids followed by exactly one curvilinear grid")
        self.lib = self.blocks[0].lib

    def enforce_bc(self):
        for b in self.blocks:
            b.enforce_bc()
        if self.topo:
            L.check(self.lib.sw4b200_grid_enforce_cart_topo(self.blocks[self.ncart - 1].h, self.blocks[self.ncart].h))

    def step(self, f=None, ftt=None, record=False):
        """f, ftt: per-grid lists of source amplitudes (or None)"""
        n = len(self.blocks)
        f = f if f is not None else [None] * n
        ftt = ftt if ftt is not None else [None] * n
        for g, b in enumerate(self.blocks):
            b.predictor(f[g])
        self.enforce_bc()
        for g, b in enumerate(self.blocks):
            b.corrector(ftt[g])
        self.enforce_bc()
        rec = [b.record() for b in self.blocks] if record else None
        for b in self.blocks:
            b.cycle()
        return rec

    def sync(self):
        for b in self.blocks:
            b.sync()
