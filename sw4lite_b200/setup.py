"""Problem setup for synthetic runs (bench.py, smoke, examples): what the reference's host code
(parser + setupRun, EW.C:1865-2146, 4636-4868, 5041-5146; SuperGrid.C:108-198) hands to the time
loop, restated for Cartesian single-grid problems.  Host-side numpy only; nothing here is on the
hot path.  Parity of these arrays (supergrid dc/str/corner, dt, windows) with the reference's own set-up is checked in
tests/test_setup.py against oracle/_ref."""
import numpy as np

from .solver import bStressFree, bSuperGrid, bProcessor, boundary_windows


def _psi0(xi):
    """C5 polynomial taper, SuperGrid::Psi0 (SuperGrid.C:170-192)"""
    xi = np.asarray(xi, dtype=np.float64)
    # (products left to right, as the reference writes them: the stretching 1-(1-epsL)*psi amplifies the last bits of psi)
    f = xi * xi * xi * xi * xi * xi * (462 - 1980 * xi + 3465 * xi * xi - 3080 * xi * xi * xi + 1386 * xi * xi * xi * xi
                                       - 252 * xi * xi * xi * xi * xi)
    return np.where(xi <= 0, 0.0, np.where(xi >= 1, 1.0, f))


def supergrid_1d(x, left, right, x0, x1, width, epsL=1e-4, cmin=0.33):
    """(dc, str, corner) at coordinates x: SuperGrid::dampingCoeff / stretching / cornerTaper
    (SuperGrid.C:108-198) with trans_width = width/2."""
    x = np.asarray(x, dtype=np.float64)
    tw = 0.5 * width
    psi = np.zeros_like(x); damp = np.zeros_like(x); lin = np.zeros_like(x)
    if left:
        m = x < x0 + width
        psi = np.where(m, _psi0((x0 + width - x) / width), psi)
        damp = np.where(m, _psi0((x0 + width - x) / tw), damp)
        lin = np.where(m, (x0 + width - x) / width, lin)
    if right:
        m = (x > x1 - width) & ~((x < x0 + width) if left else np.zeros_like(x, dtype=bool))
        psi = np.where(m, _psi0((x - (x1 - width)) / width), psi)
        damp = np.where(m, _psi0((x - (x1 - width)) / tw), damp)
        lin = np.where(m, (x - (x1 - width)) / width, lin)
    stretch = 1 - (1 - epsL) * psi
    return damp / stretch, stretch, 1.0 - (1.0 - cmin) * lin


def c6smoothbump(freq, t):
    x = t * freq
    return np.where((x < 0) | (x > 1), 0.0, 51480 * (x * (1 - x)) ** 7)


def c6smoothbump_tt(freq, t):
    x = t * freq
    v = 51480 * freq * freq * 7 * (6 * (1 - 2 * x) ** 2 * (x * (1 - x)) ** 5 - 2 * (x * (1 - x)) ** 6)
    return np.where((x < 0) | (x > 1), 0.0, v)


class CartesianProblem:
    """A single Cartesian grid: free surface on top (k=1), supergrid layers on the other five
    sides (the reference's default_bcs, EW.C:6077-6082, + `supergrid gp=`), uniform or layered
    material, point forces with a C6SmoothBump time function."""

    def __init__(self, nx, ny, nz, h, vp=4000.0, vs=2000.0, rho=2600.0, gp=30, cfl=1.3, beta=0.02, corder=1,
                 free_surface=True, layers=None, free_bottom=False):
        self.nx, self.ny, self.nz, self.h = nx, ny, nz, float(h)
        self.corder = corder
        self.bounds = (-1, nx + 2, -1, ny + 2, -1, nz + 2)
        self.ni, self.nj, self.nk = nx + 4, ny + 4, nz + 4
        self.npts = self.ni * self.nj * self.nk
        self.bctype = [bSuperGrid] * 6
        if free_surface:
            self.bctype[4] = bStressFree
        if free_bottom:                 # (a stress-free bottom: the SBP closure of side 5, rhs4sg_rev.C:602-855)
            self.bctype[5] = bStressFree
        self.onesided = [0, 0, 0, 0, 1 if free_surface else 0, 1 if free_bottom else 0]
        self.wind = boundary_windows(self.bounds, self.bctype)
        self.beta = beta
        self.gp = gp
        # material (ghost points get the same constant; `layers` = [(z_top, vp, vs, rho), ...])
        z = (np.arange(-1, nz + 3) - 1) * self.h
        vpk = np.full(self.nk, float(vp)); vsk = np.full(self.nk, float(vs)); rhk = np.full(self.nk, float(rho))
        for (ztop, lvp, lvs, lrho) in (layers or []):
            m = z >= ztop
            vpk[m], vsk[m], rhk[m] = lvp, lvs, lrho
        muk = rhk * vsk ** 2
        lak = rhk * vpk ** 2 - 2 * muk
        # the medium is layered: keep the per-plane profiles, materialise 3-D fields only on demand
        self.muk, self.lak, self.rhk = muk, lak, rhk
        # supergrid 1-D arrays (EW::assign_supergrid_damping_arrays, EW.C:4671-4801)
        width = gp * self.h
        xs = (np.arange(-1, nx + 3) - 1) * self.h
        ys = (np.arange(-1, ny + 3) - 1) * self.h
        self.dcx, self.strx, self.cox = supergrid_1d(xs, True, True, 0.0, (nx - 1) * self.h, width)
        self.dcy, self.stry, self.coy = supergrid_1d(ys, True, True, 0.0, (ny - 1) * self.h, width)
        self.dcz, self.strz, self.coz = supergrid_1d(z, not free_surface, not free_bottom, 0.0, (nz - 1) * self.h, width)
        # time step (EW::computeDT, EW.C:5041-5066): dt = cfl*h/sqrt(max (4mu+la)/rho)
        self.dt = cfl * self.h / np.sqrt(np.max((4 * muk + lak) / rhk))
        self.sources = []  # (i,j,k, fx,fy,fz, freq, t0)

    def set_end_time(self, tmax):
        """`time t=`: the reference shortens dt so that a whole number of steps reaches tmax (EW.C:5138-5145)"""
        self.nsteps = max(1, int(tmax / self.dt + 0.5))     # (rounded to the nearest integer: dt may grow slightly)
        self.dt = tmax / self.nsteps
        return self.nsteps

    @classmethod
    def loh1(cls, h, corder=1):
        """the LOH.1 layer-over-half-space benchmark of the reference (tests/loh1/LOH.1-h100.in, -h50.in: BASELINE.json config 3):
        30 km x 30 km x 17 km, free surface, supergrid gp=30 on the other five sides, 1 km layer over the half-space with the
        averaged interface block at z=1000, t=9 s.  The discretised moment source and the receiver come from
        tests/golden/loh1-h*-setup.npz (generated by the reference's own source discretisation)."""
        nx = int(round(30000.0 / h)) + 1
        nz = int(round(17000.0 / h)) + 1
        prob = cls(nx, nx, nz, h=h, vp=4000.0, vs=2000.0, rho=2600.0, gp=30, cfl=1.3, beta=0.02, corder=corder,
                   layers=[(999.0, 4630.76, 2437.56, 2650.0), (1001.0 + 1e-6, 6000.0, 3464.0, 2700.0)])
        prob.set_end_time(9.0)
        return prob

    @property
    def mu(self):
        return np.repeat(self.muk, self.ni * self.nj)

    @property
    def la(self):
        return np.repeat(self.lak, self.ni * self.nj)

    @property
    def rho(self):
        return np.repeat(self.rhk, self.ni * self.nj)

    def add_point_force(self, i, j, k, f, freq, t0=0.0):
        self.sources.append((int(i), int(j), int(k), float(f[0]), float(f[1]), float(f[2]), float(freq), float(t0)))

    def source_points(self):
        return np.array([[s[0], s[1], s[2]] for s in self.sources], dtype=np.int32).reshape(-1, 3)

    def forces(self, t, tt=False):
        out = np.zeros((len(self.sources), 3))
        for n, s in enumerate(self.sources):
            g = (c6smoothbump_tt if tt else c6smoothbump)(s[6], t - s[7])
            out[n] = np.array(s[3:6]) * g
        return out

    def slab(self, rank, nranks):
        """z-slab `rank` of `nranks`: interior planes split evenly (the reference's decomp1d rule,
        EW.C:2931-2985, applied to k), 2 halo planes towards each neighbour.  Returns
        (bounds, onesided, bctype, halo_lo, halo_hi)."""
        from .slabs import slab_range
        k0, k1 = slab_range(self.nz, rank, nranks)
        bounds = list(self.bounds); bounds[4], bounds[5] = k0 - 2, k1 + 2
        onesided = list(self.onesided); bctype = list(self.bctype)
        halo_lo, halo_hi = rank > 0, rank < nranks - 1
        if halo_lo:
            onesided[4] = 0; bctype[4] = bProcessor
        if halo_hi:
            onesided[5] = 0; bctype[5] = bProcessor
        return tuple(bounds), onesided, bctype, halo_lo, halo_hi

    def make_block(self, device=0, rank=0, nranks=1, comm=False):
        """comm: the library's communicator spans the `nranks` slabs (lib.comm_init): register the z-neighbours"""
        from .solver import GridBlock
        bounds, onesided, bctype, halo_lo, halo_hi = self.slab(rank, nranks)
        g = GridBlock(self.corder, bounds, (self.nx, self.ny, self.nz), self.h, self.dt, onesided,
                      bctype, boundary_windows(bounds, bctype), sg_order=4, beta=self.beta, device=device,
                      halo_lo=halo_lo, halo_hi=halo_hi)
        k0 = bounds[4] - self.bounds[4]
        ks = slice(k0, k0 + g.nk)
        for name in ("strx", "stry", "dcx", "dcy", "cox", "coy"):
            g.upload(name, getattr(self, name))
        for name in ("strz", "dcz", "coz"):
            g.upload(name, getattr(self, name)[ks])
        g.fill_profile("mu", self.muk[ks]); g.fill_profile("lambda", self.lak[ks]); g.fill_profile("rho", self.rhk[ks])
        if comm and nranks > 1:
            g.set_neighbours(rank - 1 if halo_lo else None, rank + 1 if halo_hi else None)
        g.src_sel = [n for n, s in enumerate(self.sources) if bounds[4] + 2 <= s[2] <= bounds[5] - 2]
        if g.src_sel:
            g.set_source_points(self.source_points()[g.src_sel])
        return g
