"""TEST INFRASTRUCTURE ONLY (oracle/): ctypes access to oracle/_ref/libsw4ref.so, the
UNMODIFIED reference CPU kernels + a steppable reference `EW` object (see oracle/ref_shim.C).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (sw4lite_b200/) never does.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libsw4ref.so")
EXE = os.path.join(HERE, "_ref", "sw4lite_ref")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def available():
    return os.path.exists(LIB)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB)
        _lib.ref_ew_create.restype = C.c_void_p
        _lib.ref_ew_create.argtypes = [C.c_char_p, C.c_char_p]
        _lib.ref_ew_int.argtypes = [C.c_void_p, C.c_char_p]
        _lib.ref_ew_double.restype = C.c_double
        _lib.ref_ew_double.argtypes = [C.c_void_p, C.c_char_p]
        _lib.ref_ew_grid_ints.argtypes = [C.c_void_p, C.c_int, _ip]
        _lib.ref_ew_grid_h.restype = C.c_double
        _lib.ref_ew_grid_h.argtypes = [C.c_void_p, C.c_int]
        _lib.ref_ew_grid_zmin.restype = C.c_double
        _lib.ref_ew_grid_zmin.argtypes = [C.c_void_p, C.c_int]
        _lib.ref_ew_array.restype = C.c_void_p
        _lib.ref_ew_array.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        _lib.ref_ew_point_sources.argtypes = [C.c_void_p, _ip, _dp]
        _lib.ref_ew_identsources.argtypes = [C.c_void_p, _ip]
        _lib.ref_ew_eval_forces.argtypes = [C.c_void_p, C.c_double, C.c_int, _dp]
        _lib.ref_ew_receivers.argtypes = [C.c_void_p, _ip, _ip]
        _lib.ref_ew_bc_forcing.argtypes = [C.c_void_p, C.c_double]
        _lib.ref_ew_step_phases.argtypes = [C.c_void_p, C.c_int, C.c_int]
        _lib.ref_ew_step.argtypes = [C.c_void_p]
        _lib.ref_ew_cycle.argtypes = [C.c_void_p]
        _lib.ref_ew_pointsource_error.argtypes = [C.c_void_p, C.c_double, C.POINTER(_dp), _dp]
        _lib.ref_ew_write_receivers.argtypes = [C.c_void_p, C.c_char_p]
    return _lib


def _d(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def _i(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_ip)


def num_threads():
    return lib().ref_num_threads()


def set_num_threads(n):
    """OpenMP team size of the reference kernels (omp_set_num_threads)"""
    lib().ref_set_num_threads(C.c_int(int(n)))


# ---------------------------------------------------------------- kernel level
def get_stencil_coefficients():
    acof = np.zeros(384); ghcof = np.zeros(6); bope = np.zeros(48); sbop = np.zeros(5)
    lib().ref_get_stencil_coefficients(_d(acof), _d(ghcof), _d(bope), _d(sbop))
    return acof, ghcof, bope, sbop


def rhs4sg(corder, b, nk, onesided, acof, bope, ghcof, lu, u, mu, la, h, strx, stry, strz):
    os_ = np.ascontiguousarray(onesided, dtype=np.int32)
    lib().ref_rhs4sg(C.c_int(corder), *[C.c_int(int(x)) for x in b], C.c_int(nk), _i(os_), _d(acof), _d(bope),
                     _d(ghcof), _d(lu), _d(u), _d(mu), _d(la), C.c_double(h), _d(strx), _d(stry), _d(strz))


def rhs4sgcurv(corder, b, u, mu, la, met, jac, lu, onesided, acof, bope, ghcof, strx, stry):
    os_ = np.ascontiguousarray(onesided, dtype=np.int32)
    lib().ref_rhs4sgcurv(C.c_int(corder), *[C.c_int(int(x)) for x in b], _d(u), _d(mu), _d(la), _d(met), _d(jac),
                         _d(lu), _i(os_), _d(acof), _d(bope), _d(ghcof), _d(strx), _d(stry))


def predfort(corder, b, up, u, um, lu, fo, rho, dt2):
    lib().ref_predfort(C.c_int(corder), *[C.c_int(int(x)) for x in b], _d(up), _d(u), _d(um), _d(lu), _d(fo),
                       _d(rho), C.c_double(dt2))


def corrfort(corder, b, up, lu, fo, rho, dt4):
    lib().ref_corrfort(C.c_int(corder), *[C.c_int(int(x)) for x in b], _d(up), _d(lu), _d(fo), _d(rho),
                       C.c_double(dt4))


def dpdmtfort(corder, b, up, u, um, u2, dt2i):
    lib().ref_dpdmtfort(C.c_int(corder), *[C.c_int(int(x)) for x in b], _d(up), _d(u), _d(um), _d(u2),
                        C.c_double(dt2i))


def addsgd(corder, order, b, up, u, um, rho, dcx, dcy, dcz, strx, stry, strz, cox, coy, coz, beta):
    lib().ref_addsgd(C.c_int(corder), C.c_int(order), *[C.c_int(int(x)) for x in b], _d(up), _d(u), _d(um), _d(rho),
                     _d(dcx), _d(dcy), _d(dcz), _d(strx), _d(stry), _d(strz), _d(cox), _d(coy), _d(coz),
                     C.c_double(beta))


def addsgdc(corder, order, b, up, u, um, rho, dcx, dcy, strx, stry, jac, cox, coy, beta):
    lib().ref_addsgdc(C.c_int(corder), C.c_int(order), *[C.c_int(int(x)) for x in b], _d(up), _d(u), _d(um),
                      _d(rho), _d(dcx), _d(dcy), _d(strx), _d(stry), _d(jac), _d(cox), _d(coy), C.c_double(beta))


def bcfortsg(corder, b, wind, nx, ny, nz, u, h, bccnd, sbop, mu, la, t, bforce, strx, stry):
    w = np.ascontiguousarray(wind, dtype=np.int32)
    bc = np.ascontiguousarray(bccnd, dtype=np.int32)
    bf = [(_d(x) if x is not None else None) for x in bforce]
    lib().ref_bcfortsg(C.c_int(corder), *[C.c_int(int(x)) for x in b], _i(w), C.c_int(nx), C.c_int(ny), C.c_int(nz),
                       _d(u), C.c_double(h), _i(bc), _d(sbop), _d(mu), _d(la), C.c_double(t), *bf, _d(strx), _d(stry))


def freesurfcurvisg(corder, b, nz, side, u, mu, la, met, sbop, forcing, strx, stry):
    lib().ref_freesurfcurvisg(C.c_int(corder), *[C.c_int(int(x)) for x in b], C.c_int(nz), C.c_int(side), _d(u),
                              _d(mu), _d(la), _d(met), _d(sbop), _d(forcing), _d(strx), _d(stry))


# ---------------------------------------------------------------- EW level
class RefGrid:
    pass


class RefEW:
    """A reference EW object set up by the reference's own parser/setupRun from an .in file."""

    def __init__(self, infile, workdir):
        self.h = lib().ref_ew_create(os.path.abspath(infile).encode(), os.path.abspath(workdir).encode())
        if not self.h:
            raise RuntimeError("ref_ew_create failed")
        self.workdir = os.path.abspath(workdir)
        gi = lambda w: lib().ref_ew_int(self.h, w.encode())
        gd = lambda w: lib().ref_ew_double(self.h, w.encode())
        self.ngrids = gi("ngrids"); self.ncart = gi("ncart"); self.nsteps = gi("nsteps")
        self.corder = gi("corder"); self.topo = gi("topo"); self.sgorder = gi("sgorder")
        self.usesg = gi("usesg"); self.pointsourcetest = gi("pointsourcetest")
        self.dt = gd("dt"); self.tstart = gd("tstart"); self.beta = gd("beta")
        self.grids = []
        for g in range(self.ngrids):
            ints = np.zeros(9 + 6 + 6 + 36 + 6, dtype=np.int32)
            lib().ref_ew_grid_ints(self.h, g, _i(ints))
            G = RefGrid()
            G.bounds = tuple(int(x) for x in ints[0:6])
            G.nx, G.ny, G.nz = (int(x) for x in ints[6:9])
            G.onesided = ints[9:15].copy(); G.bctype = ints[15:21].copy()
            G.wind = ints[21:57].copy(); G.nbcpts = ints[57:63].copy()
            G.h = lib().ref_ew_grid_h(self.h, g); G.zmin = lib().ref_ew_grid_zmin(self.h, g)
            ib, ie, jb, je, kb, ke = G.bounds
            G.ni, G.nj, G.nk = ie - ib + 1, je - jb + 1, ke - kb + 1
            G.npts = G.ni * G.nj * G.nk
            self.grids.append(G)
        self.acof = self._arr("acof", 0, 384).copy(); self.bope = self._arr("bope", 0, 48).copy()
        self.ghcof = self._arr("ghcof", 0, 6).copy(); self.sbop = self._arr("sbop", 0, 5).copy()

    def _arr(self, name, g, n):
        p = lib().ref_ew_array(self.h, name.encode(), g)
        if not p:
            return None
        return np.ctypeslib.as_array(C.cast(p, _dp), shape=(n,))

    def array(self, name, g=0):
        """numpy VIEW of a reference-owned host array (re-fetch after every step: pointers rotate)."""
        G = self.grids[g]
        n = {"U": 3, "Um": 3, "Up": 3, "F": 3, "Lu": 3, "Uacc": 3, "mu": 1, "lambda": 1, "rho": 1,
             "metric": 4, "jac": 1}.get(name)
        if n is not None:
            return self._arr(name, g, n * G.npts)
        if name in ("strx", "dcx", "cox"):
            return self._arr(name, g, G.ni)
        if name in ("stry", "dcy", "coy"):
            return self._arr(name, g, G.nj)
        if name in ("strz", "dcz", "coz"):
            return self._arr(name, g, G.nk)
        if name.startswith("bforce"):
            s = int(name[6])
            if G.nbcpts[s] == 0:
                return None
            return self._arr(name, g, 3 * int(G.nbcpts[s]))
        raise KeyError(name)

    @property
    def t(self):
        return lib().ref_ew_double(self.h, b"t")

    def point_sources(self):
        n = lib().ref_ew_int(self.h, b"npointsources")
        idx = np.zeros((max(n, 1), 4), dtype=np.int32); f = np.zeros((max(n, 1), 3))
        nu = lib().ref_ew_int(self.h, b"nunique")
        ident = np.zeros(max(nu + 1, 1), dtype=np.int32)
        if n > 0:
            lib().ref_ew_point_sources(self.h, _i(idx), _d(f))
            lib().ref_ew_identsources(self.h, _i(ident))
        return idx[:n], f[:n], ident[:nu + 1] if n > 0 else ident[:0]

    def eval_forces(self, t, tt):
        n = lib().ref_ew_int(self.h, b"npointsources")
        f = np.zeros((max(n, 1), 3))
        if n > 0:
            lib().ref_ew_eval_forces(self.h, C.c_double(t), C.c_int(1 if tt else 0), _d(f))
        return f[:n]

    def receivers(self):
        n = lib().ref_ew_int(self.h, b"nrec")
        out = np.zeros((max(n, 1), 4), dtype=np.int32); mode = np.zeros(max(n, 1), dtype=np.int32)
        if n > 0:
            lib().ref_ew_receivers(self.h, _i(out), _i(mode))
        return out[:n], mode[:n]

    def bc_forcing(self, t):
        lib().ref_ew_bc_forcing(self.h, C.c_double(t))

    def step(self):
        lib().ref_ew_step(self.h)

    def step_phases(self, a, b):
        lib().ref_ew_step_phases(self.h, a, b)

    def cycle(self):
        lib().ref_ew_cycle(self.h)

    def pointsource_error(self, t, u_per_grid):
        arrs = [np.ascontiguousarray(u, dtype=np.float64) for u in u_per_grid]
        ptrs = (_dp * len(arrs))(*[_d(a) for a in arrs])
        out = np.zeros(3)
        rc = lib().ref_ew_pointsource_error(self.h, C.c_double(t), ptrs, _d(out))
        if rc != 0:
            raise RuntimeError("not a point source test")
        return out

    def write_receivers(self):
        return lib().ref_ew_write_receivers(self.h, self.workdir.encode())
