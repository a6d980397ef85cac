"""TEST INFRASTRUCTURE ONLY (oracle/): ctypes access to oracle/libsw4oracle.so, the CPU
restatement of the reference algorithm (oracle/sw4_oracle.c).  Same call signatures as the
kernel-level functions of oracle/refshim.py so that tests can run either.  The product package
never imports this."""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libsw4oracle.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lib = None


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            from . import build_port
            build_port.build(verbose=False)
        _lib = C.CDLL(LIB)
    return _lib


def _d(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def _i(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_ip)


def _b(b):
    return [C.c_int(int(x)) for x in b]


def get_stencil_coefficients():
    acof = np.zeros(384); ghcof = np.zeros(6); bope = np.zeros(48); sbop = np.zeros(5)
    lib().oracle_stencil_coefficients(_d(acof), _d(ghcof), _d(bope), _d(sbop))
    return acof, ghcof, bope, sbop


def rhs4sg(corder, b, nk, onesided, acof, bope, ghcof, lu, u, mu, la, h, strx, stry, strz):
    os_ = np.ascontiguousarray(onesided, dtype=np.int32)
    lib().oracle_rhs4sg(C.c_int(corder), *_b(b), C.c_int(nk), _i(os_), _d(acof), _d(bope), _d(ghcof), _d(lu), _d(u),
                        _d(mu), _d(la), C.c_double(h), _d(strx), _d(stry), _d(strz))


def predfort(corder, b, up, u, um, lu, fo, rho, dt2):
    lib().oracle_predfort(C.c_int(corder), *_b(b), _d(up), _d(u), _d(um), _d(lu), _d(fo), _d(rho), C.c_double(dt2))


def corrfort(corder, b, up, lu, fo, rho, dt4):
    lib().oracle_corrfort(C.c_int(corder), *_b(b), _d(up), _d(lu), _d(fo), _d(rho), C.c_double(dt4))


def dpdmtfort(corder, b, up, u, um, u2, dt2i):
    lib().oracle_dpdmtfort(*_b(b), _d(up), _d(u), _d(um), _d(u2), C.c_double(dt2i))


def addsgd(corder, order, b, up, u, um, rho, dcx, dcy, dcz, strx, stry, strz, cox, coy, coz, beta):
    lib().oracle_addsgd(C.c_int(corder), C.c_int(order), *_b(b), _d(up), _d(u), _d(um), _d(rho), _d(dcx), _d(dcy),
                        _d(dcz), _d(strx), _d(stry), _d(strz), _d(cox), _d(coy), _d(coz), C.c_double(beta))


def bcfortsg(corder, b, wind, nx, ny, nz, u, h, bccnd, sbop, mu, la, t, bforce, strx, stry):
    w = np.ascontiguousarray(wind, dtype=np.int32)
    bc = np.ascontiguousarray(bccnd, dtype=np.int32)
    ptrs = (_dp * 6)(*[(_d(x) if x is not None else None) for x in bforce])
    lib().oracle_bcfortsg(C.c_int(corder), *_b(b), _i(w), C.c_int(nx), C.c_int(ny), C.c_int(nz), _d(u),
                          C.c_double(h), _i(bc), _d(sbop), _d(mu), _d(la), ptrs, _d(strx), _d(stry))
