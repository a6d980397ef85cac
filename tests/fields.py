"""Synthetic inputs shared by the parity tests.

`Box` describes one grid block with the reference's Fortran-style inclusive global bounds
(ghost points included).  Fields are flat float64 arrays in the reference layouts:
corder=1 -> (i,j,k,c) component-slowest; corder=0 -> (c,i,j,k) component-fastest.
"""
import numpy as np


class Box:
    def __init__(self, ni, nj, nk, ifirst=-1, jfirst=-1, kfirst=-1):
        self.ifirst, self.jfirst, self.kfirst = ifirst, jfirst, kfirst
        self.ni, self.nj, self.nk = ni, nj, nk
        self.ilast, self.jlast, self.klast = ifirst + ni - 1, jfirst + nj - 1, kfirst + nk - 1
        self.npts = ni * nj * nk

    @property
    def bounds(self):
        return (self.ifirst, self.ilast, self.jfirst, self.jlast, self.kfirst, self.klast)

    def coords(self, h):
        x = (np.arange(self.ifirst, self.ilast + 1) - 1) * h
        y = (np.arange(self.jfirst, self.jlast + 1) - 1) * h
        z = (np.arange(self.kfirst, self.klast + 1) - 1) * h
        return np.meshgrid(z, y, x, indexing="ij")  # arrays shaped (nk,nj,ni): i fastest


def pack3(comps, corder):
    """three (nk,nj,ni) arrays -> flat reference layout"""
    if corder:
        return np.ascontiguousarray(np.stack([c.ravel() for c in comps]).ravel())
    return np.ascontiguousarray(np.stack([c.ravel() for c in comps], axis=1).ravel())


def unpack3(a, box, corder):
    if corder:
        return a.reshape(3, box.nk, box.nj, box.ni)
    return np.moveaxis(a.reshape(box.nk, box.nj, box.ni, 3), 3, 0)


def harness_fields(box, h, corder=1):
    """analytic fields of the reference kernel harness (tests/testil/testil.C:411-420)"""
    z, y, x = box.coords(h)
    la = np.cos(x) * np.sin(3 * y) ** 2 * np.cos(z)
    mu = np.sin(3 * x) * np.sin(y) * np.sin(z)
    rho = x ** 3 + 1 + y ** 2 + z ** 2
    u = np.cos(x * x) * np.sin(x * y) * z * z
    v = np.sin(x) * np.cos(y * y) * np.sin(z)
    w = np.cos(x * y) * np.sin(z * y)
    return dict(u=pack3([u, v, w], corder), mu=mu.ravel().copy(), la=la.ravel().copy(), rho=rho.ravel().copy())


def random_fields(box, seed=12345, corder=1, smooth=False):
    """seeded random fields with ranges bracketing what the reference setup produces
    (mu,la in (1,3), rho in (1,2), stretch in (0.5,1), dc in (0,0.05), corner in (0.33,1))"""
    r = np.random.default_rng(seed)
    shp = (box.nk, box.nj, box.ni)
    f = {}
    for name in ("u", "um", "up", "fo", "lu"):
        f[name] = pack3([r.uniform(-1, 1, shp) for _ in range(3)], corder)
    f["mu"] = r.uniform(1, 3, shp).ravel()
    f["la"] = r.uniform(1, 3, shp).ravel()
    f["rho"] = r.uniform(1, 2, shp).ravel()
    for d, n in (("x", box.ni), ("y", box.nj), ("z", box.nk)):
        f["str" + d] = r.uniform(0.5, 1, n)
        f["dc" + d] = r.uniform(0, 0.05, n)
        f["co" + d] = r.uniform(0.33, 1, n)
    return f


def relerr(a, b):
    d = np.max(np.abs(a - b))
    s = np.max(np.abs(b))
    return d / s if s > 0 else d
