"""TEST INFRASTRUCTURE: one reference time step on the CPU, sequenced from the oracle's unfused
kernels exactly like the reference's CPU branch of EW::timesteploop (EW.C:2527-2763):
Force -> rhs4sg -> predfort -> bcfortsg -> Force_tt -> dpdmtfort -> rhs4sg -> corrfort -> addsgd ->
bcfortsg -> cycle.  Used by smoke() and the gpu parity tests for synthetic CartesianProblem blocks
(the pointsource test uses the reference's own EW object instead)."""
import numpy as np


def oracle():
    from oracle import refshim, port
    return refshim if refshim.available() else port


class OracleStepper:
    def __init__(self, prob, bounds=None, onesided=None, bctype=None, wind=None):
        """prob: sw4lite_b200.setup.CartesianProblem (corder 1 or 0).  bounds/onesided/bctype/wind
        override the block description (used for z-slabs of the problem)."""
        self.O = oracle()
        self.p = prob
        self.corder = prob.corder
        self.bounds = tuple(bounds or prob.bounds)
        ib, ie, jb, je, kb, ke = self.bounds
        self.ni, self.nj, self.nk = ie - ib + 1, je - jb + 1, ke - kb + 1
        self.npts = self.ni * self.nj * self.nk
        self.onesided = list(onesided if onesided is not None else prob.onesided)
        self.bctype = list(bctype if bctype is not None else prob.bctype)
        from sw4lite_b200.solver import boundary_windows
        self.wind = np.asarray(wind if wind is not None else boundary_windows(self.bounds, self.bctype), dtype=np.int32)
        self.acof, self.ghcof, self.bope, self.sbop = self.O.get_stencil_coefficients()
        k0 = kb - prob.bounds[4]
        sl = slice(k0 * self.ni * self.nj, (k0 + self.nk) * self.ni * self.nj)
        self.mu = np.ascontiguousarray(prob.mu[sl]); self.la = np.ascontiguousarray(prob.la[sl])
        self.rho = np.ascontiguousarray(prob.rho[sl])
        self.strz = np.ascontiguousarray(prob.strz[k0:k0 + self.nk]); self.dcz = np.ascontiguousarray(prob.dcz[k0:k0 + self.nk])
        self.coz = np.ascontiguousarray(prob.coz[k0:k0 + self.nk])
        n3 = 3 * self.npts
        self.U = np.zeros(n3); self.Um = np.zeros(n3); self.Up = np.zeros(n3)
        self.F = np.zeros(n3); self.Lu = np.zeros(n3); self.Uacc = np.zeros(n3)
        self.bforce = []
        for s in range(6):
            w = self.wind[6 * s:6 * s + 6]
            n = max(0, (w[1] - w[0] + 1)) * max(0, (w[3] - w[2] + 1)) * max(0, (w[5] - w[4] + 1))
            self.bforce.append(np.zeros(3 * n) if self.bctype[s] in (0, 1, 2) and n > 0 else None)
        self.src = [s for s in prob.sources if kb <= s[2] <= ke]

    def _force(self, f):
        """dense F with the point forces f[(nsrc,3)] (EW::Force, EW.C:3085-3123)"""
        self.F[:] = 0
        ib, ie, jb, je, kb, ke = self.bounds
        for n, s in enumerate(self.p.sources):
            i, j, k = s[0], s[1], s[2]
            if not (kb <= k <= ke):
                continue
            q = (i - ib) + self.ni * (j - jb) + self.ni * self.nj * (k - kb)
            for c in range(3):
                idx = q + c * self.npts if self.corder else 3 * q + c
                self.F[idx] += f[n][c]

    def _rhs(self, u):
        p = self.p
        self.O.rhs4sg(self.corder, self.bounds, p.nz, self.onesided, self.acof, self.bope, self.ghcof, self.Lu, u,
                      self.mu, self.la, p.h, p.strx, p.stry, self.strz)

    def _bc(self):
        p = self.p
        bf = [b if b is not None else None for b in self.bforce]
        bc = [t if t in (0, 1, 2, 3) else 99 for t in self.bctype]
        self.O.bcfortsg(self.corder, self.bounds, self.wind, p.nx, p.ny, p.nz, self.Up, p.h, bc, self.sbop, self.mu,
                        self.la, 0.0, bf, p.strx, p.stry)

    def predictor(self, f):
        p = self.p
        self._force(f if f is not None else [])
        self._rhs(self.U)
        self.O.predfort(self.corder, self.bounds, self.Up, self.U, self.Um, self.Lu, self.F, self.rho, p.dt ** 2)

    def corrector(self, ftt):
        p = self.p
        self._force(ftt if ftt is not None else [])
        self.O.dpdmtfort(self.corder, self.bounds, self.Up, self.U, self.Um, self.Uacc, 1.0 / p.dt ** 2)
        self._rhs(self.Uacc)
        self.O.corrfort(self.corder, self.bounds, self.Up, self.Lu, self.F, self.rho, p.dt ** 4)
        if p.beta != 0:
            self.O.addsgd(self.corder, 4, self.bounds, self.Up, self.U, self.Um, self.rho, p.dcx, p.dcy, self.dcz,
                          p.strx, p.stry, self.strz, p.cox, p.coy, self.coz, p.beta)

    def enforce_bc(self):
        self._bc()

    def cycle(self):
        self.Um, self.U, self.Up = self.U, self.Up, self.Um

    def step(self, f=None, ftt=None):
        self.predictor(f)
        self.enforce_bc()
        self.corrector(ftt)
        self.enforce_bc()
        self.cycle()
