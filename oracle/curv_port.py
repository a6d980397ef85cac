"""TEST INFRASTRUCTURE ONLY (oracle/): numpy restatement of the reference's curvilinear-grid kernels
for the hot path.  Same call signatures as the kernel-level functions of oracle/refshim.py.  Pinned against
the reference build (oracle/_ref) and tests/golden/curvilinear_small.npz (tests/test_oracle_curvilinear.py).
The product package never imports this.

Restated algorithm (reference files under /root/reference/src):
* rhs4sgcurv / rhs4sgcurv_rev (rhs4sgcurv.C:34-1406, rhs4sgcurv_rev.C:34-1395):
    lu_c = (strx*stry/jac) * r_c, r_c = sum of
      pp, qq terms   : G_p(cof) u_c * istry, G_q(cof) u_c * istrx with cof = (2mu+la | mu) met1^2 str   (:594-640)
      rr terms       : G_r of the 6 entries of the symmetric coefficient matrix                         (:641-716)
      pq, qp         : D0_q( mu met1^2 D0_p . ), D0_p( la met1^2 D0_q . )                               (:717-750)
      pr, rp, qr, rq : D0_r( coefficient * D0_p/q . ), D0_p/q( coefficient * D0_r . )                   (:751-860)
    G(a) f = 1/6 sum_j w_j(a) (f_j - f_0) is the SBP variable-coefficient second difference (mux1..mux4),
    D0 the centred first difference c2 (f_{+2}-f_{-2}) + c1 (f_{+1}-f_{-1}), c1=2/3, c2=-1/12 (:52-57).
    Rows k=1..6 when onesided[4]==1 (:93-584): G_r -> sum_q [sum_m acof(k,q,m) coef(m)] u(q) + ghcof(k) coef(1) u(0),
    every D0_r -> sum_q bope(k,q) . (q).
  Here the sums are organised by the outer difference direction (see sw4lite_b200/csrc/curvilinear.cu for the
  algebra): with a = (met2 strx, met3 stry, met4), e = met1 strx, f = met1 stry the coefficient of
  D_a( . D_b u_d ) in equation c is  la A_a[c] A_b[d] + mu (delta_cd A_a.A_b + A_a[d] A_b[c]).
* addsgd4cfort / addsgd6cfort (ew-cfromfort.C:1160-1322), freesurfcurvisg (curvilinear-c.C:465-618),
  enforceCartTopo (EW.C:3504-3531).
"""
import numpy as np

C1, C2 = 2.0 / 3, -1.0 / 12


def _unpack(a, shape, nc, corder):
    nk, nj, ni = shape
    a = np.asarray(a)
    if corder:
        return a.reshape(nc, nk, nj, ni).copy()
    return np.moveaxis(a.reshape(nk, nj, ni, nc), 3, 0).copy()


def _pack(a, corder):
    if corder:
        return np.ascontiguousarray(a).ravel()
    return np.ascontiguousarray(np.moveaxis(a, 0, 3)).ravel()


def _shift(a, axis, m):
    """a(index + m) along axis, zero where the index leaves the array"""
    out = np.zeros_like(a)
    n = a.shape[axis]
    src = [slice(None)] * a.ndim; dst = [slice(None)] * a.ndim
    if m >= 0:
        src[axis] = slice(m, n); dst[axis] = slice(0, n - m)
    else:
        src[axis] = slice(0, n + m); dst[axis] = slice(-m, n)
    out[tuple(dst)] = a[tuple(src)]
    return out


def _d0(a, axis):
    return C2 * (_shift(a, axis, 2) - _shift(a, axis, -2)) + C1 * (_shift(a, axis, 1) - _shift(a, axis, -1))


def _weights(c, axis):
    cm2, cm1, cp1, cp2 = (_shift(c, axis, m) for m in (-2, -1, 1, 2))
    return (cm1 - 0.75 * (c + cm2), cm2 + cp1 + 3 * (c + cm1), cm1 + cp2 + 3 * (cp1 + c), cp1 - 0.75 * (c + cp2))


def _G(c, f, axis):
    w = _weights(c, axis)
    return (w[0] * (_shift(f, axis, -2) - f) + w[1] * (_shift(f, axis, -1) - f) + w[2] * (_shift(f, axis, 1) - f) +
            w[3] * (_shift(f, axis, 2) - f)) / 6


def rhs4sgcurv(corder, b, u, mu, la, met, jac, lu, onesided, acof, bope, ghcof, strx, stry):
    ib, ie, jb, je, kb, ke = b
    ni, nj, nk = ie - ib + 1, je - jb + 1, ke - kb + 1
    shp = (nk, nj, ni)
    U = _unpack(u, shp, 3, corder); MET = _unpack(met, shp, 4, corder)
    M = np.asarray(mu).reshape(shp); L = np.asarray(la).reshape(shp); J = np.asarray(jac).reshape(shp)
    sx = np.asarray(strx)[None, None, :] * np.ones(shp); sy = np.asarray(stry)[None, :, None] * np.ones(shp)
    m1, m2, m3, m4 = MET
    top = int(onesided[4]) == 1
    A = lambda k, q, m: acof[(k - 1) + 6 * (q - 1) + 48 * (m - 1)]
    B = lambda k, q: bope[(k - 1) + 6 * (q - 1)]
    AX_K, AX_J, AX_I = 0, 1, 2

    def dr(f):
        """r-difference of a scalar field: centred, or the one-sided SBP sums on rows 1..6"""
        out = _d0(f, AX_K)
        if top:
            for k in range(1, 7):
                out[k - kb] = sum(B(k, q) * f[q - kb] for q in range(1, 9))
        return out

    dp = [_d0(U[c], AX_I) for c in range(3)]
    dq = [_d0(U[c], AX_J) for c in range(3)]
    drU = [dr(U[c]) for c in range(3)]
    l2m = 2 * M + L
    a1, a2, a3 = m2 * sx, m3 * sy, m4
    e, f = m1 * sx, m1 * sy
    # outer p: sx(i) [ G_p(coef) u_c + D0_p X_c ]; the coefficient carries the stretch of its own point
    X = [L * m1 * m1 * sy * dq[1] + l2m * m1 * a1 * drU[0] + L * m1 * a2 * drU[1] + L * m1 * a3 * drU[2],
         M * m1 * m1 * sy * dq[0] + M * m1 * a2 * drU[0] + M * m1 * a1 * drU[1],
         M * m1 * a3 * drU[0] + M * m1 * a1 * drU[2]]
    cp_l, cp_m = l2m * m1 * m1 * sx, M * m1 * m1 * sx
    rp = [_G(cp_l if c == 0 else cp_m, U[c], AX_I) + _d0(X[c], AX_I) for c in range(3)]
    # outer q
    Y = [M * m1 * m1 * sx * dp[1] + M * m1 * a2 * drU[0] + M * m1 * a1 * drU[1],
         L * m1 * m1 * sx * dp[0] + L * m1 * a1 * drU[0] + l2m * m1 * a2 * drU[1] + L * m1 * a3 * drU[2],
         M * m1 * a3 * drU[1] + M * m1 * a2 * drU[2]]
    cq_l, cq_m = l2m * m1 * m1 * sy, M * m1 * m1 * sy
    rq = [_G(cq_l if c == 1 else cq_m, U[c], AX_J) + _d0(Y[c], AX_J) for c in range(3)]
    # outer r
    Z = [l2m * a1 * e * dp[0] + M * a2 * e * dp[1] + M * a3 * e * dp[2] + M * a2 * f * dq[0] + L * a1 * f * dq[1],
         L * a2 * e * dp[0] + M * a1 * e * dp[1] + M * a1 * f * dq[0] + l2m * a2 * f * dq[1] + M * a3 * f * dq[2],
         L * a3 * e * dp[0] + M * a1 * e * dp[2] + L * a3 * f * dq[1] + M * a2 * f * dq[2]]
    lm = M + L
    N = {(0, 0): l2m * a1 * a1 + M * (a2 * a2 + a3 * a3), (1, 1): l2m * a2 * a2 + M * (a1 * a1 + a3 * a3),
         (2, 2): l2m * a3 * a3 + M * (a1 * a1 + a2 * a2), (0, 1): lm * a1 * a2, (0, 2): lm * a1 * a3, (1, 2): lm * a2 * a3}
    Ncd = lambda c, d: N[(min(c, d), max(c, d))]
    rr = [sum(_G(Ncd(c, d), U[d], AX_K) for d in range(3)) + dr(Z[c]) for c in range(3)]
    if top:
        for k in range(1, 7):
            for c in range(3):
                acc = np.zeros((nj, ni))
                for q in range(1, 9):
                    acc = acc + B(k, q) * Z[c][q - kb]
                    for d in range(3):
                        coef = sum(A(k, q, m) * Ncd(c, d)[m - kb] for m in range(1, 9))
                        acc = acc + coef * U[d][q - kb]
                for d in range(3):
                    acc = acc + ghcof[k - 1] * Ncd(c, d)[1 - kb] * U[d][0 - kb]
                rr[c][k - kb] = acc
    out = _unpack(lu, shp, 3, corder)
    k0 = 2
    I = (slice(k0, nk - 2), slice(2, nj - 2), slice(2, ni - 2))
    for c in range(3):
        r = (sx * rp[c] + sy * rq[c] + rr[c]) / J
        out[c][I] = r[I]
    lu[:] = _pack(out, corder)


def addsgdc(corder, order, b, up, u, um, rho, dcx, dcy, strx, stry, jac, cox, coy, beta):
    if beta == 0:
        return
    ib, ie, jb, je, kb, ke = b
    ni, nj, nk = ie - ib + 1, je - jb + 1, ke - kb + 1
    shp = (nk, nj, ni)
    UP = _unpack(up, shp, 3, corder); D = _unpack(u, shp, 3, corder) - _unpack(um, shp, 3, corder)
    R = np.asarray(rho).reshape(shp); J = np.asarray(jac).reshape(shp)
    ones = np.ones(shp)
    wx = R * np.asarray(dcx)[None, None, :] * J; wy = R * np.asarray(dcy)[None, :, None] * J
    prex = (np.asarray(strx)[None, None, :] * np.asarray(coy)[None, :, None]) * ones
    prey = (np.asarray(stry)[None, :, None] * np.asarray(cox)[None, None, :]) * ones
    w = 2 if order == 4 else 3
    I = (slice(w, nk - w), slice(w, nj - w), slice(w, ni - w))
    for c in range(3):
        tot = 0
        for axis, wgt, pre in ((2, wx, prex), (1, wy, prey)):
            d = D[c]
            if order == 4:
                d2 = _shift(d, axis, 1) - 2 * d + _shift(d, axis, -1)
                e = wgt * d2
                term = _shift(e, axis, 1) - 2 * e + _shift(e, axis, -1)
            else:
                # third difference centred at i+1/2, weight (w_{i+1}+w_i), then -1/2 times its third difference
                t3 = _shift(d, axis, 2) - 3 * _shift(d, axis, 1) + 3 * d - _shift(d, axis, -1)      # T(i+1/2)
                av = _shift(wgt, axis, 1) + wgt
                g = av * t3
                term = -0.5 * (_shift(g, axis, 1) - 3 * g + 3 * _shift(g, axis, -1) - _shift(g, axis, -2))
            tot = tot + pre * term
        UP[c][I] -= (beta / (R * J) * tot)[I]
    up[:] = _pack(UP, corder)


def freesurfcurvisg(corder, b, nz, side, u, mu, la, met, sbop, forcing, strx, stry):
    ib, ie, jb, je, kb, ke = b
    ni, nj, nk = ie - ib + 1, je - jb + 1, ke - kb + 1
    shp = (nk, nj, ni)
    U = _unpack(u, shp, 3, corder); MET = _unpack(met, shp, 4, corder)
    k, kl = (1, 1) if side == 5 else (nz, -1)
    lk = k - kb
    M = np.asarray(mu).reshape(shp)[lk]; L = np.asarray(la).reshape(shp)[lk]
    m1, m2, m3, m4 = (MET[c][lk] for c in range(4))
    sx = np.asarray(strx)[None, :] * np.ones((nj, ni)); sy = np.asarray(stry)[:, None] * np.ones((nj, ni))
    F = np.asarray(forcing).reshape(nj, ni, 3)
    dp = [_d0(U[c][lk], 1) for c in range(3)]; dq = [_d0(U[c][lk], 0) for c in range(3)]
    l2m = 2 * M + L
    rhs = [l2m * m2 * m1 * dp[0] * sx / sy + M * m3 * m1 * dp[1] + M * m4 * m1 * dp[2] / sy + M * m3 * m1 * dq[0] * sy / sx +
           L * m2 * m1 * dq[1] - F[:, :, 0],
           L * m3 * m1 * dp[0] + M * m2 * m1 * dp[1] * sx / sy + M * m2 * m1 * dq[0] + l2m * m3 * m1 * dq[1] * sy / sx +
           M * m4 * m1 * dq[2] / sx - F[:, :, 1],
           L * m4 * m1 * dp[0] / sy + M * m2 * m1 * dp[2] * sx / sy + M * m3 * m1 * dq[2] * sy / sx + L * m4 * m1 * dq[1] / sx -
           F[:, :, 2]]
    xoy = np.sqrt(sx / sy); yox = 1 / xoy; isq = xoy / sx
    av = [m2 * xoy, m3 * yox, m4 * isq]
    ac = sx / sy * m2 * m2 + sy / sx * m3 * m3 + m4 * m4 / (sx * sy)
    bc = 1 / (M * ac)
    cc = (M + L) / l2m * bc / ac
    dc = cc * (av[0] * rhs[0] + av[1] * rhs[1] + av[2] * rhs[2])
    I = (slice(2, nj - 2), slice(2, ni - 2))
    for c in range(3):
        s = sbop[1] * U[c][lk] + sbop[2] * U[c][lk + kl] + sbop[3] * U[c][lk + 2 * kl] + sbop[4] * U[c][lk + 3 * kl]
        U[c][lk - kl][I] = (-(s + bc * rhs[c] - dc * av[c]) / sbop[0])[I]
    u[:] = _pack(U, corder)


def enforce_cart_topo(corder, ucart, bcart, ucurv, bcurv):
    """EW.C:3504-3531: Cartesian ghost planes <- curvilinear planes kEnd-4+q; curvilinear planes kEnd-q <- Cartesian"""
    shc = (bcart[5] - bcart[4] + 1, bcart[3] - bcart[2] + 1, bcart[1] - bcart[0] + 1)
    sht = (bcurv[5] - bcurv[4] + 1, bcurv[3] - bcurv[2] + 1, bcurv[1] - bcurv[0] + 1)
    UC = _unpack(ucart, shc, 3, corder); UT = _unpack(ucurv, sht, 3, corder)
    nkt = sht[0]
    for q in range(2):
        UC[:, q] = UT[:, nkt - 1 - 4 + q]
    for q in range(3):
        UT[:, nkt - 1 - q] = UC[:, 4 - q]
    ucart[:] = _pack(UC, corder); ucurv[:] = _pack(UT, corder)
