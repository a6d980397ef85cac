/*
 * TEST INFRASTRUCTURE ONLY (oracle/).  CPU restatement ("port") of the reference algorithm for
 * SW4's explicit elastic time step.  It is the checker for the CUDA path; it is never shipped,
 * never on the product path, and only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load it.
 *
 * Parity status: PINNED.  Every function here is checked in tests/test_oracle.py against the
 * reference itself compiled in this container (oracle/_ref/libsw4ref.so, built from
 * /root/reference/src by oracle/build_ref.py) and against committed golden vectors that were
 * generated from that reference (tests/golden/, script tests/golden/make_golden.py).
 *
 * The restatement is written from the operator algebra of the scheme,
 * not transcribed from the reference loops:
 *    D0 f   = (f[-2] - 8 f[-1] + 8 f[+1] - f[+2])/12
 *    G(a) f = (1/6) sum_m w_m(a) (f[m]-f[0]),  m = -2,-1,+1,+2
 *               w_-2 = a[-1] - 3/4 (a[0]+a[-2]),   w_-1 = a[-2]+a[+1]+3(a[0]+a[-1])
 *               w_+1 = a[-1]+a[+2]+3(a[+1]+a[0]),  w_+2 = a[+1] - 3/4 (a[0]+a[+2])
 * Floating-point evaluation order therefore differs from the reference at the 1e-16 level.
 *
 * Index conventions follow the reference CPU kernels: inclusive Fortran-style global bounds
 * ifirst..klast (ghost points included), `nk` = global number of interior k points.
 * Layout: corder=1 -> (i,j,k,c), component stride npts (Sarray.C:756-766);
 *         corder=0 -> (c,i,j,k), component fastest     (Sarray.C:767-777).
 */
#include <stddef.h>
#include <math.h>
#include "../sw4lite_b200/csrc/sbp4_tables.h"

#define ACOF(k, q, m) acof[((k)-1) + 6 * ((q)-1) + 48 * ((m)-1)]
#define BOPE(k, q) bope[((k)-1) + 6 * ((q)-1)]

typedef struct
{
   int ifirst, jfirst, kfirst;
   ptrdiff_t ni, nij, npts;
   ptrdiff_t sc, sp; /* component stride, point stride */
} layout_t;

static layout_t make_layout( int corder, int ifirst, int ilast, int jfirst, int jlast, int kfirst, int klast )
{
   layout_t L;
   L.ifirst = ifirst; L.jfirst = jfirst; L.kfirst = kfirst;
   L.ni = ilast - ifirst + 1;
   L.nij = L.ni * (jlast - jfirst + 1);
   L.npts = L.nij * (klast - kfirst + 1);
   if( corder ) { L.sc = L.npts; L.sp = 1; }
   else         { L.sc = 1;      L.sp = 3; }
   return L;
}
static inline ptrdiff_t pt( const layout_t* L, int i, int j, int k )
{
   return (i - L->ifirst) + L->ni * (ptrdiff_t)(j - L->jfirst) + L->nij * (ptrdiff_t)(k - L->kfirst);
}
#define V(a, c, p) (a)[(c)*L.sc + L.sp * (p)]

/* EW::GetStencilCoefficients, ew-cfromfort.C:1487-1944 */
void oracle_stencil_coefficients( double* acof, double* ghcof, double* bope, double* sbop )
{
   int n;
   for( n = 0; n < 384; n++ ) acof[n] = 0;
   for( n = 0; n < 48; n++ ) bope[n] = 0;
   for( n = 0; n < 6; n++ ) ghcof[n] = 0;
   for( n = 0; n < (int)(sizeof(SW4B200_ACOF_NZ) / sizeof(SW4B200_ACOF_NZ[0])); n++ )
      acof[SW4B200_ACOF_NZ[n].idx] = SW4B200_ACOF_NZ[n].num / SW4B200_ACOF_NZ[n].den;
   for( n = 0; n < (int)(sizeof(SW4B200_BOPE_NZ) / sizeof(SW4B200_BOPE_NZ[0])); n++ )
      bope[SW4B200_BOPE_NZ[n].idx] = SW4B200_BOPE_NZ[n].num / SW4B200_BOPE_NZ[n].den;
   for( n = 0; n < (int)(sizeof(SW4B200_GHCOF_NZ) / sizeof(SW4B200_GHCOF_NZ[0])); n++ )
      ghcof[SW4B200_GHCOF_NZ[n].idx] = SW4B200_GHCOF_NZ[n].num / SW4B200_GHCOF_NZ[n].den;
   for( n = 0; n < 5; n++ )
      sbop[n] = SW4B200_SBOP[n].num / SW4B200_SBOP[n].den;
}

static inline void weights4( const double a[5], double w[4] )
{
   w[0] = a[1] - 0.75 * (a[2] + a[0]);
   w[1] = a[0] + a[3] + 3 * (a[2] + a[1]);
   w[2] = a[1] + a[4] + 3 * (a[3] + a[2]);
   w[3] = a[3] - 0.75 * (a[2] + a[4]);
}
/* sum_m w_m (f[m]-f[0]) for a line of 5 values f[0..4] centred on f[2] */
static inline double gsum( const double w[4], const double f[5] )
{
   return w[0] * (f[0] - f[2]) + w[1] * (f[1] - f[2]) + w[2] * (f[3] - f[2]) + w[3] * (f[4] - f[2]);
}
static inline double d0( double fm2, double fm1, double fp1, double fp2 )
{
   return (fm2 - fp2 + 8 * (fp1 - fm1)) * (1.0 / 12);
}

/*
 * rhs4sg / rhs4sg_rev (rhs4sg.C:38-849, rhs4sg_rev.C:44-864): lu = (1/h^2) L(u).
 * Interior rows k1..k2, low-k closure rows 1..6 if onesided[4], high-k rows nk-5..nk if onesided[5].
 */
void oracle_rhs4sg( int corder, int ifirst, int ilast, int jfirst, int jlast, int kfirst, int klast,
		    int nk, const int* onesided, const double* acof, const double* bope, const double* ghcof,
		    double* lu, const double* u, const double* mu, const double* la, double h,
		    const double* strx, const double* stry, const double* strz )
{
   const layout_t L = make_layout( corder, ifirst, ilast, jfirst, jlast, kfirst, klast );
   const double cof = 1.0 / (h * h);
   const ptrdiff_t di = 1, dj = L.ni, dk = L.nij;
   int k1 = kfirst + 2, k2 = klast - 2;
   if( onesided[4] == 1 ) k1 = 7;
   if( onesided[5] == 1 ) k2 = nk - 6;
#define SX(i) strx[(i)-ifirst]
#define SY(j) stry[(j)-jfirst]
#define SZ(k) strz[(k)-kfirst]
   int k;
#pragma omp parallel for
   for( k = k1; k <= k2; k++ )
      for( int j = jfirst + 2; j <= jlast - 2; j++ )
	 for( int i = ifirst + 2; i <= ilast - 2; i++ )
	 {
	    const ptrdiff_t p = pt( &L, i, j, k );
	    double r[3] = {0, 0, 0};
	    /* second-derivative terms: direction d, stride st, stretch s */
	    for( int d = 0; d < 3; d++ )
	    {
	       const ptrdiff_t st = d == 0 ? di : (d == 1 ? dj : dk);
	       double am[5], bm[5], wm[4], wb[4], f[5];
	       for( int m = -2; m <= 2; m++ )
	       {
		  double s = d == 0 ? SX( i + m ) : (d == 1 ? SY( j + m ) : SZ( k + m ));
		  am[m + 2] = mu[p + m * st] * s;
		  bm[m + 2] = (2 * mu[p + m * st] + la[p + m * st]) * s;
	       }
	       weights4( am, wm );
	       weights4( bm, wb );
	       const double s0 = d == 0 ? SX( i ) : (d == 1 ? SY( j ) : SZ( k ));
	       for( int c = 0; c < 3; c++ )
	       {
		  for( int m = -2; m <= 2; m++ ) f[m + 2] = V( u, c, p + m * st );
		  r[c] += (1.0 / 6) * s0 * gsum( c == d ? wb : wm, f );
	       }
	    }
	    /* mixed terms: for a != b,  L_a += s_a s_b [ D0_a( la D0_b u_b ) + D0_b( mu D0_a u_b ) ] */
	    for( int a = 0; a < 3; a++ )
	       for( int b = 0; b < 3; b++ )
	       {
		  if( a == b ) continue;
		  const ptrdiff_t sa = a == 0 ? di : (a == 1 ? dj : dk);
		  const ptrdiff_t sb = b == 0 ? di : (b == 1 ? dj : dk);
		  const double fa = a == 0 ? SX( i ) : (a == 1 ? SY( j ) : SZ( k ));
		  const double fb = b == 0 ? SX( i ) : (b == 1 ? SY( j ) : SZ( k ));
		  double t1[5], t2[5];
		  for( int m = -2; m <= 2; m++ )
		  {
		     if( m == 0 ) { t1[2] = t2[2] = 0; continue; }
		     ptrdiff_t q = p + m * sa;
		     t1[m + 2] = la[q] * d0( V( u, b, q - 2 * sb ), V( u, b, q - sb ), V( u, b, q + sb ), V( u, b, q + 2 * sb ) );
		     q = p + m * sb;
		     t2[m + 2] = mu[q] * d0( V( u, b, q - 2 * sa ), V( u, b, q - sa ), V( u, b, q + sa ), V( u, b, q + 2 * sa ) );
		  }
		  r[a] += fa * fb * (d0( t1[0], t1[1], t1[3], t1[4] ) + d0( t2[0], t2[1], t2[3], t2[4] ));
	       }
	    for( int c = 0; c < 3; c++ ) V( lu, c, p ) = cof * r[c];
	 }

   /* one-sided closures.  side 0: rows k=1..6 read planes 1..8 and ghost plane 0;
      side 1: rows nk-5..nk mirrored (plane nk-q+1, ghost nk+1, sign flip on bope terms). */
   for( int side = 0; side < 2; side++ )
   {
      if( onesided[4 + side] != 1 ) continue;
#pragma omp parallel for
      for( int kb = 1; kb <= 6; kb++ )
      {
	 const int kk = side == 0 ? kb : nk - kb + 1;	    /* this row */
	 const int kg = side == 0 ? 0 : nk + 1;		    /* ghost plane */
	 const int kbnd = side == 0 ? 1 : nk;		    /* boundary plane */
	 const double sgn = side == 0 ? 1.0 : -1.0;
#define KQ(q) (side == 0 ? (q) : nk - (q) + 1)
	 for( int j = jfirst + 2; j <= jlast - 2; j++ )
	    for( int i = ifirst + 2; i <= ilast - 2; i++ )
	    {
	       const ptrdiff_t p = pt( &L, i, j, kk );
	       double r[3] = {0, 0, 0};
	       /* x and y second derivatives: centred, as in the interior */
	       for( int d = 0; d < 2; d++ )
	       {
		  const ptrdiff_t st = d == 0 ? di : dj;
		  double am[5], bm[5], wm[4], wb[4], f[5];
		  for( int m = -2; m <= 2; m++ )
		  {
		     double s = d == 0 ? SX( i + m ) : SY( j + m );
		     am[m + 2] = mu[p + m * st] * s;
		     bm[m + 2] = (2 * mu[p + m * st] + la[p + m * st]) * s;
		  }
		  weights4( am, wm );
		  weights4( bm, wb );
		  const double s0 = d == 0 ? SX( i ) : SY( j );
		  for( int c = 0; c < 3; c++ )
		  {
		     for( int m = -2; m <= 2; m++ ) f[m + 2] = V( u, c, p + m * st );
		     r[c] += (1.0 / 6) * s0 * gsum( c == d ? wb : wm, f );
		  }
	       }
	       /* z second derivative with the SBP boundary operator, no strz, no 1/6 */
	       for( int q = 1; q <= 8; q++ )
	       {
		  double mucof = 0, lap2mu = 0;
		  for( int m = 1; m <= 8; m++ )
		  {
		     const ptrdiff_t pm = pt( &L, i, j, KQ( m ) );
		     mucof += ACOF( kb, q, m ) * mu[pm];
		     lap2mu += ACOF( kb, q, m ) * (2 * mu[pm] + la[pm]);
		  }
		  const ptrdiff_t pq = pt( &L, i, j, KQ( q ) );
		  r[0] += mucof * V( u, 0, pq );
		  r[1] += mucof * V( u, 1, pq );
		  r[2] += lap2mu * V( u, 2, pq );
	       }
	       {
		  const ptrdiff_t pb = pt( &L, i, j, kbnd ), pg = pt( &L, i, j, kg );
		  r[0] += ghcof[kb - 1] * mu[pb] * V( u, 0, pg );
		  r[1] += ghcof[kb - 1] * mu[pb] * V( u, 1, pg );
		  r[2] += ghcof[kb - 1] * (2 * mu[pb] + la[pb]) * V( u, 2, pg );
	       }
	       /* xy mixed terms: centred */
	       for( int a = 0; a < 2; a++ )
	       {
		  const int b = 1 - a;
		  const ptrdiff_t sa = a == 0 ? di : dj, sb = b == 0 ? di : dj;
		  double t1[5], t2[5];
		  for( int m = -2; m <= 2; m++ )
		  {
		     if( m == 0 ) { t1[2] = t2[2] = 0; continue; }
		     ptrdiff_t q = p + m * sa;
		     t1[m + 2] = la[q] * d0( V( u, b, q - 2 * sb ), V( u, b, q - sb ), V( u, b, q + sb ), V( u, b, q + 2 * sb ) );
		     q = p + m * sb;
		     t2[m + 2] = mu[q] * d0( V( u, b, q - 2 * sa ), V( u, b, q - sa ), V( u, b, q + sa ), V( u, b, q + 2 * sa ) );
		  }
		  r[a] += SX( i ) * SY( j ) * (d0( t1[0], t1[1], t1[3], t1[4] ) + d0( t2[0], t2[1], t2[3], t2[4] ));
	       }
	       /* terms with one z derivative: D0z -> sgn * sum_q bope(kb,q) f(KQ(q)); a = x or y */
	       for( int a = 0; a < 2; a++ )
	       {
		  const ptrdiff_t sa = a == 0 ? di : dj;
		  const double fa = a == 0 ? SX( i ) : SY( j );
		  /* L_a += s_a [ D0_a( la Bz w ) + Bz( mu D0_a w ) ] ;  L_3 += s_a [ D0_a( mu Bz u_a ) + Bz( la D0_a u_a ) ] */
		  double bw[5], bu[5];
		  for( int m = -2; m <= 2; m++ )
		  {
		     bw[m + 2] = bu[m + 2] = 0;
		     if( m == 0 ) continue;
		     for( int q = 1; q <= 8; q++ )
		     {
			const ptrdiff_t pq = pt( &L, i, j, KQ( q ) ) + m * sa;
			bw[m + 2] += sgn * BOPE( kb, q ) * V( u, 2, pq );
			bu[m + 2] += sgn * BOPE( kb, q ) * V( u, a, pq );
		     }
		     bw[m + 2] *= la[p + m * sa];
		     bu[m + 2] *= mu[p + m * sa];
		  }
		  double zmw = 0, zlu = 0;
		  for( int q = 1; q <= 8; q++ )
		  {
		     const ptrdiff_t pq = pt( &L, i, j, KQ( q ) );
		     zmw += sgn * BOPE( kb, q ) * (mu[pq] * d0( V( u, 2, pq - 2 * sa ), V( u, 2, pq - sa ), V( u, 2, pq + sa ), V( u, 2, pq + 2 * sa ) ));
		     zlu += sgn * BOPE( kb, q ) * (la[pq] * d0( V( u, a, pq - 2 * sa ), V( u, a, pq - sa ), V( u, a, pq + sa ), V( u, a, pq + 2 * sa ) ));
		  }
		  r[a] += fa * (d0( bw[0], bw[1], bw[3], bw[4] ) + zmw);
		  r[2] += fa * (d0( bu[0], bu[1], bu[3], bu[4] ) + zlu);
	       }
	       for( int c = 0; c < 3; c++ ) V( lu, c, p ) = cof * r[c];
	    }
#undef KQ
      }
   }
#undef SX
#undef SY
#undef SZ
}

/* EW::predfort, ew-cfromfort.C:70-113: up = 2u - um + dt^2/rho (lu + fo), all points */
void oracle_predfort( int corder, int ib, int ie, int jb, int je, int kb, int ke, double* up,
		      const double* u, const double* um, const double* lu, const double* fo,
		      const double* rho, double dt2 )
{
   const layout_t L = make_layout( corder, ib, ie, jb, je, kb, ke );
   ptrdiff_t p;
#pragma omp parallel for
   for( p = 0; p < L.npts; p++ )
   {
      const double f = dt2 / rho[p];
      for( int c = 0; c < 3; c++ )
	 V( up, c, p ) = 2 * V( u, c, p ) - V( um, c, p ) + f * (V( lu, c, p ) + V( fo, c, p ));
   }
}

/* EW::corrfort, ew-cfromfort.C:40-67: up += dt^4/(12 rho) (lu + fo) */
void oracle_corrfort( int corder, int ib, int ie, int jb, int je, int kb, int ke, double* up,
		      const double* lu, const double* fo, const double* rho, double dt4 )
{
   const layout_t L = make_layout( corder, ib, ie, jb, je, kb, ke );
   const double dt4i12 = dt4 / 12;
   ptrdiff_t p;
#pragma omp parallel for
   for( p = 0; p < L.npts; p++ )
   {
      const double f = dt4i12 / rho[p];
      for( int c = 0; c < 3; c++ )
	 V( up, c, p ) += f * (V( lu, c, p ) + V( fo, c, p ));
   }
}

/* EW::dpdmtfort, ew-cfromfort.C:116-141: u2 = (up - 2u + um)/dt^2 */
void oracle_dpdmtfort( int ib, int ie, int jb, int je, int kb, int ke, const double* up,
		       const double* u, const double* um, double* u2, double dt2i )
{
   const ptrdiff_t n = 3 * (ptrdiff_t)(ie - ib + 1) * (je - jb + 1) * (ke - kb + 1);
   ptrdiff_t p;
#pragma omp parallel for
   for( p = 0; p < n; p++ )
      u2[p] = dt2i * (up[p] - 2 * u[p] + um[p]);
}

/*
 * EW::addsgd4fort(_indrev), ew-cfromfort.C:748-1056, and addsgd6fort(_indrev), :852-1160:
 *   up -= beta/rho * sum_d  s_d c_d' c_d''  Dd( rho dc_d Dd (u-um) )
 * order 4: Dd = delta^2 (3-point), applied twice          -> 5-point stencil per direction
 * order 6: third difference of ( (rho dc)-average * third difference ), 7-point stencil per direction
 */
void oracle_addsgd( int corder, int order, int ifirst, int ilast, int jfirst, int jlast, int kfirst, int klast,
		    double* up, const double* u, const double* um, const double* rho,
		    const double* dcx, const double* dcy, const double* dcz,
		    const double* strx, const double* stry, const double* strz,
		    const double* cox, const double* coy, const double* coz, double beta )
{
   if( beta == 0 ) return;
   const layout_t L = make_layout( corder, ifirst, ilast, jfirst, jlast, kfirst, klast );
   const int w = order == 6 ? 3 : 2;
   int k;
#pragma omp parallel for
   for( k = kfirst + w; k <= klast - w; k++ )
      for( int j = jfirst + w; j <= jlast - w; j++ )
	 for( int i = ifirst + w; i <= ilast - w; i++ )
	 {
	    const ptrdiff_t p = pt( &L, i, j, k );
	    const double birho = beta / rho[p];
	    for( int c = 0; c < 3; c++ )
	    {
	       double tot = 0;
	       for( int d = 0; d < 3; d++ )
	       {
		  const ptrdiff_t st = d == 0 ? 1 : (d == 1 ? L.ni : L.nij);
		  const double* dc = d == 0 ? dcx + (i - ifirst) : (d == 1 ? dcy + (j - jfirst) : dcz + (k - kfirst));
		  const double pre = d == 0 ? strx[i - ifirst] * coy[j - jfirst] * coz[k - kfirst]
					    : (d == 1 ? stry[j - jfirst] * cox[i - ifirst] * coz[k - kfirst]
						      : strz[k - kfirst] * cox[i - ifirst] * coy[j - jfirst]);
		  double dl[7];
		  for( int m = -w; m <= w; m++ )
		     dl[m + 3] = V( u, c, p + m * st ) - V( um, c, p + m * st );
		  double s;
		  if( order == 4 )
		  {
		     /* delta^2( rho dc delta^2 d ) */
		     double e[3];
		     for( int m = -1; m <= 1; m++ )
			e[m + 1] = rho[p + m * st] * dc[m] * (dl[m + 4] - 2 * dl[m + 3] + dl[m + 2]);
		     s = e[2] - 2 * e[1] + e[0];
		  }
		  else
		  {
		     /* third difference of ( A_{m+1/2} * third difference of d ), A_{m+1/2} = (rho dc)_{m+1}+(rho dc)_m;
			the reference adds 0.5*beta/rho times it (ew-cfromfort.C:1095-1140) */
		     double acc = 0;
		     static const double cw[4] = {-1, 3, -3, 1};
		     for( int m = -2; m <= 1; m++ )
		     {
			const double A = rho[p + (m + 1) * st] * dc[m + 1] + rho[p + m * st] * dc[m];
			const double T = dl[m + 5] - 3 * dl[m + 4] + 3 * dl[m + 3] - dl[m + 2];
			acc += cw[m + 2] * A * T;
		     }
		     s = -0.5 * acc;
		  }
		  tot += pre * s;
	       }
	       V( up, c, p ) -= birho * tot;
	    }
	 }
}

/*
 * EW::bcfortsg(_indrev), ew-cfromfort.C:205-745.  Per side s (0..5 = low-i, high-i, low-j,
 * high-j, low-k, high-k) with window wind[6s..6s+5] = {i0,i1,j0,j1,k0,k1}:
 *   bccnd 1 (Dirichlet) / 2 (SuperGrid): u(window) = bforce_s (AoS 3*npoints, i fastest)
 *   bccnd 3 (Periodic): copy from the opposite side, offset nx / ny / nz points
 *   bccnd 0 (StressFree), sides 4,5 only: ghost plane from the traction-free condition
 */
void oracle_bcfortsg( int corder, int ib, int ie, int jb, int je, int kb, int ke, const int* wind,
		      int nx, int ny, int nz, double* u, double h, const int* bccnd, const double* sbop,
		      const double* mu, const double* la, double* const* bforce,
		      const double* strx, const double* stry )
{
   const layout_t L = make_layout( corder, ib, ie, jb, je, kb, ke );
   const double d4a = 2.0 / 3.0, d4b = -1.0 / 12.0;
   for( int s = 0; s < 6; s++ )
   {
      const int* w = wind + 6 * s;
      if( bccnd[s] == 1 || bccnd[s] == 2 )
      {
	 ptrdiff_t qq = 0;
	 for( int k = w[4]; k <= w[5]; k++ )
	    for( int j = w[2]; j <= w[3]; j++ )
	       for( int i = w[0]; i <= w[1]; i++, qq++ )
	       {
		  const ptrdiff_t p = pt( &L, i, j, k );
		  for( int c = 0; c < 3; c++ ) V( u, c, p ) = bforce[s][3 * qq + c];
	       }
      }
      else if( bccnd[s] == 3 )
      {
	 const ptrdiff_t off = s == 0 ? nx : (s == 1 ? -nx : (s == 2 ? L.ni * (ptrdiff_t)ny : (s == 3 ? -L.ni * (ptrdiff_t)ny
			       : (s == 4 ? L.nij * (ptrdiff_t)nz : -L.nij * (ptrdiff_t)nz))));
	 for( int k = w[4]; k <= w[5]; k++ )
	    for( int j = w[2]; j <= w[3]; j++ )
	       for( int i = w[0]; i <= w[1]; i++ )
	       {
		  const ptrdiff_t p = pt( &L, i, j, k );
		  for( int c = 0; c < 3; c++ ) V( u, c, p ) = V( u, c, p + off );
	       }
      }
      else if( bccnd[s] == 0 && (s == 4 || s == 5) )
      {
	 const int k = s == 4 ? 1 : nz, kl = s == 4 ? 1 : -1;
	 const double* bf = bforce[s];
	 for( int j = jb + 2; j <= je - 2; j++ )
	    for( int i = ib + 2; i <= ie - 2; i++ )
	    {
	       const ptrdiff_t qq = (i - ib) + L.ni * (ptrdiff_t)(j - jb);
	       const ptrdiff_t p = pt( &L, i, j, k );
	       const double sx = strx[i - ib], sy = stry[j - jb];
	       const double wx = sx * (d4a * (V( u, 2, p + 1 ) - V( u, 2, p - 1 )) + d4b * (V( u, 2, p + 2 ) - V( u, 2, p - 2 )));
	       const double ux = sx * (d4a * (V( u, 0, p + 1 ) - V( u, 0, p - 1 )) + d4b * (V( u, 0, p + 2 ) - V( u, 0, p - 2 )));
	       const double wy = sy * (d4a * (V( u, 2, p + L.ni ) - V( u, 2, p - L.ni )) + d4b * (V( u, 2, p + 2 * L.ni ) - V( u, 2, p - 2 * L.ni )));
	       const double vy = sy * (d4a * (V( u, 1, p + L.ni ) - V( u, 1, p - L.ni )) + d4b * (V( u, 1, p + 2 * L.ni ) - V( u, 1, p - 2 * L.ni )));
	       double uz = 0, vz = 0, wz = 0;
	       for( int q = 1; q <= 4; q++ )
	       {
		  const ptrdiff_t pq = p + L.nij * (ptrdiff_t)(kl * (q - 1));
		  uz += sbop[q] * V( u, 0, pq );
		  vz += sbop[q] * V( u, 1, pq );
		  wz += sbop[q] * V( u, 2, pq );
	       }
	       const ptrdiff_t pg = p - L.nij * (ptrdiff_t)kl;
	       V( u, 0, pg ) = (-uz - kl * wx + kl * h * bf[3 * qq] / mu[p]) / sbop[0];
	       V( u, 1, pg ) = (-vz - kl * wy + kl * h * bf[3 * qq + 1] / mu[p]) / sbop[0];
	       V( u, 2, pg ) = (-wz + (-kl * la[p] * (ux + vy) + kl * h * bf[3 * qq + 2]) / (2 * mu[p] + la[p])) / sbop[0];
	    }
      }
   }
}
