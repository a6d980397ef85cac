// Curvilinear-grid operators (the top grid under topography), general CUDA kernels: one thread per grid
// point, neighbours through L1/L2, any layout (AoS `corder=0` / SoA `corder=1`).
//
// Reference semantics: rhs4sgcurv.C:34-1406 / rhs4sgcurv_rev.C:34-1395 (L(u) with metric terms),
// ew-cfromfort.C:1160-1486 (addsgd4c/6c), curvilinear-c.C:465-764 (freesurfcurvisg), EW.C:3504-3531
// (enforceCartTopo).
//
// The reference writes the operator out term by term (2126 operations per point).  Here it is evaluated
// from its structure.  With the scaled contravariant vectors of the mapping x=x(p), y=y(q), z=z(p,q,r)
//    A_p = (m1 sx, 0, 0),  A_q = (0, m1 sy, 0),  A_r = (m2 sx, m3 sy, m4)     (m = met(1..4), s = stretch)
// and the isotropic stiffness, the coefficient of  D_a( . D_b u_d )  in equation c is
//    N^{ab}_{cd} = la A_a[c] A_b[d] + mu ( delta_cd A_a.A_b + A_a[d] A_b[c] )
// and  J L_c(u) = sum_{a,b,d} D_a( N^{ab}_{cd} D_b u_d ),  where D_a(N D_a) is the 4th-order SBP
// variable-coefficient second difference (the G operator of rhs4sg) and D_a(N D_b), a != b, the product of
// centred first differences.  One stretch factor of an outer p (q) difference is taken at the output point,
// as the reference does (rhs4sgcurv_rev.C:594-1392).  Rows k=1..6 under a free surface replace every
// r-difference by the one-sided SBP sums (acof/ghcof/bope; rhs4sgcurv_rev.C:93-584).
// The result agrees with the reference to rounding (different association); tests bound it by 1e-12.
#include "common.cuh"

namespace sw4b200 {

#ifndef ACOF
#define ACOF( k, q, m ) c_acof[( (k)-1 ) + 6 * ( (q)-1 ) + 48 * ( (m)-1 )]
#define BOPE( k, q ) c_bope[( (k)-1 ) + 6 * ( (q)-1 )]
#endif

namespace {

struct CurvArgs
{
   Block b;
   const double *u, *mu, *la, *met, *jac;
   double* lu;
   const double *strx, *stry;
   long long msc, msp; // metric: component stride, point stride
   int kfirst;
};

__device__ __forceinline__ void cweights4( const double a[5], double w[4] )
{
   w[0] = a[1] - 0.75 * ( a[2] + a[0] );
   w[1] = a[0] + a[3] + 3 * ( a[2] + a[1] );
   w[2] = a[1] + a[4] + 3 * ( a[3] + a[2] );
   w[3] = a[3] - 0.75 * ( a[2] + a[4] );
}

struct Mat6 { double n11, n22, n33, n12, n13, n23; };

// kb: closure row (1..6) when CLOSURE, unused otherwise.  p: point index; li, lj, lk: local (array) indices.
template <bool CLOSURE>
__device__ void curv_point( const CurvArgs& a, long long p, int li, int lj, int lk, int kb, double r[3] )
{
   const Block& b = a.b;
   const long long sc = b.sc, sp = b.sp, dj = b.ni, dk = b.nij;
   const double c1 = 2.0 / 3, c2 = -1.0 / 12, i6 = 1.0 / 6;
   auto U = [&]( int c, long long q ) { return a.u[c * sc + sp * q]; };
   auto MET = [&]( int c, long long q ) { return a.met[c * a.msc + a.msp * q]; };
   auto d0 = [&]( int c, long long q, long long st ) {
      return c2 * ( U( c, q + 2 * st ) - U( c, q - 2 * st ) ) + c1 * ( U( c, q + st ) - U( c, q - st ) );
   };
   // plane of global index kq seen from a point of plane lk
   auto plane_off = [&]( int kq ) { return dk * ( kq - a.kfirst - lk ); };
   auto dr = [&]( int c, long long q ) {
      if( !CLOSURE ) return d0( c, q, dk );
      double s = 0;
#pragma unroll
      for( int m = 1; m <= 8; m++ ) s += BOPE( kb, m ) * U( c, q + plane_off( m ) );
      return s;
   };
   const double sxi = a.strx[li], syj = a.stry[lj];
   double rp[3], rq[3], rr[3] = { 0, 0, 0 };

   // ---- outer p
   {
      double cm[5], cl[5], wm[4], wl[4];
      double X[5][3];
#pragma unroll
      for( int m = -2; m <= 2; m++ )
      {
	 const long long q = p + m;
	 const double M = a.mu[q], L = a.la[q], m1 = MET( 0, q ), s = a.strx[li + m];
	 const double t = m1 * m1 * s;
	 cm[m + 2] = M * t;
	 cl[m + 2] = ( 2 * M + L ) * t;
	 if( m != 0 )
	 {
	    const double m2 = MET( 1, q ), m3 = MET( 2, q ), m4 = MET( 3, q );
	    const double dq0 = d0( 0, q, dj ), dq1 = d0( 1, q, dj );
	    const double r0 = dr( 0, q ), r1 = dr( 1, q ), r2 = dr( 2, q );
	    const double m12s = m1 * m2 * s, m13y = m1 * m3 * syj, m11y = m1 * m1 * syj, m14 = m1 * m4;
	    X[m + 2][0] = L * m11y * dq1 + ( 2 * M + L ) * m12s * r0 + L * m13y * r1 + L * m14 * r2;
	    X[m + 2][1] = M * m11y * dq0 + M * m13y * r0 + M * m12s * r1;
	    X[m + 2][2] = M * m14 * r0 + M * m12s * r2;
	 }
      }
      cweights4( cl, wl );
      cweights4( cm, wm );
#pragma unroll
      for( int c = 0; c < 3; c++ )
      {
	 const double* w = c == 0 ? wl : wm;
	 const double u0 = U( c, p );
	 const double g = w[0] * ( U( c, p - 2 ) - u0 ) + w[1] * ( U( c, p - 1 ) - u0 ) + w[2] * ( U( c, p + 1 ) - u0 ) +
			  w[3] * ( U( c, p + 2 ) - u0 );
	 rp[c] = i6 * g + ( c2 * ( X[4][c] - X[0][c] ) + c1 * ( X[3][c] - X[1][c] ) );
      }
   }
   // ---- outer q
   {
      double cm[5], cl[5], wm[4], wl[4];
      double Y[5][3];
#pragma unroll
      for( int m = -2; m <= 2; m++ )
      {
	 const long long q = p + m * dj;
	 const double M = a.mu[q], L = a.la[q], m1 = MET( 0, q ), s = a.stry[lj + m];
	 const double t = m1 * m1 * s;
	 cm[m + 2] = M * t;
	 cl[m + 2] = ( 2 * M + L ) * t;
	 if( m != 0 )
	 {
	    const double m2 = MET( 1, q ), m3 = MET( 2, q ), m4 = MET( 3, q );
	    const double dp0 = d0( 0, q, 1 ), dp1 = d0( 1, q, 1 );
	    const double r0 = dr( 0, q ), r1 = dr( 1, q ), r2 = dr( 2, q );
	    const double m13s = m1 * m3 * s, m12x = m1 * m2 * sxi, m11x = m1 * m1 * sxi, m14 = m1 * m4;
	    Y[m + 2][0] = M * m11x * dp1 + M * m13s * r0 + M * m12x * r1;
	    Y[m + 2][1] = L * m11x * dp0 + L * m12x * r0 + ( 2 * M + L ) * m13s * r1 + L * m14 * r2;
	    Y[m + 2][2] = M * m14 * r1 + M * m13s * r2;
	 }
      }
      cweights4( cl, wl );
      cweights4( cm, wm );
#pragma unroll
      for( int c = 0; c < 3; c++ )
      {
	 const double* w = c == 1 ? wl : wm;
	 const double u0 = U( c, p );
	 const double g = w[0] * ( U( c, p - 2 * dj ) - u0 ) + w[1] * ( U( c, p - dj ) - u0 ) +
			  w[2] * ( U( c, p + dj ) - u0 ) + w[3] * ( U( c, p + 2 * dj ) - u0 );
	 rq[c] = i6 * g + ( c2 * ( Y[4][c] - Y[0][c] ) + c1 * ( Y[3][c] - Y[1][c] ) );
      }
   }
   // ---- outer r
   // coefficient matrix N^{rr} and cross-term vector Z at a point of the column
   auto nrr = [&]( long long q ) {
      const double M = a.mu[q], L = a.la[q];
      const double a1 = MET( 1, q ) * sxi, a2 = MET( 2, q ) * syj, a3 = MET( 3, q );
      const double l2m = 2 * M + L, lm = M + L;
      Mat6 n;
      n.n11 = l2m * a1 * a1 + M * ( a2 * a2 + a3 * a3 );
      n.n22 = l2m * a2 * a2 + M * ( a1 * a1 + a3 * a3 );
      n.n33 = l2m * a3 * a3 + M * ( a1 * a1 + a2 * a2 );
      n.n12 = lm * a1 * a2;
      n.n13 = lm * a1 * a3;
      n.n23 = lm * a2 * a3;
      return n;
   };
   auto zvec = [&]( long long q, double Z[3] ) {
      const double M = a.mu[q], L = a.la[q], m1 = MET( 0, q );
      const double a1 = MET( 1, q ) * sxi, a2 = MET( 2, q ) * syj, a3 = MET( 3, q );
      const double e = m1 * sxi, f = m1 * syj, l2m = 2 * M + L;
      const double dp0 = d0( 0, q, 1 ), dp1 = d0( 1, q, 1 ), dp2 = d0( 2, q, 1 );
      const double dq0 = d0( 0, q, dj ), dq1 = d0( 1, q, dj ), dq2 = d0( 2, q, dj );
      Z[0] = l2m * a1 * e * dp0 + M * a2 * e * dp1 + M * a3 * e * dp2 + M * a2 * f * dq0 + L * a1 * f * dq1;
      Z[1] = L * a2 * e * dp0 + M * a1 * e * dp1 + M * a1 * f * dq0 + l2m * a2 * f * dq1 + M * a3 * f * dq2;
      Z[2] = L * a3 * e * dp0 + M * a1 * e * dp2 + L * a3 * f * dq1 + M * a2 * f * dq2;
   };
   if( !CLOSURE )
   {
      double n11[5], n22[5], n33[5], n12[5], n13[5], n23[5];
      double Z[5][3];
#pragma unroll
      for( int m = -2; m <= 2; m++ )
      {
	 const long long q = p + m * dk;
	 const Mat6 n = nrr( q );
	 n11[m + 2] = n.n11; n22[m + 2] = n.n22; n33[m + 2] = n.n33;
	 n12[m + 2] = n.n12; n13[m + 2] = n.n13; n23[m + 2] = n.n23;
	 if( m != 0 ) zvec( q, Z[m + 2] );
      }
      double w11[4], w22[4], w33[4], w12[4], w13[4], w23[4];
      cweights4( n11, w11 ); cweights4( n22, w22 ); cweights4( n33, w33 );
      cweights4( n12, w12 ); cweights4( n13, w13 ); cweights4( n23, w23 );
      double du[3][4];
#pragma unroll
      for( int c = 0; c < 3; c++ )
      {
	 const double u0 = U( c, p );
	 du[c][0] = U( c, p - 2 * dk ) - u0; du[c][1] = U( c, p - dk ) - u0;
	 du[c][2] = U( c, p + dk ) - u0; du[c][3] = U( c, p + 2 * dk ) - u0;
      }
      auto g = [&]( const double w[4], int d ) {
	 return w[0] * du[d][0] + w[1] * du[d][1] + w[2] * du[d][2] + w[3] * du[d][3];
      };
      rr[0] = i6 * ( g( w11, 0 ) + g( w12, 1 ) + g( w13, 2 ) );
      rr[1] = i6 * ( g( w12, 0 ) + g( w22, 1 ) + g( w23, 2 ) );
      rr[2] = i6 * ( g( w13, 0 ) + g( w23, 1 ) + g( w33, 2 ) );
#pragma unroll
      for( int c = 0; c < 3; c++ ) rr[c] += c2 * ( Z[4][c] - Z[0][c] ) + c1 * ( Z[3][c] - Z[1][c] );
   }
   else
   {
      Mat6 n[8];
#pragma unroll
      for( int m = 1; m <= 8; m++ ) n[m - 1] = nrr( p + plane_off( m ) );
      for( int q = 1; q <= 8; q++ )
      {
	 Mat6 s = { 0, 0, 0, 0, 0, 0 };
#pragma unroll
	 for( int m = 1; m <= 8; m++ )
	 {
	    const double ac = ACOF( kb, q, m );
	    s.n11 += ac * n[m - 1].n11; s.n22 += ac * n[m - 1].n22; s.n33 += ac * n[m - 1].n33;
	    s.n12 += ac * n[m - 1].n12; s.n13 += ac * n[m - 1].n13; s.n23 += ac * n[m - 1].n23;
	 }
	 const long long pq = p + plane_off( q );
	 const double u0 = U( 0, pq ), u1 = U( 1, pq ), u2 = U( 2, pq );
	 rr[0] += s.n11 * u0 + s.n12 * u1 + s.n13 * u2;
	 rr[1] += s.n12 * u0 + s.n22 * u1 + s.n23 * u2;
	 rr[2] += s.n13 * u0 + s.n23 * u1 + s.n33 * u2;
	 const double bq = BOPE( kb, q );
	 if( bq != 0 )
	 {
	    double Z[3];
	    zvec( pq, Z );
	    rr[0] += bq * Z[0]; rr[1] += bq * Z[1]; rr[2] += bq * Z[2];
	 }
      }
      // ghost plane k=0
      const double gh = c_ghcof[kb - 1];
      if( gh != 0 )
      {
	 const long long pg = p + plane_off( 0 );
	 const double u0 = U( 0, pg ), u1 = U( 1, pg ), u2 = U( 2, pg );
	 rr[0] += gh * ( n[0].n11 * u0 + n[0].n12 * u1 + n[0].n13 * u2 );
	 rr[1] += gh * ( n[0].n12 * u0 + n[0].n22 * u1 + n[0].n23 * u2 );
	 rr[2] += gh * ( n[0].n13 * u0 + n[0].n23 * u1 + n[0].n33 * u2 );
      }
   }
   const double ij = 1.0 / a.jac[p];
#pragma unroll
   for( int c = 0; c < 3; c++ ) r[c] = ( sxi * rp[c] + syj * rq[c] + rr[c] ) * ij;
}

template <bool CLOSURE>
__global__ void __launch_bounds__( 128 ) k_rhs4sgcurv( CurvArgs a, int k_lo, int k_hi )
{
   const Block& b = a.b;
   const int li = 2 + blockIdx.x * blockDim.x + threadIdx.x;
   const int lj = 2 + blockIdx.y * blockDim.y + threadIdx.y;
   const int k = k_lo + blockIdx.z * blockDim.z + threadIdx.z;
   if( li > b.nil - 3 || lj > b.nj - 3 || k > k_hi ) return;
   const int lk = k - b.kfirst;
   const long long p = (long long)li + (long long)b.ni * lj + b.nij * lk;
   double r[3];
   curv_point<CLOSURE>( a, p, li, lj, lk, k, r );
#pragma unroll
   for( int c = 0; c < 3; c++ ) a.lu[c * b.sc + b.sp * p] = r[c];
}

// ---- interior rows in two sweeps with shared fluxes (no point evaluates another point's first differences).
// With  X_c = sum_{b != p, d} N^{pb}_{cd} D_b u_d  (the part of the p-flux that is differenced by the centred D_p), and Y, Z
// likewise for q and r, the operator is
//    J L_c = sx [ G_p(N^{pp}) u + D0_p X ]_c + sy [ G_q(N^{qq}) u + D0_q Y ]_c + [ G_r(N^{rr}) u + D0_r Z ]_c .
// k_curv_flux evaluates X, Y, Z once per point (9 first differences of u,v,w and the metric products) into scratch arrays;
// k_curv_apply adds the three variable-coefficient second differences and the centred differences of the stored fluxes and
// finishes with the Jacobian.  The one-thread-per-point kernel above re-evaluates the fluxes of its 12 neighbours
// (about 1100 fp64 instructions per point); the two sweeps need about 600, for 72 B per point of scratch traffic each way.
// The free-surface closure rows k=1..6 keep the kernel above.
struct CurvFlux
{
   double* x[3]; // X_c, Y_c, Z_c: one value per grid point (point-indexed, whatever the layout of u)
   double* y[3];
   double* z[3];
};

__global__ void __launch_bounds__( 128 ) k_curv_flux( CurvArgs a, CurvFlux f, int k_lo, int k_hi )
{
   const Block& b = a.b;
   const int li = blockIdx.x * blockDim.x + threadIdx.x;
   const int lj = blockIdx.y * blockDim.y + threadIdx.y;
   const int k = k_lo - 2 + blockIdx.z * blockDim.z + threadIdx.z; // planes k_lo-2 .. k_hi+2 (Z is differenced in r)
   if( li > b.nil - 1 || lj > b.nj - 1 || k > k_hi + 2 ) return;
   const int lk = k - b.kfirst;
   if( lk < 0 || lk > b.nk - 1 ) return;
   const long long sc = b.sc, sp = b.sp, dj = b.ni, dk = b.nij;
   const long long p = (long long)li + dj * lj + dk * lk;
   const double c1 = 2.0 / 3, c2 = -1.0 / 12;
   auto U = [&]( int c, long long q ) { return a.u[c * sc + sp * q]; };
   auto d0 = [&]( int c, long long st ) {
      return c2 * ( U( c, p + 2 * st ) - U( c, p - 2 * st ) ) + c1 * ( U( c, p + st ) - U( c, p - st ) );
   };
   const bool iin = li >= 2 && li <= b.nil - 3, jin = lj >= 2 && lj <= b.nj - 3, kin = lk >= 2 && lk <= b.nk - 3 && k >= k_lo && k <= k_hi;
   const double M = a.mu[p], L = a.la[p], l2m = 2 * M + L;
   const double m1 = a.met[0 * a.msc + a.msp * p], m2 = a.met[1 * a.msc + a.msp * p], m3 = a.met[2 * a.msc + a.msp * p],
		m4 = a.met[3 * a.msc + a.msp * p];
   const double sx = a.strx[li], sy = a.stry[lj];
   double dp[3] = { 0, 0, 0 }, dq[3] = { 0, 0, 0 }, dr[3] = { 0, 0, 0 };
   if( iin )
   {
#pragma unroll
      for( int c = 0; c < 3; c++ ) dp[c] = d0( c, 1 );
   }
   if( jin )
   {
#pragma unroll
      for( int c = 0; c < 3; c++ ) dq[c] = d0( c, dj );
   }
   if( kin )
   {
#pragma unroll
      for( int c = 0; c < 3; c++ ) dr[c] = d0( c, dk );
   }
   if( jin && kin ) // X: differenced in p by the points (i +- 1,2, j, k)
   {
      const double m12s = m1 * m2 * sx, m13y = m1 * m3 * sy, m11y = m1 * m1 * sy, m14 = m1 * m4;
      f.x[0][p] = L * m11y * dq[1] + l2m * m12s * dr[0] + L * m13y * dr[1] + L * m14 * dr[2];
      f.x[1][p] = M * m11y * dq[0] + M * m13y * dr[0] + M * m12s * dr[1];
      f.x[2][p] = M * m14 * dr[0] + M * m12s * dr[2];
   }
   if( iin && kin ) // Y: differenced in q
   {
      const double m13s = m1 * m3 * sy, m12x = m1 * m2 * sx, m11x = m1 * m1 * sx, m14 = m1 * m4;
      f.y[0][p] = M * m11x * dp[1] + M * m13s * dr[0] + M * m12x * dr[1];
      f.y[1][p] = L * m11x * dp[0] + L * m12x * dr[0] + l2m * m13s * dr[1] + L * m14 * dr[2];
      f.y[2][p] = M * m14 * dr[1] + M * m13s * dr[2];
   }
   if( iin && jin ) // Z: differenced in r (in-plane differences only: defined on the ghost planes too)
   {
      const double a1 = m2 * sx, a2 = m3 * sy, a3 = m4, e = m1 * sx, g = m1 * sy;
      f.z[0][p] = l2m * a1 * e * dp[0] + M * a2 * e * dp[1] + M * a3 * e * dp[2] + M * a2 * g * dq[0] + L * a1 * g * dq[1];
      f.z[1][p] = L * a2 * e * dp[0] + M * a1 * e * dp[1] + M * a1 * g * dq[0] + l2m * a2 * g * dq[1] + M * a3 * g * dq[2];
      f.z[2][p] = L * a3 * e * dp[0] + M * a1 * e * dp[2] + L * a3 * g * dq[1] + M * a2 * g * dq[2];
   }
}

__global__ void __launch_bounds__( 128 ) k_curv_apply( CurvArgs a, CurvFlux f, int k_lo, int k_hi )
{
   const Block& b = a.b;
   const int li = 2 + blockIdx.x * blockDim.x + threadIdx.x;
   const int lj = 2 + blockIdx.y * blockDim.y + threadIdx.y;
   const int k = k_lo + blockIdx.z * blockDim.z + threadIdx.z;
   if( li > b.nil - 3 || lj > b.nj - 3 || k > k_hi ) return;
   const int lk = k - b.kfirst;
   const long long sc = b.sc, sp = b.sp, dj = b.ni, dk = b.nij;
   const long long p = (long long)li + dj * lj + dk * lk;
   const double c1 = 2.0 / 3, c2 = -1.0 / 12, i6 = 1.0 / 6;
   auto U = [&]( int c, long long q ) { return a.u[c * sc + sp * q]; };
   auto MET = [&]( int c, long long q ) { return a.met[c * a.msc + a.msp * q]; };
   const double sxi = a.strx[li], syj = a.stry[lj];
   const double u0[3] = { U( 0, p ), U( 1, p ), U( 2, p ) };
   double r[3];
   // ---- p and q: G(mu m1^2 s) on two components, G((2mu+la) m1^2 s) on the own one, plus the centred difference of the flux
   auto inplane = [&]( long long st, const double* str, int own, double* const* flux, double out[3] ) {
      double cm[5], cl[5], wm[4], wl[4];
#pragma unroll
      for( int m = -2; m <= 2; m++ )
      {
	 const long long q = p + m * st;
	 const double M = a.mu[q], L = a.la[q], m1 = MET( 0, q );
	 const double t = m1 * m1 * str[m];
	 cm[m + 2] = M * t;
	 cl[m + 2] = ( 2 * M + L ) * t;
      }
      cweights4( cl, wl );
      cweights4( cm, wm );
#pragma unroll
      for( int c = 0; c < 3; c++ )
      {
	 const double* w = c == own ? wl : wm;
	 const double g = w[0] * ( U( c, p - 2 * st ) - u0[c] ) + w[1] * ( U( c, p - st ) - u0[c] ) + w[2] * ( U( c, p + st ) - u0[c] ) +
			  w[3] * ( U( c, p + 2 * st ) - u0[c] );
	 out[c] = i6 * g + ( c2 * ( flux[c][p + 2 * st] - flux[c][p - 2 * st] ) + c1 * ( flux[c][p + st] - flux[c][p - st] ) );
      }
   };
   double rp[3], rq[3];
   inplane( 1, a.strx + li, 0, f.x, rp );
   inplane( dj, a.stry + lj, 1, f.y, rq );
   // ---- r: the full coefficient matrix N^{rr} = la A A^T + mu (|A|^2 I + A A^T), A = (m2 sx, m3 sy, m4)
   {
      double n11[5], n22[5], n33[5], n12[5], n13[5], n23[5];
#pragma unroll
      for( int m = -2; m <= 2; m++ )
      {
	 const long long q = p + m * dk;
	 const double M = a.mu[q], L = a.la[q];
	 const double a1 = MET( 1, q ) * sxi, a2 = MET( 2, q ) * syj, a3 = MET( 3, q );
	 const double l2m = 2 * M + L, lm = M + L;
	 n11[m + 2] = l2m * a1 * a1 + M * ( a2 * a2 + a3 * a3 );
	 n22[m + 2] = l2m * a2 * a2 + M * ( a1 * a1 + a3 * a3 );
	 n33[m + 2] = l2m * a3 * a3 + M * ( a1 * a1 + a2 * a2 );
	 n12[m + 2] = lm * a1 * a2;
	 n13[m + 2] = lm * a1 * a3;
	 n23[m + 2] = lm * a2 * a3;
      }
      double w11[4], w22[4], w33[4], w12[4], w13[4], w23[4];
      cweights4( n11, w11 ); cweights4( n22, w22 ); cweights4( n33, w33 );
      cweights4( n12, w12 ); cweights4( n13, w13 ); cweights4( n23, w23 );
      double du[3][4];
#pragma unroll
      for( int c = 0; c < 3; c++ )
      {
	 du[c][0] = U( c, p - 2 * dk ) - u0[c]; du[c][1] = U( c, p - dk ) - u0[c];
	 du[c][2] = U( c, p + dk ) - u0[c]; du[c][3] = U( c, p + 2 * dk ) - u0[c];
      }
      auto g = [&]( const double w[4], int d ) { return w[0] * du[d][0] + w[1] * du[d][1] + w[2] * du[d][2] + w[3] * du[d][3]; };
      r[0] = i6 * ( g( w11, 0 ) + g( w12, 1 ) + g( w13, 2 ) );
      r[1] = i6 * ( g( w12, 0 ) + g( w22, 1 ) + g( w23, 2 ) );
      r[2] = i6 * ( g( w13, 0 ) + g( w23, 1 ) + g( w33, 2 ) );
#pragma unroll
      for( int c = 0; c < 3; c++ )
	 r[c] += c2 * ( f.z[c][p + 2 * dk] - f.z[c][p - 2 * dk] ) + c1 * ( f.z[c][p + dk] - f.z[c][p - dk] );
   }
   const double ij = 1.0 / a.jac[p];
#pragma unroll
   for( int c = 0; c < 3; c++ ) a.lu[c * sc + sp * p] = ( sxi * rp[c] + syj * rq[c] + r[c] ) * ij;
}

// supergrid damping on the curvilinear grid: x and y terms only, weights rho*dc*jac, prefactor beta/(rho*jac)
__device__ __forceinline__ double sgdc_dir( int order, const double* __restrict__ u, const double* __restrict__ um,
					    const double* __restrict__ rho, const double* __restrict__ jac,
					    long long q, long long p, long long sp, long long st, const double* dc )
{
   const long long su = sp * st;
   if( order == 4 )
   {
      double d[5];
#pragma unroll
      for( int m = 0; m < 5; m++ ) d[m] = u[q + ( m - 2 ) * su] - um[q + ( m - 2 ) * su];
      const double e0 = rho[p - st] * dc[-1] * jac[p - st] * ( d[2] - 2 * d[1] + d[0] );
      const double e1 = rho[p] * dc[0] * jac[p] * ( d[3] - 2 * d[2] + d[1] );
      const double e2 = rho[p + st] * dc[1] * jac[p + st] * ( d[4] - 2 * d[3] + d[2] );
      return e2 - 2 * e1 + e0;
   }
   double d[7];
#pragma unroll
   for( int m = 0; m < 7; m++ ) d[m] = u[q + ( m - 3 ) * su] - um[q + ( m - 3 ) * su];
   double acc = 0;
   const double cw[4] = { -1, 3, -3, 1 };
#pragma unroll
   for( int m = -2; m <= 1; m++ )
   {
      const double A = rho[p + ( m + 1 ) * st] * dc[m + 1] * jac[p + ( m + 1 ) * st] + rho[p + m * st] * dc[m] * jac[p + m * st];
      const double T = d[m + 5] - 3 * d[m + 4] + 3 * d[m + 3] - d[m + 2];
      acc += cw[m + 2] * A * T;
   }
   return -0.5 * acc;
}

__global__ void k_addsgdc( int order, Block b, double* __restrict__ up, const double* __restrict__ u,
			   const double* __restrict__ um, const double* __restrict__ rho,
			   const double* __restrict__ dcx, const double* __restrict__ dcy,
			   const double* __restrict__ strx, const double* __restrict__ stry,
			   const double* __restrict__ jac, const double* __restrict__ cox,
			   const double* __restrict__ coy, double beta )
{
   const int w = order == 6 ? 3 : 2;
   const int ii = w + blockIdx.x * blockDim.x + threadIdx.x;
   const int jj = w + blockIdx.y * blockDim.y + threadIdx.y;
   const int kk = w + blockIdx.z * blockDim.z + threadIdx.z;
   if( ii > b.nil - 1 - w || jj > b.nj - 1 - w || kk > b.nk - 1 - w ) return;
   const long long p = (long long)ii + (long long)b.ni * jj + b.nij * kk;
   const double irhoj = beta / ( rho[p] * jac[p] );
   const double prex = strx[ii] * coy[jj], prey = stry[jj] * cox[ii];
#pragma unroll
   for( int c = 0; c < 3; c++ )
   {
      const long long q = c * b.sc + b.sp * p;
      const double s = prex * sgdc_dir( order, u, um, rho, jac, q, p, b.sp, 1LL, dcx + ii ) +
		       prey * sgdc_dir( order, u, um, rho, jac, q, p, b.sp, (long long)b.ni, dcy + jj );
      up[q] -= irhoj * s;
   }
}

// traction-free ghost plane of the curvilinear grid: side 5 (k=1, ghost k=0) or 6 (k=nz, ghost k=nz+1)
__global__ void k_freesurfcurvisg( CurvArgs a, double* __restrict__ u, int k, int kl, const double* __restrict__ forcing )
{
   const Block& b = a.b;
   const int li = 2 + blockIdx.x * blockDim.x + threadIdx.x;
   const int lj = 2 + blockIdx.y * blockDim.y + threadIdx.y;
   if( li > b.nil - 3 || lj > b.nj - 3 ) return;
   const long long sc = b.sc, sp = b.sp, dj = b.ni, dk = b.nij;
   const double c1 = 2.0 / 3, c2 = -1.0 / 12;
   const long long qq = (long long)li + (long long)b.ni * lj;
   const long long p = qq + dk * ( k - b.kfirst );
   auto U = [&]( int c, long long q ) { return u[c * sc + sp * q]; };
   auto d0 = [&]( int c, long long st ) {
      return c2 * ( U( c, p + 2 * st ) - U( c, p - 2 * st ) ) + c1 * ( U( c, p + st ) - U( c, p - st ) );
   };
   const double M = a.mu[p], L = a.la[p];
   const double m1 = a.met[0 * a.msc + a.msp * p], m2 = a.met[1 * a.msc + a.msp * p];
   const double m3 = a.met[2 * a.msc + a.msp * p], m4 = a.met[3 * a.msc + a.msp * p];
   const double sx = a.strx[li], sy = a.stry[lj], isx = 1 / sx, isy = 1 / sy;
   const double dp0 = d0( 0, 1 ), dp1 = d0( 1, 1 ), dp2 = d0( 2, 1 );
   const double dq0 = d0( 0, dj ), dq1 = d0( 1, dj ), dq2 = d0( 2, dj );
   const long long qb = (long long)li + (long long)b.nil * lj; // forcing(3,i,j) of the surface window
   const double f0 = forcing ? forcing[3 * qb] : 0.0, f1 = forcing ? forcing[3 * qb + 1] : 0.0, f2 = forcing ? forcing[3 * qb + 2] : 0.0;
   // tangential part of the normal traction, divided by sx*sy (curvilinear-c.C:520-573)
   const double rhs1 = ( 2 * M + L ) * m2 * m1 * dp0 * sx * isy + M * m3 * m1 * dp1 + M * m4 * m1 * dp2 * isy +
		       M * m3 * m1 * dq0 * isx * sy + L * m2 * m1 * dq1 - f0;
   const double rhs2 = L * m3 * m1 * dp0 + M * m2 * m1 * dp1 * sx * isy + M * m2 * m1 * dq0 +
		       ( 2 * M + L ) * m3 * m1 * dq1 * sy * isx + M * m4 * m1 * dq2 * isx - f1;
   const double rhs3 = L * m4 * m1 * dp0 * isy + M * m2 * m1 * dp2 * sx * isy + M * m3 * m1 * dq2 * sy * isx +
		       L * m4 * m1 * dq1 * isx - f2;
   // normal system (mu |a|^2 I + (la+mu) a a^T) solved with Sherman-Morrison, a = scaled A_r
   const double xoysqrt = sqrt( sx * isy ), yoxsqrt = 1 / xoysqrt, isqrtxy = isx * xoysqrt;
   const double ac = sx * isy * m2 * m2 + sy * isx * m3 * m3 + m4 * m4 * isy * isx;
   const double bc = 1 / ( M * ac );
   const double cc = ( M + L ) / ( 2 * M + L ) * bc / ac;
   const double dc = cc * ( xoysqrt * m2 * rhs1 + yoxsqrt * m3 * rhs2 + isqrtxy * m4 * rhs3 );
   const double s0i = 1 / c_sbop[0];
   const double rhs[3] = { rhs1, rhs2, rhs3 };
   const double av[3] = { m2 * xoysqrt, m3 * yoxsqrt, m4 * isqrtxy };
#pragma unroll
   for( int c = 0; c < 3; c++ )
   {
      const double s = c_sbop[1] * U( c, p ) + c_sbop[2] * U( c, p + kl * dk ) + c_sbop[3] * U( c, p + 2 * kl * dk ) +
		       c_sbop[4] * U( c, p + 3 * kl * dk );
      u[c * sc + sp * ( p - kl * dk )] = -s0i * ( s + bc * rhs[c] - dc * av[c] );
   }
}

// interface injection between the top Cartesian grid (bc) and the curvilinear grid (bt) above it
__global__ void k_enforce_cart_topo( Block bc, Block bt, double* __restrict__ ucart, double* __restrict__ ucurv )
{
   const long long nij = bc.nij;
   for( long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nij; t += (long long)gridDim.x * blockDim.x )
   {
#pragma unroll
      for( int c = 0; c < 3; c++ )
      {
	 // Cartesian ghost planes kStart+q <- curvilinear planes kEnd-4+q, q=0,1
#pragma unroll
	 for( int q = 0; q < 2; q++ )
	    ucart[c * bc.sc + bc.sp * ( t + nij * q )] = ucurv[c * bt.sc + bt.sp * ( t + nij * ( bt.nk - 1 - 4 + q ) )];
	 // curvilinear planes kEnd-q <- Cartesian planes kStart+4-q, q=0,1,2
#pragma unroll
	 for( int q = 0; q <= 2; q++ )
	    ucurv[c * bt.sc + bt.sp * ( t + nij * ( bt.nk - 1 - q ) )] = ucart[c * bc.sc + bc.sp * ( t + nij * ( 4 - q ) )];
      }
   }
}

} // namespace

static void curv_args( CurvArgs& a, const Block& b, const double* u, const double* mu, const double* la,
		       const double* met, const double* jac, double* lu, const double* strx, const double* stry )
{
   a.b = b; a.u = u; a.mu = mu; a.la = la; a.met = met; a.jac = jac; a.lu = lu; a.strx = strx; a.stry = stry;
   a.kfirst = b.kfirst;
   if( b.sp == 1 ) { a.msc = b.npts; a.msp = 1; }
   else            { a.msc = 1;      a.msp = 4; }
}

int launch_rhs4sgcurv( const Block& b, const double* u, const double* mu, const double* la, const double* met,
		       const double* jac, double* lu, int onesided4, const double* strx, const double* stry,
		       cudaStream_t st, double* flux_scratch )
{
   if( b.nil < 5 || b.nj < 5 || b.nk < 5 ) return 0;
   CurvArgs a;
   curv_args( a, b, u, mu, la, met, jac, lu, strx, stry );
   int kstart = b.kfirst + 2;
   const int kend = b.klast - 2;
   dim3 bs( 32, 4, 1 );
   ProfScope prof( "rhs4sgcurv", st );
   if( onesided4 )
   {
      if( b.kfirst > 0 || b.klast < 8 )
	 return set_error( "rhs4sgcurv: the free-surface closure needs planes 0..8 inside the block (k range %d:%d)", b.kfirst, b.klast );
      dim3 gs( ( b.nil - 4 + bs.x - 1 ) / bs.x, ( b.nj - 4 + bs.y - 1 ) / bs.y, 6 );
      k_rhs4sgcurv<true><<<gs, bs, 0, st>>>( a, 1, 6 );
      count_launch();
      kstart = 7;
   }
   if( kend >= kstart && flux_scratch )
   {
      // interior rows in two sweeps: fluxes once per point (9 * npts doubles of scratch), then the operator
      CurvFlux f;
      for( int c = 0; c < 3; c++ )
      {
	 f.x[c] = flux_scratch + (size_t)c * b.npts;
	 f.y[c] = flux_scratch + (size_t)( 3 + c ) * b.npts;
	 f.z[c] = flux_scratch + (size_t)( 6 + c ) * b.npts;
      }
      dim3 gf( ( b.nil + bs.x - 1 ) / bs.x, ( b.nj + bs.y - 1 ) / bs.y, kend - kstart + 5 );
      k_curv_flux<<<gf, bs, 0, st>>>( a, f, kstart, kend );
      dim3 gs( ( b.nil - 4 + bs.x - 1 ) / bs.x, ( b.nj - 4 + bs.y - 1 ) / bs.y, kend - kstart + 1 );
      k_curv_apply<<<gs, bs, 0, st>>>( a, f, kstart, kend );
      count_launch( 2 );
   }
   else if( kend >= kstart )
   {
      dim3 gs( ( b.nil - 4 + bs.x - 1 ) / bs.x, ( b.nj - 4 + bs.y - 1 ) / bs.y, kend - kstart + 1 );
      k_rhs4sgcurv<false><<<gs, bs, 0, st>>>( a, kstart, kend );
      count_launch();
   }
   return check_launch( "k_rhs4sgcurv" );
}

int launch_addsgdc( int order, const Block& b, double* up, const double* u, const double* um, const double* rho,
		    const double* dcx, const double* dcy, const double* strx, const double* stry,
		    const double* jac, const double* cox, const double* coy, double beta, cudaStream_t st )
{
   if( beta == 0 ) return 0;
   const int w = order == 6 ? 3 : 2;
   if( b.nil <= 2 * w || b.nj <= 2 * w || b.nk <= 2 * w ) return 0;
   ProfScope prof( "addsgdc", st );
   dim3 bs( 32, 4, 2 );
   dim3 gs( ( b.nil - 2 * w + bs.x - 1 ) / bs.x, ( b.nj - 2 * w + bs.y - 1 ) / bs.y, ( b.nk - 2 * w + bs.z - 1 ) / bs.z );
   k_addsgdc<<<gs, bs, 0, st>>>( order, b, up, u, um, rho, dcx, dcy, strx, stry, jac, cox, coy, beta );
   count_launch();
   return check_launch( "k_addsgdc" );
}

int launch_freesurfcurvisg( const Block& b, int nz, int side, double* u, const double* mu, const double* la,
			    const double* met, const double* forcing, const double* strx, const double* stry,
			    cudaStream_t st )
{
   if( side != 5 && side != 6 ) return set_error( "freesurfcurvisg: side must be 5 (k=1) or 6 (k=nz)" );
   if( b.nil < 5 || b.nj < 5 ) return 0;
   CurvArgs a;
   curv_args( a, b, u, mu, la, met, 0, 0, strx, stry );
   dim3 bs( 32, 8 );
   dim3 gs( ( b.nil - 4 + 31 ) / 32, ( b.nj - 4 + 7 ) / 8 );
   k_freesurfcurvisg<<<gs, bs, 0, st>>>( a, u, side == 5 ? 1 : nz, side == 5 ? 1 : -1, forcing );
   count_launch();
   return check_launch( "k_freesurfcurvisg" );
}

int launch_enforce_cart_topo( int corder, double* ucart, const Block& bc, double* ucurv, const Block& bt,
			      cudaStream_t st )
{
   (void)corder;
   if( bc.nil != bt.nil || bc.ni != bt.ni || bc.nj != bt.nj ) return set_error( "enforce_cart_topo: the two grids must share their i,j extents" );
   if( bt.nk < 5 || bc.nk < 5 ) return set_error( "enforce_cart_topo: grids too thin" );
   long long g = ( bc.nij + 255 ) / 256;
   if( g > 148 * 8 ) g = 148 * 8;
   k_enforce_cart_topo<<<(int)g, 256, 0, st>>>( bc, bt, ucart, ucurv );
   count_launch();
   return check_launch( "k_enforce_cart_topo" );
}

} // namespace sw4b200
