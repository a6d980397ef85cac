// Cartesian-grid operators, general (layout-agnostic) CUDA kernels: one thread per grid point,
// neighbours through L1/L2.  These are the correctness baseline for every layout/option
// (AoS and SoA, both closures, 6th-order damping, dense forcing) and the path used for the
// thin SBP-closure rows and the ghost shell.  The throughput path for the SoA interior is
// rhs4sg_fast.cu.
//
// Reference semantics: rhs4sg.C:38-849 / rhs4sg_rev.C:44-864 (L(u)), ew-cfromfort.C:40-141
// (corrector, predictor, dpdmt), :748-1160 (supergrid damping), :205-745 (bcfortsg).
#include "common.cuh"
#include "tma.cuh"
#include "sbp4_constexpr.h"
#include <cstdlib>

namespace sw4b200 {

#define ACOF( k, q, m ) c_acof[( (k)-1 ) + 6 * ( (q)-1 ) + 48 * ( (m)-1 )]
#define BOPE( k, q ) c_bope[( (k)-1 ) + 6 * ( (q)-1 )]

namespace {

template <int MODE>
struct Acc
{
   const double *u, *um, *up;
   long long sc, sp;
   double dt2i;
   __device__ __forceinline__ double operator()( int c, long long p ) const
   {
      const long long q = c * sc + sp * p;
      if( MODE == MODE_CORR )
	 return dt2i * ( up[q] - 2 * u[q] + um[q] );
      return u[q];
   }
};

__device__ __forceinline__ void weights4( const double a[5], double w[4] )
{
   w[0] = a[1] - 0.75 * ( a[2] + a[0] );
   w[1] = a[0] + a[3] + 3 * ( a[2] + a[1] );
   w[2] = a[1] + a[4] + 3 * ( a[3] + a[2] );
   w[3] = a[3] - 0.75 * ( a[2] + a[4] );
}
__device__ __forceinline__ double gsum( const double w[4], const double f[5] )
{
   return w[0] * ( f[0] - f[2] ) + w[1] * ( f[1] - f[2] ) + w[2] * ( f[3] - f[2] ) + w[3] * ( f[4] - f[2] );
}
__device__ __forceinline__ double d0( double fm2, double fm1, double fp1, double fp2 )
{
   return ( fm2 - fp2 + 8 * ( fp1 - fm1 ) ) * ( 1.0 / 12 );
}

// second-derivative terms in direction with stride st; s[5] = stretch at the 5 points; comp d is "normal"
template <class A>
__device__ __forceinline__ void second_derivative( const A& U, const double* __restrict__ mu,
						   const double* __restrict__ la, long long p, long long st,
						   const double s[5], int d, double r[3] )
{
   double am[5], bm[5], wm[4], wb[4], f[5];
#pragma unroll
   for( int m = 0; m < 5; m++ )
   {
      const double mm = mu[p + ( m - 2 ) * st];
      am[m] = mm * s[m];
      bm[m] = ( 2 * mm + la[p + ( m - 2 ) * st] ) * s[m];
   }
   weights4( am, wm );
   weights4( bm, wb );
#pragma unroll
   for( int c = 0; c < 3; c++ )
   {
#pragma unroll
      for( int m = 0; m < 5; m++ ) f[m] = U( c, p + ( m - 2 ) * st );
      r[c] += ( 1.0 / 6 ) * s[2] * gsum( c == d ? wb : wm, f );
   }
}

// s_a s_b [ D0_a( la D0_b u_b ) + D0_b( mu D0_a u_b ) ] added to r_a
template <class A>
__device__ __forceinline__ double mixed_pair( const A& U, const double* __restrict__ mu,
					      const double* __restrict__ la, long long p, long long sa,
					      long long sb, int b )
{
   double t1[5], t2[5];
#pragma unroll
   for( int m = 0; m < 5; m++ )
   {
      if( m == 2 ) continue;
      long long q = p + ( m - 2 ) * sa;
      t1[m] = la[q] * d0( U( b, q - 2 * sb ), U( b, q - sb ), U( b, q + sb ), U( b, q + 2 * sb ) );
      q = p + ( m - 2 ) * sb;
      t2[m] = mu[q] * d0( U( b, q - 2 * sa ), U( b, q - sa ), U( b, q + sa ), U( b, q + 2 * sa ) );
   }
   return d0( t1[0], t1[1], t1[3], t1[4] ) + d0( t2[0], t2[1], t2[3], t2[4] );
}

template <class A>
__device__ void rhs_interior_point( const A& U, const double* __restrict__ mu, const double* __restrict__ la,
				    long long p, long long dj, long long dk, const double sx[5],
				    const double sy[5], const double sz[5], double r[3] )
{
   r[0] = r[1] = r[2] = 0;
   second_derivative( U, mu, la, p, 1LL, sx, 0, r );
   second_derivative( U, mu, la, p, dj, sy, 1, r );
   second_derivative( U, mu, la, p, dk, sz, 2, r );
   r[0] += sx[2] * sy[2] * mixed_pair( U, mu, la, p, 1LL, dj, 1 ) + sx[2] * sz[2] * mixed_pair( U, mu, la, p, 1LL, dk, 2 );
   r[1] += sx[2] * sy[2] * mixed_pair( U, mu, la, p, dj, 1LL, 0 ) + sy[2] * sz[2] * mixed_pair( U, mu, la, p, dj, dk, 2 );
   r[2] += sx[2] * sz[2] * mixed_pair( U, mu, la, p, dk, 1LL, 0 ) + sy[2] * sz[2] * mixed_pair( U, mu, la, p, dk, dj, 1 );
}

// SBP closure row kb (1..6) of side (0: low-k, planes q -> k=q; 1: high-k, planes q -> nk-q+1).
// pcol = index of (i,j,k=kfirst) column base; plane(k) = pcol + dk*(k-kfirst).
template <class A>
__device__ void rhs_closure_point( const A& U, const double* __restrict__ mu, const double* __restrict__ la,
				   long long pcol, int kfirst, int nk, int side, int kb, long long dj,
				   long long dk, const double sx[5], const double sy[5], double r[3] )
{
   auto plane = [&]( int q ) -> long long {
      const int k = side == 0 ? q : nk - q + 1;
      return pcol + dk * ( k - kfirst );
   };
   const double sgn = side == 0 ? 1.0 : -1.0;
   const long long p = plane( kb );
   r[0] = r[1] = r[2] = 0;
   second_derivative( U, mu, la, p, 1LL, sx, 0, r );
   second_derivative( U, mu, la, p, dj, sy, 1, r );
   // z second derivative, boundary-modified, no strz and no 1/6
   double muq[8], laq[8];
#pragma unroll
   for( int m = 0; m < 8; m++ )
   {
      muq[m] = mu[plane( m + 1 )];
      laq[m] = 2 * muq[m] + la[plane( m + 1 )];
   }
   // (zero entries of the tables are skipped: they contribute exact zeros in the reference's sums)
#pragma unroll
   for( int q = 1; q <= 8; q++ )
   {
      double mucof = 0, lap2mu = 0;
      bool any = false;
#pragma unroll
      for( int m = 1; m <= 8; m++ )
      {
	 const double a = ACOF( kb, q, m );
	 if( a != 0 )
	 {
	    mucof += a * muq[m - 1];
	    lap2mu += a * laq[m - 1];
	    any = true;
	 }
      }
      if( !any ) continue;
      const long long pq = plane( q );
      r[0] += mucof * U( 0, pq );
      r[1] += mucof * U( 1, pq );
      r[2] += lap2mu * U( 2, pq );
   }
   if( c_ghcof[kb - 1] != 0 )
   {
      const long long pg = plane( 0 );
      const double g = c_ghcof[kb - 1];
      r[0] += g * muq[0] * U( 0, pg );
      r[1] += g * muq[0] * U( 1, pg );
      r[2] += g * laq[0] * U( 2, pg );
   }
   // xy cross terms, centred
   r[0] += sx[2] * sy[2] * mixed_pair( U, mu, la, p, 1LL, dj, 1 );
   r[1] += sx[2] * sy[2] * mixed_pair( U, mu, la, p, dj, 1LL, 0 );
   // terms with one z-derivative: D0z -> sgn*sum_q bope(kb,q) f(plane q)
#pragma unroll
   for( int a = 0; a < 2; a++ )
   {
      const long long sa = a == 0 ? 1LL : dj;
      const double fa = a == 0 ? sx[2] : sy[2];
      double bw[5] = { 0, 0, 0, 0, 0 }, bu[5] = { 0, 0, 0, 0, 0 };
      double zmw = 0, zlu = 0;
#pragma unroll
      for( int q = 1; q <= 8; q++ )
      {
	 const double bq = sgn * BOPE( kb, q );
	 if( bq == 0 ) continue;
	 const long long pq = plane( q );
	 const double wm2 = U( 2, pq - 2 * sa ), wm1 = U( 2, pq - sa ), wp1 = U( 2, pq + sa ), wp2 = U( 2, pq + 2 * sa );
	 const double um2 = U( a, pq - 2 * sa ), um1 = U( a, pq - sa ), up1 = U( a, pq + sa ), up2 = U( a, pq + 2 * sa );
	 bw[0] += bq * wm2; bw[1] += bq * wm1; bw[3] += bq * wp1; bw[4] += bq * wp2;
	 bu[0] += bq * um2; bu[1] += bq * um1; bu[3] += bq * up1; bu[4] += bq * up2;
	 zmw += bq * ( muq[q - 1] * d0( wm2, wm1, wp1, wp2 ) );
	 zlu += bq * ( la[pq] * d0( um2, um1, up1, up2 ) );
      }
#pragma unroll
      for( int m = 0; m < 5; m++ )
      {
	 if( m == 2 ) continue;
	 bw[m] *= la[p + ( m - 2 ) * sa];
	 bu[m] *= mu[p + ( m - 2 ) * sa];
      }
      r[a] += fa * ( d0( bw[0], bw[1], bw[3], bw[4] ) + zmw );
      r[2] += fa * ( d0( bu[0], bu[1], bu[3], bu[4] ) + zlu );
   }
}

// supergrid damping of one point, all three components: returns sum_d pre_d * D_d(...)
__device__ __forceinline__ double sgd_point( int order, const double* __restrict__ u, const double* __restrict__ um,
					     const double* __restrict__ rho, long long q, long long p,
					     long long sp, long long st, const double* dc, double pre )
{
   // q = index of component value at the point in u/um; point stride in u is sp*st
   const long long su = sp * st;
   if( order == 4 )
   {
      double d[5];
#pragma unroll
      for( int m = 0; m < 5; m++ ) d[m] = u[q + ( m - 2 ) * su] - um[q + ( m - 2 ) * su];
      const double e0 = rho[p - st] * dc[-1] * ( d[2] - 2 * d[1] + d[0] );
      const double e1 = rho[p] * dc[0] * ( d[3] - 2 * d[2] + d[1] );
      const double e2 = rho[p + st] * dc[1] * ( d[4] - 2 * d[3] + d[2] );
      return pre * ( e2 - 2 * e1 + e0 );
   }
   else
   {
      double d[7];
#pragma unroll
      for( int m = 0; m < 7; m++ ) d[m] = u[q + ( m - 3 ) * su] - um[q + ( m - 3 ) * su];
      double acc = 0;
      const double cw[4] = { -1, 3, -3, 1 };
#pragma unroll
      for( int m = -2; m <= 1; m++ )
      {
	 const double A = rho[p + ( m + 1 ) * st] * dc[m + 1] + rho[p + m * st] * dc[m];
	 const double T = d[m + 5] - 3 * d[m + 4] + 3 * d[m + 3] - d[m + 2];
	 acc += cw[m + 2] * A * T;
      }
      return pre * ( -0.5 * acc );
   }
}

template <int MODE>
__device__ __forceinline__ void rhs_epilogue( const RhsArgs& a, long long p, int i, int j, int k, double cof, const double r[3] )
{
   const Block& b = a.b;
   if( MODE == MODE_LU )
   {
#pragma unroll
      for( int c = 0; c < 3; c++ ) a.out[c * b.sc + b.sp * p] = cof * r[c];
   }
   else if( MODE == MODE_PRED )
   {
      const double f = ( a.dt * a.dt ) / a.rho[p];
#pragma unroll
      for( int c = 0; c < 3; c++ )
      {
	 const long long q = c * b.sc + b.sp * p;
	 const double fo = a.fo ? a.fo[q] : 0.0;
	 const double acc = cof * r[c] + fo;
	 a.out[q] = 2 * a.u[q] - a.um[q] + f * acc;
	 if( a.out2 ) a.out2[q] = acc / a.rho[p];
      }
   }
   else if( MODE == MODE_CORR_ACC )
   {
      // field = stored uacc array; in-place correction of up at the own point
      const double dt2 = a.dt * a.dt;
      const double f = ( dt2 * dt2 / 12 ) / a.rho[p];
#pragma unroll
      for( int c = 0; c < 3; c++ )
      {
	 const long long q = c * b.sc + b.sp * p;
	 const double fo = a.fo ? a.fo[q] : 0.0;
	 a.out[q] = a.up[q] + f * ( cof * r[c] + fo );
      }
   }
   else
   {
      const double dt2 = a.dt * a.dt;
      const double f = ( dt2 * dt2 / 12 ) / a.rho[p];
      double sg[3] = { 0, 0, 0 };
      const int w = a.sg_order == 6 ? 3 : 2;
      if( a.sg_order != 0 && a.beta != 0 && i >= b.ifirst + w && i <= b.ilast - w && j >= b.jfirst + w &&
	  j <= b.jlast - w && k >= b.kfirst + w && k <= b.klast - w )
      {
	 const int ii = i - b.ifirst, jj = j - b.jfirst, kk = k - b.kfirst;
	 const double prex = a.strx[ii] * a.coy[jj] * a.coz[kk];
	 const double prey = a.stry[jj] * a.cox[ii] * a.coz[kk];
	 const double prez = a.strz[kk] * a.cox[ii] * a.coy[jj];
	 const double birho = a.beta / a.rho[p];
#pragma unroll
	 for( int c = 0; c < 3; c++ )
	 {
	    const long long q = c * b.sc + b.sp * p;
	    sg[c] = birho * ( sgd_point( a.sg_order, a.u, a.um, a.rho, q, p, b.sp, 1LL, a.dcx + ii, prex ) +
			      sgd_point( a.sg_order, a.u, a.um, a.rho, q, p, b.sp, (long long)b.ni, a.dcy + jj, prey ) +
			      sgd_point( a.sg_order, a.u, a.um, a.rho, q, p, b.sp, b.nij, a.dcz + kk, prez ) );
	 }
      }
#pragma unroll
      for( int c = 0; c < 3; c++ )
      {
	 const long long q = c * b.sc + b.sp * p;
	 const double fo = a.fo ? a.fo[q] : 0.0;
	 a.out[q] = ( a.up[q] + f * ( cof * r[c] + fo ) ) - sg[c];
      }
   }
}

template <int MODE>
__global__ void __launch_bounds__( 256 ) k_rhs_v1( RhsArgs a, int k_lo, int k_hi )
{
   const Block& b = a.b;
   const int i = b.ifirst + 2 + blockIdx.x * blockDim.x + threadIdx.x;
   const int j = b.jfirst + 2 + blockIdx.y * blockDim.y + threadIdx.y;
   const int k = k_lo + blockIdx.z * blockDim.z + threadIdx.z;
   if( i > b.ilast - 2 || j > b.jlast - 2 || k > k_hi ) return;
   Acc<MODE> U = { a.u, a.um, a.up, b.sc, b.sp, 1.0 / ( a.dt * a.dt ) };
   double sx[5], sy[5], sz[5];
#pragma unroll
   for( int m = 0; m < 5; m++ )
   {
      sx[m] = a.strx[i - b.ifirst + m - 2];
      sy[m] = a.stry[j - b.jfirst + m - 2];
   }
   const long long p = pidx( b, i, j, k );
   double r[3];
   const bool low = a.onesided4 && k <= 6;
   const bool high = a.onesided5 && k >= a.nk - 5;
   if( low || high )
   {
      const long long pcol = pidx( b, i, j, b.kfirst );
      rhs_closure_point( U, a.mu, a.la, pcol, b.kfirst, a.nk, low ? 0 : 1, low ? k : a.nk - k + 1,
			 (long long)b.ni, b.nij, sx, sy, r );
   }
   else
   {
#pragma unroll
      for( int m = 0; m < 5; m++ ) sz[m] = a.strz[k - b.kfirst + m - 2];
      rhs_interior_point( U, a.mu, a.la, p, (long long)b.ni, b.nij, sx, sy, sz, r );
   }
   rhs_epilogue<MODE>( a, p, i, j, k, 1.0 / ( a.h * a.h ), r );
}

// SBP closure rows with the 9 planes they read staged in shared memory: one thread per (i,j)
// column of a 32x8 tile computes rows kb_lo..kb_hi of one side (side 0: k=kb, planes 0..8;
// side 1: k=nk-kb+1, planes nk-7..nk+1).  Same arithmetic as k_rhs_v1 (rhs_closure_point); the
// neighbours come from shared memory instead of L1/L2.
constexpr int CL_TX = 32, CL_TY = 8, CL_PX = CL_TX + 4, CL_PY = CL_TY + 4, CL_PLANE = CL_PX * CL_PY, CL_NP = 9;

template <int MODE>
__global__ void __launch_bounds__( CL_TX* CL_TY ) k_closure_staged( RhsArgs a, int side, int kb_lo, int kb_hi )
{
   extern __shared__ double sm[]; // [5 fields][9 planes][CL_PLANE]
   const Block& b = a.b;
   const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * CL_TX + tx;
   const int li0 = 2 + blockIdx.x * CL_TX, lj0 = 2 + blockIdx.y * CL_TY;
   const int kbase = side == 0 ? 0 : a.nk - 7; // global k of staged plane 0
   const double dt2i = 1.0 / ( a.dt * a.dt );
   for( int idx = tid; idx < CL_PLANE; idx += CL_TX * CL_TY )
   {
      const int sy_ = idx / CL_PX, sx_ = idx - sy_ * CL_PX;
      const int li = li0 - 2 + sx_, lj = lj0 - 2 + sy_;
      const bool inb = li < b.ni && lj < b.nj;
#pragma unroll
      for( int s = 0; s < CL_NP; s++ )
      {
	 const long long p = inb ? (long long)li + (long long)b.ni * lj + b.nij * ( kbase + s - b.kfirst ) : 0;
#pragma unroll
	 for( int c = 0; c < 3; c++ )
	 {
	    const long long q = c * b.sc + b.sp * p;
	    double v = 0;
	    if( inb ) v = MODE == MODE_CORR ? dt2i * ( a.up[q] - 2 * a.u[q] + a.um[q] ) : a.u[q];
	    sm[( c * CL_NP + s ) * CL_PLANE + idx] = v;
	 }
	 sm[( 3 * CL_NP + s ) * CL_PLANE + idx] = inb ? a.mu[p] : 0.0;
	 sm[( 4 * CL_NP + s ) * CL_PLANE + idx] = inb ? a.la[p] : 0.0;
      }
   }
   __syncthreads();
   const int li = li0 + tx, lj = lj0 + ty;
   if( li > b.nil - 3 || lj > b.nj - 3 ) return;
   Acc<MODE_LU> U = { sm, 0, 0, (long long)CL_NP * CL_PLANE, 1, 0.0 };
   const double* smu = sm + 3 * CL_NP * CL_PLANE;
   const double* sla = sm + 4 * CL_NP * CL_PLANE;
   double sx[5], sy[5];
#pragma unroll
   for( int m = 0; m < 5; m++ )
   {
      sx[m] = a.strx[li + m - 2];
      sy[m] = a.stry[lj + m - 2];
   }
   const long long pcol = ( ty + 2 ) * CL_PX + tx + 2;
   const double cof = 1.0 / ( a.h * a.h );
   for( int kb = kb_lo; kb <= kb_hi; kb++ )
   {
      const int k = side == 0 ? kb : a.nk - kb + 1;
      double r[3];
      rhs_closure_point( U, smu, sla, pcol, kbase, a.nk, side, kb, (long long)CL_PX, (long long)CL_PLANE, sx, sy, r );
      const long long p = (long long)li + (long long)b.ni * lj + b.nij * ( k - b.kfirst );
      rhs_epilogue<MODE>( a, p, li + b.ifirst, lj + b.jfirst, k, cof, r );
   }
}

// SBP closure rows, second version: the staged planes as above, but the operator is evaluated like the
// interior throughput kernel (rhs4sg_fast2.cu) -- every in-plane piece once per point, the 12 cross terms from
// 6 exchanged products per point, differenced through shared memory -- with the z pieces replaced by the
// one-sided sums of the closure (rhs4sg_rev.C:349-855): G_z -> sum_q [sum_m acof(k,q,m) a(m)] u(q) + ghcof(k) a(1) u(0),
// D0z -> +-sum_q bope(k,q) . (q), no strz, no 1/6.  Zero table entries are skipped.  One thread per column of a
// 32x8 tile, rows kb_lo..kb_hi of one side.  Agrees with rhs_closure_point to rounding (different association).
constexpr int CF_EX = 3 * CL_TY * CL_PX, CF_EY = 3 * CL_PY * CL_TX;
constexpr int CF_OPS = 6 * 4 * CL_TX * CL_TY; // epilogue operands (rho, um or up) of the own points of the 6 rows
constexpr int CF_SMEM_DOUBLES = 5 * CL_NP * CL_PLANE + CF_EX + CF_EY + CL_PX + CL_PY + CF_OPS + 2; // + one mbarrier (TMA staging)

__device__ __forceinline__ void cf_cp_async8( double* sdst, const double* gsrc, bool valid )
{
   const unsigned d = (unsigned)__cvta_generic_to_shared( sdst );
   const int sz = valid ? 8 : 0;
   asm volatile( "cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"( d ), "l"( gsrc ), "r"( sz ) : "memory" );
}

__device__ __forceinline__ double cf_d0u( double fm2, double fm1, double fp1, double fp2 )
{
   return ( fm2 - fp2 ) + 8 * ( fp1 - fm1 ); // 12 * centred first difference
}

// TMA: the 9 halo'd planes of u,v,w,mu,la arrive as ONE 36x12x9 box per field and the epilogue operands as one 32x8x6
// box per array (tma.cuh; even ni, 16-byte aligned arrays), instead of 45 + 24 cp.async per thread with their 64-bit
// address arithmetic -- that staging phase was 46 % of the kernel's time (profiles/r01h_closure_ncu.md).
template <int MODE, bool TMA>
__global__ void __launch_bounds__( CL_TX* CL_TY, 1 ) k_closure_fast( RhsArgs a, int side, int kb_lo, int kb_hi, const __grid_constant__ FastMaps maps )
{
   extern __shared__ __align__( 128 ) double sm_cl[];
   double* const sm = sm_cl;
   constexpr int TX = CL_TX, TY = CL_TY, PX = CL_PX, PY = CL_PY, PLANE = CL_PLANE, NP = CL_NP, NT = CL_TX * CL_TY;
   double* const s_f = sm;				 // [5 fields][9 planes][PLANE]
   double* const s_ex = sm + 5 * NP * PLANE;		 // [3][TY][PX]
   double* const s_ey = s_ex + CF_EX;			 // [3][PY][TX]
   double* const s_sx = s_ey + CF_EY;			 // [PX]
   double* const s_sy = s_sx + PX;			 // [PY]
   double* const s_op = s_sy + PY;			 // [6 rows][4][NT]   (TMA: [4][6 planes][NT])
   double* const s_mbar = s_op + CF_OPS;
   // operand f (0: rho, 1..3: um / up) of closure row kb at the thread's own point
   auto opidx = [&]( int kb, int f ) { return TMA ? ( f * 6 + ( side == 0 ? kb - 1 : 6 - kb ) ) * NT : ( ( kb - 1 ) * 4 + f ) * NT; };
   constexpr bool STAGED_OPS = MODE == MODE_PRED || MODE == MODE_CORR_ACC;
   const Block& b = a.b;
   const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
   const int li0 = 2 + blockIdx.x * TX, lj0 = 2 + blockIdx.y * TY;
   const int kbase = side == 0 ? 0 : a.nk - 7; // global k of staged plane 0
   if( TMA )
   {
      using namespace fast4;
      if( tid == 0 )
      {
	 mbar_init( s_mbar, 1 );
	 asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );
	 mbar_arrive_expect( s_mbar, 5 * NP * PLANE * 8 + ( STAGED_OPS ? CF_OPS * 8 : 0 ) );
	 const int c0 = li0 - 2, c1 = lj0 - 2, c2 = kbase - b.kfirst;
#pragma unroll
	 for( int f = 0; f < 3; f++ ) tma_tile<PX, PY, NP>( s_f + f * NP * PLANE, &maps.u[f], b, c0, c1, c2, s_mbar );
	 tma_tile<PX, PY, NP>( s_f + 3 * NP * PLANE, &maps.mu, b, c0, c1, c2, s_mbar );
	 tma_tile<PX, PY, NP>( s_f + 4 * NP * PLANE, &maps.la, b, c0, c1, c2, s_mbar );
	 if( STAGED_OPS )
	 {
	    const int k0 = ( side == 0 ? 1 : a.nk - 5 ) - b.kfirst; // rows 1..6 / nk-5..nk, in increasing k
	    tma_tile<TX, TY, 6>( s_op, &maps.rho, b, li0, lj0, k0, s_mbar );
#pragma unroll
	    for( int c = 0; c < 3; c++ ) tma_tile<TX, TY, 6>( s_op + ( c + 1 ) * 6 * NT, &maps.um[c], b, li0, lj0, k0, s_mbar );
	 }
      }
   }
   if( !TMA && STAGED_OPS && li0 + tx <= b.nil - 3 && lj0 + ty <= b.nj - 3 )
   {
      // epilogue operands of the thread's own points: in flight together with the planes
      for( int kb = kb_lo; kb <= kb_hi; kb++ )
      {
	 const int k = side == 0 ? kb : a.nk - kb + 1;
	 const long long p = (long long)( li0 + tx ) + (long long)b.ni * ( lj0 + ty ) + b.nij * ( k - b.kfirst );
	 double* const d = s_op + ( kb - 1 ) * 4 * NT + tid;
	 cf_cp_async8( d, a.rho + p, true );
	 const double* const src = MODE == MODE_PRED ? a.um : a.up;
#pragma unroll
	 for( int c = 0; c < 3; c++ ) cf_cp_async8( d + ( c + 1 ) * NT, src + c * b.sc + b.sp * p, true );
      }
   }
   const double dt2i = 1.0 / ( a.dt * a.dt );
   for( int idx = tid; !TMA && idx < PLANE; idx += NT )
   {
      const int sy_ = idx / PX, sx_ = idx - sy_ * PX;
      const int li = li0 - 2 + sx_, lj = lj0 - 2 + sy_;
      const bool inb = li < b.ni && lj < b.nj;
#pragma unroll
      for( int s = 0; s < NP; s++ )
      {
	 const long long p = inb ? (long long)li + (long long)b.ni * lj + b.nij * ( kbase + s - b.kfirst ) : 0;
	 if( MODE == MODE_CORR )
	 {
#pragma unroll
	    for( int c = 0; c < 3; c++ )
	    {
	       const long long q = c * b.sc + b.sp * p;
	       s_f[( c * NP + s ) * PLANE + idx] = inb ? dt2i * ( a.up[q] - 2 * a.u[q] + a.um[q] ) : 0.0;
	    }
	    s_f[( 3 * NP + s ) * PLANE + idx] = inb ? a.mu[p] : 0.0;
	    s_f[( 4 * NP + s ) * PLANE + idx] = inb ? a.la[p] : 0.0;
	 }
	 else
	 {
	    // asynchronous copies: all 45 values of the column are in flight at once (zero fill outside the block)
#pragma unroll
	    for( int c = 0; c < 3; c++ ) cf_cp_async8( s_f + ( c * NP + s ) * PLANE + idx, a.u + c * b.sc + b.sp * p, inb );
	    cf_cp_async8( s_f + ( 3 * NP + s ) * PLANE + idx, a.mu + p, inb );
	    cf_cp_async8( s_f + ( 4 * NP + s ) * PLANE + idx, a.la + p, inb );
	 }
      }
   }
   if( !TMA && MODE != MODE_CORR ) asm volatile( "cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory" );
   for( int t = tid; t < PX + PY; t += NT )
   {
      if( t < PX ) { const int li = li0 - 2 + t; s_sx[t] = li < b.ni ? a.strx[li] : 0.0; }
      else { const int lj = lj0 - 2 + ( t - PX ); s_sy[t - PX] = lj < b.nj ? a.stry[lj] : 0.0; }
   }
   __syncthreads();
   if( TMA ) fast4::mbar_wait( s_mbar, 0 ); // the boxes have landed (the barrier above published the initialised mbarrier)
   // closure row q (1..8; 0 = ghost plane) lives in staged plane q (low side) or 8-q (high side)
   const int pbase = side == 0 ? 0 : 8 * PLANE, pstr = side == 0 ? PLANE : -PLANE;
   const double sgn = side == 0 ? 1.0 : -1.0;
   const int o = ( ty + 2 ) * PX + tx + 2;
   const int li = li0 + tx, lj = lj0 + ty;
   const bool act = li <= b.nil - 3 && lj <= b.nj - 3;
   const double sx = s_sx[tx + 2], sy = s_sy[ty + 2];
   const double sxm2 = s_sx[tx], sxm1 = s_sx[tx + 1], sxp1 = s_sx[tx + 3], sxp2 = s_sx[tx + 4];
   const double sym2 = s_sy[ty], sym1 = s_sy[ty + 1], syp1 = s_sy[ty + 3], syp2 = s_sy[ty + 4];
   auto F = [&]( int f, int q, int off ) { return s_f[f * NP * PLANE + pbase + q * pstr + off]; };

   // column data shared by all rows: coefficients and the z-differenced in-plane products of rows 1..8
   double muq[8], l2q[8], g1[8], g2[8], g3[8];
#pragma unroll
   for( int q = 1; q <= 8; q++ )
   {
      const double m = F( 3, q, o ), l = F( 4, q, o );
      muq[q - 1] = m;
      l2q[q - 1] = 2 * m + l;
      const double dxw = cf_d0u( F( 2, q, o - 2 ), F( 2, q, o - 1 ), F( 2, q, o + 1 ), F( 2, q, o + 2 ) );
      const double dyw = cf_d0u( F( 2, q, o - 2 * PX ), F( 2, q, o - PX ), F( 2, q, o + PX ), F( 2, q, o + 2 * PX ) );
      const double dxu = cf_d0u( F( 0, q, o - 2 ), F( 0, q, o - 1 ), F( 0, q, o + 1 ), F( 0, q, o + 2 ) );
      const double dyv = cf_d0u( F( 1, q, o - 2 * PX ), F( 1, q, o - PX ), F( 1, q, o + PX ), F( 1, q, o + 2 * PX ) );
      g1[q - 1] = m * dxw;
      g2[q - 1] = m * dyw;
      g3[q - 1] = l * ( sx * dxu + sy * dyv );
   }
   const double cof = 1.0 / ( a.h * a.h );
#pragma unroll
   for( int kb = 1; kb <= 6; kb++ )
   {
      if( kb < kb_lo || kb > kb_hi ) continue; // uniform
      // ---- in-plane pieces of row kb
      const double u0 = F( 0, kb, o ), v0 = F( 1, kb, o ), w0 = F( 2, kb, o ), m0 = muq[kb - 1], l0 = F( 4, kb, o );
      const double uxm2 = F( 0, kb, o - 2 ), uxm1 = F( 0, kb, o - 1 ), uxp1 = F( 0, kb, o + 1 ), uxp2 = F( 0, kb, o + 2 );
      const double vxm2 = F( 1, kb, o - 2 ), vxm1 = F( 1, kb, o - 1 ), vxp1 = F( 1, kb, o + 1 ), vxp2 = F( 1, kb, o + 2 );
      const double wxm2 = F( 2, kb, o - 2 ), wxm1 = F( 2, kb, o - 1 ), wxp1 = F( 2, kb, o + 1 ), wxp2 = F( 2, kb, o + 2 );
      const double uym2 = F( 0, kb, o - 2 * PX ), uym1 = F( 0, kb, o - PX ), uyp1 = F( 0, kb, o + PX ), uyp2 = F( 0, kb, o + 2 * PX );
      const double vym2 = F( 1, kb, o - 2 * PX ), vym1 = F( 1, kb, o - PX ), vyp1 = F( 1, kb, o + PX ), vyp2 = F( 1, kb, o + 2 * PX );
      const double wym2 = F( 2, kb, o - 2 * PX ), wym1 = F( 2, kb, o - PX ), wyp1 = F( 2, kb, o + PX ), wyp2 = F( 2, kb, o + 2 * PX );
      const double dxu = cf_d0u( uxm2, uxm1, uxp1, uxp2 ), dxv = cf_d0u( vxm2, vxm1, vxp1, vxp2 );
      const double dyu = cf_d0u( uym2, uym1, uyp1, uyp2 ), dyv = cf_d0u( vym2, vym1, vyp1, vyp2 );
      double pr[3];
      {
	 const double mxm2 = F( 3, kb, o - 2 ), mxm1 = F( 3, kb, o - 1 ), mxp1 = F( 3, kb, o + 1 ), mxp2 = F( 3, kb, o + 2 );
	 const double lxm2 = F( 4, kb, o - 2 ), lxm1 = F( 4, kb, o - 1 ), lxp1 = F( 4, kb, o + 1 ), lxp2 = F( 4, kb, o + 2 );
	 const double mym2 = F( 3, kb, o - 2 * PX ), mym1 = F( 3, kb, o - PX ), myp1 = F( 3, kb, o + PX ), myp2 = F( 3, kb, o + 2 * PX );
	 const double lym2 = F( 4, kb, o - 2 * PX ), lym1 = F( 4, kb, o - PX ), lyp1 = F( 4, kb, o + PX ), lyp2 = F( 4, kb, o + 2 * PX );
	 const double amx[5] = { mxm2 * sxm2, mxm1 * sxm1, m0 * sx, mxp1 * sxp1, mxp2 * sxp2 };
	 const double alx[5] = { ( 2 * mxm2 + lxm2 ) * sxm2, ( 2 * mxm1 + lxm1 ) * sxm1, ( 2 * m0 + l0 ) * sx,
				 ( 2 * mxp1 + lxp1 ) * sxp1, ( 2 * mxp2 + lxp2 ) * sxp2 };
	 const double amy[5] = { mym2 * sym2, mym1 * sym1, m0 * sy, myp1 * syp1, myp2 * syp2 };
	 const double aly[5] = { ( 2 * mym2 + lym2 ) * sym2, ( 2 * mym1 + lym1 ) * sym1, ( 2 * m0 + l0 ) * sy,
				 ( 2 * myp1 + lyp1 ) * syp1, ( 2 * myp2 + lyp2 ) * syp2 };
	 double wmx[4], wlx[4], wmy[4], wly[4];
	 weights4( amx, wmx ); weights4( alx, wlx ); weights4( amy, wmy ); weights4( aly, wly );
	 const double fxu[5] = { uxm2, uxm1, u0, uxp1, uxp2 }, fxv[5] = { vxm2, vxm1, v0, vxp1, vxp2 }, fxw[5] = { wxm2, wxm1, w0, wxp1, wxp2 };
	 const double fyu[5] = { uym2, uym1, u0, uyp1, uyp2 }, fyv[5] = { vym2, vym1, v0, vyp1, vyp2 }, fyw[5] = { wym2, wym1, w0, wyp1, wyp2 };
	 pr[0] = sx * gsum( wlx, fxu ) + sy * gsum( wmy, fyu );
	 pr[1] = sx * gsum( wmx, fxv ) + sy * gsum( wly, fyv );
	 pr[2] = sx * gsum( wmx, fxw ) + sy * gsum( wmy, fyw );
      }
      // ---- one-sided z pieces
      double rz[3] = { 0, 0, 0 }, bz[3] = { 0, 0, 0 }, t[3] = { 0, 0, 0 };
#pragma unroll
      for( int q = 1; q <= 8; q++ )
      {
	 double mc = 0, lc = 0;
	 bool any = false;
#pragma unroll
	 for( int m = 1; m <= 8; m++ )
	 {
	    const double ac = acof_c( ( kb - 1 ) + 6 * ( q - 1 ) + 48 * ( m - 1 ) );
	    if( ac != 0 ) { mc += ac * muq[m - 1]; lc += ac * l2q[m - 1]; any = true; }
	 }
	 const double bq = sgn * bope_c( ( kb - 1 ) + 6 * ( q - 1 ) );
	 if( !any && bq == 0 ) continue; // uniform
	 const double uq = F( 0, q, o ), vq = F( 1, q, o ), wq = F( 2, q, o );
	 rz[0] += mc * uq; rz[1] += mc * vq; rz[2] += lc * wq;
	 if( bq != 0 )
	 {
	    bz[0] += bq * uq; bz[1] += bq * vq; bz[2] += bq * wq;
	    t[0] += bq * g1[q - 1]; t[1] += bq * g2[q - 1]; t[2] += bq * g3[q - 1];
	 }
      }
      {
	 const double gh = ghcof_c( kb - 1 );
	 if( gh != 0 )
	 {
	    rz[0] += gh * muq[0] * F( 0, 0, o );
	    rz[1] += gh * muq[0] * F( 1, 0, o );
	    rz[2] += gh * l2q[0] * F( 2, 0, o );
	 }
      }
      // ---- exchanged products (first differences carry a factor 12)
      const double bw12 = 12 * bz[2];
      {
	 double* const ex = s_ex + ty * PX + tx + 2;
	 double* const ey = s_ey + ( ty + 2 ) * TX + tx;
	 ex[0] = l0 * ( sy * dyv + bw12 );
	 ex[TY * PX] = m0 * dyu;
	 ex[2 * TY * PX] = m0 * ( 12 * bz[0] );
	 ey[0] = m0 * dxv;
	 ey[PY * TX] = l0 * ( sx * dxu + bw12 );
	 ey[2 * PY * TX] = m0 * ( 12 * bz[1] );
      }
      for( int hh = tid; hh < 4 * TY + 4 * TX; hh += NT )
      {
	 int sx_, sy_;
	 const bool xr = hh < 4 * TY;
	 if( xr ) { const int hx = hh & 3, row = hh >> 2; sx_ = hx < 2 ? hx : TX + hx; sy_ = row + 2; }
	 else { const int tt = hh - 4 * TY; const int hy = tt >> 5, col = tt & 31; sy_ = hy < 2 ? hy : TY + hy; sx_ = col + 2; }
	 const int oo = sy_ * PX + sx_;
	 const double hm = F( 3, kb, oo ), hl = F( 4, kb, oo );
	 double hb[3] = { 0, 0, 0 };
#pragma unroll
	 for( int q = 1; q <= 8; q++ )
	 {
	    const double bq = sgn * bope_c( ( kb - 1 ) + 6 * ( q - 1 ) );
	    if( bq != 0 )
	    {
	       hb[0] += bq * F( 0, q, oo ); hb[1] += bq * F( 1, q, oo ); hb[2] += bq * F( 2, q, oo );
	    }
	 }
	 if( xr )
	 {
	    const double hdyv = cf_d0u( F( 1, kb, oo - 2 * PX ), F( 1, kb, oo - PX ), F( 1, kb, oo + PX ), F( 1, kb, oo + 2 * PX ) );
	    const double hdyu = cf_d0u( F( 0, kb, oo - 2 * PX ), F( 0, kb, oo - PX ), F( 0, kb, oo + PX ), F( 0, kb, oo + 2 * PX ) );
	    double* const hx_ = s_ex + ( sy_ - 2 ) * PX + sx_;
	    hx_[0] = hl * ( s_sy[sy_] * hdyv + 12 * hb[2] );
	    hx_[TY * PX] = hm * hdyu;
	    hx_[2 * TY * PX] = hm * ( 12 * hb[0] );
	 }
	 else
	 {
	    const double hdxv = cf_d0u( F( 1, kb, oo - 2 ), F( 1, kb, oo - 1 ), F( 1, kb, oo + 1 ), F( 1, kb, oo + 2 ) );
	    const double hdxu = cf_d0u( F( 0, kb, oo - 2 ), F( 0, kb, oo - 1 ), F( 0, kb, oo + 1 ), F( 0, kb, oo + 2 ) );
	    double* const hy_ = s_ey + sy_ * TX + ( sx_ - 2 );
	    hy_[0] = hm * hdxv;
	    hy_[PY * TX] = hl * ( s_sx[sx_] * hdxu + 12 * hb[2] );
	    hy_[2 * PY * TX] = hm * ( 12 * hb[1] );
	 }
      }
      __syncthreads();
      if( act )
      {
	 const double* const ex = s_ex + ty * PX + tx + 2;
	 const double* const ey = s_ey + ( ty + 2 ) * TX + tx;
	 const double x1 = cf_d0u( ex[-2], ex[-1], ex[1], ex[2] );
	 const double x2 = cf_d0u( ex[TY * PX - 2], ex[TY * PX - 1], ex[TY * PX + 1], ex[TY * PX + 2] );
	 const double x3 = cf_d0u( ex[2 * TY * PX - 2], ex[2 * TY * PX - 1], ex[2 * TY * PX + 1], ex[2 * TY * PX + 2] );
	 const double y1 = cf_d0u( ey[-2 * TX], ey[-TX], ey[TX], ey[2 * TX] );
	 const double y2 = cf_d0u( ey[PY * TX - 2 * TX], ey[PY * TX - TX], ey[PY * TX + TX], ey[PY * TX + 2 * TX] );
	 const double y3 = cf_d0u( ey[2 * PY * TX - 2 * TX], ey[2 * PY * TX - TX], ey[2 * PY * TX + TX], ey[2 * PY * TX + 2 * TX] );
	 const double i6 = 1.0 / 6, i144 = 1.0 / 144;
	 double r[3];
	 r[0] = ( i6 * pr[0] + rz[0] ) + i144 * ( sx * ( x1 + sy * y1 + 12 * t[0] ) );
	 r[1] = ( i6 * pr[1] + rz[1] ) + i144 * ( sy * ( sx * x2 + y2 + 12 * t[1] ) );
	 r[2] = ( i6 * pr[2] + rz[2] ) + i144 * ( sx * x3 + sy * y3 + 12 * t[2] );
	 const int k = side == 0 ? kb : a.nk - kb + 1;
	 const long long p = (long long)li + (long long)b.ni * lj + b.nij * ( k - b.kfirst );
	 if( STAGED_OPS )
	 {
	    // rhs_epilogue<MODE> with rho, um / up taken from the staged operands and u from the staged plane
	    const double* const d = s_op + tid;
	    const double rho = d[opidx( kb, 0 )];
	    const double uk[3] = { u0, v0, w0 };
	    const double dt2 = a.dt * a.dt;
	    const double f = MODE == MODE_PRED ? dt2 / rho : ( dt2 * dt2 / 12 ) / rho;
#pragma unroll
	    for( int c = 0; c < 3; c++ )
	    {
	       const long long q = c * b.sc + b.sp * p;
	       const double fo = a.fo ? a.fo[q] : 0.0;
	       const double acc = cof * r[c] + fo;
	       if( MODE == MODE_PRED )
	       {
		  a.out[q] = 2 * uk[c] - d[opidx( kb, c + 1 )] + f * acc;
		  if( a.out2 ) a.out2[q] = acc / rho;
	       }
	       else
		  a.out[q] = d[opidx( kb, c + 1 )] + f * acc;
	    }
	 }
	 else
	    rhs_epilogue<MODE>( a, p, li + b.ifirst, lj + b.jfirst, k, cof, r );
      }
      __syncthreads(); // the exchange buffers are reused by the next row
   }
}

// the 2-point shell where L(u) is never written (stays 0 in the reference): pred/corr with lu=0.
// Enumerates the shell as 6 slabs: k-low, k-high (full planes), j-low, j-high, i-low, i-high.
template <int MODE>
__global__ void k_shell_update( RhsArgs a )
{
   const Block& b = a.b;
   const int nil = b.nil;					   // points per row (the row pitch b.ni may be one larger)
   const long long nijl = (long long)nil * b.nj;
   const long long nkplane = 2 * nijl;				   // per k side
   const long long njslab = 2LL * nil * ( b.nk - 4 );		   // per j side
   const long long nislab = 2LL * ( b.nj - 4 ) * ( b.nk - 4 );	   // per i side
   const long long total = 2 * ( nkplane + njslab + nislab );
   for( long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
	t += (long long)gridDim.x * blockDim.x )
   {
      long long s = t;
      int i, j, k;
      if( s < 2 * nkplane )
      {
	 const int side = s >= nkplane;
	 s -= side * nkplane;
	 k = (int)( s / nijl );
	 const long long rr = s % nijl;
	 j = (int)( rr / nil );
	 i = (int)( rr % nil );
	 if( side ) k += b.nk - 2;
      }
      else if( ( s -= 2 * nkplane ) < 2 * njslab )
      {
	 const int side = s >= njslab;
	 s -= side * njslab;
	 const long long per_k = 2LL * nil;
	 k = 2 + (int)( s / per_k );
	 const long long rr = s % per_k;
	 j = (int)( rr / nil );
	 i = (int)( rr % nil );
	 if( side ) j += b.nj - 2;
      }
      else
      {
	 s -= 2 * njslab;
	 const int side = s >= nislab;
	 s -= side * nislab;
	 const long long per_k = 2LL * ( b.nj - 4 );
	 k = 2 + (int)( s / per_k );
	 const long long rr = s % per_k;
	 j = 2 + (int)( rr / 2 );
	 i = (int)( rr % 2 );
	 if( side ) i += nil - 2;
      }
      const long long p = (long long)i + (long long)b.ni * j + b.nij * k;
      if( MODE == MODE_SHELL_DPDMT )
      {
	 const bool inner = i >= 2 && i < nil - 2 && j >= 2 && j < b.nj - 2;
	 if( inner && ( ( a.halo_lo && k < 2 ) || ( a.halo_hi && k >= b.nk - 2 ) ) ) continue;
	 const double dt2i = 1.0 / ( a.dt * a.dt );
#pragma unroll
	 for( int c = 0; c < 3; c++ )
	 {
	    const long long q = c * b.sc + b.sp * p;
	    a.out2[q] = dt2i * ( a.up[q] - 2 * a.u[q] + a.um[q] );
	 }
      }
      else if( MODE == MODE_PRED )
      {
	 const double f = ( a.dt * a.dt ) / a.rho[p];
#pragma unroll
	 for( int c = 0; c < 3; c++ )
	 {
	    const long long q = c * b.sc + b.sp * p;
	    const double fo = a.fo ? a.fo[q] : 0.0;
	    a.out[q] = 2 * a.u[q] - a.um[q] + f * ( 0.0 + fo );
	 }
      }
      else
      {
	 const double dt2 = a.dt * a.dt;
	 const double f = ( dt2 * dt2 / 12 ) / a.rho[p];
#pragma unroll
	 for( int c = 0; c < 3; c++ )
	 {
	    const long long q = c * b.sc + b.sp * p;
	    const double fo = a.fo ? a.fo[q] : 0.0;
	    a.out[q] = a.up[q] + f * ( 0.0 + fo );
	 }
      }
   }
}

__global__ void k_predfort( Block b, double* __restrict__ up, const double* __restrict__ u,
			    const double* __restrict__ um, const double* __restrict__ lu,
			    const double* __restrict__ fo, const double* __restrict__ rho, double dt2 )
{
   for( long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < b.npts;
	p += (long long)gridDim.x * blockDim.x )
   {
      const double f = dt2 / rho[p];
#pragma unroll
      for( int c = 0; c < 3; c++ )
      {
	 const long long q = c * b.sc + b.sp * p;
	 up[q] = 2 * u[q] - um[q] + f * ( lu[q] + ( fo ? fo[q] : 0.0 ) );
      }
   }
}

__global__ void k_corrfort( Block b, double* __restrict__ up, const double* __restrict__ lu,
			    const double* __restrict__ fo, const double* __restrict__ rho, double dt4i12 )
{
   for( long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < b.npts;
	p += (long long)gridDim.x * blockDim.x )
   {
      const double f = dt4i12 / rho[p];
#pragma unroll
      for( int c = 0; c < 3; c++ )
      {
	 const long long q = c * b.sc + b.sp * p;
	 up[q] += f * ( lu[q] + ( fo ? fo[q] : 0.0 ) );
      }
   }
}

__global__ void k_dpdmt( long long n, const double* __restrict__ up, const double* __restrict__ u,
			 const double* __restrict__ um, double* __restrict__ u2, double dt2i )
{
   for( long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n;
	p += (long long)gridDim.x * blockDim.x )
      u2[p] = dt2i * ( up[p] - 2 * u[p] + um[p] );
}

__global__ void k_addsgd( int order, Block b, Int6 box, double* __restrict__ up, const double* __restrict__ u,
			  const double* __restrict__ um, const double* __restrict__ rho,
			  const double* __restrict__ dcx, const double* __restrict__ dcy,
			  const double* __restrict__ dcz, const double* __restrict__ strx,
			  const double* __restrict__ stry, const double* __restrict__ strz,
			  const double* __restrict__ cox, const double* __restrict__ coy,
			  const double* __restrict__ coz, double beta )
{
   // box = local (array) index ranges, inclusive
   const int ii = box.v[0] + blockIdx.x * blockDim.x + threadIdx.x;
   const int jj = box.v[2] + blockIdx.y * blockDim.y + threadIdx.y;
   const int kk = box.v[4] + blockIdx.z * blockDim.z + threadIdx.z;
   if( ii > box.v[1] || jj > box.v[3] || kk > box.v[5] ) return;
   const long long p = (long long)ii + (long long)b.ni * jj + b.nij * kk;
   const double prex = strx[ii] * coy[jj] * coz[kk];
   const double prey = stry[jj] * cox[ii] * coz[kk];
   const double prez = strz[kk] * cox[ii] * coy[jj];
   const double birho = beta / rho[p];
#pragma unroll
   for( int c = 0; c < 3; c++ )
   {
      const long long q = c * b.sc + b.sp * p;
      const double s = sgd_point( order, u, um, rho, q, p, b.sp, 1LL, dcx + ii, prex ) +
		       sgd_point( order, u, um, rho, q, p, b.sp, (long long)b.ni, dcy + jj, prey ) +
		       sgd_point( order, u, um, rho, q, p, b.sp, b.nij, dcz + kk, prez );
      up[q] -= birho * s;
   }
}

// Dirichlet / supergrid ghost fill and periodic copy for one side window
__global__ void k_bc_window( Block b, int s, int i0, int i1, int j0, int j1, int k0, int k1, int type,
			     long long off, double* __restrict__ u, const double* __restrict__ bf )
{
   const long long wi = i1 - i0 + 1, wj = j1 - j0 + 1, wk = k1 - k0 + 1;
   const long long total = wi * wj * wk;
   for( long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
	t += (long long)gridDim.x * blockDim.x )
   {
      const int i = i0 + (int)( t % wi );
      const int j = j0 + (int)( ( t / wi ) % wj );
      const int k = k0 + (int)( t / ( wi * wj ) );
      const long long p = pidx( b, i, j, k );
#pragma unroll
      for( int c = 0; c < 3; c++ )
      {
	 const long long q = c * b.sc + b.sp * p;
	 if( type == 3 )
	    u[q] = u[q + b.sp * off];
	 else
	    u[q] = bf[3 * t + c];
      }
   }
}

// stress-free ghost plane of side 4 (k=1, kl=1) or 5 (k=nz, kl=-1)
__global__ void k_bc_freesurface( Block b, int k, int kl, double h, double* __restrict__ u,
				  const double* __restrict__ mu, const double* __restrict__ la,
				  const double* __restrict__ bf, const double* __restrict__ strx,
				  const double* __restrict__ stry )
{
   const int ii = 2 + blockIdx.x * blockDim.x + threadIdx.x;
   const int jj = 2 + blockIdx.y * blockDim.y + threadIdx.y;
   if( ii > b.nil - 3 || jj > b.nj - 3 ) return;
   const double d4a = 2.0 / 3.0, d4b = -1.0 / 12.0;
   const long long qq = (long long)ii + (long long)b.ni * jj;
   const long long p = qq + b.nij * ( k - b.kfirst );
   const long long sp = b.sp, sc = b.sc, dj = b.ni;
   auto U = [&]( int c, long long pp ) { return u[c * sc + sp * pp]; };
   const double sx = strx[ii], sy = stry[jj];
   const double wx = sx * ( d4a * ( U( 2, p + 1 ) - U( 2, p - 1 ) ) + d4b * ( U( 2, p + 2 ) - U( 2, p - 2 ) ) );
   const double ux = sx * ( d4a * ( U( 0, p + 1 ) - U( 0, p - 1 ) ) + d4b * ( U( 0, p + 2 ) - U( 0, p - 2 ) ) );
   const double wy = sy * ( d4a * ( U( 2, p + dj ) - U( 2, p - dj ) ) + d4b * ( U( 2, p + 2 * dj ) - U( 2, p - 2 * dj ) ) );
   const double vy = sy * ( d4a * ( U( 1, p + dj ) - U( 1, p - dj ) ) + d4b * ( U( 1, p + 2 * dj ) - U( 1, p - 2 * dj ) ) );
   double uz = 0, vz = 0, wz = 0;
#pragma unroll
   for( int q = 1; q <= 4; q++ )
   {
      const long long pq = p + b.nij * ( kl * ( q - 1 ) );
      uz += c_sbop[q] * U( 0, pq );
      vz += c_sbop[q] * U( 1, pq );
      wz += c_sbop[q] * U( 2, pq );
   }
   const long long pg = p - b.nij * kl;
   const double m = mu[p], l = la[p];
   const long long qb = (long long)ii + (long long)b.nil * jj; // the forcing array is indexed by window point (bcfortsg: qq)
   const double b0 = bf ? bf[3 * qb] : 0.0, b1 = bf ? bf[3 * qb + 1] : 0.0, b2 = bf ? bf[3 * qb + 2] : 0.0;
   u[0 * sc + sp * pg] = ( -uz - kl * wx + kl * h * b0 / m ) / c_sbop[0];
   u[1 * sc + sp * pg] = ( -vz - kl * wy + kl * h * b1 / m ) / c_sbop[0];
   u[2 * sc + sp * pg] = ( -wz + ( -kl * l * ( ux + vy ) + kl * h * b2 ) / ( 2 * m + l ) ) / c_sbop[0];
}

__global__ void k_add_point_forces( long long sc, long long sp, double* __restrict__ up,
				    const double* __restrict__ rho, int n, const long long* __restrict__ pidx_,
				    const double* __restrict__ f, double factor, double* __restrict__ up2, double factor2,
				    long long nij, int kplane_lo, int kplane_hi )
{
   const int t = blockIdx.x * blockDim.x + threadIdx.x;
   if( t >= n ) return;
   const long long p = pidx_[t];
   if( nij > 0 )
   {
      const long long kp = p / nij;
      if( kp < kplane_lo || kp > kplane_hi ) return;
   }
   const double s = factor / rho[p];
#pragma unroll
   for( int c = 0; c < 3; c++ ) up[c * sc + sp * p] += s * f[3 * t + c];
   if( up2 )
   {
      const double s2 = factor2 / rho[p];
#pragma unroll
      for( int c = 0; c < 3; c++ ) up2[c * sc + sp * p] += s2 * f[3 * t + c];
   }
}

__global__ void k_gather_points( long long sc, long long sp, const double* __restrict__ u, int n,
				 const long long* __restrict__ pidx_, double* __restrict__ out )
{
   const int t = blockIdx.x * blockDim.x + threadIdx.x;
   if( t >= n ) return;
   const long long p = pidx_[t];
#pragma unroll
   for( int c = 0; c < 3; c++ ) out[3 * t + c] = u[c * sc + sp * p];
}

// copy two k-planes (kplane, kplane+1 as local plane offsets) of a 3-component field to/from a
// buffer laid out [c][2][nj][ni]
__global__ void k_halo_copy( Block b, double* __restrict__ field, int kplane, double* __restrict__ buf, int pack )
{
   const long long n2 = 2 * b.nij;
   for( long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < 3 * n2;
	t += (long long)gridDim.x * blockDim.x )
   {
      const int c = (int)( t / n2 );
      const long long r = t % n2;
      const long long p = b.nij * kplane + r;
      const long long q = c * b.sc + b.sp * p;
      if( pack ) buf[t] = field[q];
      else field[q] = buf[t];
   }
}

__global__ void k_fill_profile( Block b, double* __restrict__ a, const double* __restrict__ prof )
{
   for( long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < b.npts;
	p += (long long)gridDim.x * blockDim.x )
      a[p] = prof[p / b.nij];
}

inline int nblocks( long long n, int bs, int cap = 148 * 16 )
{
   long long g = ( n + bs - 1 ) / bs;
   if( g > cap ) g = cap;
   if( g < 1 ) g = 1;
   return (int)g;
}

} // namespace

static int launch_rows_general( RhsMode mode, const RhsArgs& a, int k_lo, int k_hi, cudaStream_t st )
{
   const Block& b = a.b;
   if( k_hi < k_lo ) return 0;
   ProfScope prof( "rhs_v1", st );
   dim3 bs( 32, 4, 2 );
   dim3 gs( ( b.nil - 4 + bs.x - 1 ) / bs.x, ( b.nj - 4 + bs.y - 1 ) / bs.y, ( k_hi - k_lo + 1 + bs.z - 1 ) / bs.z );
   if( mode == MODE_LU ) k_rhs_v1<MODE_LU><<<gs, bs, 0, st>>>( a, k_lo, k_hi );
   else if( mode == MODE_PRED ) k_rhs_v1<MODE_PRED><<<gs, bs, 0, st>>>( a, k_lo, k_hi );
   else if( mode == MODE_CORR ) k_rhs_v1<MODE_CORR><<<gs, bs, 0, st>>>( a, k_lo, k_hi );
   else k_rhs_v1<MODE_CORR_ACC><<<gs, bs, 0, st>>>( a, k_lo, k_hi );
   count_launch();
   return check_launch( "k_rhs_v1" );
}

// set by api.cu: the runtime tables (sw4b200_copy_stencilcoefficients) equal the built-in ones
bool g_builtin_sbp_tables = true;

static int closure_generation()
{
   if( !g_builtin_sbp_tables ) return 1; // k_closure_fast has the built-in tables folded in at compile time
   static int v = -1;
   if( v < 0 )
   {
      const char* e = getenv( "SW4B200_CLOSURE_GEN" );
      v = ( e && e[0] == '1' ) ? 1 : 2;
   }
   return v;
}

template <int MODE, bool TMA>
static int launch_closure_fast_tt( const RhsArgs& a, int side, int kb_lo, int kb_hi, const FastMaps& maps, cudaStream_t st )
{
   static bool configured = false;
   const size_t smem = (size_t)CF_SMEM_DOUBLES * sizeof( double );
   if( !configured )
   {
      cudaError_t e = cudaFuncSetAttribute( k_closure_fast<MODE, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem );
      if( e != cudaSuccess ) return set_error( "k_closure_fast: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString( e ) );
      configured = true;
   }
   const Block& b = a.b;
   ProfScope prof( "closure", st );
   ProfScope prof2( TMA ? "closure_tma" : "closure_cpasync", st ); // (which staging ran: the parity tests assert it)
   dim3 bs( CL_TX, CL_TY, 1 );
   dim3 gs( ( b.nil - 4 + CL_TX - 1 ) / CL_TX, ( b.nj - 4 + CL_TY - 1 ) / CL_TY, 1 );
   k_closure_fast<MODE, TMA><<<gs, bs, smem, st>>>( a, side, kb_lo, kb_hi, maps );
   count_launch();
   return check_launch( "k_closure_fast" );
}

template <int MODE>
static int launch_closure_fast_t( const RhsArgs& a, int side, int kb_lo, int kb_hi, cudaStream_t st )
{
   const Block& b = a.b;
   FastMaps maps;
   memset( &maps, 0, sizeof( maps ) );
   // TMA staging needs tensor maps: SoA layout, even ni (row pitch a multiple of 16 bytes), 16-byte aligned arrays, arrays at
   // least one box wide; MODE_CORR forms its input on the fly and keeps the direct loads
   constexpr bool staged_ops = MODE == MODE_PRED || MODE == MODE_CORR_ACC;
   const double* const ops = MODE == MODE_PRED ? a.um : a.up;
   uintptr_t al = (uintptr_t)a.u | (uintptr_t)a.mu | (uintptr_t)a.la;
   if( staged_ops ) al |= (uintptr_t)a.rho | (uintptr_t)ops;
   const bool tma = MODE != MODE_CORR && b.sp == 1 && !( b.ni & 1 ) && !( b.npts & 1 ) && !( al & 15 ) && b.ni >= CL_PX && b.nj >= CL_PY &&
		    b.nk >= CL_NP;
   if( !tma ) return launch_closure_fast_tt<MODE, false>( a, side, kb_lo, kb_hi, maps, st );
   for( int c = 0; c < 3; c++ )
      if( make_tmap( &maps.u[c], a.u + c * b.sc, b, CL_PX, CL_PY, CL_NP ) ) return 1;
   if( make_tmap( &maps.mu, a.mu, b, CL_PX, CL_PY, CL_NP ) || make_tmap( &maps.la, a.la, b, CL_PX, CL_PY, CL_NP ) ) return 1;
   if( staged_ops )
   {
      if( make_tmap( &maps.rho, a.rho, b, CL_TX, CL_TY, 6 ) ) return 1;
      for( int c = 0; c < 3; c++ )
	 if( make_tmap( &maps.um[c], ops + c * b.sc, b, CL_TX, CL_TY, 6 ) ) return 1;
   }
   return launch_closure_fast_tt<MODE, true>( a, side, kb_lo, kb_hi, maps, st );
}

template <int MODE>
static int launch_closure_t( const RhsArgs& a, int side, int kb_lo, int kb_hi, cudaStream_t st )
{
   if( closure_generation() == 2 ) return launch_closure_fast_t<MODE>( a, side, kb_lo, kb_hi, st );
   static bool configured = false;
   const size_t smem = (size_t)5 * CL_NP * CL_PLANE * sizeof( double );
   if( !configured )
   {
      cudaError_t e = cudaFuncSetAttribute( k_closure_staged<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem );
      if( e != cudaSuccess ) return set_error( "k_closure_staged: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString( e ) );
      configured = true;
   }
   const Block& b = a.b;
   ProfScope prof( "closure", st );
   dim3 bs( CL_TX, CL_TY, 1 );
   dim3 gs( ( b.nil - 4 + CL_TX - 1 ) / CL_TX, ( b.nj - 4 + CL_TY - 1 ) / CL_TY, 1 );
   k_closure_staged<MODE><<<gs, bs, smem, st>>>( a, side, kb_lo, kb_hi );
   count_launch();
   return check_launch( "k_closure_staged" );
}

static int launch_closure( RhsMode mode, const RhsArgs& a, int side, int kb_lo, int kb_hi, cudaStream_t st )
{
   if( mode == MODE_LU ) return launch_closure_t<MODE_LU>( a, side, kb_lo, kb_hi, st );
   if( mode == MODE_PRED ) return launch_closure_t<MODE_PRED>( a, side, kb_lo, kb_hi, st );
   if( mode == MODE_CORR ) return launch_closure_t<MODE_CORR>( a, side, kb_lo, kb_hi, st );
   return launch_closure_t<MODE_CORR_ACC>( a, side, kb_lo, kb_hi, st );
}

// rows [k_lo,k_hi] of the block (interior and/or closure rows): the SBP closure rows of a side go to the
// staged kernel when the 9 planes they read lie inside the block, everything else to the general kernel
int launch_rhs_v1_rows( RhsMode mode, const RhsArgs& a, int k_lo, int k_hi, cudaStream_t st )
{
   const Block& b = a.b;
   if( k_hi < k_lo || b.nil < 5 || b.nj < 5 ) return 0;
   int lo = k_lo, hi = k_hi;
   if( a.onesided4 && lo <= 6 && b.kfirst <= 0 && b.klast >= 8 && a.nk >= 12 )
   {
      const int c1 = hi < 6 ? hi : 6;
      if( launch_closure( mode, a, 0, lo < 1 ? 1 : lo, c1, st ) ) return 1;
      lo = c1 + 1;
   }
   if( a.onesided5 && hi >= a.nk - 5 && lo <= hi && b.klast >= a.nk + 1 && b.kfirst <= a.nk - 7 && a.nk >= 12 )
   {
      const int c0 = lo > a.nk - 5 ? lo : a.nk - 5;
      // rows k=c0..hi  <->  kb = nk-k+1
      if( launch_closure( mode, a, 1, a.nk - hi + 1, a.nk - c0 + 1, st ) ) return 1;
      hi = c0 - 1;
   }
   return launch_rows_general( mode, a, lo, hi, st );
}

int launch_rhs_v1( RhsMode mode, const RhsArgs& a, cudaStream_t st )
{
   return launch_rhs_v1_rows( mode, a, a.b.kfirst + 2, a.b.klast - 2, st );
}

int launch_shell_update( RhsMode mode, const RhsArgs& a, cudaStream_t st )
{
   const Block& b = a.b;
   const long long total = 2 * ( 2LL * b.nil * b.nj + 2LL * b.nil * ( b.nk - 4 ) + 2LL * ( b.nj - 4 ) * ( b.nk - 4 ) );
   ProfScope prof( "shell", st );
   if( mode == MODE_PRED ) k_shell_update<MODE_PRED><<<nblocks( total, 256 ), 256, 0, st>>>( a );
   else if( mode == MODE_SHELL_DPDMT ) k_shell_update<MODE_SHELL_DPDMT><<<nblocks( total, 256 ), 256, 0, st>>>( a );
   else k_shell_update<MODE_CORR><<<nblocks( total, 256 ), 256, 0, st>>>( a );
   count_launch();
   return check_launch( "k_shell_update" );
}

int launch_predfort( const Block& b, double* up, const double* u, const double* um, const double* lu,
		     const double* fo, const double* rho, double dt2, cudaStream_t st )
{
   k_predfort<<<nblocks( b.npts, 256 ), 256, 0, st>>>( b, up, u, um, lu, fo, rho, dt2 );
   count_launch();
   return check_launch( "k_predfort" );
}
int launch_corrfort( const Block& b, double* up, const double* lu, const double* fo, const double* rho,
		     double dt4, cudaStream_t st )
{
   k_corrfort<<<nblocks( b.npts, 256 ), 256, 0, st>>>( b, up, lu, fo, rho, dt4 / 12 );
   count_launch();
   return check_launch( "k_corrfort" );
}
int launch_dpdmt( long long n, const double* up, const double* u, const double* um, double* u2, double dt2i,
		  cudaStream_t st )
{
   k_dpdmt<<<nblocks( n, 256 ), 256, 0, st>>>( n, up, u, um, u2, dt2i );
   count_launch();
   return check_launch( "k_dpdmt" );
}
int launch_addsgd( int order, const Block& b, double* up, const double* u, const double* um, const double* rho,
		   const double* dcx, const double* dcy, const double* dcz, const double* strx,
		   const double* stry, const double* strz, const double* cox, const double* coy,
		   const double* coz, double beta, cudaStream_t st )
{
   if( beta == 0 ) return 0;
   const int w = order == 6 ? 3 : 2;
   if( b.nil <= 2 * w || b.nj <= 2 * w || b.nk <= 2 * w ) return 0;
   Int6 box = { { w, b.nil - 1 - w, w, b.nj - 1 - w, w, b.nk - 1 - w } };
   return launch_addsgd_box( order, b, box, up, u, um, rho, dcx, dcy, dcz, strx, stry, strz, cox, coy, coz, beta, st );
}

// the same update restricted to a sub-box (local index ranges, inclusive) of the update region
int launch_addsgd_box( int order, const Block& b, const Int6& box, double* up, const double* u, const double* um,
		       const double* rho, const double* dcx, const double* dcy, const double* dcz, const double* strx,
		       const double* stry, const double* strz, const double* cox, const double* coy,
		       const double* coz, double beta, cudaStream_t st )
{
   const int nx = box.v[1] - box.v[0] + 1, ny = box.v[3] - box.v[2] + 1, nz = box.v[5] - box.v[4] + 1;
   if( beta == 0 || nx <= 0 || ny <= 0 || nz <= 0 ) return 0;
   if( order == 4 && b.sp == 1 )
      return launch_addsgd4_fast( b, box, up, u, um, rho, dcx, dcy, dcz, strx, stry, strz, cox, coy, coz, beta, st );
   ProfScope prof( "addsgd", st );
   dim3 bs( 32, 4, 2 );
   dim3 gs( ( nx + bs.x - 1 ) / bs.x, ( ny + bs.y - 1 ) / bs.y, ( nz + bs.z - 1 ) / bs.z );
   k_addsgd<<<gs, bs, 0, st>>>( order, b, box, up, u, um, rho, dcx, dcy, dcz, strx, stry, strz, cox, coy, coz, beta );
   count_launch();
   return check_launch( "k_addsgd" );
}

int launch_bcfortsg( const Block& b, const Int36& wind, int nx, int ny, int nz, double* u, double h,
		     const Int6& bccnd, const double* mu, const double* la, const Ptr6& bforce,
		     const double* strx, const double* stry, cudaStream_t st )
{
   ProfScope prof( "bc", st );
   for( int s = 0; s < 6; s++ )
   {
      const int* w = wind.v + 6 * s;
      const int type = bccnd.v[s];
      if( type == 1 || type == 2 || type == 3 )
      {
	 const long long total = (long long)( w[1] - w[0] + 1 ) * ( w[3] - w[2] + 1 ) * ( w[5] - w[4] + 1 );
	 if( total <= 0 ) continue;
	 if( type != 3 && bforce.p[s] == 0 )
	    return set_error( "bcfortsg: side %d needs a forcing array", s );
	 long long off = 0;
	 if( type == 3 )
	    off = s == 0 ? nx : ( s == 1 ? -nx : ( s == 2 ? (long long)b.ni * ny : ( s == 3 ? -(long long)b.ni * ny
				 : ( s == 4 ? b.nij * nz : -b.nij * nz ) ) ) );
	 k_bc_window<<<nblocks( total, 256 ), 256, 0, st>>>( b, s, w[0], w[1], w[2], w[3], w[4], w[5], type, off, u,
							      bforce.p[s] );
	 count_launch();
      }
      else if( type == 0 )
      {
	 if( s != 4 && s != 5 )
	    return set_error( "bcfortsg: free surface condition not implemented for side %d", s );
	 if( b.nil < 5 || b.nj < 5 ) continue;
	 dim3 bs( 32, 8 );
	 dim3 gs( ( b.nil - 4 + 31 ) / 32, ( b.nj - 4 + 7 ) / 8 );
	 k_bc_freesurface<<<gs, bs, 0, st>>>( b, s == 4 ? 1 : nz, s == 4 ? 1 : -1, h, u, mu, la, bforce.p[s], strx, stry );
	 count_launch();
      }
   }
   return check_launch( "bcfortsg" );
}

int launch_add_point_forces( int corder, long long npts, double* up, const double* rho, int n,
			     const long long* pidx_, const double* f, double factor, cudaStream_t st, double* up2,
			     double factor2, long long nij, int kplane_lo, int kplane_hi )
{
   if( n <= 0 ) return 0;
   k_add_point_forces<<<( n + 127 ) / 128, 128, 0, st>>>( corder ? npts : 1, corder ? 1 : 3, up, rho, n, pidx_, f, factor,
							   up2, factor2, nij, kplane_lo, kplane_hi );
   count_launch();
   return check_launch( "k_add_point_forces" );
}
int launch_gather_points( int corder, long long npts, const double* u, int n, const long long* pidx_,
			  double* out, cudaStream_t st )
{
   if( n <= 0 ) return 0;
   k_gather_points<<<( n + 127 ) / 128, 128, 0, st>>>( corder ? npts : 1, corder ? 1 : 3, u, n, pidx_, out );
   count_launch();
   return check_launch( "k_gather_points" );
}
// derived, time-invariant coefficient arrays of a grid block (read by the fused passes of rhs4sg_fast4.cu instead of lambda, rho):
// 2 mu + lambda in ONE rounding (2 mu is exact; the same value every kernel forms with fma( 2, mu, lambda )) and 1 / rho (IEEE division)
__global__ void k_derive_materials( long long n, const double* __restrict__ mu, const double* __restrict__ la, const double* __restrict__ rho,
				    double* __restrict__ la2, double* __restrict__ rhoi )
{
   for( long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x )
   {
      la2[t] = fma( 2.0, mu[t], la[t] );
      rhoi[t] = 1.0 / rho[t];
   }
}
int launch_derive_materials( long long n, const double* mu, const double* la, const double* rho, double* la2, double* rhoi, cudaStream_t st )
{
   k_derive_materials<<<nblocks( n, 256 ), 256, 0, st>>>( n, mu, la, rho, la2, rhoi );
   count_launch();
   return check_launch( "k_derive_materials" );
}
int launch_fill_profile( const Block& b, double* a, const double* prof, cudaStream_t st )
{
   k_fill_profile<<<nblocks( b.npts, 256 ), 256, 0, st>>>( b, a, prof );
   count_launch();
   return check_launch( "k_fill_profile" );
}
int launch_halo_copy( const Block& b, double* field, int kplane, double* buf, int pack, cudaStream_t st )
{
   k_halo_copy<<<nblocks( 6 * b.nij, 256 ), 256, 0, st>>>( b, field, kplane, buf, pack );
   count_launch();
   return check_launch( "k_halo_copy" );
}

} // namespace sw4b200
