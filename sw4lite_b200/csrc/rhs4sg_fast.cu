// Throughput path for the Cartesian interior rows (SoA layout, corder=1): 2.5-D z-marching
// kernel with cp.async-staged halo'd planes in a shared-memory ring, register-rotated z columns
// and a fused epilogue (lu / predictor / corrector).
//
// Operator: interior formulas of rhs4sg_rev.C:112-348 (reference), evaluated with common
// subexpressions shared between threads instead of recomputed per point:
//   h^2 L_c = sx Gx(a)U_c + sy Gy(a)U_c + sz Gz(a)U_c + mixed terms
// * every in-plane (x,y) piece of plane p is computed ONCE, when plane p arrives in shared memory
//   (second-difference sums Gx,Gy -> a 3-deep register pipeline; first differences D0x,D0y ->
//   z rings for the D0z( coef * D0x/y . ) terms and a 3-deep delay for the cross terms),
// * the z pieces come from per-thread register columns (5 planes of u,v,w, mu*sz, (2mu+la)*sz),
// * the 12 cross-derivative terms are built from 6 exchanged products per point
//     E1 = la (sy D0y v + sz D0z w)   E2 = mu D0y u   E3 = mu D0z u     (differenced in x)
//     E4 = mu D0x v   E5 = la (sx D0x u + sz D0z w)   E6 = mu D0z v     (differenced in y)
//   written to shared memory (the 2-wide ring around the tile is filled by helper threads from the
//   staged planes) plus 3 z rings g1 = mu D0x w, g2 = mu D0y w, g3 = la (sx D0x u + sy D0y v).
// The result agrees with the reference expression to rounding (different association); the parity
// tests bound the difference by 1e-12 relative.
//
// This file is also compiled by g++ as plain C++ (SW4B200_EMULATE) by tests/emu: the CPU test of
// the kernel's index logic and algebra.  That build is test infrastructure, never a product path.
#include "common.cuh"
#include <cstdlib>

#ifndef SW4B200_DEFAULT_FAST_GEN
#define SW4B200_DEFAULT_FAST_GEN 4
#endif

namespace sw4b200 {

enum FastEpilogue { EPI_LU = 0, EPI_PRED = 1, EPI_CORR = 2 };

struct FastArgs
{
   Block b;
   int klo, khi;     // output planes of the whole launch (global k), inclusive
   int kchunk;	     // planes per CTA in z
   const double* u[3];	 // input field (u for LU/PRED, uacc for CORR), halo'd reads
   const double *mu, *la;
   const double *strx, *stry, *strz;
   double cof6, cof144; // 1/(6 h^2), 1/(144 h^2)
   // epilogue
   double* out[3];	// LU: lu ; PRED: up ; CORR: up_out
   double* out2[3];	// PRED: uacc = (L(u)/h^2+fo)/rho (may be null)
   const double* um[3]; // PRED: um ; CORR: up (value to be corrected)
   const double* rho;
   const double* fo[3]; // dense forcing (may be null)
   double fac;		// PRED: dt^2 ; CORR: dt^4/12
};

#ifdef SW4B200_EMULATE
#define SW4_DYN_SMEM( name ) double* name = emu_shared_memory()
#else
#define SW4_DYN_SMEM( name ) extern __shared__ double name[]
#endif

namespace fast {

__device__ __forceinline__ void cp_async8( double* sdst, const double* gsrc, bool valid )
{
#if defined( __CUDA_ARCH__ )
   const unsigned d = (unsigned)__cvta_generic_to_shared( sdst );
   const int sz = valid ? 8 : 0;
   asm volatile( "cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"( d ), "l"( gsrc ), "r"( sz ) : "memory" );
#else
   *sdst = valid ? *gsrc : 0.0;
#endif
}
__device__ __forceinline__ void cp_async_commit()
{
#if defined( __CUDA_ARCH__ )
   asm volatile( "cp.async.commit_group;" ::: "memory" );
#endif
}
__device__ __forceinline__ void cp_async_wait_all()
{
#if defined( __CUDA_ARCH__ )
   asm volatile( "cp.async.wait_group 0;" ::: "memory" );
#endif
}

struct W4 { double w0, w1, w2, w3; };
// a0..a4 = coefficient at offsets -2..2 (rhs4sg_rev.C:123-148: mux1..mux4)
__device__ __forceinline__ W4 weights4( double a0, double a1, double a2, double a3, double a4 )
{
   W4 w;
   w.w0 = a1 - 0.75 * ( a2 + a0 );
   w.w1 = a0 + a3 + 3 * ( a2 + a1 );
   w.w2 = a1 + a4 + 3 * ( a3 + a2 );
   w.w3 = a3 - 0.75 * ( a2 + a4 );
   return w;
}
__device__ __forceinline__ double gsum( const W4& w, double f0, double f1, double f2, double f3, double f4 )
{
   return w.w0 * ( f0 - f2 ) + w.w1 * ( f1 - f2 ) + w.w2 * ( f3 - f2 ) + w.w3 * ( f4 - f2 );
}
// 12 * centred first difference
__device__ __forceinline__ double d0u( double fm2, double fm1, double fp1, double fp2 )
{
   return ( fm2 - fp2 ) + 8 * ( fp1 - fm1 );
}

template <int TY>
struct Cfg
{
   static constexpr int TX = 32, PX = TX + 4, PY = TY + 4, PLANE = PX * PY, NT = TX * TY, NSLOT = 6;
   static constexpr int NPT = ( PLANE + NT - 1 ) / NT;
   static constexpr int NH = 4 * TY + 4 * TX;  // helper (ring) points per plane
   static constexpr int SMEM_DOUBLES = 5 * NSLOT * PLANE + 3 * TY * PX + 3 * PY * TX + PX + PY;
};

} // namespace fast

template <int TY, int EPI>
__global__ void __launch_bounds__( 32 * TY, 1 ) k_rhs_fast( const FastArgs a )
{
   using namespace fast;
   typedef Cfg<TY> C;
   constexpr int TX = C::TX, PX = C::PX, PY = C::PY, PLANE = C::PLANE, NT = C::NT, NSLOT = C::NSLOT;
   SW4_DYN_SMEM( smem );
   double* const s_f = smem;			     // [5 fields][NSLOT][PLANE]: u,v,w,mu,la
   double* const s_ex = smem + 5 * NSLOT * PLANE;    // [3][TY][PX]  E1..E3
   double* const s_ey = s_ex + 3 * TY * PX;	     // [3][PY][TX]  E4..E6
   double* const s_sx = s_ey + 3 * PY * TX;	     // [PX] strx of the tile columns incl. ring
   double* const s_sy = s_sx + PX;		     // [PY]

   const Block& b = a.b;
   const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
   const int li0 = 2 + blockIdx.x * TX, lj0 = 2 + blockIdx.y * TY; // local (array) index of the tile's first output
   const int ka = a.klo + blockIdx.z * a.kchunk;
   const int kb = ( ka + a.kchunk - 1 < a.khi ) ? ka + a.kchunk - 1 : a.khi;
   if( ka > kb ) return;

   // staging assignments: smem point idx -> global offset inside a plane
   int goff[C::NPT];
   bool inb[C::NPT];
#pragma unroll
   for( int q = 0; q < C::NPT; q++ )
   {
      const int idx = tid + q * NT;
      const int sy_ = idx / PX, sx_ = idx - sy_ * PX;
      const int li = li0 - 2 + sx_, lj = lj0 - 2 + sy_;
      inb[q] = idx < PLANE && li < b.ni && lj < b.nj;
      goff[q] = inb[q] ? lj * b.ni + li : 0;
   }
   for( int t = tid; t < PX + PY; t += NT )
   {
      if( t < PX )
      {
	 const int li = li0 - 2 + t;
	 s_sx[t] = li < b.ni ? a.strx[li] : 0.0;
      }
      else
      {
	 const int lj = lj0 - 2 + ( t - PX );
	 s_sy[t - PX] = lj < b.nj ? a.stry[lj] : 0.0;
      }
   }
   const double* const gsrc[5] = { a.u[0], a.u[1], a.u[2], a.mu, a.la };

   auto stage = [&]( int p, int slot ) {
      const long long koff = b.nij * ( p - b.kfirst );
#pragma unroll
      for( int q = 0; q < C::NPT; q++ )
      {
	 const int idx = tid + q * NT;
	 if( idx < PLANE )
	 {
#pragma unroll
	    for( int f = 0; f < 5; f++ )
	       cp_async8( s_f + ( f * NSLOT + slot ) * PLANE + idx, gsrc[f] + ( inb[q] ? koff + goff[q] : 0 ), inb[q] );
	 }
      }
      cp_async_commit();
   };

   const int o = ( ty + 2 ) * PX + tx + 2; // own point in a staged plane
   const int li = li0 + tx, lj = lj0 + ty;
   const bool act = li <= b.ni - 3 && lj <= b.nj - 3;
   const long long gown = (long long)lj * b.ni + li; // own offset inside a plane (valid only if act)

   // register state.  Ring index m <-> plane (k-2+m) once plane p=k+2 has been processed.
   double cu[5], cv[5], cw[5], amz[5], alz[5];
   double g1[5], g2[5], g3[5];
   double pr[3][3];			    // in-plane sums of planes k, k+1, k+2: sx*Gx + sy*Gy per component
   double dyv[3], dyu[3], dxv[3], dxu[3];   // in-plane first differences of planes k..k+2
   double muk[3], lak[3];		    // own mu, la of planes k..k+2
#pragma unroll
   for( int m = 0; m < 5; m++ ) cu[m] = cv[m] = cw[m] = amz[m] = alz[m] = g1[m] = g2[m] = g3[m] = 0;
#pragma unroll
   for( int m = 0; m < 3; m++ )
   {
      pr[m][0] = pr[m][1] = pr[m][2] = 0;
      dyv[m] = dyu[m] = dxv[m] = dxu[m] = muk[m] = lak[m] = 0;
   }

   int slot = 0;
   stage( ka - 2, 0 );
   __syncthreads(); // s_sx, s_sy visible
   const double sx = s_sx[tx + 2], sy = s_sy[ty + 2];
   const double sxm2 = s_sx[tx], sxm1 = s_sx[tx + 1], sxp1 = s_sx[tx + 3], sxp2 = s_sx[tx + 4];
   const double sym2 = s_sy[ty], sym1 = s_sy[ty + 1], syp1 = s_sy[ty + 3], syp2 = s_sy[ty + 4];

   for( int p = ka - 2; p <= kb + 2; p++ )
   {
      const int k = p - 2;
      const bool do_out = k >= ka;
      cp_async_wait_all();
      __syncthreads(); // plane p visible; everybody is done with the previous step's E buffers and old planes
      const int nslot = slot + 1 == NSLOT ? 0 : slot + 1;
      if( p + 1 <= kb + 2 ) stage( p + 1, nslot );

      // epilogue operands of plane k: issue the global loads early
      double e_rho = 1, e_um[3] = { 0, 0, 0 }, e_fo[3] = { 0, 0, 0 };
      if( EPI != EPI_LU && do_out && act )
      {
	 const long long q = b.nij * ( k - b.kfirst ) + gown;
	 e_rho = a.rho[q];
#pragma unroll
	 for( int c = 0; c < 3; c++ )
	 {
	    e_um[c] = a.um[c][q];
	    if( a.fo[0] ) e_fo[c] = a.fo[c][q];
	 }
      }

      // ---- rotate the rings
#pragma unroll
      for( int m = 0; m < 4; m++ )
      {
	 cu[m] = cu[m + 1]; cv[m] = cv[m + 1]; cw[m] = cw[m + 1];
	 amz[m] = amz[m + 1]; alz[m] = alz[m + 1];
	 g1[m] = g1[m + 1]; g2[m] = g2[m + 1]; g3[m] = g3[m + 1];
      }
#pragma unroll
      for( int m = 0; m < 2; m++ )
      {
	 pr[m][0] = pr[m + 1][0]; pr[m][1] = pr[m + 1][1]; pr[m][2] = pr[m + 1][2];
	 dyv[m] = dyv[m + 1]; dyu[m] = dyu[m + 1]; dxv[m] = dxv[m + 1]; dxu[m] = dxu[m + 1];
	 muk[m] = muk[m + 1]; lak[m] = lak[m + 1];
      }

      // ---- in-plane pieces of plane p
      {
	 const double* const pu = s_f + ( 0 * NSLOT + slot ) * PLANE + o;
	 const double* const pv = s_f + ( 1 * NSLOT + slot ) * PLANE + o;
	 const double* const pw = s_f + ( 2 * NSLOT + slot ) * PLANE + o;
	 const double* const pm = s_f + ( 3 * NSLOT + slot ) * PLANE + o;
	 const double* const pl = s_f + ( 4 * NSLOT + slot ) * PLANE + o;
	 const double u0 = pu[0], v0 = pv[0], w0 = pw[0], m0 = pm[0], l0 = pl[0];
	 const double szp = a.strz[p - b.kfirst];
	 cu[4] = u0; cv[4] = v0; cw[4] = w0;
	 amz[4] = m0 * szp; alz[4] = ( 2 * m0 + l0 ) * szp;
	 muk[2] = m0; lak[2] = l0;
	 // x direction
	 const double uxm2 = pu[-2], uxm1 = pu[-1], uxp1 = pu[1], uxp2 = pu[2];
	 const double vxm2 = pv[-2], vxm1 = pv[-1], vxp1 = pv[1], vxp2 = pv[2];
	 const double wxm2 = pw[-2], wxm1 = pw[-1], wxp1 = pw[1], wxp2 = pw[2];
	 const double dxu_ = d0u( uxm2, uxm1, uxp1, uxp2 );
	 const double dxv_ = d0u( vxm2, vxm1, vxp1, vxp2 );
	 const double dxw_ = d0u( wxm2, wxm1, wxp1, wxp2 );
	 // y direction
	 const double uym2 = pu[-2 * PX], uym1 = pu[-PX], uyp1 = pu[PX], uyp2 = pu[2 * PX];
	 const double vym2 = pv[-2 * PX], vym1 = pv[-PX], vyp1 = pv[PX], vyp2 = pv[2 * PX];
	 const double wym2 = pw[-2 * PX], wym1 = pw[-PX], wyp1 = pw[PX], wyp2 = pw[2 * PX];
	 const double dyu_ = d0u( uym2, uym1, uyp1, uyp2 );
	 const double dyv_ = d0u( vym2, vym1, vyp1, vyp2 );
	 const double dyw_ = d0u( wym2, wym1, wyp1, wyp2 );
	 dxu[2] = dxu_; dxv[2] = dxv_; dyu[2] = dyu_; dyv[2] = dyv_;
	 g1[4] = m0 * dxw_;
	 g2[4] = m0 * dyw_;
	 g3[4] = l0 * ( sx * dxu_ + sy * dyv_ );
	 if( p >= ka )
	 {
	    const double mxm2 = pm[-2], mxm1 = pm[-1], mxp1 = pm[1], mxp2 = pm[2];
	    const double lxm2 = pl[-2], lxm1 = pl[-1], lxp1 = pl[1], lxp2 = pl[2];
	    const W4 wmx = weights4( mxm2 * sxm2, mxm1 * sxm1, m0 * sx, mxp1 * sxp1, mxp2 * sxp2 );
	    const W4 wlx = weights4( ( 2 * mxm2 + lxm2 ) * sxm2, ( 2 * mxm1 + lxm1 ) * sxm1, ( 2 * m0 + l0 ) * sx,
				     ( 2 * mxp1 + lxp1 ) * sxp1, ( 2 * mxp2 + lxp2 ) * sxp2 );
	    const double mym2 = pm[-2 * PX], mym1 = pm[-PX], myp1 = pm[PX], myp2 = pm[2 * PX];
	    const double lym2 = pl[-2 * PX], lym1 = pl[-PX], lyp1 = pl[PX], lyp2 = pl[2 * PX];
	    const W4 wmy = weights4( mym2 * sym2, mym1 * sym1, m0 * sy, myp1 * syp1, myp2 * syp2 );
	    const W4 wly = weights4( ( 2 * mym2 + lym2 ) * sym2, ( 2 * mym1 + lym1 ) * sym1, ( 2 * m0 + l0 ) * sy,
				     ( 2 * myp1 + lyp1 ) * syp1, ( 2 * myp2 + lyp2 ) * syp2 );
	    pr[2][0] = sx * gsum( wlx, uxm2, uxm1, u0, uxp1, uxp2 ) + sy * gsum( wmy, uym2, uym1, u0, uyp1, uyp2 );
	    pr[2][1] = sx * gsum( wmx, vxm2, vxm1, v0, vxp1, vxp2 ) + sy * gsum( wly, vym2, vym1, v0, vyp1, vyp2 );
	    pr[2][2] = sx * gsum( wmx, wxm2, wxm1, w0, wxp1, wxp2 ) + sy * gsum( wmy, wym2, wym1, w0, wyp1, wyp2 );
	 }
      }

      // ---- z pieces of plane k and the exchanged products
      double szk = 0, rz[3] = { 0, 0, 0 }, t1 = 0, t2 = 0, t3 = 0;
      if( do_out )
      {
	 szk = a.strz[k - b.kfirst];
	 const W4 wmz = weights4( amz[0], amz[1], amz[2], amz[3], amz[4] );
	 const W4 wlz = weights4( alz[0], alz[1], alz[2], alz[3], alz[4] );
	 rz[0] = pr[0][0] + szk * gsum( wmz, cu[0], cu[1], cu[2], cu[3], cu[4] );
	 rz[1] = pr[0][1] + szk * gsum( wmz, cv[0], cv[1], cv[2], cv[3], cv[4] );
	 rz[2] = pr[0][2] + szk * gsum( wlz, cw[0], cw[1], cw[2], cw[3], cw[4] );
	 const double dzu = d0u( cu[0], cu[1], cu[3], cu[4] );
	 const double dzv = d0u( cv[0], cv[1], cv[3], cv[4] );
	 const double dzw = d0u( cw[0], cw[1], cw[3], cw[4] );
	 t1 = d0u( g1[0], g1[1], g1[3], g1[4] );
	 t2 = d0u( g2[0], g2[1], g2[3], g2[4] );
	 t3 = d0u( g3[0], g3[1], g3[3], g3[4] );
	 const double mk = muk[0], lk = lak[0];
	 double* const ex = s_ex + ty * PX + tx + 2;
	 double* const ey = s_ey + ( ty + 2 ) * TX + tx;
	 ex[0] = lk * ( sy * dyv[0] + szk * dzw );
	 ex[TY * PX] = mk * dyu[0];
	 ex[2 * TY * PX] = mk * dzu;
	 ey[0] = mk * dxv[0];
	 ey[PY * TX] = lk * ( sx * dxu[0] + szk * dzw );
	 ey[2 * PY * TX] = mk * dzv;

	 // ring of width 2 around the tile: recompute the products from the staged planes
	 const int sk = slot - 2 < 0 ? slot - 2 + NSLOT : slot - 2; // slot of plane k
	 const int skm1 = sk - 1 < 0 ? sk - 1 + NSLOT : sk - 1, skm2 = skm1 - 1 < 0 ? skm1 - 1 + NSLOT : skm1 - 1;
	 const int skp1 = sk + 1 == NSLOT ? 0 : sk + 1;
	 for( int hh = tid; hh < C::NH; hh += NT )
	 {
	    if( hh < 4 * TY )
	    {
	       const int hx = hh & 3, row = hh >> 2;
	       const int sx_ = hx < 2 ? hx : TX + hx, sy_ = row + 2;
	       const int oo = sy_ * PX + sx_;
	       const double* const qu = s_f + ( 0 * NSLOT + sk ) * PLANE + oo;
	       const double* const qv = s_f + ( 1 * NSLOT + sk ) * PLANE + oo;
	       const double hm = s_f[( 3 * NSLOT + sk ) * PLANE + oo], hl = s_f[( 4 * NSLOT + sk ) * PLANE + oo];
	       const double hdyv = d0u( qv[-2 * PX], qv[-PX], qv[PX], qv[2 * PX] );
	       const double hdyu = d0u( qu[-2 * PX], qu[-PX], qu[PX], qu[2 * PX] );
	       const double hdzu = d0u( s_f[( 0 * NSLOT + skm2 ) * PLANE + oo], s_f[( 0 * NSLOT + skm1 ) * PLANE + oo],
					s_f[( 0 * NSLOT + skp1 ) * PLANE + oo], s_f[( 0 * NSLOT + slot ) * PLANE + oo] );
	       const double hdzw = d0u( s_f[( 2 * NSLOT + skm2 ) * PLANE + oo], s_f[( 2 * NSLOT + skm1 ) * PLANE + oo],
					s_f[( 2 * NSLOT + skp1 ) * PLANE + oo], s_f[( 2 * NSLOT + slot ) * PLANE + oo] );
	       double* const hx_ = s_ex + row * PX + sx_;
	       hx_[0] = hl * ( s_sy[sy_] * hdyv + szk * hdzw );
	       hx_[TY * PX] = hm * hdyu;
	       hx_[2 * TY * PX] = hm * hdzu;
	    }
	    else
	    {
	       const int t = hh - 4 * TY;
	       const int hy = t >> 5, col = t & 31;
	       const int sy_ = hy < 2 ? hy : TY + hy, sx_ = col + 2;
	       const int oo = sy_ * PX + sx_;
	       const double* const qu = s_f + ( 0 * NSLOT + sk ) * PLANE + oo;
	       const double* const qv = s_f + ( 1 * NSLOT + sk ) * PLANE + oo;
	       const double hm = s_f[( 3 * NSLOT + sk ) * PLANE + oo], hl = s_f[( 4 * NSLOT + sk ) * PLANE + oo];
	       const double hdxv = d0u( qv[-2], qv[-1], qv[1], qv[2] );
	       const double hdxu = d0u( qu[-2], qu[-1], qu[1], qu[2] );
	       const double hdzv = d0u( s_f[( 1 * NSLOT + skm2 ) * PLANE + oo], s_f[( 1 * NSLOT + skm1 ) * PLANE + oo],
					s_f[( 1 * NSLOT + skp1 ) * PLANE + oo], s_f[( 1 * NSLOT + slot ) * PLANE + oo] );
	       const double hdzw = d0u( s_f[( 2 * NSLOT + skm2 ) * PLANE + oo], s_f[( 2 * NSLOT + skm1 ) * PLANE + oo],
					s_f[( 2 * NSLOT + skp1 ) * PLANE + oo], s_f[( 2 * NSLOT + slot ) * PLANE + oo] );
	       double* const hy_ = s_ey + sy_ * TX + col;
	       hy_[0] = hm * hdxv;
	       hy_[PY * TX] = hl * ( s_sx[sx_] * hdxu + szk * hdzw );
	       hy_[2 * PY * TX] = hm * hdzv;
	    }
	 }
      }
      __syncthreads();
      if( do_out && act )
      {
	 const double* const ex = s_ex + ty * PX + tx + 2;
	 const double* const ey = s_ey + ( ty + 2 ) * TX + tx;
	 const double x1 = d0u( ex[-2], ex[-1], ex[1], ex[2] );
	 const double x2 = d0u( ex[TY * PX - 2], ex[TY * PX - 1], ex[TY * PX + 1], ex[TY * PX + 2] );
	 const double x3 = d0u( ex[2 * TY * PX - 2], ex[2 * TY * PX - 1], ex[2 * TY * PX + 1], ex[2 * TY * PX + 2] );
	 const double y1 = d0u( ey[-2 * TX], ey[-TX], ey[TX], ey[2 * TX] );
	 const double y2 = d0u( ey[PY * TX - 2 * TX], ey[PY * TX - TX], ey[PY * TX + TX], ey[PY * TX + 2 * TX] );
	 const double y3 = d0u( ey[2 * PY * TX - 2 * TX], ey[2 * PY * TX - TX], ey[2 * PY * TX + TX], ey[2 * PY * TX + 2 * TX] );
	 // lu_c = cof * ( rz_c/6 + mixed_c/144 )
	 double r[3];
	 r[0] = a.cof6 * rz[0] + a.cof144 * ( sx * ( x1 + sy * y1 + szk * t1 ) );
	 r[1] = a.cof6 * rz[1] + a.cof144 * ( sy * ( sx * x2 + y2 + szk * t2 ) );
	 r[2] = a.cof6 * rz[2] + a.cof144 * ( szk * ( sx * x3 + sy * y3 + t3 ) );
	 const long long q = b.nij * ( k - b.kfirst ) + gown;
	 if( EPI == EPI_LU )
	 {
#pragma unroll
	    for( int c = 0; c < 3; c++ ) a.out[c][q] = r[c];
	 }
	 else if( EPI == EPI_PRED )
	 {
	    const double rinv = 1.0 / e_rho; // one division per point; dt^2/rho and acc/rho are formed from it
	    const double f = a.fac * rinv;
	    const double uk[3] = { cu[2], cv[2], cw[2] };
#pragma unroll
	    for( int c = 0; c < 3; c++ )
	    {
	       const double acc = r[c] + e_fo[c];
	       a.out[c][q] = 2 * uk[c] - e_um[c] + f * acc;
	       if( a.out2[0] ) a.out2[c][q] = acc * rinv;
	    }
	 }
	 else
	 {
	    const double f = a.fac / e_rho;
#pragma unroll
	    for( int c = 0; c < 3; c++ ) a.out[c][q] = e_um[c] + f * ( r[c] + e_fo[c] );
	 }
      }
      slot = nslot;
   }
}

#ifndef SW4B200_EMULATE
namespace {
template <int TY, int EPI>
int launch_fast_t( const FastArgs& a, cudaStream_t st )
{
   typedef fast::Cfg<TY> C;
   static bool configured = false;
   const size_t smem = C::SMEM_DOUBLES * sizeof( double );
   if( !configured )
   {
      cudaError_t e = cudaFuncSetAttribute( k_rhs_fast<TY, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem );
      if( e != cudaSuccess ) return set_error( "k_rhs_fast: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString( e ) );
      configured = true;
   }
   const Block& b = a.b;
   dim3 bs( C::TX, TY, 1 );
   dim3 gs( ( b.ni - 4 + C::TX - 1 ) / C::TX, ( b.nj - 4 + TY - 1 ) / TY, ( a.khi - a.klo + 1 + a.kchunk - 1 ) / a.kchunk );
   ProfScope prof( EPI == EPI_PRED ? "rhs_fast_pred" : ( EPI == EPI_CORR ? "rhs_fast_corr" : "rhs_fast_lu" ), st );
   k_rhs_fast<TY, EPI><<<gs, bs, smem, st>>>( a );
   count_launch();
   return check_launch( "k_rhs_fast" );
}
} // namespace

// planes per CTA in z: enough CTAs to fill the 148 SMs a few times over, chunks not shorter than 16
int fast_kchunk( const Block& b, int nplanes, int ty )
{
   const long long tiles = (long long)( ( b.ni - 4 + 31 ) / 32 ) * ( ( b.nj - 4 + ty - 1 ) / ty );
   const long long want = 148LL * 6; // CTAs
   long long nch = ( want + tiles - 1 ) / tiles;
   if( nch < 1 ) nch = 1;
   long long kc = ( nplanes + nch - 1 ) / nch;
   if( kc < 16 ) kc = 16;
   if( kc > nplanes ) kc = nplanes;
   return (int)kc;
}

int launch_fast2( int epi, FastArgs a, cudaStream_t st );
int launch_fast3( int variant, int epi, const FastArgs& a, cudaStream_t st );
int launch_fast4( int epi, const FastArgs& a, cudaStream_t st );

// Generation of the interior kernel.  Default: the fourth generation (rhs4sg_fast4.cu: x-pair register blocking, z state
// in tensor memory, TMA-staged planes; grids whose rows are not 16-byte aligned fall back to the second generation
// inside launch_fast4).  SW4B200_FAST_GEN=1 / 2 select the earlier generations, 3000+10*TY+TMODE a variant of the third
// (3122, 3121, 3082, 3081: one point per thread with the z state in tensor memory); all are kept for A/B measurements.
static int fast_generation()
{
   static int v = -1;
   if( v < 0 )
   {
      const char* e = getenv( "SW4B200_FAST_GEN" );
      v = e ? atoi( e ) : SW4B200_DEFAULT_FAST_GEN;
      if( v == 3 ) v = 3122;
      if( v != 1 && v != 2 && v != 4 && v != 3122 && v != 3121 && v != 3082 && v != 3081 ) v = SW4B200_DEFAULT_FAST_GEN;
   }
   return v;
}

int launch_fast( int epi, FastArgs a, cudaStream_t st )
{
   if( a.khi < a.klo ) return 0;
   if( fast_generation() == 4 ) return launch_fast4( epi, a, st );
   if( fast_generation() >= 3000 ) return launch_fast3( fast_generation() - 3000, epi, a, st );
   if( fast_generation() == 2 ) return launch_fast2( epi, a, st );
   if( a.kchunk <= 0 ) a.kchunk = fast_kchunk( a.b, a.khi - a.klo + 1, 8 );
   switch( epi )
   {
   case EPI_LU: return launch_fast_t<8, EPI_LU>( a, st );
   case EPI_PRED: return launch_fast_t<8, EPI_PRED>( a, st );
   default: return launch_fast_t<8, EPI_CORR>( a, st );
   }
}
#endif

} // namespace sw4b200
