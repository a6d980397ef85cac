// Throughput path for the Cartesian interior rows, second generation (SoA layout, corder=1).
//
// Same operator algebra and data flow as rhs4sg_fast.cu (reference formulas rhs4sg_rev.C:112-348 with
// the common subexpressions shared between threads), restructured for the instruction issue limits
// that ncu showed on the first version (profiles/r01a_k_rhs_fast_ncu.md):
//  * the z-march is unrolled by the depth of the shared-memory plane ring (6): every register ring
//    index and every ring slot offset is a compile-time constant, so no ring-rotation moves and no
//    slot arithmetic are executed;
//  * ONE block barrier per plane instead of two: the exchanged cross-term products E1..E6 of plane k
//    are double buffered and differenced one step later (while plane k+1's work is in flight), so the
//    barrier that publishes the newly staged plane also publishes them;
//  * the steady-state step is branch free (prologue planes run the same code, their stores are
//    predicated off), which lets ptxas interleave the independent in-plane and z-column work.
//
// Also compiled by g++ (SW4B200_EMULATE) for the CPU check of the kernel source (tests/emu).
#include "fast_common.cuh"

namespace sw4b200 {

namespace fast2 {

using fast::cp_async8;
using fast::cp_async_commit;
using fast::cp_async_wait_all;
using fast::W4;
using fast::weights4;
using fast::gsum;
using fast::d0u;

template <int TY>
struct Cfg
{
   static constexpr int TX = 32, PX = TX + 4, PY = TY + 4, PLANE = PX * PY, NT = TX * TY, NSLOT = 6;
   static constexpr int NPT = ( PLANE + NT - 1 ) / NT;
   static constexpr int NH = 4 * TY + 4 * TX; // helper (ring) points per plane
   static constexpr int EX = 3 * TY * PX, EY = 3 * PY * TX;
   static constexpr int OPS = 4 * NT; // epilogue operands of one plane: rho, um[3] of the own points
   static constexpr int SZMAX = 1024; // strz of the planes a CTA marches through (kchunk + 6 <= SZMAX)
   static constexpr int SMEM_DOUBLES = 5 * NSLOT * PLANE + 2 * ( EX + EY ) + PX + PY + 2 * OPS + SZMAX;
};

// per-thread register state.  6-rings: plane q lives at physical index (q - p0) % 6 (5 planes are
// live at any time); 3-rings likewise modulo 3.
struct State
{
   double cu[6], cv[6], cw[6], amz[6], alz[6], g1[6], g2[6], g3[6];
   double pr[3][3];			  // in-plane sums sx*Gx + sy*Gy per component
   double dyv[3], dyu[3], dxv[3], dxu[3]; // in-plane first differences
   double muk[3], lak[3];
   double rp[3];			  // result of the previous plane, still lacking the exchanged cross terms
};

template <int TY>
struct Ctx
{
   typedef Cfg<TY> C;
   double *s_f, *s_ex, *s_ey, *s_sx, *s_sy, *s_op, *s_sz;
   int p0;
   int goff[C::NPT];
   bool inb[C::NPT];
   int tid, tx, ty, o;
   int ka, kb, pend;
   bool act;
   long long gown;
   double sx, sy, sxm2, sxm1, sxp1, sxp2, sym2, sym1, syp1, syp2;
};

// stage plane p into ring slot `slot`; with an epilogue also the own-point operands (rho, um) of plane p-3,
// the plane that the step handling plane p finishes, into operand buffer `ob`
template <int TY, int EPI>
__device__ __forceinline__ void stage( const FastArgs& a, const Ctx<TY>& c, int p, int slot, int ob )
{
   typedef Cfg<TY> C;
   if( p > c.pend + 1 ) return;
   if( EPI != EPI_LU )
   {
      const int kq = p - 3;
      if( c.act && kq >= c.ka && kq <= c.kb )
      {
	 const long long q = a.b.nij * ( kq - a.b.kfirst ) + c.gown;
	 double* const d = c.s_op + ob * C::OPS + c.tid;
	 cp_async8( d, a.rho + q, true );
#pragma unroll
	 for( int m = 0; m < 3; m++ ) cp_async8( d + ( m + 1 ) * C::NT, a.um[m] + q, true );
      }
   }
   const long long koff = a.b.nij * ( p - a.b.kfirst );
   const double* const gsrc[5] = { a.u[0], a.u[1], a.u[2], a.mu, a.la };
#pragma unroll
   for( int q = 0; q < C::NPT; q++ )
   {
      const int idx = c.tid + q * C::NT;
      if( idx < C::PLANE && p <= c.pend ) // (the step after the last plane only finishes plane kb)
      {
#pragma unroll
	 for( int f = 0; f < 5; f++ )
	    cp_async8( c.s_f + ( f * C::NSLOT + slot ) * C::PLANE + idx, gsrc[f] + ( c.inb[q] ? koff + c.goff[q] : 0 ), c.inb[q] );
      }
   }
   cp_async_commit();
}

// One step of the march: plane p has been staged into ring slot S.  Does the in-plane work of plane p,
// the z work of plane k=p-2 (publishing its exchanged products in E buffer S&1) and finishes plane k-1
// (reading E buffer (S+1)&1).
template <int S, int TY, int EPI>
__device__ __forceinline__ void step( const FastArgs& a, Ctx<TY>& c, State& s, const int p )
{
   typedef Cfg<TY> C;
   constexpr int TX = C::TX, PX = C::PX, PY = C::PY, PLANE = C::PLANE, NT = C::NT, NSLOT = C::NSLOT;
   // ring positions of planes p, p-1, ..., p-4 and of the 3-deep delays
   constexpr int R0 = S, R1 = ( S + 5 ) % 6, R2 = ( S + 4 ) % 6, R3 = ( S + 3 ) % 6, R4 = ( S + 2 ) % 6;
   constexpr int T0 = S % 3, T2 = ( S + 1 ) % 3; // T0: plane p (= k+2), T2: plane k
   constexpr int EB = S & 1, EF = ( S + 1 ) & 1;
   const Block& b = a.b;
   const int k = p - 2, kf = p - 3;

   cp_async_wait_all();
   __syncthreads(); // plane p and the E products of plane k-1 are visible; slot of plane p-5 is free
   stage<TY, EPI>( a, c, p + 1, ( S + 1 ) % NSLOT, ( S + 1 ) & 1 );

   // epilogue operands of plane kf (staged one step ago by this thread itself); dense forcing, if any, from global
   const bool fin = c.act && kf >= c.ka && kf <= c.kb;
   const long long qf = fin ? b.nij * ( kf - b.kfirst ) + c.gown : 0;
   double e_fo[3] = { 0, 0, 0 };
   if( EPI != EPI_LU && a.fo[0] )
   {
#pragma unroll
      for( int m = 0; m < 3; m++ ) e_fo[m] = a.fo[m][qf];
   }

   const double sx = c.sx, sy = c.sy;
   // ---- in-plane pieces of plane p
   {
      const double* const pu = c.s_f + ( 0 * NSLOT + S ) * PLANE + c.o;
      const double* const pv = c.s_f + ( 1 * NSLOT + S ) * PLANE + c.o;
      const double* const pw = c.s_f + ( 2 * NSLOT + S ) * PLANE + c.o;
      const double* const pm = c.s_f + ( 3 * NSLOT + S ) * PLANE + c.o;
      const double* const pl = c.s_f + ( 4 * NSLOT + S ) * PLANE + c.o;
      const double u0 = pu[0], v0 = pv[0], w0 = pw[0], m0 = pm[0], l0 = pl[0];
      const double szp = c.s_sz[p - c.p0];
      s.cu[R0] = u0; s.cv[R0] = v0; s.cw[R0] = w0;
      s.amz[R0] = m0 * szp; s.alz[R0] = ( 2 * m0 + l0 ) * szp;
      s.muk[T0] = m0; s.lak[T0] = l0;
      const double uxm2 = pu[-2], uxm1 = pu[-1], uxp1 = pu[1], uxp2 = pu[2];
      const double vxm2 = pv[-2], vxm1 = pv[-1], vxp1 = pv[1], vxp2 = pv[2];
      const double wxm2 = pw[-2], wxm1 = pw[-1], wxp1 = pw[1], wxp2 = pw[2];
      const double dxu_ = d0u( uxm2, uxm1, uxp1, uxp2 );
      const double dxv_ = d0u( vxm2, vxm1, vxp1, vxp2 );
      const double dxw_ = d0u( wxm2, wxm1, wxp1, wxp2 );
      const double uym2 = pu[-2 * PX], uym1 = pu[-PX], uyp1 = pu[PX], uyp2 = pu[2 * PX];
      const double vym2 = pv[-2 * PX], vym1 = pv[-PX], vyp1 = pv[PX], vyp2 = pv[2 * PX];
      const double wym2 = pw[-2 * PX], wym1 = pw[-PX], wyp1 = pw[PX], wyp2 = pw[2 * PX];
      const double dyu_ = d0u( uym2, uym1, uyp1, uyp2 );
      const double dyv_ = d0u( vym2, vym1, vyp1, vyp2 );
      const double dyw_ = d0u( wym2, wym1, wyp1, wyp2 );
      s.dxu[T0] = dxu_; s.dxv[T0] = dxv_; s.dyu[T0] = dyu_; s.dyv[T0] = dyv_;
      s.g1[R0] = m0 * dxw_;
      s.g2[R0] = m0 * dyw_;
      s.g3[R0] = l0 * ( sx * dxu_ + sy * dyv_ );
      const double mxm2 = pm[-2], mxm1 = pm[-1], mxp1 = pm[1], mxp2 = pm[2];
      const double lxm2 = pl[-2], lxm1 = pl[-1], lxp1 = pl[1], lxp2 = pl[2];
      const W4 wmx = weights4( mxm2 * c.sxm2, mxm1 * c.sxm1, m0 * sx, mxp1 * c.sxp1, mxp2 * c.sxp2 );
      const W4 wlx = weights4( ( 2 * mxm2 + lxm2 ) * c.sxm2, ( 2 * mxm1 + lxm1 ) * c.sxm1, ( 2 * m0 + l0 ) * sx,
			       ( 2 * mxp1 + lxp1 ) * c.sxp1, ( 2 * mxp2 + lxp2 ) * c.sxp2 );
      const double mym2 = pm[-2 * PX], mym1 = pm[-PX], myp1 = pm[PX], myp2 = pm[2 * PX];
      const double lym2 = pl[-2 * PX], lym1 = pl[-PX], lyp1 = pl[PX], lyp2 = pl[2 * PX];
      const W4 wmy = weights4( mym2 * c.sym2, mym1 * c.sym1, m0 * sy, myp1 * c.syp1, myp2 * c.syp2 );
      const W4 wly = weights4( ( 2 * mym2 + lym2 ) * c.sym2, ( 2 * mym1 + lym1 ) * c.sym1, ( 2 * m0 + l0 ) * sy,
			       ( 2 * myp1 + lyp1 ) * c.syp1, ( 2 * myp2 + lyp2 ) * c.syp2 );
      s.pr[T0][0] = sx * gsum( wlx, uxm2, uxm1, u0, uxp1, uxp2 ) + sy * gsum( wmy, uym2, uym1, u0, uyp1, uyp2 );
      s.pr[T0][1] = sx * gsum( wmx, vxm2, vxm1, v0, vxp1, vxp2 ) + sy * gsum( wly, vym2, vym1, v0, vyp1, vyp2 );
      s.pr[T0][2] = sx * gsum( wmx, wxm2, wxm1, w0, wxp1, wxp2 ) + sy * gsum( wmy, wym2, wym1, w0, wyp1, wyp2 );
   }

   // ---- z pieces of plane k and its exchanged products
   const double szk = k >= c.p0 ? c.s_sz[k - c.p0] : 0.0;
   double rnew[3];
   {
      const W4 wmz = weights4( s.amz[R4], s.amz[R3], s.amz[R2], s.amz[R1], s.amz[R0] );
      const W4 wlz = weights4( s.alz[R4], s.alz[R3], s.alz[R2], s.alz[R1], s.alz[R0] );
      const double rz0 = s.pr[T2][0] + szk * gsum( wmz, s.cu[R4], s.cu[R3], s.cu[R2], s.cu[R1], s.cu[R0] );
      const double rz1 = s.pr[T2][1] + szk * gsum( wmz, s.cv[R4], s.cv[R3], s.cv[R2], s.cv[R1], s.cv[R0] );
      const double rz2 = s.pr[T2][2] + szk * gsum( wlz, s.cw[R4], s.cw[R3], s.cw[R2], s.cw[R1], s.cw[R0] );
      const double dzu = d0u( s.cu[R4], s.cu[R3], s.cu[R1], s.cu[R0] );
      const double dzv = d0u( s.cv[R4], s.cv[R3], s.cv[R1], s.cv[R0] );
      const double dzw = d0u( s.cw[R4], s.cw[R3], s.cw[R1], s.cw[R0] );
      const double t1 = d0u( s.g1[R4], s.g1[R3], s.g1[R1], s.g1[R0] );
      const double t2 = d0u( s.g2[R4], s.g2[R3], s.g2[R1], s.g2[R0] );
      const double t3 = d0u( s.g3[R4], s.g3[R3], s.g3[R1], s.g3[R0] );
      const double mk = s.muk[T2], lk = s.lak[T2];
      double* const ex = c.s_ex + EB * C::EX + c.ty * PX + c.tx + 2;
      double* const ey = c.s_ey + EB * C::EY + ( c.ty + 2 ) * TX + c.tx;
      const double lzw = szk * dzw;
      ex[0] = lk * ( sy * s.dyv[T2] + lzw );
      ex[TY * PX] = mk * s.dyu[T2];
      ex[2 * TY * PX] = mk * dzu;
      ey[0] = mk * s.dxv[T2];
      ey[PY * TX] = lk * ( sx * s.dxu[T2] + lzw );
      ey[2 * PY * TX] = mk * dzv;
      const double c144z = a.cof144 * szk;
      rnew[0] = a.cof6 * rz0 + ( c144z * sx ) * t1;
      rnew[1] = a.cof6 * rz1 + ( c144z * sy ) * t2;
      rnew[2] = a.cof6 * rz2 + c144z * t3;

      // ring of width 2 around the tile: the same products recomputed from the staged planes
      // (plane k sits in slot R2, k-2..k+2 in R4..R0)
      for( int hh = c.tid; hh < C::NH; hh += NT )
      {
	 if( hh < 4 * TY )
	 {
	    const int hx = hh & 3, row = hh >> 2;
	    const int sx_ = hx < 2 ? hx : TX + hx, sy_ = row + 2;
	    const int oo = sy_ * PX + sx_;
	    const double* const qu = c.s_f + ( 0 * NSLOT + R2 ) * PLANE + oo;
	    const double* const qv = c.s_f + ( 1 * NSLOT + R2 ) * PLANE + oo;
	    const double hm = c.s_f[( 3 * NSLOT + R2 ) * PLANE + oo], hl = c.s_f[( 4 * NSLOT + R2 ) * PLANE + oo];
	    const double hdyv = d0u( qv[-2 * PX], qv[-PX], qv[PX], qv[2 * PX] );
	    const double hdyu = d0u( qu[-2 * PX], qu[-PX], qu[PX], qu[2 * PX] );
	    const double hdzu = d0u( c.s_f[( 0 * NSLOT + R4 ) * PLANE + oo], c.s_f[( 0 * NSLOT + R3 ) * PLANE + oo],
				     c.s_f[( 0 * NSLOT + R1 ) * PLANE + oo], c.s_f[( 0 * NSLOT + R0 ) * PLANE + oo] );
	    const double hdzw = d0u( c.s_f[( 2 * NSLOT + R4 ) * PLANE + oo], c.s_f[( 2 * NSLOT + R3 ) * PLANE + oo],
				     c.s_f[( 2 * NSLOT + R1 ) * PLANE + oo], c.s_f[( 2 * NSLOT + R0 ) * PLANE + oo] );
	    double* const hx_ = c.s_ex + EB * C::EX + row * PX + sx_;
	    hx_[0] = hl * ( c.s_sy[sy_] * hdyv + szk * hdzw );
	    hx_[TY * PX] = hm * hdyu;
	    hx_[2 * TY * PX] = hm * hdzu;
	 }
	 else
	 {
	    const int t = hh - 4 * TY;
	    const int hy = t >> 5, col = t & 31;
	    const int sy_ = hy < 2 ? hy : TY + hy, sx_ = col + 2;
	    const int oo = sy_ * PX + sx_;
	    const double* const qu = c.s_f + ( 0 * NSLOT + R2 ) * PLANE + oo;
	    const double* const qv = c.s_f + ( 1 * NSLOT + R2 ) * PLANE + oo;
	    const double hm = c.s_f[( 3 * NSLOT + R2 ) * PLANE + oo], hl = c.s_f[( 4 * NSLOT + R2 ) * PLANE + oo];
	    const double hdxv = d0u( qv[-2], qv[-1], qv[1], qv[2] );
	    const double hdxu = d0u( qu[-2], qu[-1], qu[1], qu[2] );
	    const double hdzv = d0u( c.s_f[( 1 * NSLOT + R4 ) * PLANE + oo], c.s_f[( 1 * NSLOT + R3 ) * PLANE + oo],
				     c.s_f[( 1 * NSLOT + R1 ) * PLANE + oo], c.s_f[( 1 * NSLOT + R0 ) * PLANE + oo] );
	    const double hdzw = d0u( c.s_f[( 2 * NSLOT + R4 ) * PLANE + oo], c.s_f[( 2 * NSLOT + R3 ) * PLANE + oo],
				     c.s_f[( 2 * NSLOT + R1 ) * PLANE + oo], c.s_f[( 2 * NSLOT + R0 ) * PLANE + oo] );
	    double* const hy_ = c.s_ey + EB * C::EY + sy_ * TX + col;
	    hy_[0] = hm * hdxv;
	    hy_[PY * TX] = hl * ( c.s_sx[sx_] * hdxu + szk * hdzw );
	    hy_[2 * PY * TX] = hm * hdzv;
	 }
      }
   }

   // ---- finish plane kf = k-1: difference its exchanged products (published by this step's barrier)
   {
      const double* const ex = c.s_ex + EF * C::EX + c.ty * PX + c.tx + 2;
      const double* const ey = c.s_ey + EF * C::EY + ( c.ty + 2 ) * TX + c.tx;
      const double x1 = d0u( ex[-2], ex[-1], ex[1], ex[2] );
      const double x2 = d0u( ex[TY * PX - 2], ex[TY * PX - 1], ex[TY * PX + 1], ex[TY * PX + 2] );
      const double x3 = d0u( ex[2 * TY * PX - 2], ex[2 * TY * PX - 1], ex[2 * TY * PX + 1], ex[2 * TY * PX + 2] );
      const double y1 = d0u( ey[-2 * TX], ey[-TX], ey[TX], ey[2 * TX] );
      const double y2 = d0u( ey[PY * TX - 2 * TX], ey[PY * TX - TX], ey[PY * TX + TX], ey[PY * TX + 2 * TX] );
      const double y3 = d0u( ey[2 * PY * TX - 2 * TX], ey[2 * PY * TX - TX], ey[2 * PY * TX + TX], ey[2 * PY * TX + 2 * TX] );
      double e_rho = 1, e_um[3] = { 0, 0, 0 };
      if( EPI != EPI_LU )
      {
	 const double* const d = c.s_op + ( S & 1 ) * C::OPS + c.tid;
	 e_rho = d[0];
	 e_um[0] = d[NT]; e_um[1] = d[2 * NT]; e_um[2] = d[3 * NT];
      }
      double r[3];
      r[0] = s.rp[0] + ( a.cof144 * sx ) * ( x1 + sy * y1 );
      r[1] = s.rp[1] + ( a.cof144 * sy ) * ( sx * x2 + y2 );
      r[2] = s.rp[2] + ( a.cof144 * ( kf >= c.p0 ? c.s_sz[kf - c.p0] : 0.0 ) ) * ( sx * x3 + sy * y3 );
      if( fin )
      {
	 if( EPI == EPI_LU )
	 {
#pragma unroll
	    for( int m = 0; m < 3; m++ ) a.out[m][qf] = r[m];
	 }
	 else if( EPI == EPI_PRED )
	 {
	    const double rinv = 1.0 / e_rho; // one division per point; dt^2/rho and acc/rho are formed from it
	    const double f = a.fac * rinv;
	    const double uk[3] = { s.cu[R3], s.cv[R3], s.cw[R3] }; // plane kf = p-3
#pragma unroll
	    for( int m = 0; m < 3; m++ )
	    {
	       const double acc = r[m] + e_fo[m];
	       a.out[m][qf] = 2 * uk[m] - e_um[m] + f * acc;
	       if( a.out2[0] ) a.out2[m][qf] = acc * rinv;
	    }
	 }
	 else
	 {
	    const double f = a.fac / e_rho;
#pragma unroll
	    for( int m = 0; m < 3; m++ ) a.out[m][qf] = e_um[m] + f * ( r[m] + e_fo[m] );
	 }
      }
   }
   s.rp[0] = rnew[0]; s.rp[1] = rnew[1]; s.rp[2] = rnew[2];
}

} // namespace fast2

template <int TY, int EPI>
__global__ void __launch_bounds__( 32 * TY, 1 ) k_rhs_fast2( const FastArgs a )
{
   using namespace fast2;
   typedef fast2::Cfg<TY> C;
   constexpr int TX = C::TX, PX = C::PX, PY = C::PY, PLANE = C::PLANE, NT = C::NT, NSLOT = C::NSLOT;
   SW4_DYN_SMEM( smem );
   fast2::Ctx<TY> c;
   c.s_f = smem;				    // [5 fields][NSLOT][PLANE]: u,v,w,mu,la
   c.s_ex = smem + 5 * NSLOT * PLANE;		    // [2][3][TY][PX]  E1..E3, double buffered
   c.s_ey = c.s_ex + 2 * C::EX;			    // [2][3][PY][TX]  E4..E6
   c.s_sx = c.s_ey + 2 * C::EY;			    // [PX] strx of the tile columns incl. ring
   c.s_sy = c.s_sx + PX;			    // [PY]
   c.s_op = c.s_sy + PY;			    // [2][4][NT] rho, um of the own points of the plane being finished
   c.s_sz = c.s_op + 2 * C::OPS;		    // [SZMAX] strz of planes ka-2 .. kb+3

   const Block& b = a.b;
   c.tx = threadIdx.x; c.ty = threadIdx.y; c.tid = c.ty * TX + c.tx;
   const int li0 = 2 + blockIdx.x * TX, lj0 = 2 + blockIdx.y * TY; // local (array) index of the tile's first output
   c.ka = a.klo + blockIdx.z * a.kchunk;
   c.kb = ( c.ka + a.kchunk - 1 < a.khi ) ? c.ka + a.kchunk - 1 : a.khi;
   if( c.ka > c.kb ) return;
   c.pend = c.kb + 2;

#pragma unroll
   for( int q = 0; q < C::NPT; q++ )
   {
      const int idx = c.tid + q * NT;
      const int sy_ = idx / PX, sx_ = idx - sy_ * PX;
      const int li = li0 - 2 + sx_, lj = lj0 - 2 + sy_;
      c.inb[q] = idx < PLANE && li < b.ni && lj < b.nj;
      c.goff[q] = c.inb[q] ? lj * b.ni + li : 0;
   }
   for( int t = c.tid; t < PX + PY; t += NT )
   {
      if( t < PX )
      {
	 const int li = li0 - 2 + t;
	 c.s_sx[t] = li < b.ni ? a.strx[li] : 0.0;
      }
      else
      {
	 const int lj = lj0 - 2 + ( t - PX );
	 c.s_sy[t - PX] = lj < b.nj ? a.stry[lj] : 0.0;
      }
   }
   c.p0 = c.ka - 2;
   for( int t = c.tid; t <= c.kb + 3 - c.p0; t += NT )
   {
      int kp = c.p0 + t - b.kfirst;
      kp = kp > b.nk - 1 ? b.nk - 1 : kp;
      c.s_sz[t] = a.strz[kp];
   }
   c.o = ( c.ty + 2 ) * PX + c.tx + 2; // own point in a staged plane
   const int li = li0 + c.tx, lj = lj0 + c.ty;
   c.act = li <= b.nil - 3 && lj <= b.nj - 3;
   c.gown = (long long)lj * b.ni + li; // own offset inside a plane (used only if act)

   fast2::State s;
#pragma unroll
   for( int m = 0; m < 6; m++ ) s.cu[m] = s.cv[m] = s.cw[m] = s.amz[m] = s.alz[m] = s.g1[m] = s.g2[m] = s.g3[m] = 0;
#pragma unroll
   for( int m = 0; m < 3; m++ )
   {
      s.pr[m][0] = s.pr[m][1] = s.pr[m][2] = 0;
      s.dyv[m] = s.dyu[m] = s.dxv[m] = s.dxu[m] = s.muk[m] = s.lak[m] = 0;
      s.rp[m] = 0;
   }
   int p = c.ka - 2;
   const int plast = c.kb + 3; // one extra step finishes plane kb
   {
      int ph0 = p % 6;
      ph0 = ph0 < 0 ? ph0 + 6 : ph0;
      fast2::stage<TY, EPI>( a, c, p, ph0, ph0 & 1 );
   }
   __syncthreads(); // s_sx, s_sy visible
   c.sx = c.s_sx[c.tx + 2]; c.sy = c.s_sy[c.ty + 2];
   c.sxm2 = c.s_sx[c.tx]; c.sxm1 = c.s_sx[c.tx + 1]; c.sxp1 = c.s_sx[c.tx + 3]; c.sxp2 = c.s_sx[c.tx + 4];
   c.sym2 = c.s_sy[c.ty]; c.sym1 = c.s_sy[c.ty + 1]; c.syp1 = c.s_sy[c.ty + 3]; c.syp2 = c.s_sy[c.ty + 4];

   // Plane p is always handled by the copy of the step with S = p mod 6 (p = global k index), whatever rows the
   // launch covers: a z-slab run then executes, for every plane, exactly the instruction sequence of the
   // undivided run and stays bit-identical to it.  (Slot 0 was used for the first plane: renumber.)
   int ph = p % 6;
   ph = ph < 0 ? ph + 6 : ph;
   for( ;; )
   {
      switch( ph )
      {
      case 0: fast2::step<0, TY, EPI>( a, c, s, p ); if( ++p > plast ) return;
      case 1: fast2::step<1, TY, EPI>( a, c, s, p ); if( ++p > plast ) return;
      case 2: fast2::step<2, TY, EPI>( a, c, s, p ); if( ++p > plast ) return;
      case 3: fast2::step<3, TY, EPI>( a, c, s, p ); if( ++p > plast ) return;
      case 4: fast2::step<4, TY, EPI>( a, c, s, p ); if( ++p > plast ) return;
      default: fast2::step<5, TY, EPI>( a, c, s, p ); if( ++p > plast ) return;
      }
      ph = 0;
   }
}

#ifndef SW4B200_EMULATE
namespace {
template <int TY, int EPI>
int launch_fast2_t( const FastArgs& a, cudaStream_t st )
{
   typedef fast2::Cfg<TY> C;
   static bool configured = false;
   const size_t smem = C::SMEM_DOUBLES * sizeof( double );
   if( !configured )
   {
      cudaError_t e = cudaFuncSetAttribute( k_rhs_fast2<TY, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem );
      if( e != cudaSuccess ) return set_error( "k_rhs_fast2: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString( e ) );
      configured = true;
   }
   const Block& b = a.b;
   dim3 bs( C::TX, TY, 1 );
   dim3 gs( ( b.nil - 4 + C::TX - 1 ) / C::TX, ( b.nj - 4 + TY - 1 ) / TY, ( a.khi - a.klo + 1 + a.kchunk - 1 ) / a.kchunk );
   ProfScope prof( EPI == EPI_PRED ? "rhs_fast2_pred" : ( EPI == EPI_CORR ? "rhs_fast2_corr" : "rhs_fast2_lu" ), st );
   k_rhs_fast2<TY, EPI><<<gs, bs, smem, st>>>( a );
   count_launch();
   return check_launch( "k_rhs_fast2" );
}
} // namespace

int launch_fast2( int epi, FastArgs a, cudaStream_t st )
{
   if( a.khi < a.klo ) return 0;
   if( a.kchunk <= 0 ) a.kchunk = fast_kchunk( a.b, a.khi - a.klo + 1, 8 );
   if( a.kchunk > fast2::Cfg<8>::SZMAX - 6 ) a.kchunk = fast2::Cfg<8>::SZMAX - 6;
   switch( epi )
   {
   case EPI_LU: return launch_fast2_t<8, EPI_LU>( a, st );
   case EPI_PRED: return launch_fast2_t<8, EPI_PRED>( a, st );
   default: return launch_fast2_t<8, EPI_CORR>( a, st );
   }
}
#endif

} // namespace sw4b200
