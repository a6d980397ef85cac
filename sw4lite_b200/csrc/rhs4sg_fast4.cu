// Throughput path for the Cartesian interior rows, fourth generation (SoA layout, corder=1):
// register blocking over x-pairs with the z state of both points in tensor memory.
//
// Same operator algebra and one-barrier-per-plane z-march as rhs4sg_fast2.cu (reference formulas
// rhs4sg_rev.C:112-348 with the common subexpressions shared between threads).  ncu on those kernels
// (profiles/r01c_fast3_ncu.md) showed a kernel that waits: issue slots 43 %, fp64 pipe 40 %, shared-memory pipe 67 %,
// 8 warps per SM; a third of the instructions were integer/address work, and 12 warps only moved the stall to the
// memory-instruction queue.  The cure for a latency-bound stencil is more independent work per thread and fewer
// memory instructions per point, i.e. register blocking -- which the z rings and delay lines (70 doubles per point)
// made impossible.  With those in tensor memory (rhs4sg_fast3.cu) it fits:
//   * every thread owns TWO x-adjacent points of a 32x16 tile (256 threads, 16 pairs per row, 2 rows per warp);
//   * every shared-memory access is a 16-byte LDS.128/STS.128 on an aligned pair: the six x values i-2..i+3 of a field
//     serve both points (3 loads instead of 10), y neighbours and the exchanged cross products come as pairs;
//     62 shared-memory instructions per pair and plane instead of 2 x 103;
//   * x-direction coefficient products (mu sx, (2mu+la) sx at i-2..i+3) are shared by the pair;
//   * per plane and point 10 doubles go to tensor memory (tcgen05.st) when plane p arrives and come back
//     (tcgen05.ld) when plane p-2 is finished: g1,g2,g3 (read at p+1,p+3,p+4), the in-plane sums pr[3] and the
//     in-plane parts of the exchanged products la sy D0y v, mu D0y u, mu D0x v, la sx D0x u.  Registers keep only
//     the five-plane rings of u,v,w, mu sz, (2mu+la) sz of the two points;
//   * mu and la are needed in shared memory only for the plane that arrives (2 slots instead of 6; the ring points
//     of the tile keep a 3-deep side copy), which is what lets a 32x16 tile fit in 227 KB.
//   * planes arrive by TMA tile loads (cp.async.bulk.tensor.3d on 3-D tensor maps of the arrays: one request per
//     field and plane, issued by one thread, completing on an mbarrier per plane parity, out-of-array elements
//     zero-filled): no LDGSTS, no per-thread address arithmetic;
//   * the march has two step bodies (even / odd planes) over six-deep register arrays shifted by two: half the
//     register moves of a shift register, and a third of the code of the six-fold unrolled ring (which overflowed
//     the instruction cache).
// Measured history and what was tried on top: profiles/r01d_k_rhs_fast4_ncu.md.
// sz is folded into the z-type exchanged products (E3 = mu sz D0z u, E6 = mu sz D0z v) and la sz = (2mu+la)sz - 2 mu sz
// comes from the register rings, so mu, la themselves need no delay line.
//
// Tensor-memory layout: 8 warps, warp w owns lanes 32 (w%4)..+31 and columns 256 (w/4)..+255; a thread's record of
// plane p sits in ring slot p mod 6, 40 columns (A,B = the two points):
//    0-11 g1A g1B g2A g2B g3A g3B | 12-39 pr0A pr0B pr1A pr1B pr2A pr2B e1A e1B e2A e2B e4A e4B e5A e5B
//
// Also compiled by g++ (SW4B200_EMULATE) for the CPU check of the kernel source (tests/emu).
#include "fast_common.cuh"
#include "tmem.cuh"
#include "tma.cuh"
#ifndef SW4B200_EMULATE
#include <mutex>
#endif
#include <type_traits>

namespace sw4b200 {

#if defined( SW4B200_EMULATE )
#define F4SM( c ) ( ( c ).sm )
#else
extern __shared__ __align__( 128 ) double smem_f4[];
#define F4SM( c ) smem_f4
#endif

#if !defined( SW4B200_EMULATE )
// thread blocks of k_rhs_fast4 that took the march without / with stretching factors (sw4b200_profile_read "tiles_plain",
// "tiles_general"; one atomic per thread block)
__device__ unsigned long long g_f4_tiles[2] = { 0, 0 };
int read_f4_tiles( long long out[2], bool reset )
{
   unsigned long long v[2] = { 0, 0 };
   if( cudaMemcpyFromSymbol( v, g_f4_tiles, sizeof( v ) ) != cudaSuccess ) return set_error( "reading the tile counters failed" );
   out[0] = (long long)v[0]; out[1] = (long long)v[1];
   if( reset )
   {
      const unsigned long long z[2] = { 0, 0 };
      if( cudaMemcpyToSymbol( g_f4_tiles, z, sizeof( z ) ) != cudaSuccess ) return set_error( "resetting the tile counters failed" );
   }
   return 0;
}
#endif
namespace fast4 {
// DER (template parameter of the kernel): the staged "lambda" plane and the "rho" operand hold the derived, time-invariant
// arrays 2 mu + lambda and 1 / rho of a grid block (FastArgs::la2, rhoi).  2 mu + lambda is then read instead of formed at
// every use (3 + 5 + 1 times per point), lambda itself is recovered where the operator needs it (own point: one fma; 2 mu is
// exact, so lambda comes back to within one rounding of 2 mu + lambda), and the epilogue multiplies by the stored reciprocal.
template <bool DER>
__device__ __forceinline__ double bsum( double m, double l ) { return DER ? l : 2 * m + l; } // 2 mu + lambda

using fast::W4;
using fast::weights4;
using fast::gsum;
using fast::d0u;
using tmem::Tm;
using tmem::TmVal;
using tmem::tm_get;

#if defined( SW4B200_EMULATE )
struct D2 { double x, y; };
#else
typedef double2 D2;
#endif
// 1/x without the division's slow-path branch: single-precision seed and two Newton steps (error < 1 ulp of the
// double result for x in the normal range of a float -- x is a density)
__device__ __forceinline__ double rcp_nr( double x )
{
#if defined( __CUDA_ARCH__ )
   double y = (double)__frcp_rn( (float)x );
   double e = fma( -x, y, 1.0 );
   y = fma( y, e, y );
   e = fma( -x, y, 1.0 );
   return fma( y, e, y );
#else
   return 1.0 / x;
#endif
}
__device__ __forceinline__ D2 ld2( const double* p ) { return *reinterpret_cast<const D2*>( p ); }
__device__ __forceinline__ void st2( double* p, double x, double y )
{
   D2 v;
   v.x = x; v.y = y;
   *reinterpret_cast<D2*>( p ) = v;
}

template <int TY>
struct Cfg
{
   static constexpr int TXP = 16, TX = 2 * TXP, PX = TX + 4, PY = TY + 4, PLANE = PX * PY, NT = TXP * TY, NSLOT = 6;
   static constexpr int NH = 4 * TY + 4 * TX; // ring points of the tile per plane
   static constexpr int EX = 3 * TY * PX, EY = 3 * PY * TX;
   static constexpr int OPS = 4 * TX * TY; // epilogue operands of one plane: rho, um[3] of the own points
   static constexpr int SZMAX = 512;	   // strz of the planes a CTA marches through (kchunk + 6 <= SZMAX)
   // shared memory, in doubles (every region starts on a 16-byte boundary)
   static constexpr int O_UVW = 0;			      // [3][NSLOT][PLANE]
   static constexpr int O_ML = O_UVW + 3 * NSLOT * PLANE;     // [2][2][PLANE]   mu, la of the arriving plane
   static constexpr int O_OP = O_ML + 4 * PLANE;	      // [2][4][TX*TY]   rho, um of the own points (TMA destinations first:
   static constexpr int O_EX = O_OP + 2 * OPS;		      // [2][3][TY][PX]  E1..E3, double buffered        128-byte aligned)
   static constexpr int O_EY = O_EX + 2 * EX;		      // [2][3][PY][TX]  E4..E6
   static constexpr int O_HML = O_EY + 2 * EY;		      // [3][2][NH]      mu, la of the ring points, planes p-2..p
   static constexpr int O_SX = O_HML + 6 * NH;		      // [PX]
   static constexpr int O_SY = O_SX + PX;		      // [PY]
   static constexpr int O_SZ = O_SY + PY;		      // [SZMAX]
   static constexpr int O_MBAR = O_SZ + SZMAX;		      // two mbarriers (even / odd planes), 16 bytes apart
   static constexpr int SMEM_DOUBLES = O_MBAR + 4 + 2;	      // + the tensor-memory base address
   static_assert( ( PLANE % 16 ) == 0 && ( O_ML % 16 ) == 0 && ( O_OP % 16 ) == 0 && ( ( TX * TY ) % 16 ) == 0, "TMA destinations: 128-byte aligned" );
   static constexpr int REC = 40;			      // tensor-memory columns per plane record
   static constexpr int COLS = 256;			      // columns per warp of a lane quadrant (8 warps)
   static_assert( NT == 256, "8 warps: two per tensor-memory lane quadrant" );
   static_assert( NSLOT * REC <= COLS, "tensor-memory strip too small" );
   static_assert( ( PLANE % 2 ) == 0 && ( EX % 2 ) == 0 && ( EY % 2 ) == 0 && ( NH % 2 ) == 0 && ( ( PX + PY ) % 2 ) == 0, "16-byte alignment" );
};

// per-thread register state ([..][2]: the two points): five-plane shift registers in arrays of six.  The march has TWO
// step bodies: planes with even k use positions [4]..[0] for planes p..p-4, planes with odd k [5]..[1], and the
// registers are shifted by two after the odd one: 20 double moves per step instead of 40.  (The rings of the earlier
// generations needed a step copy per ring position, 6 x 1600 instructions: more than the instruction cache holds.)
// A plane is always handled by the body of its parity, so z-slab runs execute the undivided run's instruction
// sequence for every plane and stay bit-identical to it.
struct State
{
   double cu[6][2], cv[6][2], cw[6][2], amz[6][2], alz[6][2];
   double rp[3][2]; // result of the previous plane, still lacking the exchanged cross terms
};

template <int TY>
struct Ctx
{
   typedef Cfg<TY> C;
   double* sm; // (device code addresses the shared array by name, F4SM: no generic-to-shared conversions)
   int p0;
   int li0, lj0; // local (array) index of the tile's first output
   int tid, txh, ty, o;
   int ka, kb, pend;
   bool act;  // the left point of the pair is inside the interior
   bool act2; // SPLIT only (rows padded to an even pitch, odd number of points per row): the right point is inside too; without
	      // padding pairs start at even i and the interior ends at an odd i, so a pair is never split by the boundary
   long long gown; // offset of the left point inside a plane
};

// run-time (CTA-uniform) offsets of one step: [j] = plane p-j
struct Ph
{
   int slot;   // p mod 6
   int o[5];   // PLANE * slot of plane p-j in the u,v,w ring
   int c[5];   // REC * slot: tensor-memory column of the record of plane p-j
   int par;    // p & 1: E buffer written, mu/la slot, operand buffer
   int wpar;   // bit b: phase parity to wait for on mbarrier b
   int t0, t2; // 2 NH (p mod 3), 2 NH ((p-2) mod 3): side copies of the ring points' mu, la
};

// stage plane p (one thread): TMA tile loads completing on the mbarrier of the plane's parity: the PX x PY boxes of u,v,w
// into ring slot `slot`, of mu,la into slot `par`; with an epilogue also the TX x TY boxes of the own-point operands
// (rho, um) of plane p-3, the plane that the step handling plane p finishes, into operand buffer `par`.
// Requires ni even and 16-byte aligned arrays (tensor-map strides; launch_fast4 checks, other grids take the cp.async
// kernel rhs4sg_fast2.cu).
template <int TY, int EPI>
__device__ __forceinline__ void stage( const FastArgs& a, const FastMaps& maps, const Ctx<TY>& c, int p, int slot, int par )
{
   typedef Cfg<TY> C;
   if( p > c.pend + 1 || c.tid != 0 ) return;
   const Block& b = a.b;
   double* const mbar = F4SM( c ) + C::O_MBAR + 2 * par;
   const int kq = p - 3;
   const bool plane = p <= c.pend; // (the step after the last plane only finishes plane kb)
   const bool ops = EPI != EPI_LU && kq >= c.ka && kq <= c.kb;
#if !defined( SW4B200_EMULATE )
   mbar_arrive_expect( mbar, ( plane ? 5 * C::PLANE * 8 : 0 ) + ( ops ? 4 * C::TX * TY * 8 : 0 ) );
#endif
   if( plane )
   {
      const int c0 = c.li0 - 2, c1 = c.lj0 - 2, c2 = p - b.kfirst;
#pragma unroll
      for( int f = 0; f < 3; f++ )
	 tma_tile<C::PX, C::PY>( F4SM( c ) + C::O_UVW + ( f * C::NSLOT + slot ) * C::PLANE, &maps.u[f], b, c0, c1, c2, mbar );
      tma_tile<C::PX, C::PY>( F4SM( c ) + C::O_ML + ( 0 * 2 + par ) * C::PLANE, &maps.mu, b, c0, c1, c2, mbar );
      tma_tile<C::PX, C::PY>( F4SM( c ) + C::O_ML + ( 1 * 2 + par ) * C::PLANE, &maps.la, b, c0, c1, c2, mbar );
   }
   if( ops )
   {
      double* const d = F4SM( c ) + C::O_OP + par * C::OPS;
      tma_tile<C::TX, TY>( d, &maps.rho, b, c.li0, c.lj0, kq - b.kfirst, mbar );
#pragma unroll
      for( int m = 0; m < 3; m++ ) tma_tile<C::TX, TY>( d + ( m + 1 ) * C::TX * TY, &maps.um[m], b, c.li0, c.lj0, kq - b.kfirst, mbar );
   }
#if defined( SW4B200_EMULATE )
   mbar_arrive_expect( mbar, 0 ); // (after the synchronous copies)
#endif
}

__device__ __forceinline__ double pick( const D2& v, int t ) { return t ? v.y : v.x; }

// One step of the march: plane p has been staged into ring slot S.  Does the in-plane work of plane p,
// the z work of plane k=p-2 (publishing its exchanged products in E buffer S&1) and finishes plane k-1
// (reading E buffer (S+1)&1).
// NOSTR: strx = 1 and stry = 1 on the whole tile (every tile away from the supergrid layers, nine tenths of a production grid):
// the stretching factors are compile-time ones and their 22 multiplications per point disappear.  x * 1.0 and fma( x, 1.0, y )
// are exact, so this body returns the bits of the general one.
template <int TY, int EPI, int SPLIT, bool DER, bool NOSTR, int H>
__device__ __forceinline__ void step( const FastArgs& a, const FastMaps& maps, Ctx<TY>& c, State& s, Tm& tm, const int p, const Ph& ph )
{
   typedef Cfg<TY> C;
   constexpr int TX = C::TX, PX = C::PX, PY = C::PY, PLANE = C::PLANE, NT = C::NT, NSLOT = C::NSLOT, NH = C::NH;
   // shift-register positions of planes p, p-1, ..., p-4
   constexpr int R0 = 4 + H, R1 = 3 + H, R2 = 2 + H, R3 = 1 + H, R4 = H;
   const int EB = ph.par, EF = ph.par ^ 1, ML = ph.par;
   const Block& b = a.b;
   const int k = p - 2, kf = p - 3;

   mbar_wait( F4SM( c ) + C::O_MBAR + 2 * ph.par, ( ph.wpar >> ph.par ) & 1 ); // the boxes of plane p have landed
   __syncthreads(); // the E products of plane k-1 are visible; slot of plane p-5 is free
   stage<TY, EPI>( a, maps, c, p + 1, ph.slot == NSLOT - 1 ? 0 : ph.slot + 1, ph.par ^ 1 );
   tm.wait_st(); // the records stored by the earlier steps (long done) are readable

   const bool kfin = kf >= c.ka && kf <= c.kb;
   const bool fin = c.act && kfin;
   const long long qf = kfin ? b.nij * ( kf - b.kfirst ) + c.gown : 0;

   // strx at i-2..i+3 (i = left point), stry at j-2..j+2: re-read every step instead of held in 22 registers
   double csx[6], csy[5];
   if( NOSTR )
   {
#pragma unroll
      for( int j = 0; j < 6; j++ ) csx[j] = 1.0;
#pragma unroll
      for( int j = 0; j < 5; j++ ) csy[j] = 1.0;
   }
   else
   {
      const D2 s0 = ld2( F4SM( c ) + C::O_SX + 2 * c.txh ), s1 = ld2( F4SM( c ) + C::O_SX + 2 * c.txh + 2 ), s2 = ld2( F4SM( c ) + C::O_SX + 2 * c.txh + 4 );
      csx[0] = s0.x; csx[1] = s0.y; csx[2] = s1.x; csx[3] = s1.y; csx[4] = s2.x; csx[5] = s2.y;
#pragma unroll
      for( int j = 0; j < 5; j++ ) csy[j] = F4SM( c )[C::O_SY + c.ty + j];
   }
   const double syo = csy[2];
   const double sxo[2] = { csx[2], csx[3] };
   // ---- finish plane kf = k-1: difference its exchanged products (published by this step's barrier), one component
   // at a time (7 pair loads in flight)
   auto finish = [&]() {
      const double* const ex = F4SM( c ) + C::O_EX + EF * C::EX + c.ty * PX + 2 * c.txh + 2;
      const double* const ey = F4SM( c ) + C::O_EY + EF * C::EY + ( c.ty + 2 ) * TX + 2 * c.txh;
      const double* const d = F4SM( c ) + C::O_OP + ph.par * C::OPS + 2 * c.tid;
      double fr[2] = { 0, 0 }, rinv[2] = { 0, 0 }; // PRED: dt^2/rho, 1/rho ; CORR: dt^4/(12 rho)
      if( EPI != EPI_LU )
      {
	 const D2 e_rho = ld2( d );
#pragma unroll
	 for( int t = 0; t < 2; t++ )
	 {
	    const double rh = fin ? pick( e_rho, t ) : 1.0;
	    if( EPI == EPI_PRED )
	    {
	       rinv[t] = DER ? rh : rcp_nr( rh ); // one reciprocal per point; dt^2/rho and acc/rho are formed from it
	       fr[t] = a.fac * rinv[t];
	    }
	    else
	       fr[t] = a.fac * ( DER ? rh : rcp_nr( rh ) );
	 }
      }
      constexpr int RF = R3; // plane kf = p-3
      const double* const ring[3] = { &s.cu[RF][0], &s.cv[RF][0], &s.cw[RF][0] };
#pragma unroll
      for( int m = 0; m < 3; m++ )
      {
	 const D2 l = ld2( ex + m * TY * PX - 2 ), o = ld2( ex + m * TY * PX ), r = ld2( ex + m * TY * PX + 2 );
	 const D2 m2 = ld2( ey + m * PY * TX - 2 * TX ), m1 = ld2( ey + m * PY * TX - TX ), p1 = ld2( ey + m * PY * TX + TX ),
		  p2 = ld2( ey + m * PY * TX + 2 * TX );
	 const double xd[2] = { d0u( l.x, l.y, o.y, r.x ), d0u( l.y, o.x, r.x, r.y ) };
	 const double yd[2] = { d0u( m2.x, m1.x, p1.x, p2.x ), d0u( m2.y, m1.y, p1.y, p2.y ) };
	 D2 e_um;
	 e_um.x = e_um.y = 0;
	 if( EPI != EPI_LU ) e_um = ld2( d + ( m + 1 ) * TX * TY );
	 double o1[2], o2[2] = { 0, 0 };
#pragma unroll
	 for( int t = 0; t < 2; t++ )
	 {
	    double r_;
	    if( m == 0 ) r_ = s.rp[0][t] + ( a.cof144 * sxo[t] ) * ( xd[t] + syo * yd[t] );
	    else if( m == 1 ) r_ = s.rp[1][t] + ( a.cof144 * syo ) * ( sxo[t] * xd[t] + yd[t] );
	    else r_ = s.rp[2][t] + a.cof144 * ( sxo[t] * xd[t] + syo * yd[t] );
	    // (dense forcing and a predictor without the acceleration output take the cp.async kernel: launch_fast4)
	    if( EPI == EPI_LU )
	       o1[t] = r_;
	    else if( EPI == EPI_PRED )
	    {
	       o1[t] = 2 * ring[m][t] - pick( e_um, t ) + fr[t] * r_;
	       o2[t] = r_ * rinv[t];
	    }
	    else
	       o1[t] = pick( e_um, t ) + fr[t] * r_;
	 }
	 // the pair is 16-byte aligned in global memory too (even ni, even li): one store per array
	 if( fin )
	 {
	    if( SPLIT && !c.act2 )
	    {
	       a.out[m][qf] = o1[0];
	       if( EPI == EPI_PRED ) a.out2[m][qf] = o2[0];
	    }
	    else
	    {
	       st2( a.out[m] + qf, o1[0], o1[1] );
	       if( EPI == EPI_PRED ) st2( a.out2[m] + qf, o2[0], o2[1] );
	    }
	 }
      }
   };
   double g1n[2], g2n[2], g3n[2]; // g products of plane p
   // ---- in-plane pieces of plane p: an x pass and a y pass over the fields, so that only the weights of one
   // direction (2 points x 2 coefficient sets) are live next to the neighbours of one field
   {
      double* const sm = F4SM( c );
      const double* const pf[3] = { sm + C::O_UVW + 0 * NSLOT * PLANE + ph.o[0] + c.o, sm + C::O_UVW + 1 * NSLOT * PLANE + ph.o[0] + c.o,
				    sm + C::O_UVW + 2 * NSLOT * PLANE + ph.o[0] + c.o };
      const double* const pm = sm + C::O_ML + ( 0 * 2 + ML ) * PLANE + c.o;
      const double* const pl = sm + C::O_ML + ( 1 * 2 + ML ) * PLANE + c.o;
      const double szp = sm[C::O_SZ + p - c.p0];
      double rec[14]; // pr0A pr0B pr1A pr1B pr2A pr2B e1A e1B e2A e2B e4A e4B e5A e5B
      double m0[2], l0[2], b0[2], q0[3][2], dx[3][2], dy[3][2];
      {
	 // x-direction coefficients mu sx, (2mu+la) sx at i-2..i+3, shared by the two points
	 double axm[6], axl[6];
	 const D2 ma = ld2( pm - 2 ), mb = ld2( pm ), mc = ld2( pm + 2 ), la = ld2( pl - 2 ), lb = ld2( pl ), lc = ld2( pl + 2 );
	 D2 xa[3], xb[3], xc[3];
#pragma unroll
	 for( int f = 0; f < 3; f++ ) { xa[f] = ld2( pf[f] - 2 ); xb[f] = ld2( pf[f] ); xc[f] = ld2( pf[f] + 2 ); }
	 {
	    m0[0] = mb.x; m0[1] = mb.y;
	    b0[0] = bsum<DER>( mb.x, lb.x ); b0[1] = bsum<DER>( mb.y, lb.y );
	    l0[0] = DER ? fma( -2.0, mb.x, lb.x ) : lb.x; l0[1] = DER ? fma( -2.0, mb.y, lb.y ) : lb.y;
	    axm[0] = ma.x * csx[0]; axm[1] = ma.y * csx[1]; axm[2] = mb.x * csx[2]; axm[3] = mb.y * csx[3]; axm[4] = mc.x * csx[4]; axm[5] = mc.y * csx[5];
	    axl[0] = bsum<DER>( ma.x, la.x ) * csx[0]; axl[1] = bsum<DER>( ma.y, la.y ) * csx[1]; axl[2] = b0[0] * csx[2];
	    axl[3] = b0[1] * csx[3]; axl[4] = bsum<DER>( mc.x, lc.x ) * csx[4]; axl[5] = bsum<DER>( mc.y, lc.y ) * csx[5];
	 }
	 W4 wmx[2], wlx[2];
#pragma unroll
	 for( int t = 0; t < 2; t++ )
	 {
	    wmx[t] = weights4( axm[t], axm[t + 1], axm[t + 2], axm[t + 3], axm[t + 4] );
	    wlx[t] = weights4( axl[t], axl[t + 1], axl[t + 2], axl[t + 3], axl[t + 4] );
	 }
#pragma unroll
	 for( int f = 0; f < 3; f++ )
	 {
	    const double x[6] = { xa[f].x, xa[f].y, xb[f].x, xb[f].y, xc[f].x, xc[f].y };
#pragma unroll
	    for( int t = 0; t < 2; t++ )
	    {
	       q0[f][t] = x[2 + t];
	       dx[f][t] = d0u( x[t], x[t + 1], x[t + 3], x[t + 4] );
	       rec[2 * f + t] = sxo[t] * gsum( f == 0 ? wlx[t] : wmx[t], x[t], x[t + 1], x[2 + t], x[t + 3], x[t + 4] );
	    }
	 }
      }
#pragma unroll
      for( int t = 0; t < 2; t++ )
      {
	 s.cu[R0][t] = q0[0][t]; s.cv[R0][t] = q0[1][t]; s.cw[R0][t] = q0[2][t];
	 s.amz[R0][t] = m0[t] * szp; s.alz[R0][t] = b0[t] * szp;
      }
      {
	 W4 wmy[2], wly[2];
	 {
	    const D2 ma = ld2( pm - 2 * PX ), mb = ld2( pm - PX ), mc = ld2( pm + PX ), md = ld2( pm + 2 * PX );
	    const D2 la = ld2( pl - 2 * PX ), lb = ld2( pl - PX ), lc = ld2( pl + PX ), ld = ld2( pl + 2 * PX );
#pragma unroll
	    for( int t = 0; t < 2; t++ )
	    {
	       const double mym2 = pick( ma, t ), mym1 = pick( mb, t ), myp1 = pick( mc, t ), myp2 = pick( md, t );
	       const double lym2 = pick( la, t ), lym1 = pick( lb, t ), lyp1 = pick( lc, t ), lyp2 = pick( ld, t );
	       wmy[t] = weights4( mym2 * csy[0], mym1 * csy[1], m0[t] * syo, myp1 * csy[3], myp2 * csy[4] );
	       wly[t] = weights4( bsum<DER>( mym2, lym2 ) * csy[0], bsum<DER>( mym1, lym1 ) * csy[1], b0[t] * syo, bsum<DER>( myp1, lyp1 ) * csy[3],
				  bsum<DER>( myp2, lyp2 ) * csy[4] );
	    }
	 }
#pragma unroll
	 for( int f = 0; f < 3; f++ )
	 {
	    const D2 ya = ld2( pf[f] - 2 * PX ), yb = ld2( pf[f] - PX ), yc = ld2( pf[f] + PX ), yd = ld2( pf[f] + 2 * PX );
#pragma unroll
	    for( int t = 0; t < 2; t++ )
	    {
	       const double ym2 = pick( ya, t ), ym1 = pick( yb, t ), yp1 = pick( yc, t ), yp2 = pick( yd, t );
	       dy[f][t] = d0u( ym2, ym1, yp1, yp2 );
	       rec[2 * f + t] += syo * gsum( f == 1 ? wly[t] : wmy[t], ym2, ym1, q0[f][t], yp1, yp2 );
	    }
	 }
      }
#pragma unroll
      for( int t = 0; t < 2; t++ )
      {
	 const double e1p = ( l0[t] * syo ) * dy[1][t];	   // la sy D0y v   (in-plane part of E1)
	 const double e5p = ( l0[t] * sxo[t] ) * dx[0][t]; // la sx D0x u   (in-plane part of E5)
	 g1n[t] = m0[t] * dx[2][t];
	 g2n[t] = m0[t] * dy[2][t];
	 g3n[t] = e1p + e5p;
	 rec[6 + t] = e1p;
	 rec[8 + t] = m0[t] * dy[0][t];	 // E2 = mu D0y u
	 rec[10 + t] = m0[t] * dx[1][t]; // E4 = mu D0x v
	 rec[12 + t] = e5p;
      }
      // pair-wise stores (two doubles of a pair are neighbours in registers anyway; 16-register operands had to be packed
      // with 40 moves per step)
      tm.template st<0, 2>( g1n, ph.c[0] ); tm.template st<4, 2>( g2n, ph.c[0] ); tm.template st<8, 2>( g3n, ph.c[0] );
#pragma unroll
      for( int m = 0; m < 7; m++ ) tm.st2i( 12 + 4 * m, rec + 2 * m, ph.c[0] );
   }

   // ---- tensor-memory loads for the z work of plane k (the records were retired by the wait at the top of the
   // step): issued here, consumed after the finish phase, so that their latency -- all eight warps ask at the
   // same time -- is covered by the finish phase
   TmVal tg[18], tq[14];
   tm.template ld<0, 4>( tg, ph.c[4] ); tm.template ld<8, 2>( tg + 4, ph.c[4] );	// g1 g2 | g3 of plane k-2
   tm.template ld<0, 4>( tg + 6, ph.c[3] ); tm.template ld<8, 2>( tg + 10, ph.c[3] );	// k-1
   tm.template ld<0, 4>( tg + 12, ph.c[1] ); tm.template ld<8, 2>( tg + 16, ph.c[1] ); // k+1
   tm.template ld<12, 2>( tq, ph.c[2] );	 // pr0A pr0B of plane k
   tm.template ld<16, 8>( tq + 2, ph.c[2] );	 // pr1 pr2 e1 e2
   tm.template ld<32, 4>( tq + 10, ph.c[2] ); // e4 e5

   const double szk = k >= c.p0 ? F4SM( c )[C::O_SZ + k - c.p0] : 0.0;
   auto helper = [&]() {
      // ring of width 2 around the tile: the same products recomputed from the staged planes
      // (plane k sits in slot R2, k-2..k+2 in R4..R0); mu, la of the ring points from the side copy
      for( int hh = c.tid; hh < NH; hh += NT )
      {
	 const double* const sf = F4SM( c ) + C::O_UVW;
	 double* const hml = F4SM( c ) + C::O_HML;
	 int sx_, sy_;
	 if( hh < 4 * TY )
	 {
	    const int hx = hh & 3, row = hh >> 2;
	    sx_ = hx < 2 ? hx : TX + hx; sy_ = row + 2;
	 }
	 else
	 {
	    const int t = hh - 4 * TY;
	    const int hy = t >> 5, col = t & 31;
	    sy_ = hy < 2 ? hy : TY + hy; sx_ = col + 2;
	 }
	 const int oo = sy_ * PX + sx_;
	 hml[ph.t0 + hh] = F4SM( c )[C::O_ML + ( 0 * 2 + ML ) * PLANE + oo];
	 hml[ph.t0 + NH + hh] = DER ? fma( -2.0, hml[ph.t0 + hh], F4SM( c )[C::O_ML + ( 1 * 2 + ML ) * PLANE + oo] )
				    : F4SM( c )[C::O_ML + ( 1 * 2 + ML ) * PLANE + oo];
	 const double hm = hml[ph.t2 + hh], hl = hml[ph.t2 + NH + hh];
	 const double* const qu = sf + 0 * NSLOT * PLANE + ph.o[2] + oo;
	 const double* const qv = sf + 1 * NSLOT * PLANE + ph.o[2] + oo;
	 const double hdzw = d0u( sf[2 * NSLOT * PLANE + ph.o[4] + oo], sf[2 * NSLOT * PLANE + ph.o[3] + oo],
				  sf[2 * NSLOT * PLANE + ph.o[1] + oo], sf[2 * NSLOT * PLANE + ph.o[0] + oo] );
	 if( hh < 4 * TY )
	 {
	    const double hdyv = d0u( qv[-2 * PX], qv[-PX], qv[PX], qv[2 * PX] );
	    const double hdyu = d0u( qu[-2 * PX], qu[-PX], qu[PX], qu[2 * PX] );
	    const double hdzu = d0u( sf[0 * NSLOT * PLANE + ph.o[4] + oo], sf[0 * NSLOT * PLANE + ph.o[3] + oo],
				     sf[0 * NSLOT * PLANE + ph.o[1] + oo], sf[0 * NSLOT * PLANE + ph.o[0] + oo] );
	    double* const hx_ = F4SM( c ) + C::O_EX + EB * C::EX + ( sy_ - 2 ) * PX + sx_;
	    hx_[0] = hl * ( ( NOSTR ? 1.0 : F4SM( c )[C::O_SY + sy_] ) * hdyv + szk * hdzw );
	    hx_[TY * PX] = hm * hdyu;
	    hx_[2 * TY * PX] = ( hm * szk ) * hdzu;
	 }
	 else
	 {
	    const double hdxv = d0u( qv[-2], qv[-1], qv[1], qv[2] );
	    const double hdxu = d0u( qu[-2], qu[-1], qu[1], qu[2] );
	    const double hdzv = d0u( sf[1 * NSLOT * PLANE + ph.o[4] + oo], sf[1 * NSLOT * PLANE + ph.o[3] + oo],
				     sf[1 * NSLOT * PLANE + ph.o[1] + oo], sf[1 * NSLOT * PLANE + ph.o[0] + oo] );
	    double* const hy_ = F4SM( c ) + C::O_EY + EB * C::EY + sy_ * TX + ( sx_ - 2 );
	    hy_[0] = hm * hdxv;
	    hy_[PY * TX] = hl * ( ( NOSTR ? 1.0 : F4SM( c )[C::O_SX + sx_] ) * hdxu + szk * hdzw );
	    hy_[2 * PY * TX] = ( hm * szk ) * hdzv;
	 }
      }
      };
   finish();
   // ---- z pieces of plane k and its exchanged products
   double rnew[3][2];
   {
      tm.template wait_ld<18>( tg );
      tm.template wait_ld<14>( tq );
      double t1[2], t2[2], t3[2];
#pragma unroll
      for( int t = 0; t < 2; t++ )
      {
	 t1[t] = d0u( tm_get( tg[0 + t] ), tm_get( tg[6 + t] ), tm_get( tg[12 + t] ), g1n[t] );
	 t2[t] = d0u( tm_get( tg[2 + t] ), tm_get( tg[8 + t] ), tm_get( tg[14 + t] ), g2n[t] );
	 t3[t] = d0u( tm_get( tg[4 + t] ), tm_get( tg[10 + t] ), tm_get( tg[16 + t] ), g3n[t] );
      }
      const double c144z = a.cof144 * szk;
      double e1[2], e2[2], e3[2], e4[2], e5[2], e6[2];
#pragma unroll
      for( int t = 0; t < 2; t++ )
      {
	 const double prk0 = tm_get( tq[0 + t] ), prk1 = tm_get( tq[2 + t] ), prk2 = tm_get( tq[4 + t] );
	 const double e1k = tm_get( tq[6 + t] ), e2k = tm_get( tq[8 + t] ), e4k = tm_get( tq[10 + t] ), e5k = tm_get( tq[12 + t] );
	 const W4 wmz = weights4( s.amz[R4][t], s.amz[R3][t], s.amz[R2][t], s.amz[R1][t], s.amz[R0][t] );
	 const W4 wlz = weights4( s.alz[R4][t], s.alz[R3][t], s.alz[R2][t], s.alz[R1][t], s.alz[R0][t] );
	 const double rz0 = prk0 + szk * gsum( wmz, s.cu[R4][t], s.cu[R3][t], s.cu[R2][t], s.cu[R1][t], s.cu[R0][t] );
	 const double rz1 = prk1 + szk * gsum( wmz, s.cv[R4][t], s.cv[R3][t], s.cv[R2][t], s.cv[R1][t], s.cv[R0][t] );
	 const double rz2 = prk2 + szk * gsum( wlz, s.cw[R4][t], s.cw[R3][t], s.cw[R2][t], s.cw[R1][t], s.cw[R0][t] );
	 const double dzu = d0u( s.cu[R4][t], s.cu[R3][t], s.cu[R1][t], s.cu[R0][t] );
	 const double dzv = d0u( s.cv[R4][t], s.cv[R3][t], s.cv[R1][t], s.cv[R0][t] );
	 const double dzw = d0u( s.cw[R4][t], s.cw[R3][t], s.cw[R1][t], s.cw[R0][t] );
	 const double msz = s.amz[R2][t];		   // mu sz of plane k
	 const double lzw = ( s.alz[R2][t] - 2 * msz ) * dzw; // la sz D0z w
	 e1[t] = e1k + lzw;
	 e2[t] = e2k;
	 e3[t] = msz * dzu;
	 e4[t] = e4k;
	 e5[t] = e5k + lzw;
	 e6[t] = msz * dzv;
	 rnew[0][t] = a.cof6 * rz0 + ( c144z * sxo[t] ) * t1[t];
	 rnew[1][t] = a.cof6 * rz1 + ( c144z * syo ) * t2[t];
	 rnew[2][t] = a.cof6 * rz2 + c144z * t3[t];
      }
      double* const ex = F4SM( c ) + C::O_EX + EB * C::EX + c.ty * PX + 2 * c.txh + 2;
      double* const ey = F4SM( c ) + C::O_EY + EB * C::EY + ( c.ty + 2 ) * TX + 2 * c.txh;
      st2( ex, e1[0], e1[1] );
      st2( ex + TY * PX, e2[0], e2[1] );
      st2( ex + 2 * TY * PX, e3[0], e3[1] );
      st2( ey, e4[0], e4[1] );
      st2( ey + PY * TX, e5[0], e5[1] );
      st2( ey + 2 * PY * TX, e6[0], e6[1] );

   }
   helper();

#pragma unroll
   for( int m = 0; m < 3; m++ ) { s.rp[m][0] = rnew[m][0]; s.rp[m][1] = rnew[m][1]; }
}

} // namespace fast4

// KIND: which tiles of the grid this launch serves -- 0: all, with the general march; 1: the plain tiles (strx = stry = 1 on
// the whole tile), with the march compiled without stretching factors; 2: the other tiles, with the general march.  A thread
// block of a tile of the other kind leaves at once.  The fused passes of a grid block are a launch of kind 2 followed by one
// of kind 1 (two kernels rather than one with both marches: each keeps its own register allocation -- both marches in one
// kernel cost the few spills that, with no L1 left beside 227 KB of shared memory, made the corrector pass 10 % slower).
template <int TY, int EPI, int SPLIT, bool DER, int KIND>
__global__ void __launch_bounds__( 16 * TY, 1 ) k_rhs_fast4( const FastArgs a, const SW4_GRID_CONSTANT FastMaps maps )
{
   using namespace fast4;
   typedef fast4::Cfg<TY> C;
   constexpr int TX = C::TX, PX = C::PX, PY = C::PY, PLANE = C::PLANE, NT = C::NT;
#if defined( SW4B200_EMULATE )
   SW4_DYN_SMEM( smem );
#else
   double* const smem = smem_f4;
#endif
   fast4::Ctx<TY> c;
   c.sm = smem;

   const Block& b = a.b;
   c.tid = threadIdx.x;
   asm volatile( "" : "+r"( c.tid ) ); // (keeps ptxas from re-reading SR_TID in every step: a long-scoreboard wait)
   c.txh = c.tid & 15; c.ty = c.tid >> 4;
   const int li0 = 2 + blockIdx.x * TX, lj0 = 2 + blockIdx.y * TY; // local (array) index of the tile's first output
   c.li0 = li0; c.lj0 = lj0;
   c.ka = a.klo + blockIdx.z * a.kchunk;
   c.kb = ( c.ka + a.kchunk - 1 < a.khi ) ? c.ka + a.kchunk - 1 : a.khi;
   if( c.ka > c.kb ) return;
   c.pend = c.kb + 2;

   bool plain = true; // no stretching on this tile: strx = stry = 1 at all its (in-array) points
   for( int t = c.tid; t < PX + PY; t += NT )
   {
      if( t < PX )
      {
	 const int li = li0 - 2 + t;
	 const double v = li < b.ni ? a.strx[li] : 0.0;
	 smem[C::O_SX + t] = v;
	 plain = plain && ( li >= b.nil || v == 1.0 ); // (nil: points per row; the pad column of a padded block does not count)
      }
      else
      {
	 const int lj = lj0 - 2 + ( t - PX );
	 const double v = lj < b.nj ? a.stry[lj] : 0.0;
	 smem[C::O_SY + t - PX] = v;
	 plain = plain && ( lj >= b.nj || v == 1.0 );
      }
   }
   if( KIND != 0 )
   {
      const bool nostr = __syncthreads_and( plain ) != 0;
      if( nostr != ( KIND == 1 ) ) return; // a tile of the other launch
#if defined( SW4B200_EMULATE )
      if( KIND == 1 && c.tid == 0 ) emu::g_nostr_ctas++;
#else
      if( c.tid == 0 ) atomicAdd( &g_f4_tiles[KIND == 1 ? 0 : 1], 1ULL );
#endif
   }

   fast4::Tm tm;
   if( c.tid == 0 )
   {
      mbar_init( smem + C::O_MBAR, 1 );
      mbar_init( smem + C::O_MBAR + 2, 1 );
#if !defined( SW4B200_EMULATE )
      asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );
#endif
   }
#if !defined( SW4B200_EMULATE )
   // all 512 columns of tensor memory: one CTA per SM (shared memory), so nobody else can want them
   uint32_t* const tm_slot = reinterpret_cast<uint32_t*>( smem + C::O_MBAR + 4 );
   if( c.tid < 32 )
   {
      asm volatile( "tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"( (uint32_t)__cvta_generic_to_shared( tm_slot ) )
		    : "memory" );
      asm volatile( "tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory" );
   }
#endif

   c.p0 = c.ka - 2;
   for( int t = c.tid; t <= c.kb + 3 - c.p0; t += NT )
   {
      int kp = c.p0 + t - b.kfirst;
      kp = kp > b.nk - 1 ? b.nk - 1 : kp;
      smem[C::O_SZ + t] = a.strz[kp];
   }
   for( int t = c.tid; t < 6 * C::NH; t += NT ) smem[C::O_HML + t] = 0.0;
   c.o = ( c.ty + 2 ) * PX + 2 * c.txh + 2; // left own point in a staged plane
   const int li = li0 + 2 * c.txh, lj = lj0 + c.ty;
   c.act = li <= b.nil - 3 && lj <= b.nj - 3;
   c.act2 = li + 1 <= b.nil - 3;
   c.gown = (long long)lj * b.ni + li;

   fast4::State s;
#pragma unroll
   for( int m = 0; m < 6; m++ )
#pragma unroll
      for( int t = 0; t < 2; t++ ) s.cu[m][t] = s.cv[m][t] = s.cw[m][t] = s.amz[m][t] = s.alz[m][t] = 0;
#pragma unroll
   for( int m = 0; m < 3; m++ ) s.rp[m][0] = s.rp[m][1] = 0;

#if !defined( SW4B200_EMULATE )
   asm volatile( "tcgen05.fence::before_thread_sync;" ::: "memory" );
   __syncthreads(); // tensor-memory address, s_sx, s_sy visible
   asm volatile( "tcgen05.fence::after_thread_sync;" ::: "memory" );
   const uint32_t tm_alloc = *tm_slot;
   {
      const int w = c.tid >> 5; // warp w owns lanes 32 (w%4)..+31 and the (w/4)-th column range
      tm.base = tm_alloc + ( (uint32_t)( ( w & 3 ) * 32 ) << 16 ) + (uint32_t)( ( w >> 2 ) * C::COLS );
   }
#else
   __syncthreads();
#endif
   {
      // records of the planes before the first one read as zero, like the register rings above
      const double z[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
      tm.template st<0 * C::REC, 8>( z ); tm.template st<0 * C::REC + 16, 8>( z ); tm.template st<0 * C::REC + 32, 4>( z );
      tm.template st<1 * C::REC, 8>( z ); tm.template st<1 * C::REC + 16, 8>( z ); tm.template st<1 * C::REC + 32, 4>( z );
      tm.template st<2 * C::REC, 8>( z ); tm.template st<2 * C::REC + 16, 8>( z ); tm.template st<2 * C::REC + 32, 4>( z );
      tm.template st<3 * C::REC, 8>( z ); tm.template st<3 * C::REC + 16, 8>( z ); tm.template st<3 * C::REC + 32, 4>( z );
      tm.template st<4 * C::REC, 8>( z ); tm.template st<4 * C::REC + 16, 8>( z ); tm.template st<4 * C::REC + 32, 4>( z );
      tm.template st<5 * C::REC, 8>( z ); tm.template st<5 * C::REC + 16, 8>( z ); tm.template st<5 * C::REC + 32, 4>( z );
      tm.wait_st();
   }

   const int plast = c.kb + 3; // one extra step finishes plane kb
   fast4::Ph ph;
   {
      const int p = c.ka - 2;
      int m6 = p % 6, m3 = p % 3;
      m6 = m6 < 0 ? m6 + 6 : m6; m3 = m3 < 0 ? m3 + 3 : m3;
      ph.slot = m6;
#pragma unroll
      for( int j = 0; j < 5; j++ )
      {
	 const int sj = ( m6 + 6 - j ) % 6;
	 ph.o[j] = PLANE * sj; ph.c[j] = C::REC * sj;
      }
      ph.par = p & 1;
      ph.wpar = 0;
      ph.t0 = 2 * C::NH * m3; ph.t2 = 2 * C::NH * ( ( m3 + 1 ) % 3 );
      fast4::stage<TY, EPI>( a, maps, c, p, ph.slot, ph.par );
   }
   auto next_plane = [&]() {
      ph.slot = ph.slot == 5 ? 0 : ph.slot + 1;
#pragma unroll
      for( int j = 4; j > 0; j-- ) { ph.o[j] = ph.o[j - 1]; ph.c[j] = ph.c[j - 1]; }
      ph.o[0] = PLANE * ph.slot; ph.c[0] = C::REC * ph.slot;
      ph.wpar ^= 1 << ph.par; // the barrier just consumed flips its phase
      ph.par ^= 1;
      ph.t2 = ( ph.t0 == 0 ? 2 : ( ph.t0 == 2 * C::NH ? 0 : 1 ) ) * 2 * C::NH; // slot of plane (p+1)-2 = slot of p+2 mod 3
      ph.t0 = ph.t0 == 4 * C::NH ? 0 : ph.t0 + 2 * C::NH;
   };
   auto shift2 = [&]() {
#pragma unroll
      for( int t = 0; t < 2; t++ )
#pragma unroll
	 for( int j = 0; j < 4; j++ )
	 {
	    s.cu[j][t] = s.cu[j + 2][t]; s.cv[j][t] = s.cv[j + 2][t]; s.cw[j][t] = s.cw[j + 2][t];
	    s.amz[j][t] = s.amz[j + 2][t]; s.alz[j][t] = s.alz[j + 2][t];
	 }
   };
   constexpr bool NS = KIND == 1;
   int p = c.ka - 2;
   if( p & 1 )
   {
      fast4::step<TY, EPI, SPLIT, DER, NS, 1>( a, maps, c, s, tm, p, ph );
      shift2(); next_plane(); p++;
   }
   while( p <= plast )
   {
      fast4::step<TY, EPI, SPLIT, DER, NS, 0>( a, maps, c, s, tm, p, ph );
      next_plane(); p++;
      if( p > plast ) break;
      fast4::step<TY, EPI, SPLIT, DER, NS, 1>( a, maps, c, s, tm, p, ph );
      shift2(); next_plane(); p++;
   }
#if !defined( SW4B200_EMULATE )
   // every warp is done with its strip (its loads were waited for) before the columns go back
   asm volatile( "tcgen05.fence::before_thread_sync;" ::: "memory" );
   __syncthreads();
   if( c.tid < 32 ) asm volatile( "tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"( tm_alloc ) : "memory" );
#endif
}

#ifndef SW4B200_EMULATE
namespace {
template <int TY, int EPI, int SPLIT, bool DER, int KIND>
int launch_fast4_t( FastArgs a, cudaStream_t st )
{
   typedef fast4::Cfg<TY> C;
   static bool configured = false;
   const size_t smem = C::SMEM_DOUBLES * sizeof( double );
   if( !configured )
   {
      cudaError_t e = cudaFuncSetAttribute( k_rhs_fast4<TY, EPI, SPLIT, DER, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem );
      if( e != cudaSuccess ) return set_error( "k_rhs_fast4: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString( e ) );
      configured = true;
   }
   if( a.kchunk <= 0 ) a.kchunk = fast_kchunk( a.b, a.khi - a.klo + 1, TY );
   if( a.kchunk > C::SZMAX - 6 ) a.kchunk = C::SZMAX - 6;
   const Block& b = a.b;
   FastMaps maps;
   memset( &maps, 0, sizeof( maps ) );
   for( int f = 0; f < 3; f++ )
      if( make_tmap( &maps.u[f], a.u[f], b, C::PX, C::PY ) ) return 1;
   if( make_tmap( &maps.mu, a.mu, b, C::PX, C::PY ) || make_tmap( &maps.la, DER ? a.la2 : a.la, b, C::PX, C::PY ) ) return 1;
   if( EPI != EPI_LU )
   {
      if( make_tmap( &maps.rho, DER ? a.rhoi : a.rho, b, C::TX, TY ) ) return 1;
      for( int f = 0; f < 3; f++ )
	 if( make_tmap( &maps.um[f], a.um[f], b, C::TX, TY ) ) return 1;
   }
   dim3 bs( C::NT, 1, 1 );
   dim3 gs( ( b.nil - 4 + C::TX - 1 ) / C::TX, ( b.nj - 4 + TY - 1 ) / TY, ( a.khi - a.klo + 1 + a.kchunk - 1 ) / a.kchunk );
   k_rhs_fast4<TY, EPI, SPLIT, DER, KIND><<<gs, bs, smem, st>>>( a, maps );
   count_launch();
   return check_launch( "k_rhs_fast4" );
}
} // namespace

int launch_fast2( int epi, FastArgs a, cudaStream_t st );

namespace {
// the two launches of a fused pass (plain tiles, tiles with stretching: disjoint outputs, read-only inputs): the second on the
// auxiliary stream, so that its thread blocks fill the SMs the first leaves idle in its last wave instead of adding a tail
int launch_pair( int ( *plain )( FastArgs, cudaStream_t ), int ( *general )( FastArgs, cudaStream_t ), const FastArgs& a, cudaStream_t st )
{
   // (the library has ONE auxiliary stream and one pair of fork / join events: host threads driving different grids take turns)
   static std::mutex mtx;
   std::lock_guard<std::mutex> lock( mtx );
   cudaStream_t aux = aux_fork( st );
   if( plain( a, st ) ) return 1;
   if( !aux ) return general( a, st );
   const int rc = general( a, aux );
   return aux_join( st ) || rc;
}
} // namespace

int launch_fast4( int epi, const FastArgs& a, cudaStream_t st )
{
   if( a.khi < a.klo ) return 0;
   // the tensor maps need 16-byte aligned rows: an even row pitch and 16-byte aligned arrays.  Caller-owned arrays with odd ni
   // (the operator-level entry points on the reference's own layout) take the cp.async kernel of the second generation; grid
   // blocks (sw4b200_grid_create) pad their rows to an even pitch and stay here.
   uintptr_t al = (uintptr_t)a.u[0] | (uintptr_t)a.u[1] | (uintptr_t)a.u[2] | (uintptr_t)a.mu | (uintptr_t)a.la;
   if( epi != EPI_LU ) al |= (uintptr_t)a.rho | (uintptr_t)a.um[0] | (uintptr_t)a.um[1] | (uintptr_t)a.um[2] | (uintptr_t)a.la2 | (uintptr_t)a.rhoi;
   al |= (uintptr_t)a.out[0] | (uintptr_t)a.out[1] | (uintptr_t)a.out[2];
   if( epi == EPI_PRED && a.out2[0] ) al |= (uintptr_t)a.out2[0] | (uintptr_t)a.out2[1] | (uintptr_t)a.out2[2];
   if( ( a.b.ni & 1 ) || ( al & 15 ) ) return launch_fast2( epi, a, st );
   if( a.b.ni < fast4::Cfg<16>::PX || a.b.nj < fast4::Cfg<16>::PY ) return launch_fast2( epi, a, st ); // (arrays narrower than one TMA box)
   // dense forcing arrays and a predictor without the stored acceleration are not on the time-stepping path (forcing is
   // injected sparsely, pass A always stores uacc): the operator-level calls that use them take the cp.async kernel too
   if( epi != EPI_LU && a.fo[0] ) return launch_fast2( epi, a, st );
   if( epi == EPI_PRED && !a.out2[0] ) return launch_fast2( epi, a, st );
   // the fused passes of this kernel exist for the derived coefficient arrays of a grid block only
   if( epi != EPI_LU && ( !a.la2 || !a.rhoi ) ) return launch_fast2( epi, a, st );
   // grid blocks with an odd number of points per row are allocated with rows padded to an even pitch (api.cu): the last pair of
   // a row is then split by the boundary (SPLIT variant: that pair stores its left point only).
   // The fused passes: the plain tiles and, beside them, the tiles with stretching (see KIND above)
   ProfScope prof( epi == EPI_PRED ? "rhs_fast_pred" : ( epi == EPI_CORR ? "rhs_fast_corr" : "rhs_fast_lu" ), st );
   if( a.b.nil != a.b.ni )
      switch( epi )
      {
      case EPI_LU: return launch_fast4_t<16, EPI_LU, 1, false, 0>( a, st );
      case EPI_PRED: return launch_pair( launch_fast4_t<16, EPI_PRED, 1, true, 1>, launch_fast4_t<16, EPI_PRED, 1, true, 2>, a, st );
      default: return launch_pair( launch_fast4_t<16, EPI_CORR, 1, true, 1>, launch_fast4_t<16, EPI_CORR, 1, true, 2>, a, st );
      }
   switch( epi )
   {
   case EPI_LU: return launch_fast4_t<16, EPI_LU, 0, false, 0>( a, st );
   case EPI_PRED: return launch_pair( launch_fast4_t<16, EPI_PRED, 0, true, 1>, launch_fast4_t<16, EPI_PRED, 0, true, 2>, a, st );
   default: return launch_pair( launch_fast4_t<16, EPI_CORR, 0, true, 1>, launch_fast4_t<16, EPI_CORR, 0, true, 2>, a, st );
   }
}
#endif

} // namespace sw4b200
