// A thread's private strip of the SM's tensor memory (tcgen05.st / tcgen05.ld, 32x32b shape: one lane per thread, N
// 32-bit columns per instruction), used by the stencil kernels as a spill-free store for per-thread z state.
// Layout contract: 512 columns x 128 lanes x 32 bit per SM; warp w may only touch lanes 32 (w%4) .. +31.
// Under SW4B200_EMULATE (tests/emu) the strip is a per-thread array.
#pragma once
#include "common.cuh"

namespace sw4b200 {
namespace tmem {

// ---- a thread's private strip of tensor memory ------------------------------------------------------------
#if defined( SW4B200_EMULATE )
struct TmVal { double v; };
__device__ __forceinline__ double tm_get( const TmVal& x ) { return x.v; }
struct Tm
{
   double mem[256];
   template <int COL, int N> // N doubles to columns COL..COL+2N-1
   __device__ __forceinline__ void st( const double* v, int off = 0 ) { for( int i = 0; i < N; i++ ) mem[( COL + off ) / 2 + i] = v[i]; }
   template <int COL, int N>
   __device__ __forceinline__ void ld( TmVal* r, int off = 0 ) const { for( int i = 0; i < N; i++ ) r[i].v = mem[( COL + off ) / 2 + i]; }
   __device__ __forceinline__ void st2i( int col, const double* v, int off = 0 ) { mem[( col + off ) / 2] = v[0]; mem[( col + off ) / 2 + 1] = v[1]; }
   __device__ __forceinline__ void wait_st() const {}
   template <int N>
   __device__ __forceinline__ void wait_ld( TmVal* ) const {}
};
#else
struct TmVal { uint32_t lo, hi; };
__device__ __forceinline__ double tm_get( const TmVal& x ) { return __hiloint2double( (int)x.hi, (int)x.lo ); }
#define SW4_LO( x ) "r"( __double2loint( x ) )
#define SW4_HI( x ) "r"( __double2hiint( x ) )
struct Tm
{
   uint32_t base; // tensor-memory address of the strip: (first lane of the warp's quadrant) << 16 | first column
   template <int COL, int N>
   __device__ __forceinline__ void st( const double* v, int off = 0 ) // off: run-time column offset (warp uniform)
   {
      static_assert( N == 1 || N == 2 || N == 4 || N == 8, "st: 1, 2, 4 or 8 doubles" );
      const uint32_t ta = base + COL + off;
      if constexpr( N == 1 )
	 asm volatile( "tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"( ta ), SW4_LO( v[0] ), SW4_HI( v[0] ) : "memory" );
      else if constexpr( N == 2 )
	 asm volatile( "tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"( ta ), SW4_LO( v[0] ), SW4_HI( v[0] ),
		       SW4_LO( v[1] ), SW4_HI( v[1] )
		       : "memory" );
      else if constexpr( N == 4 )
	 asm volatile( "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"( ta ), SW4_LO( v[0] ),
		       SW4_HI( v[0] ), SW4_LO( v[1] ), SW4_HI( v[1] ), SW4_LO( v[2] ), SW4_HI( v[2] ), SW4_LO( v[3] ), SW4_HI( v[3] )
		       : "memory" );
      else
	 asm volatile( "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"( ta ),
		       SW4_LO( v[0] ), SW4_HI( v[0] ), SW4_LO( v[1] ), SW4_HI( v[1] ), SW4_LO( v[2] ), SW4_HI( v[2] ), SW4_LO( v[3] ),
		       SW4_HI( v[3] ), SW4_LO( v[4] ), SW4_HI( v[4] ), SW4_LO( v[5] ), SW4_HI( v[5] ), SW4_LO( v[6] ), SW4_HI( v[6] ),
		       SW4_LO( v[7] ), SW4_HI( v[7] )
		       : "memory" );
   }
   // two doubles to columns col..col+3 (col: a constant after unrolling)
   __device__ __forceinline__ void st2i( int col, const double* v, int off = 0 )
   {
      const uint32_t ta = base + col + off;
      asm volatile( "tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"( ta ), SW4_LO( v[0] ), SW4_HI( v[0] ), SW4_LO( v[1] ),
		    SW4_HI( v[1] )
		    : "memory" );
   }
   template <int COL, int N>
   __device__ __forceinline__ void ld( TmVal* r, int off = 0 ) const
   {
      static_assert( N == 1 || N == 2 || N == 4 || N == 8, "ld: 1, 2, 4 or 8 doubles" );
      const uint32_t ta = base + COL + off;
      if constexpr( N == 1 )
	 asm volatile( "tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"( r[0].lo ), "=r"( r[0].hi ) : "r"( ta ) );
      else if constexpr( N == 2 )
	 asm volatile( "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
		       : "=r"( r[0].lo ), "=r"( r[0].hi ), "=r"( r[1].lo ), "=r"( r[1].hi )
		       : "r"( ta ) );
      else if constexpr( N == 4 )
	 asm volatile( "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		       : "=r"( r[0].lo ), "=r"( r[0].hi ), "=r"( r[1].lo ), "=r"( r[1].hi ), "=r"( r[2].lo ), "=r"( r[2].hi ),
			 "=r"( r[3].lo ), "=r"( r[3].hi )
		       : "r"( ta ) );
      else
	 asm volatile( "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
		       : "=r"( r[0].lo ), "=r"( r[0].hi ), "=r"( r[1].lo ), "=r"( r[1].hi ), "=r"( r[2].lo ), "=r"( r[2].hi ),
			 "=r"( r[3].lo ), "=r"( r[3].hi ), "=r"( r[4].lo ), "=r"( r[4].hi ), "=r"( r[5].lo ), "=r"( r[5].hi ),
			 "=r"( r[6].lo ), "=r"( r[6].hi ), "=r"( r[7].lo ), "=r"( r[7].hi )
		       : "r"( ta ) );
   }
   // stores issued so far by this warp have been performed (a later tcgen05.ld sees them)
   __device__ __forceinline__ void wait_st() const { asm volatile( "tcgen05.wait::st.sync.aligned;" ::: "memory" ); }
   // loads issued so far have landed in their registers; the empty asm statements tie the N values to this
   // point so that no use of them can be moved above the wait
   template <int N>
   __device__ __forceinline__ void wait_ld( TmVal* r ) const
   {
      asm volatile( "tcgen05.wait::ld.sync.aligned;" ::: "memory" );
#pragma unroll
      for( int i = 0; i < N; i++ ) asm volatile( "" : "+r"( r[i].lo ), "+r"( r[i].hi ) );
   }
};
#endif

} // namespace tmem
} // namespace sw4b200
