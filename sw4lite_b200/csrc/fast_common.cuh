// Declarations shared by the throughput kernels of the Cartesian interior rows (rhs4sg_fast2.cu: cp.async staging,
// rhs4sg_fast4.cu: TMA staging + tensor memory): launch arguments, the 1-D building blocks of the operator
// (reference formulas rhs4sg_rev.C:112-348) and the dispatcher.  Also compiled by g++ (SW4B200_EMULATE, tests/emu).
#pragma once
#include "common.cuh"
#include <cstdlib>

namespace sw4b200 {
enum FastEpilogue { EPI_LU = 0, EPI_PRED = 1, EPI_CORR = 2 };

struct FastArgs
{
   Block b;
   int klo, khi;     // output planes of the whole launch (global k), inclusive
   int kchunk;	     // planes per CTA in z
   const double* u[3];	 // input field (u for LU/PRED, uacc for CORR), halo'd reads
   const double *mu, *la;
   // derived, time-invariant coefficient arrays of a grid block (api.cu keeps them next to mu, la, rho): 2 mu + lambda and
   // 1 / rho.  The TMA kernel's fused passes read THESE instead of lambda and rho (13 fp64 instructions per point less);
   // null for caller-owned arrays, whose fused passes take the cp.async kernel
   const double *la2, *rhoi;
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
} // namespace fast

#ifndef SW4B200_EMULATE
// planes per CTA in z: enough CTAs to fill the 148 SMs a few times over, chunks not shorter than 16
inline int fast_kchunk( const Block& b, int nplanes, int ty )
{
   const long long tiles = (long long)( ( b.nil - 4 + 31 ) / 32 ) * ( ( b.nj - 4 + ty - 1 ) / ty );
   const long long want = 148LL * 6; // CTAs
   long long nch = ( want + tiles - 1 ) / tiles;
   if( nch < 1 ) nch = 1;
   long long kc = ( nplanes + nch - 1 ) / nch;
   if( kc < 16 ) kc = 16;
   if( kc > nplanes ) kc = nplanes;
   return (int)kc;
}

int launch_fast2( int epi, FastArgs a, cudaStream_t st );
int launch_fast4( int epi, const FastArgs& a, cudaStream_t st );
// interior rows [a.klo,a.khi]: the TMA kernel (rhs4sg_fast4.cu), which hands grids it cannot stage to the cp.async
// kernel (rhs4sg_fast2.cu).  SW4B200_FAST_GEN=2 forces the cp.async kernel (A/B measurements).
inline int launch_fast( int epi, const FastArgs& a, cudaStream_t st )
{
   if( a.khi < a.klo ) return 0;
   static int gen = -1;
   if( gen < 0 )
   {
      const char* e = getenv( "SW4B200_FAST_GEN" );
      gen = ( e && atoi( e ) == 2 ) ? 2 : 4;
   }
   return gen == 2 ? launch_fast2( epi, a, st ) : launch_fast4( epi, a, st );
}
#endif

} // namespace sw4b200
