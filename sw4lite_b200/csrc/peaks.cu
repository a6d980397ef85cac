// Measured co-bound of the stencil kernels: the fp64 FMA issue rate of this GPU (BASELINE.md section 3 leaves it "to be
// measured by the builder").  A register-only kernel: every thread runs 8 independent chains of dependent DFMAs, 32 warps
// per SM, so the fp64 pipe is the only thing that can limit it.  bench.py reports the result beside the HBM peak.
#include "common.cuh"

namespace sw4b200 {
namespace {
__global__ void __launch_bounds__( 256 ) k_fp64_fma_peak( double* out, int iters, double b, double c )
{
   double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
   for( int i = 0; i < iters; i++ )
   {
#pragma unroll
      for( int u = 0; u < 8; u++ )
      {
	 a0 = fma( a0, b, c ); a1 = fma( a1, b, c ); a2 = fma( a2, b, c ); a3 = fma( a3, b, c );
	 a4 = fma( a4, b, c ); a5 = fma( a5, b, c ); a6 = fma( a6, b, c ); a7 = fma( a7, b, c );
      }
   }
   out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = ( a0 + a1 ) + ( a2 + a3 ) + ( a4 + a5 ) + ( a6 + a7 );
}
} // namespace

// DFMA warp-instructions per second and the equivalent TFLOP/s (2 flop per lane), best of `reps` launches
int measure_fp64_peak( double* tflops, double* fma_per_s, cudaStream_t st )
{
   int dev = 0, sms = 148;
   cudaGetDevice( &dev );
   cudaDeviceGetAttribute( &sms, cudaDevAttrMultiProcessorCount, dev );
   const int blocks = sms * 4, threads = 256, iters = 4096;
   double* out = 0;
   if( cudaMalloc( (void**)&out, (size_t)blocks * threads * sizeof( double ) ) != cudaSuccess ) return set_error( "measure_fp64_peak: cudaMalloc failed" );
   cudaEvent_t e0, e1;
   cudaEventCreate( &e0 ); cudaEventCreate( &e1 );
   double best = 0;
   for( int r = 0; r < 6; r++ )
   {
      cudaEventRecord( e0, st );
      k_fp64_fma_peak<<<blocks, threads, 0, st>>>( out, iters, 1.0000001, 1e-9 );
      cudaEventRecord( e1, st );
      cudaEventSynchronize( e1 );
      float ms = 0;
      cudaEventElapsedTime( &ms, e0, e1 );
      const double fmas = (double)blocks * threads * iters * 64.0;
      if( r > 0 && ms > 0 && fmas / ( ms * 1e-3 ) > best ) best = fmas / ( ms * 1e-3 );
   }
   count_launch( 6 );
   cudaEventDestroy( e0 ); cudaEventDestroy( e1 );
   cudaFree( out );
   if( check_launch( "k_fp64_fma_peak" ) ) return 1;
   *fma_per_s = best;
   *tflops = 2 * best / 1e12;
   return 0;
}
} // namespace sw4b200
