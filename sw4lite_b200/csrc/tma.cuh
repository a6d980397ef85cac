// TMA tile loads on mbarriers for the stencil kernels (sm_100a): tensor-map type, the descriptors of one launch, the
// device-side mbarrier / cp.async.bulk.tensor wrappers and the host-side encoder.  Shared by rhs4sg_fast4.cu (interior
// rows) and cart_v1.cu (closure rows).  Under SW4B200_EMULATE (tests/emu) a tile load is a zero-filling copy and an
// mbarrier a triple of atomics.
#pragma once
#include "common.cuh"
#include <cstring>
#ifdef SW4B200_EMULATE
#include <atomic>
#include <thread>
#else
#include <cuda.h> // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, libcuda is not linked)
#endif

namespace sw4b200 {

// TMA descriptors of the arrays one launch reads: 3-D tensors (ni, nj, nk) of doubles, boxes of one tile plane
#if defined( SW4B200_EMULATE )
struct TMap { const double* base; };
#define SW4_GRID_CONSTANT
#else
typedef CUtensorMap TMap;
#define SW4_GRID_CONSTANT __grid_constant__
#endif
struct FastMaps
{
   TMap u[3], mu, la; // boxes PX x PY x 1 (tile plane with its ring)
   TMap rho, um[3];   // boxes TX x TY x 1 (own points)
};

namespace fast4 {

// ---- TMA tile loads (cp.async.bulk.tensor, SASS UTMALDG) completing on an mbarrier: one request moves a whole
// BX x BY plane box from global to shared memory -- no registers, no LSU queue, no per-thread address arithmetic,
// out-of-array elements arrive as zeros
#if defined( SW4B200_EMULATE )
// (emulation: an mbarrier is 16 bytes = {arrivals per phase, pending arrivals, phase}; one OS thread per CUDA thread)
struct EmuBar { std::atomic<int> count, pending, phase; };
__device__ __forceinline__ void mbar_init( double* m, int count )
{
   EmuBar* b = reinterpret_cast<EmuBar*>( m );
   b->count.store( count ); b->pending.store( count ); b->phase.store( 0 );
}
__device__ __forceinline__ void mbar_arrive_expect( double* m, int ) // (the emulated tile loads are synchronous and precede it)
{
   EmuBar* b = reinterpret_cast<EmuBar*>( m );
   if( b->pending.fetch_sub( 1 ) == 1 )
   {
      b->pending.store( b->count.load() );
      b->phase.fetch_add( 1 );
   }
}
__device__ __forceinline__ void mbar_arrive( double* m ) { mbar_arrive_expect( m, 0 ); }
__device__ __forceinline__ void mbar_wait( double* m, int parity )
{
   EmuBar* b = reinterpret_cast<EmuBar*>( m );
   while( ( b->phase.load() & 1 ) == parity ) std::this_thread::yield();
}
template <int BX, int BY, int BZ = 1>
__device__ __forceinline__ void tma_tile( double* dst, const TMap* map, const Block& b, int c0, int c1, int c2z, double* )
{
   for( int z = 0; z < BZ; z++ )
   for( int y = 0; y < BY; y++ )
      for( int x = 0; x < BX; x++ )
      {
	 const int i = c0 + x, j = c1 + y;
	 const int c2 = c2z + z;
	 dst[( z * BY + y ) * BX + x] = ( i >= 0 && i < b.ni && j >= 0 && j < b.nj && c2 >= 0 && c2 < b.nk ) ? map->base[b.nij * c2 + (long long)j * b.ni + i] : 0.0;
      }
}
#else
__device__ __forceinline__ void mbar_init( double* mbar, int count )
{
   asm volatile( "mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"( (uint32_t)__cvta_generic_to_shared( mbar ) ), "r"( count ) : "memory" );
}
__device__ __forceinline__ void mbar_arrive_expect( double* mbar, int bytes )
{
   asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"( (uint32_t)__cvta_generic_to_shared( mbar ) ), "r"( bytes )
		 : "memory" );
}
// plain arrival (release at CTA scope: what the thread, and the threads it synchronised with, wrote to shared memory before
// is visible to whoever sees the phase complete)
__device__ __forceinline__ void mbar_arrive( double* mbar )
{
   asm volatile( "mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"( (uint32_t)__cvta_generic_to_shared( mbar ) ) : "memory" );
}
__device__ __forceinline__ void mbar_wait( double* mbar, int parity )
{
   asm volatile( "{\n\t"
		 ".reg .pred P1;\n\t"
		 "SW4_MBAR_WAIT:\n\t"
		 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
		 "@P1 bra SW4_MBAR_DONE;\n\t"
		 "bra SW4_MBAR_WAIT;\n\t"
		 "SW4_MBAR_DONE:\n\t"
		 "}" ::"r"( (uint32_t)__cvta_generic_to_shared( mbar ) ),
		 "r"( parity )
		 : "memory" );
}
// dst 128-byte aligned; the box dimensions are those of the descriptor
template <int BX, int BY, int BZ = 1>
__device__ __forceinline__ void tma_tile( double* dst, const TMap* map, const Block&, int c0, int c1, int c2, double* mbar )
{
   asm volatile( "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
		     (uint32_t)__cvta_generic_to_shared( dst ) ),
		 "l"( map ), "r"( c0 ), "r"( c1 ), "r"( c2 ), "r"( (uint32_t)__cvta_generic_to_shared( mbar ) )
		 : "memory" );
}
#endif

} // namespace fast4

#ifndef SW4B200_EMULATE
// tensor map of one array of the block: dims (ni, nj, nk), box bx x by x bz, no swizzle, zeros outside the array
inline int make_tmap( TMap* m, const double* base, const Block& b, int bx, int by, int bz = 1 )
{
   typedef CUresult ( *Encode )( CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
				 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill );
   static Encode encode = 0;
   if( !encode )
   {
      void* fn = 0;
      cudaDriverEntryPointQueryResult qres;
      if( cudaGetDriverEntryPoint( "cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres ) != cudaSuccess || !fn )
	 return set_error( "cuTensorMapEncodeTiled is not available from the driver" );
      encode = (Encode)fn;
   }
   const cuuint64_t dims[3] = { (cuuint64_t)b.ni, (cuuint64_t)b.nj, (cuuint64_t)b.nk };
   const cuuint64_t strides[2] = { (cuuint64_t)b.ni * 8, (cuuint64_t)b.nij * 8 };
   const cuuint32_t box[3] = { (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz };
   const cuuint32_t estr[3] = { 1, 1, 1 };
   const CUresult r = encode( m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
			      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE );
   if( r != CUDA_SUCCESS ) return set_error( "cuTensorMapEncodeTiled failed (%d)", (int)r );
   return 0;
}

#endif

} // namespace sw4b200
