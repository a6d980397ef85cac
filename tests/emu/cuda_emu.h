// TEST INFRASTRUCTURE ONLY: a minimal CUDA-on-CPU shim so that the kernel source of
// sw4lite_b200/csrc/*.cu can be compiled by g++ and executed with one OS thread per CUDA thread
// (one thread block at a time, __syncthreads = a barrier).  It lets the CPU test-suite check the
// kernels' index logic and algebra against the oracle without a GPU.  Never used by the product.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cmath>
#include <atomic>
#include <barrier>
#include <thread>
#include <vector>
#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__( ... )
#define __constant__ static
#define __restrict__

struct dim3 { unsigned x, y, z; dim3( unsigned a = 1, unsigned b = 1, unsigned c = 1 ) : x( a ), y( b ), z( c ) {} };
typedef void* cudaStream_t;

namespace emu {
extern thread_local dim3 t_threadIdx;
extern dim3 g_blockIdx, g_blockDim, g_gridDim;
extern double* g_smem;
extern std::barrier<>* g_barrier;
}
#define threadIdx ( emu::t_threadIdx )
#define blockIdx ( emu::g_blockIdx )
#define blockDim ( emu::g_blockDim )
#define gridDim ( emu::g_gridDim )
inline void __syncthreads() { emu::g_barrier->arrive_and_wait(); }
// block-wide AND of a predicate (threads that have left the block do not vote)
namespace emu { extern std::atomic<int> g_vote_false; extern int g_force_general, g_nostr_ctas; }
inline int __syncthreads_and( int pred )
{
   if( !pred ) emu::g_vote_false.fetch_add( 1 );
   emu::g_barrier->arrive_and_wait();
   const int r = emu::g_vote_false.load() == 0;
   emu::g_barrier->arrive_and_wait();
   if( threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0 ) emu::g_vote_false.store( 0 );
   emu::g_barrier->arrive_and_wait();
   return r;
}
inline double* emu_shared_memory() { return emu::g_smem; }

namespace emu {
// run kernel body `fn` for every block of `grid`, `block` threads each, smem_doubles of shared memory
inline void launch( dim3 grid, dim3 block, size_t smem_doubles, const std::function<void()>& fn )
{
   std::vector<double> smem( smem_doubles + 16 );
   g_smem = smem.data();
   g_blockDim = block; g_gridDim = grid;
   const unsigned nt = block.x * block.y * block.z;
   for( unsigned bz = 0; bz < grid.z; bz++ )
      for( unsigned by = 0; by < grid.y; by++ )
	 for( unsigned bx = 0; bx < grid.x; bx++ )
	 {
	    g_blockIdx = dim3( bx, by, bz );
	    for( auto& v : smem ) v = std::nan( "" ); // uninitialised shared memory must not be relied upon
	    std::barrier<> bar( nt );
	    g_barrier = &bar;
	    std::vector<std::thread> th;
	    th.reserve( nt );
	    for( unsigned t = 0; t < nt; t++ )
	       th.emplace_back( [&, t]() {
		  t_threadIdx = dim3( t % block.x, ( t / block.x ) % block.y, t / ( block.x * block.y ) );
		  fn();
		  // a thread that returned early must not block the others
		  g_barrier->arrive_and_drop();
	       } );
	    for( auto& x : th ) x.join();
	 }
}
}
