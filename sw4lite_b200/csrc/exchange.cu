// Halo exchange of z-slab grid blocks inside the library: replaces the reference's MPI halo swap of the new solution
// (EW::communicate_array, EW.C:3247-3317; communicate_arrayCU_X/Y + pack/unpack_HaloArrayCU_*, EW_cuda.C:1515-1997, called at
// EW.C:2599-2608 and :2731-2740) for one process per GPU on one NVSwitch node.
//
// A k-plane of an (i,j,k,c) block is one contiguous run of ni*nj doubles per component (of 3*ni*nj doubles in the (c,i,j,k)
// layout), so the two face planes a neighbour needs are sent STRAIGHT from the field arrays into the neighbour's halo planes:
// ncclSend / ncclRecv grouped per exchange on the library's communication stream -- no pack / unpack kernels, no staging
// buffers (the reference packs into dev_SideEdge_Send buffers and goes through the host unless MPI is CUDA-aware).  NCCL moves
// the planes GPU to GPU over NVLink.  The exchange is asynchronous to the compute stream: begin() makes the communication
// stream wait for the face rows (an event), end() makes the compute stream wait for the received planes; the bulk rows of the
// slab are computed in between (sw4b200_grid_*_part).
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy already loaded by the process, e.g. PyTorch's, else the
// system's); single-GPU users of the library never need it.
#include "common.cuh"
#include <nccl.h> // types only
#include <dlfcn.h>
#include <cstdio>

namespace sw4b200 {
namespace {

struct NcclApi
{
   void* h;
   decltype( &ncclGetUniqueId ) GetUniqueId;
   decltype( &ncclCommInitRank ) CommInitRank;
   decltype( &ncclCommDestroy ) CommDestroy;
   decltype( &ncclSend ) Send;
   decltype( &ncclRecv ) Recv;
   decltype( &ncclAllReduce ) AllReduce;
   decltype( &ncclGroupStart ) GroupStart;
   decltype( &ncclGroupEnd ) GroupEnd;
   decltype( &ncclGetErrorString ) GetErrorString;
};
NcclApi g_nccl = { 0 };
ncclComm_t g_comm = 0;
int g_rank = 0, g_nranks = 1;

int nccl_load()
{
   if( g_nccl.h ) return 0;
   const char* names[] = { "libnccl.so.2", "libnccl.so" };
   void* h = 0;
   for( const char* n : names )
      if( ( h = dlopen( n, RTLD_NOW | RTLD_GLOBAL ) ) ) break;
   if( !h ) return set_error( "NCCL is not available (dlopen libnccl.so.2: %s)", dlerror() );
#define SW4_SYM( field, name )                                                                   \
   g_nccl.field = (decltype( g_nccl.field ))dlsym( h, name );                                    \
   if( !g_nccl.field ) return set_error( "NCCL: symbol %s not found", name );
   SW4_SYM( GetUniqueId, "ncclGetUniqueId" )
   SW4_SYM( CommInitRank, "ncclCommInitRank" )
   SW4_SYM( CommDestroy, "ncclCommDestroy" )
   SW4_SYM( Send, "ncclSend" )
   SW4_SYM( Recv, "ncclRecv" )
   SW4_SYM( AllReduce, "ncclAllReduce" )
   SW4_SYM( GroupStart, "ncclGroupStart" )
   SW4_SYM( GroupEnd, "ncclGroupEnd" )
   SW4_SYM( GetErrorString, "ncclGetErrorString" )
#undef SW4_SYM
   g_nccl.h = h;
   return 0;
}

#define NCCL_OK( call )                                                                                         \
   do                                                                                                           \
   {                                                                                                            \
      ncclResult_t r_ = ( call );                                                                               \
      if( r_ != ncclSuccess ) return set_error( "%s: %s", #call, g_nccl.GetErrorString( r_ ) );                 \
   } while( 0 )

} // namespace

int comm_unique_id( void* out128 )
{
   if( nccl_load() ) return 1;
   ncclUniqueId id;
   NCCL_OK( g_nccl.GetUniqueId( &id ) );
   memcpy( out128, &id, sizeof( id ) );
   return 0;
}

int comm_init( int rank, int nranks, const void* id128 )
{
   if( nranks < 1 || rank < 0 || rank >= nranks ) return set_error( "comm_init: rank %d of %d", rank, nranks );
   if( g_comm ) return set_error( "comm_init: the communicator exists already" );
   g_rank = rank; g_nranks = nranks;
   if( nranks == 1 ) return 0;
   if( nccl_load() ) return 1;
   ncclUniqueId id;
   memcpy( &id, id128, sizeof( id ) );
   NCCL_OK( g_nccl.CommInitRank( &g_comm, nranks, id, rank ) );
   return 0;
}

int comm_finalize()
{
   if( g_comm ) { g_nccl.CommDestroy( g_comm ); g_comm = 0; }
   g_rank = 0; g_nranks = 1;
   return 0;
}

// small host-value reductions over the ranks (dt, error norms, timings: EW.C:5134, 4606-4608): op 0 = max, 1 = sum, 2 = min
int comm_allreduce( double* v, int n, int op, cudaStream_t st )
{
   if( g_nranks == 1 ) return 0;
   if( !g_comm ) return set_error( "comm_allreduce: sw4b200_comm_init has not been called" );
   double* d = 0;
   if( cudaMalloc( (void**)&d, n * sizeof( double ) ) != cudaSuccess ) return set_error( "comm_allreduce: cudaMalloc failed" );
   cudaMemcpyAsync( d, v, n * sizeof( double ), cudaMemcpyHostToDevice, st );
   const ncclRedOp_t o = op == 0 ? ncclMax : ( op == 1 ? ncclSum : ncclMin );
   ncclResult_t r = g_nccl.AllReduce( d, d, n, ncclDouble, o, g_comm, st );
   cudaMemcpyAsync( v, d, n * sizeof( double ), cudaMemcpyDeviceToHost, st );
   cudaStreamSynchronize( st );
   cudaFree( d );
   if( r != ncclSuccess ) return set_error( "ncclAllReduce: %s", g_nccl.GetErrorString( r ) );
   return 0;
}

int comm_rank() { return g_rank; }
int comm_size() { return g_nranks; }

// the face planes of `field` (3 components) towards both neighbours: sends the two interior planes next to a halo face,
// receives into the two halo planes; inside an open NCCL group
int exchange_field( const Block& b, double* field, int peer_lo, int peer_hi, cudaStream_t st )
{
   const long long nij = b.nij;
   const int nrun = b.sp == 1 ? 3 : 1;			      // contiguous runs per face: one per component, or one in all
   const size_t count = (size_t)( b.sp == 1 ? 2 : 6 ) * nij;  // doubles per run
   for( int side = 0; side < 2; side++ )
   {
      const int peer = side == 0 ? peer_lo : peer_hi;
      if( peer < 0 ) continue;
      const long long ksend = side == 0 ? 2 : b.nk - 4, krecv = side == 0 ? 0 : b.nk - 2;
      for( int r = 0; r < nrun; r++ )
      {
	 double* const base = field + ( b.sp == 1 ? r * b.sc : 0 );
	 const long long pl = b.sp == 1 ? nij : 3 * nij;
	 NCCL_OK( g_nccl.Send( base + pl * ksend, count, ncclDouble, peer, g_comm, st ) );
	 NCCL_OK( g_nccl.Recv( base + pl * krecv, count, ncclDouble, peer, g_comm, st ) );
      }
   }
   return 0;
}

int exchange_group_start()
{
   if( !g_comm ) return set_error( "grid_exchange: sw4b200_comm_init has not been called" );
   NCCL_OK( g_nccl.GroupStart() );
   return 0;
}
int exchange_group_end()
{
   NCCL_OK( g_nccl.GroupEnd() );
   return 0;
}

} // namespace sw4b200
