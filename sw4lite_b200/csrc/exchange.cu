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
#include <cstdlib>

namespace sw4b200 {
namespace {

struct NcclApi
{
   void* h;
   decltype( &ncclGetUniqueId ) GetUniqueId;
   decltype( &ncclCommInitRank ) CommInitRank;
   decltype( &ncclCommInitRankConfig ) CommInitRankConfig; // (optional)
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
   g_nccl.CommInitRankConfig = (decltype( g_nccl.CommInitRankConfig ))dlsym( h, "ncclCommInitRankConfig" );
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
   // The transfer kernels share the SMs with the bulk rows of the slab, whose thread blocks fill a whole SM each: every
   // block NCCL runs takes an SM away from them for the length of the transfer.  A face exchange is 0.2-0.4 GB per neighbour
   // under >= 15 ms of bulk work, so a few blocks are enough (default 8; SW4B200_NCCL_MAX_CTAS overrides, 0 = NCCL's own
   // choice, 32 on an NVSwitch box).
   int max_ctas = 8;
   if( const char* e = getenv( "SW4B200_NCCL_MAX_CTAS" ) ) max_ctas = atoi( e );
   if( g_nccl.CommInitRankConfig && max_ctas > 0 )
   {
      ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
      cfg.maxCTAs = max_ctas;
      cfg.minCTAs = 1;
      NCCL_OK( g_nccl.CommInitRankConfig( &g_comm, nranks, id, rank, &cfg ) );
   }
   else
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

// ---- peer-to-peer transport: the copy engines move the face planes, no SM is involved --------------------------------
// Every rank maps its neighbours' solution arrays into its own address space (CUDA IPC) and PUSHES its face planes straight
// into their halo planes with cudaMemcpyAsync on the communication stream: NVLink copies by the DMA engines, which neither
// wait for nor take away an SM from the bulk rows (the NCCL transfer runs thread blocks: it starts late and slows the stencil
// kernel down, which matters most for thin slabs).  Two counters per direction order it, each written into the other
// rank's memory by a one-thread kernel and awaited by a one-thread spinning kernel:
//   ready[n]    receiver -> sender: "my face rows of phase n are done", i.e. nothing of mine reads the halo planes of the
//               previous exchange any more (their last readers -- the stencil of the face rows, the ghost-shell update, the
//               damping of the face rows -- all belong to the face-row part that precedes my own exchange n);
//   arrived[n]  sender -> receiver, behind the copies: the halo planes of exchange n are complete; the receiver's compute
//               stream waits for it before the boundary conditions.
// This is the ordering a receive posted after the face rows gives the NCCL path.  The handles travel once, at set-up, over
// the NCCL communicator.
struct P2PInfo // what one rank tells a neighbour
{
   cudaIpcMemHandle_t mem[4]; // U, Um, Up, Uacc allocations (in creation order: pointer rotation keeps them in step on all ranks)
   cudaIpcMemHandle_t flags;  // counters written by the neighbours: arrived[0,1] (by the low / high neighbour), ready[2,3]
   long long nk, nij, npts;
   int valid;
};
struct P2PLink
{
   bool on;
   double* mine[4];
   unsigned long long* flags; // my counters (device memory)
   struct Peer { int rank; double* base[4]; unsigned long long* flags; long long nk; } peer[2];
   unsigned long long posted;	  // exchanges started so far (every rank counts the same)
};

namespace {
__global__ void k_set_flag( unsigned long long* flag, unsigned long long value )
{
   __threadfence_system();
   *reinterpret_cast<volatile unsigned long long*>( flag ) = value;
   __threadfence_system();
}
__global__ void k_wait_flag( const unsigned long long* flag, unsigned long long value )
{
   const volatile unsigned long long* f = flag;
   const long long t0 = clock64();
   while( *f < value )
   {
      __nanosleep( 200 );
      // a neighbour that died must not leave this GPU spinning for ever: give up after 2^38 cycles (more than two minutes)
      // with a device-side trap, which the next CUDA call of this process reports as an error
      if( clock64() - t0 > ( 1LL << 38 ) ) __trap();
   }
   __threadfence_system();
}
} // namespace

// map the neighbours' arrays; returns 0 and sets l.on if the peer-to-peer transport is usable on every rank involved
int p2p_setup( P2PLink& l, double* const mine[4], const Block& b, int peer_lo, int peer_hi, cudaStream_t st )
{
   l.on = false;
   l.posted = 0;
   l.flags = 0;
   for( int s = 0; s < 2; s++ ) { l.peer[s].rank = s == 0 ? peer_lo : peer_hi; l.peer[s].flags = 0; for( int a = 0; a < 4; a++ ) l.peer[s].base[a] = 0; }
   for( int a = 0; a < 4; a++ ) l.mine[a] = mine[a];
   if( !g_comm ) return 0;
   P2PInfo me;
   memset( &me, 0, sizeof( me ) );
   me.valid = 1;
   if( cudaMalloc( (void**)&l.flags, 4 * sizeof( unsigned long long ) ) != cudaSuccess ) me.valid = 0;
   else cudaMemset( l.flags, 0, 4 * sizeof( unsigned long long ) );
   for( int a = 0; a < 4 && me.valid; a++ )
      if( cudaIpcGetMemHandle( &me.mem[a], mine[a] ) != cudaSuccess ) me.valid = 0;
   if( me.valid && cudaIpcGetMemHandle( &me.flags, l.flags ) != cudaSuccess ) me.valid = 0;
   cudaGetLastError();
   me.nk = b.nk; me.nij = b.nij; me.npts = b.npts;
   // swap the records with both neighbours
   char* dbuf = 0;
   if( cudaMalloc( (void**)&dbuf, 3 * sizeof( P2PInfo ) ) != cudaSuccess ) return set_error( "p2p_setup: cudaMalloc failed" );
   cudaMemcpyAsync( dbuf, &me, sizeof( me ), cudaMemcpyHostToDevice, st );
   NCCL_OK( g_nccl.GroupStart() );
   for( int s = 0; s < 2; s++ )
      if( l.peer[s].rank >= 0 )
      {
	 NCCL_OK( g_nccl.Send( dbuf, sizeof( P2PInfo ), ncclChar, l.peer[s].rank, g_comm, st ) );
	 NCCL_OK( g_nccl.Recv( dbuf + ( 1 + s ) * sizeof( P2PInfo ), sizeof( P2PInfo ), ncclChar, l.peer[s].rank, g_comm, st ) );
      }
   NCCL_OK( g_nccl.GroupEnd() );
   P2PInfo other[2];
   cudaMemcpyAsync( other, dbuf + sizeof( P2PInfo ), 2 * sizeof( P2PInfo ), cudaMemcpyDeviceToHost, st );
   cudaStreamSynchronize( st );
   cudaFree( dbuf );
   bool ok_ = me.valid != 0;
   for( int s = 0; s < 2 && ok_; s++ )
   {
      if( l.peer[s].rank < 0 ) continue;
      const P2PInfo& o = other[s];
      if( !o.valid || o.nij != b.nij ) { ok_ = false; break; }
      l.peer[s].nk = o.nk;
      for( int a = 0; a < 4 && ok_; a++ )
	 if( cudaIpcOpenMemHandle( (void**)&l.peer[s].base[a], o.mem[a], cudaIpcMemLazyEnablePeerAccess ) != cudaSuccess ) ok_ = false;
      if( ok_ && cudaIpcOpenMemHandle( (void**)&l.peer[s].flags, o.flags, cudaIpcMemLazyEnablePeerAccess ) != cudaSuccess ) ok_ = false;
   }
   cudaGetLastError();
   // all ranks of the communicator must agree (a rank that cannot map its neighbour makes both fall back to NCCL)
   double agree = ok_ ? 1.0 : 0.0;
   if( comm_allreduce( &agree, 1, 2, st ) ) return 1;
   l.on = agree > 0.5;
   return 0;
}

void p2p_release( P2PLink& l )
{
   for( int s = 0; s < 2; s++ )
   {
      for( int a = 0; a < 4; a++ )
	 if( l.peer[s].base[a] ) { cudaIpcCloseMemHandle( l.peer[s].base[a] ); l.peer[s].base[a] = 0; }
      if( l.peer[s].flags ) { cudaIpcCloseMemHandle( l.peer[s].flags ); l.peer[s].flags = 0; }
   }
   if( l.flags ) { cudaFree( l.flags ); l.flags = 0; }
   l.on = false;
   cudaGetLastError();
}

// push the face planes of `field` (one of my four arrays) into the neighbours' halo planes
int p2p_push_field( P2PLink& l, const Block& b, const double* field, cudaStream_t st )
{
   int id = -1;
   for( int a = 0; a < 4; a++ )
      if( l.mine[a] == field ) id = a;
   if( id < 0 ) return set_error( "p2p exchange: the field is not one of the block's solution arrays" );
   const long long nij = b.nij;
   const int nrun = b.sp == 1 ? 3 : 1;
   const size_t bytes = (size_t)( b.sp == 1 ? 2 : 6 ) * nij * sizeof( double );
   const long long pl = b.sp == 1 ? nij : 3 * nij;
   for( int side = 0; side < 2; side++ )
   {
      const P2PLink::Peer& p = l.peer[side];
      if( p.rank < 0 ) continue;
      // my low face -> the low neighbour's HIGH halo planes; my high face -> the high neighbour's LOW halo planes
      const long long ksend = side == 0 ? 2 : b.nk - 4, krecv = side == 0 ? p.nk - 2 : 0;
      const long long peer_npts = nij * p.nk;
      for( int r = 0; r < nrun; r++ )
      {
	 const double* src = field + ( b.sp == 1 ? r * b.sc : 0 ) + pl * ksend;
	 double* dst = p.base[id] + ( b.sp == 1 ? r * peer_npts : 0 ) + pl * krecv;
	 if( cudaMemcpyAsync( dst, src, bytes, cudaMemcpyDefault, st ) != cudaSuccess )
	    return set_error( "p2p exchange: cudaMemcpyAsync to the neighbour failed: %s", cudaGetErrorString( cudaGetLastError() ) );
      }
   }
   return 0;
}

// start of exchange n on the communication stream (behind the event of my face rows): tell the neighbours that their
// pushes may land, and wait for the same word from them
int p2p_open( P2PLink& l, cudaStream_t st )
{
   l.posted++;
   for( int side = 0; side < 2; side++ )
      if( l.peer[side].rank >= 0 )
	 k_set_flag<<<1, 1, 0, st>>>( l.peer[side].flags + 2 + ( side == 0 ? 1 : 0 ), l.posted ); // I am its high (low) neighbour
   for( int side = 0; side < 2; side++ )
      if( l.peer[side].rank >= 0 ) k_wait_flag<<<1, 1, 0, st>>>( l.flags + 2 + side, l.posted );
   count_launch();
   return check_launch( "k_set_flag / k_wait_flag" );
}

// behind the copies of one exchange: tell the neighbours that their halo planes are complete
int p2p_signal( P2PLink& l, cudaStream_t st )
{
   for( int side = 0; side < 2; side++ )
      if( l.peer[side].rank >= 0 )
	 k_set_flag<<<1, 1, 0, st>>>( l.peer[side].flags + ( side == 0 ? 1 : 0 ), l.posted ); // I am its high (low) neighbour
   count_launch();
   return check_launch( "k_set_flag" );
}

// on the compute stream: wait until both neighbours have delivered the planes of the current exchange
int p2p_wait( P2PLink& l, cudaStream_t st )
{
   for( int side = 0; side < 2; side++ )
      if( l.peer[side].rank >= 0 ) k_wait_flag<<<1, 1, 0, st>>>( l.flags + side, l.posted );
   count_launch();
   return check_launch( "k_wait_flag" );
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
