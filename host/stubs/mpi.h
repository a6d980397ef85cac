/*
 * TEST INFRASTRUCTURE ONLY (oracle/): single-rank stand-in for <mpi.h>.
 *
 * The reference (geodynamics/sw4lite) needs MPI to compile; this image has none.
 * This header gives the ~45 MPI symbols the reference uses the semantics of a
 * one-rank communicator: collectives copy send->recv, neighbours are
 * MPI_PROC_NULL, point-to-point to self goes through a small in-process mailbox.
 * It is written from the MPI standard's API, not from any MPI implementation,
 * and is used only to build oracle/_ref (the reference CPU oracle).
 */
#ifndef SW4B200_ORACLE_MPI_STUB_H
#define SW4B200_ORACLE_MPI_STUB_H

#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <ctime>
#include <unistd.h>
#include <vector>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Group;
typedef int MPI_Request;
typedef int MPI_Aint;
struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR; };

#define MPI_SUCCESS 0
#define MPI_COMM_WORLD 1
#define MPI_COMM_NULL 0
#define MPI_PROC_NULL (-2)
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_MAX_PROCESSOR_NAME 256
#define MPI_WTIME_IS_GLOBAL 4
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)

/* datatypes: value = size in bytes for the basic ones; derived types get ids >= 1000 */
#define MPI_CHAR 1
#define MPI_BYTE 1
#define MPI_INT 4
#define MPI_FLOAT 4004 /* distinct id, size 4 */
#define MPI_DOUBLE 8
#define MPI_LONG_LONG 8008
#define MPI_FLOAT_SW4 MPI_DOUBLE

#define MPI_MAX 1
#define MPI_MIN 2
#define MPI_SUM 3

namespace sw4b200_mpistub {
struct Msg { int tag; std::vector<char> data; };
inline std::vector<Msg>& mailbox() { static std::vector<Msg> m; return m; }
/* receives posted before their message was sent (Parallel_IO::write_array posts MPI_Irecv, then MPI_Send to itself) */
struct PendingRecv { void* buf; size_t nbytes; int tag; bool done; };
inline std::vector<PendingRecv>& pending() { static std::vector<PendingRecv> p; return p; }
struct VecType { int count, blocklen, stride; MPI_Datatype base; };
inline std::vector<VecType>& vtypes() { static std::vector<VecType> v; return v; }
inline size_t type_size(MPI_Datatype t)
{
   if( t == MPI_FLOAT ) return 4;
   if( t == MPI_LONG_LONG ) return 8;
   if( t >= 1000 && t < 4000 )
   {
      const VecType& v = vtypes()[t-1000];
      return (size_t)v.count*v.blocklen*type_size(v.base);
   }
   return (size_t)t;
}
}

inline int MPI_Init( int*, char*** ) { return MPI_SUCCESS; }
inline int MPI_Finalize() { return MPI_SUCCESS; }
inline int MPI_Comm_rank( MPI_Comm, int* r ) { *r = 0; return MPI_SUCCESS; }
inline int MPI_Comm_size( MPI_Comm, int* s ) { *s = 1; return MPI_SUCCESS; }
inline int MPI_Barrier( MPI_Comm ) { return MPI_SUCCESS; }
inline int MPI_Abort( MPI_Comm, int code ) { fprintf(stderr,"MPI_Abort(%d) [stub]\n",code); exit(code ? code : 1); return 0; }
inline double MPI_Wtime()
{
   struct timespec ts; clock_gettime( CLOCK_MONOTONIC, &ts );
   return ts.tv_sec + 1e-9*ts.tv_nsec;
}
inline double MPI_Wtick() { return 1e-9; }
inline int MPI_Get_processor_name( char* name, int* len )
{
   if( gethostname( name, MPI_MAX_PROCESSOR_NAME ) != 0 ) strcpy( name, "localhost" );
   *len = (int)strlen(name); return MPI_SUCCESS;
}
inline int MPI_Comm_get_attr( MPI_Comm, int, void* val, int* flag ) { (void)val; *flag = 0; return MPI_SUCCESS; }

/* collectives over one rank: copy */
inline int MPI_Allreduce( const void* s, void* r, int n, MPI_Datatype t, MPI_Op, MPI_Comm )
{ if( s != r ) memcpy( r, s, n*sw4b200_mpistub::type_size(t) ); return MPI_SUCCESS; }
inline int MPI_Reduce( const void* s, void* r, int n, MPI_Datatype t, MPI_Op, int, MPI_Comm )
{ if( s != r ) memcpy( r, s, n*sw4b200_mpistub::type_size(t) ); return MPI_SUCCESS; }
inline int MPI_Bcast( void*, int, MPI_Datatype, int, MPI_Comm ) { return MPI_SUCCESS; }
inline int MPI_Gather( const void* s, int n, MPI_Datatype t, void* r, int, MPI_Datatype, int, MPI_Comm )
{ if( s != r ) memcpy( r, s, n*sw4b200_mpistub::type_size(t) ); return MPI_SUCCESS; }
inline int MPI_Allgather( const void* s, int n, MPI_Datatype t, void* r, int, MPI_Datatype, MPI_Comm )
{ if( s != r ) memcpy( r, s, n*sw4b200_mpistub::type_size(t) ); return MPI_SUCCESS; }

/* topology: 1x1 process grid, no neighbours */
inline int MPI_Cart_create( MPI_Comm, int, const int*, const int*, int, MPI_Comm* c ) { *c = 2; return MPI_SUCCESS; }
inline int MPI_Cart_get( MPI_Comm, int nd, int* dims, int* periods, int* coords )
{ for( int d=0; d<nd; d++ ){ dims[d]=1; periods[d]=0; coords[d]=0; } return MPI_SUCCESS; }
inline int MPI_Cart_shift( MPI_Comm, int, int, int* lo, int* hi ) { *lo = MPI_PROC_NULL; *hi = MPI_PROC_NULL; return MPI_SUCCESS; }
inline int MPI_Comm_split( MPI_Comm, int, int, MPI_Comm* c ) { *c = 3; return MPI_SUCCESS; }
inline int MPI_Comm_group( MPI_Comm, MPI_Group* g ) { *g = 1; return MPI_SUCCESS; }
inline int MPI_Group_incl( MPI_Group, int, const int*, MPI_Group* g ) { *g = 2; return MPI_SUCCESS; }
inline int MPI_Group_free( MPI_Group* ) { return MPI_SUCCESS; }
inline int MPI_Comm_create( MPI_Comm, MPI_Group, MPI_Comm* c ) { *c = 4; return MPI_SUCCESS; }

/* derived datatypes (only MPI_Type_vector is used) */
inline int MPI_Type_vector( int count, int blocklen, int stride, MPI_Datatype base, MPI_Datatype* nt )
{
   sw4b200_mpistub::VecType v = {count, blocklen, stride, base};
   sw4b200_mpistub::vtypes().push_back(v);
   *nt = 1000 + (int)sw4b200_mpistub::vtypes().size() - 1;
   return MPI_SUCCESS;
}
inline int MPI_Type_commit( MPI_Datatype* ) { return MPI_SUCCESS; }

/* point to point: neighbours are always MPI_PROC_NULL in the halo code; messages
   to self (rank 0) are queued by tag (Parallel_IO does this with basic types). */
inline int MPI_Send( const void* buf, int n, MPI_Datatype t, int dest, int tag, MPI_Comm )
{
   if( dest == MPI_PROC_NULL ) return MPI_SUCCESS;
   {
      std::vector<sw4b200_mpistub::PendingRecv>& pr = sw4b200_mpistub::pending();
      for( size_t i=0; i<pr.size(); i++ )
	 if( !pr[i].done && ( pr[i].tag == MPI_ANY_TAG || pr[i].tag == tag ) )
	 {
	    size_t nb = n*sw4b200_mpistub::type_size(t);
	    if( pr[i].nbytes < nb ) nb = pr[i].nbytes;
	    memcpy( pr[i].buf, buf, nb );
	    pr[i].done = true;
	    return MPI_SUCCESS;
	 }
   }
   sw4b200_mpistub::Msg m; m.tag = tag;
   m.data.assign( (const char*)buf, (const char*)buf + n*sw4b200_mpistub::type_size(t) );
   sw4b200_mpistub::mailbox().push_back( m );
   return MPI_SUCCESS;
}
inline int MPI_Recv( void* buf, int n, MPI_Datatype t, int src, int tag, MPI_Comm, MPI_Status* st )
{
   if( src == MPI_PROC_NULL ) return MPI_SUCCESS;
   std::vector<sw4b200_mpistub::Msg>& mb = sw4b200_mpistub::mailbox();
   for( size_t i=0; i<mb.size(); i++ )
      if( tag == MPI_ANY_TAG || mb[i].tag == tag )
      {
	 size_t nb = n*sw4b200_mpistub::type_size(t);
	 if( mb[i].data.size() < nb ) nb = mb[i].data.size();
	 memcpy( buf, mb[i].data.data(), nb );
	 if( st ){ st->MPI_SOURCE = 0; st->MPI_TAG = mb[i].tag; st->MPI_ERROR = 0; }
	 mb.erase( mb.begin()+i );
	 return MPI_SUCCESS;
      }
   fprintf( stderr, "MPI stub: MPI_Recv with no matching message (tag %d)\n", tag );
   exit(3);
   return 1;
}
inline int MPI_Isend( const void* buf, int n, MPI_Datatype t, int dest, int tag, MPI_Comm c, MPI_Request* r )
{ *r = 0; return MPI_Send( buf, n, t, dest, tag, c ); }
inline int MPI_Irecv( void* buf, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request* r )
{
   *r = 0;
   if( src == MPI_PROC_NULL ) return MPI_SUCCESS;
   std::vector<sw4b200_mpistub::Msg>& mb = sw4b200_mpistub::mailbox();
   for( size_t i=0; i<mb.size(); i++ )
      if( tag == MPI_ANY_TAG || mb[i].tag == tag ) return MPI_Recv( buf, n, t, src, tag, c, 0 );
   /* not sent yet: completed by the matching MPI_Send */
   sw4b200_mpistub::PendingRecv p = { buf, n*sw4b200_mpistub::type_size(t), tag, false };
   sw4b200_mpistub::pending().push_back( p );
   *r = (int)sw4b200_mpistub::pending().size();
   return MPI_SUCCESS;
}
inline int MPI_Wait( MPI_Request* r, MPI_Status* )
{
   if( r && *r > 0 && *r <= (int)sw4b200_mpistub::pending().size() && !sw4b200_mpistub::pending()[*r-1].done )
   {
      fprintf( stderr, "MPI stub: MPI_Wait on a receive whose message was never sent (tag %d)\n", sw4b200_mpistub::pending()[*r-1].tag );
      exit(3);
   }
   return MPI_SUCCESS;
}
inline int MPI_Waitall( int n, MPI_Request* r, MPI_Status* ) { for( int i=0; i<n; i++ ) MPI_Wait( r+i, 0 ); return MPI_SUCCESS; }
inline int MPI_Sendrecv( const void* sb, int sn, MPI_Datatype st, int dest, int stag,
			 void* rb, int rn, MPI_Datatype rt, int src, int rtag, MPI_Comm c, MPI_Status* status )
{
   if( dest != MPI_PROC_NULL ) MPI_Send( sb, sn, st, dest, stag, c );
   if( src  != MPI_PROC_NULL ) MPI_Recv( rb, rn, rt, src, rtag, c, status );
   return MPI_SUCCESS;
}
#endif
