// Shared declarations of the sw4b200 CUDA library (sm_100a only).
#pragma once
#ifndef SW4B200_EMULATE
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <stddef.h>

namespace sw4b200 {

// SBP closure tables, device constant memory (reference: device-routines.C:41-44 dev_acof...).
// acof(k,q,m) = c_acof[(k-1)+6*(q-1)+48*(m-1)], bope(k,q) = c_bope[(k-1)+6*(q-1)].
// The library is built as ONE translation unit (sw4b200.cu includes every .cu), so these are
// plain definitions.
__constant__ double c_acof[384];
__constant__ double c_bope[48];
__constant__ double c_ghcof[6];
__constant__ double c_sbop[5];

// Geometry of one block in the reference's index convention (bounds include ghost points).
struct Block
{
   int ifirst, ilast, jfirst, jlast, kfirst, klast;
   int ni, nj, nk;	 // allocated extents (ni = row pitch in doubles)
   int nil;		 // points per row, ilast-ifirst+1: ni, or ni-1 when the block's rows are padded to an even pitch
			 // (grid blocks with odd ni: 16-byte aligned rows for the TMA-staged kernels); the pad column is
			 // never written by a kernel and never read into a result
   long long nij, npts;	 // plane size, block size (of the allocation)
   long long sc, sp;	 // component stride, point stride (corder=1: npts,1 ; corder=0: 1,3)
};

inline Block make_block( int corder, int ifirst, int ilast, int jfirst, int jlast, int kfirst, int klast, bool pad_even = false )
{
   Block b;
   b.ifirst = ifirst; b.ilast = ilast; b.jfirst = jfirst; b.jlast = jlast; b.kfirst = kfirst; b.klast = klast;
   b.ni = ilast - ifirst + 1; b.nj = jlast - jfirst + 1; b.nk = klast - kfirst + 1;
   b.nil = b.ni;
   if( pad_even && corder && ( b.ni & 1 ) ) b.ni++;
   b.nij = (long long)b.ni * b.nj;
   b.npts = b.nij * b.nk;
   if( corder ) { b.sc = b.npts; b.sp = 1; }
   else         { b.sc = 1;      b.sp = 3; }
   return b;
}

__host__ __device__ inline long long pidx( const Block& b, int i, int j, int k )
{
   return (long long)(i - b.ifirst) + (long long)b.ni * (j - b.jfirst) + b.nij * (k - b.kfirst);
}

struct Int6 { int v[6]; };
struct Int36 { int v[36]; };
struct Ptr6 { const double* p[6]; };

extern bool g_builtin_sbp_tables; // cart_v1.cu

// error plumbing (api.cu)
int set_error( const char* fmt, ... );
int check_launch( const char* what );
void count_launch( int n = 1 );
cudaStream_t as_stream( void* s );

// ---- launch wrappers implemented in the .cu files -------------------------------------------
// MODE_CORR forms uacc=(up-2u+um)/dt^2 on the fly and writes out-of-place; MODE_CORR_ACC reads a stored
// uacc array (in `u`) and corrects `up` in place; MODE_SHELL_DPDMT: uacc on the 2-point shell only.
enum RhsMode { MODE_LU = 0, MODE_PRED = 1, MODE_CORR = 2, MODE_CORR_ACC = 3, MODE_SHELL_DPDMT = 4 };

// optional per-kernel device timing (sw4b200_profile_*): records an event pair around a launch
struct ProfScope
{
   ProfScope( const char* name, cudaStream_t st );
   ~ProfScope();
   int slot;
   cudaStream_t st;
};

struct RhsArgs
{
   Block b;
   int nk;	    // global interior size in k
   int onesided4, onesided5;
   double* out;	    // lu (MODE_LU) or up_out (PRED/CORR)
   double* out2;    // PRED: uacc = (L(u)/h^2+fo)/rho (optional); SHELL_DPDMT: uacc
   const double *u, *um, *up; // PRED: u,um ; CORR: up,u,um (uacc formed on the fly); LU: u
   const double *mu, *la, *rho, *fo;
   const double *la2, *rhoi; // derived arrays of a grid block: 2 mu + lambda, 1 / rho (null: caller-owned arrays)
   const double *strx, *stry, *strz;
   double h, dt;
   // supergrid damping fused into MODE_CORR when sg_order != 0
   const double *dcx, *dcy, *dcz, *cox, *coy, *coz;
   double beta;
   int sg_order;
   int halo_lo, halo_hi; // SHELL_DPDMT: the k ghost planes of that side are halos whose interior is not to be touched
};

int launch_rhs_v1( RhsMode mode, const RhsArgs& a, cudaStream_t st );
int launch_rhs_v1_rows( RhsMode mode, const RhsArgs& a, int k_lo, int k_hi, cudaStream_t st );
int launch_shell_update( RhsMode mode, const RhsArgs& a, cudaStream_t st );

int launch_predfort( const Block& b, double* up, const double* u, const double* um, const double* lu,
		     const double* fo, const double* rho, double dt2, cudaStream_t st );
int launch_corrfort( const Block& b, double* up, const double* lu, const double* fo, const double* rho,
		     double dt4, cudaStream_t st );
int launch_dpdmt( long long n, const double* up, const double* u, const double* um, double* u2, double dt2i,
		  cudaStream_t st );
int launch_addsgd( int order, const Block& b, double* up, const double* u, const double* um, const double* rho,
		   const double* dcx, const double* dcy, const double* dcz,
		   const double* strx, const double* stry, const double* strz,
		   const double* cox, const double* coy, const double* coz, double beta, cudaStream_t st );
int launch_addsgd_box( int order, const Block& b, const Int6& box, double* up, const double* u, const double* um,
		       const double* rho, const double* dcx, const double* dcy, const double* dcz, const double* strx,
		       const double* stry, const double* strz, const double* cox, const double* coy,
		       const double* coz, double beta, cudaStream_t st );
int launch_addsgd4_fast( const Block& b, const Int6& box, double* up, const double* u, const double* um,
			 const double* rho, const double* dcx, const double* dcy, const double* dcz, const double* strx,
			 const double* stry, const double* strz, const double* cox, const double* coy, const double* coz,
			 double beta, cudaStream_t st );
int launch_addsgd4_zonly( const Block& b, const Int6& box, double* up, const double* u, const double* um, const double* rho,
			  const double* dcz, const double* strz, const double* cox, const double* coy, double beta, cudaStream_t st );
int launch_bcfortsg( const Block& b, const Int36& wind, int nx, int ny, int nz, double* u, double h,
		     const Int6& bccnd, const double* mu, const double* la, const Ptr6& bforce,
		     const double* strx, const double* stry, cudaStream_t st );
int launch_add_point_forces( int corder, long long npts, double* up, const double* rho, int n,
			     const long long* pidx, const double* f, double factor, cudaStream_t st,
			     double* up2 = 0, double factor2 = 0, long long nij = 0, int kplane_lo = 0, int kplane_hi = 0 );
int launch_gather_points( int corder, long long npts, const double* u, int n, const long long* pidx,
			  double* out, cudaStream_t st );
cudaStream_t aux_fork( cudaStream_t st ); // api.cu: the library's auxiliary stream, ordered after st (0: not available)
int aux_join( cudaStream_t st );
int launch_fill_profile( const Block& b, double* a, const double* prof, cudaStream_t st );
int launch_derive_materials( long long n, const double* mu, const double* la, const double* rho, double* la2, double* rhoi, cudaStream_t st );
int launch_halo_copy( const Block& b, double* field, int kplane, double* buf, int pack, cudaStream_t st );

// curvilinear (curvilinear.cu)
int launch_rhs4sgcurv( const Block& b, const double* u, const double* mu, const double* la, const double* met,
		       const double* jac, double* lu, int onesided4, const double* strx, const double* stry,
		       cudaStream_t st, double* flux_scratch = 0 ); // flux_scratch: 9*npts doubles (two-sweep interior rows) or null
int launch_addsgdc( int order, const Block& b, double* up, const double* u, const double* um, const double* rho,
		    const double* dcx, const double* dcy, const double* strx, const double* stry,
		    const double* jac, const double* cox, const double* coy, double beta, cudaStream_t st );
int launch_freesurfcurvisg( const Block& b, int nz, int side, double* u, const double* mu, const double* la,
			    const double* met, const double* forcing, const double* strx, const double* stry,
			    cudaStream_t st );
int launch_enforce_cart_topo( int corder, double* ucart, const Block& bc, double* ucurv, const Block& bt,
			      cudaStream_t st );

} // namespace sw4b200
