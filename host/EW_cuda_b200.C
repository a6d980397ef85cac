// Host binding: the reference's GPU operator surface re-targeted to libsw4b200.so.
//
// sw4lite reaches its GPU path through EW::*CU members that EW::timesteploop branches to under
// m_cuobj->has_gpu() (reference EW.C:2536-2842).  The reference defines them in src/EW_cuda.C (+ the kernels
// of src/device-routines.C and the EWCuda class of src/EWCuda.C).  This translation unit defines the SAME
// members, with the same signatures (src/EW.h:86-127,296-330,494-504, src/EWCuda.h:39-53), on top of the
// C-ABI of include/sw4b200.h.  host/build_host.py compiles the reference's own host sources where they lie
// under /root/reference/src (main, parser, set-up, sources, receivers, error norms -- unmodified, never copied)
// with -DSW4_CUDA, leaves out EW_cuda.C / device-routines.C / EWCuda.C, and links this file and libsw4b200.so
// instead: `sw4lite_b200 file.in` is then the reference program running on the B200 kernels of this repository.
//
// The device-resident state of every grid lives in a sw4b200_grid object (the grid-block level of the C-ABI): rows padded
// to an even pitch (so the reference's odd-ni grids run on the TMA-staged kernels), stored acceleration, sparse forcing,
// box-decomposed supergrid damping.  One time step is two fused passes per grid (sw4b200_grid_predictor_dev /
// _corrector_dev) plus the boundary conditions; the reference's Sarray device copies are only written when the host is
// about to read them (checkpoints, the final error norm).
//
// There is no CPU fallback here: when libsw4b200 cannot initialise a device the program stops.
#include "mpi.h"
#include "sw4.h"
#include "EW.h"
#include "EWCuda.h"
#include "Sarray.h"
#include "GridPointSource.h"
#include "TimeSeries.h"
#include "Source.h"
#include "CheckPoint.h"
#include "../include/sw4b200.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>
using namespace std;

namespace {

void ok( int rc, const char* what )
{
   if( rc != 0 )
   {
      fprintf( stderr, "sw4lite_b200: %s failed: %s\n", what, sw4b200_last_error() );
      exit( 1 );
   }
}
#define B200( call ) ok( ( call ), #call )

void* stream0() { return sw4b200_stream( 0 ); }

// unique source points per grid and their amplitudes (EW::Force sums the sources that share a grid point,
// reference EW.C:3092-3121; the dense F array is never touched: forcing is injected sparsely)
struct GridSources
{
   int n;
   vector<int> ijk; // the unique source points of the grid
   long long* d_idx; // flat indices in the reference's (unpadded) layout, for the operator-level calls
   double *d_f, *h_f;
};
vector<GridSources> g_src;
bool g_src_built = false;

vector<double*> g_lu; // scratch L(u) of the curvilinear grid (unfused operator calls only)

// one device-resident grid block per grid; the time loop's phases run on them
vector<sw4b200_grid*> g_grid;
bool g_imported = false;  // the initial U, Um have been moved into the grid blocks
bool g_corr_done = false; // the corrector of the current step has run (the next enforceBCCU finishes the step)
double g_t_step = 0;	  // time at the start of the current step (ForceCU)
vector<Sarray*> g_sUp, g_sU; // the reference's arrays holding Up / U of the current step (export targets)

// read-only view of CheckPoint::timeToWrite (CheckPoint.C:205-230), which itself advances the check point's state
struct CheckPointPeek : public CheckPoint
{
   static bool due( const CheckPoint* c, float_sw4 time, int cycle, float_sw4 dt )
   {
      const CheckPointPeek* p = static_cast<const CheckPointPeek*>( c );
      bool do_it = cycle == p->mWritingCycle;
      if( p->mCycleInterval != 0 && cycle % p->mCycleInterval == 0 && time >= p->mStartTime ) do_it = true;
      if( p->mTime > 0.0 && p->mTime <= time + dt * 0.5 && !p->m_time_done ) do_it = true;
      if( p->mTimeInterval != 0.0 && p->mNextTime <= time + dt * 0.5 && time >= p->mStartTime ) do_it = true;
      return do_it;
   }
};

double** g_bforce_host_ptrs( vector<double**>& dev_BCForcing, int g ) { return dev_BCForcing[g]; }

} // namespace

// --------------------------------------------------------------------------------------------- EWCuda
EWCuda::EWCuda( int ndevice, int nstream ) : m_nstream( nstream ), m_ndevice( ndevice ), m_active_gpu( 0 )
{
   m_stream = 0;
   if( ndevice > 0 && nstream > 0 )
   {
      // one ordered stream: the reference's two-stream choreography (boundary strips on stream 0, centre on
      // stream 1, EW.C:2561-2566) has no counterpart here, every phase is one call
      m_stream = new cudaStream_t[nstream];
      for( int s = 0; s < nstream; s++ ) m_stream[s] = (cudaStream_t)stream0();
   }
}
EWCuda::~EWCuda() { delete[] m_stream; }
void EWCuda::reset_gpu()
{
   if( m_ndevice > 0 ) { sw4b200_sync_device(); sw4b200_finalize(); }
}
void EWCuda::initialize_gpu( int myrank )
{
   if( m_ndevice > 0 )
   {
      m_active_gpu = myrank % m_ndevice;
      B200( sw4b200_init( m_active_gpu ) );
      for( int s = 0; s < m_nstream; s++ ) m_stream[s] = (cudaStream_t)stream0();
   }
}
void EWCuda::sync_stream( int st ) { if( m_ndevice > 0 ) B200( sw4b200_sync_stream( 0 ) ); }
void EWCuda::sync_device() { if( m_ndevice > 0 ) B200( sw4b200_sync_device() ); }

extern "C" void setupgpu( int verbose ) { (void)verbose; }

// --------------------------------------------------------------------------------------------- set-up
void EW::find_cuda_device()
{
   m_ndevice = sw4b200_device_count();
   if( m_ndevice <= 0 )
   {
      cerr << "sw4lite_b200: no CUDA device; this build has no CPU path" << endl;
      exit( 1 );
   }
   if( m_myrank == 0 ) cout << m_ndevice << " CUDA device(s), " << sw4b200_version() << endl;
   B200( sw4b200_init( m_myrank % m_ndevice ) );
   for( int d = 0; d < 3; d++ ) m_gpu_gridsize[d] = 1;
   m_cuobj = new EWCuda( m_ndevice, 2 );
}

void EW::setupSBPCoeff()
{
   if( mVerbose >= 1 && m_myrank == 0 ) cout << "Setting up SBP boundary stencils" << endl;
   GetStencilCoefficients( m_acof, m_ghcof, m_bope, m_sbop );
   B200( sw4b200_copy_stencilcoefficients( m_acof, m_ghcof, m_bope, m_sbop ) );
}

static double* to_device( const double* h, size_t n )
{
   double* d = (double*)sw4b200_malloc( n * sizeof( double ) );
   if( !d ) ok( 1, "sw4b200_malloc" );
   B200( sw4b200_memcpy_h2d( d, h, n * sizeof( double ), 0 ) );
   B200( sw4b200_sync_stream( 0 ) );
   return d;
}

void EW::copy_supergrid_arrays_to_device()
{
   const int ng = mNumberOfGrids;
   dev_sg_str_x.resize( ng ); dev_sg_str_y.resize( ng ); dev_sg_str_z.resize( ng );
   dev_sg_dc_x.resize( ng ); dev_sg_dc_y.resize( ng ); dev_sg_dc_z.resize( ng );
   dev_sg_corner_x.resize( ng ); dev_sg_corner_y.resize( ng ); dev_sg_corner_z.resize( ng );
   if( m_ndevice <= 0 ) return;
   for( int g = 0; g < ng; g++ )
   {
      const size_t ni = m_iEnd[g] - m_iStart[g] + 1, nj = m_jEnd[g] - m_jStart[g] + 1, nk = m_kEnd[g] - m_kStart[g] + 1;
      dev_sg_str_x[g] = to_device( m_sg_str_x[g], ni ); dev_sg_str_y[g] = to_device( m_sg_str_y[g], nj );
      dev_sg_str_z[g] = to_device( m_sg_str_z[g], nk );
      dev_sg_dc_x[g] = to_device( m_sg_dc_x[g], ni ); dev_sg_dc_y[g] = to_device( m_sg_dc_y[g], nj );
      dev_sg_dc_z[g] = to_device( m_sg_dc_z[g], nk );
      dev_sg_corner_x[g] = to_device( m_sg_corner_x[g], ni ); dev_sg_corner_y[g] = to_device( m_sg_corner_y[g], nj );
      dev_sg_corner_z[g] = to_device( m_sg_corner_z[g], nk );
   }
}

// materials, metric and the supergrid arrays are uploaded into the grid blocks by setup_device_communication_array (the last
// set-up call before the time loop, EW.C:2524); the reference's own device copies are not needed
void EW::copy_material_to_device() {}

void EW::copy_bcforcing_arrays_to_device()
{
   if( m_ndevice <= 0 ) return;
   dev_BCForcing.resize( mNumberOfGrids );
   for( int g = 0; g < mNumberOfGrids; g++ )
   {
      dev_BCForcing[g] = new float_sw4*[6];
      for( int side = 0; side < 6; side++ )
      {
	 dev_BCForcing[g][side] = NULL;
	 if( m_bcType[g][side] == bStressFree || m_bcType[g][side] == bDirichlet || m_bcType[g][side] == bSuperGrid )
	 {
	    // zero forcing (what cartesian_bc_forcingCU memsets every step, EW_cuda.C:1100-1137); Dirichlet sides of the
	    // point-source test are refreshed by cartesian_bc_forcingCU below
	    const size_t nb = 3 * (size_t)m_NumberOfBCPoints[g][side] * sizeof( double );
	    dev_BCForcing[g][side] = (double*)sw4b200_malloc( nb );
	    if( !dev_BCForcing[g][side] ) ok( 1, "sw4b200_malloc" );
	    B200( sw4b200_memset_zero( dev_BCForcing[g][side], nb, 0 ) );
	 }
      }
   }
}
// boundary types and windows are passed by value to sw4b200_bcfortsg: nothing to copy
void EW::copy_bctype_arrays_to_device() {}
void EW::copy_bndrywindow_arrays_to_device() {}
static void build_sources( EW* ew );

// Called right before the time loop (EW.C:2524), when dt, the boundary windows, the supergrid arrays and the sources are
// final: build the device-resident grid blocks.  (The x-y halo swap the reference sets up here has no neighbour: one rank per
// GPU in this build; multi-GPU runs decompose in z through sw4b200_grid_exchange_*, host/slab_driver.C.)
void EW::setup_device_communication_array()
{
   if( m_ndevice <= 0 ) return;
   build_sources( this );
   g_grid.assign( mNumberOfGrids, (sw4b200_grid*)0 );
   for( int g = 0; g < mNumberOfGrids; g++ )
   {
      sw4b200_grid_desc d;
      memset( &d, 0, sizeof( d ) );
      d.corder = Sarray::m_corder ? 1 : 0;
      d.ifirst = m_iStart[g]; d.ilast = m_iEnd[g]; d.jfirst = m_jStart[g]; d.jlast = m_jEnd[g]; d.kfirst = m_kStart[g]; d.klast = m_kEnd[g];
      d.nx = m_global_nx[g]; d.ny = m_global_ny[g]; d.nz = m_global_nz[g];
      d.h = mGridSize[g]; d.dt = mDt;
      for( int s = 0; s < 6; s++ ) { d.onesided[s] = m_onesided[g][s]; d.bctype[s] = (int)m_bcType[g][s]; }
      for( int s = 0; s < 36; s++ ) d.wind[s] = m_BndryWindow[g][s];
      d.sg_order = m_use_supergrid ? m_sg_damping_order : 0;
      d.beta = m_use_supergrid ? m_supergrid_damping_coefficient : 0.0;
      d.curvilinear = topographyExists() && g == mNumberOfGrids - 1;
      sw4b200_grid* G = sw4b200_grid_create( &d );
      if( !G ) ok( 1, "sw4b200_grid_create" );
      g_grid[g] = G;
      B200( sw4b200_grid_upload( G, "mu", mMu[g].c_ptr() ) );
      B200( sw4b200_grid_upload( G, "lambda", mLambda[g].c_ptr() ) );
      B200( sw4b200_grid_upload( G, "rho", mRho[g].c_ptr() ) );
      if( d.curvilinear )
      {
	 B200( sw4b200_grid_upload( G, "jac", mJ.c_ptr() ) );
	 B200( sw4b200_grid_upload( G, "metric", mMetric.c_ptr() ) );
      }
      B200( sw4b200_grid_upload( G, "strx", m_sg_str_x[g] ) ); B200( sw4b200_grid_upload( G, "stry", m_sg_str_y[g] ) );
      B200( sw4b200_grid_upload( G, "strz", m_sg_str_z[g] ) );
      B200( sw4b200_grid_upload( G, "dcx", m_sg_dc_x[g] ) ); B200( sw4b200_grid_upload( G, "dcy", m_sg_dc_y[g] ) );
      B200( sw4b200_grid_upload( G, "dcz", m_sg_dc_z[g] ) );
      B200( sw4b200_grid_upload( G, "cox", m_sg_corner_x[g] ) ); B200( sw4b200_grid_upload( G, "coy", m_sg_corner_y[g] ) );
      B200( sw4b200_grid_upload( G, "coz", m_sg_corner_z[g] ) );
      if( g_src[g].n > 0 ) B200( sw4b200_grid_set_source_points( G, g_src[g].n, g_src[g].ijk.data() ) );
   }
}
void EW::pack_HaloArrayCU( Sarray&, int, int ) {}
void EW::unpack_HaloArrayCU( Sarray&, int, int ) {}
void EW::communicate_arrayCU( Sarray&, int, int ) {}
void EW::pack_HaloArrayCU_X( Sarray&, int, int ) {}
void EW::unpack_HaloArrayCU_X( Sarray&, int, int ) {}
void EW::communicate_arrayCU_X( Sarray&, int, int ) {}
void EW::pack_HaloArrayCU_Y( Sarray&, int, int ) {}
void EW::unpack_HaloArrayCU_Y( Sarray&, int, int ) {}
void EW::communicate_arrayCU_Y( Sarray&, int, int ) {}

bool EW::check_for_nan_GPU( vector<Sarray>& a_U, int verbose, string name )
{
   (void)a_U; (void)verbose;
   // the state lives in the grid blocks: whatever array the loop asks about, the three time levels are checked
   bool clean = true;
   const char* fields[3] = { "Um", "U", "Up" };
   for( int g = 0; g < mNumberOfGrids && !g_grid.empty(); g++ )
      for( int f = 0; f < 3; f++ )
      {
	 vector<double> h( sw4b200_grid_array_size( g_grid[g], fields[f] ) );
	 B200( sw4b200_grid_download( g_grid[g], fields[f], h.data() ) );
	 size_t cnt = 0, first = 0;
	 for( size_t q = 0; q < h.size(); q++ )
	    if( h[q] != h[q] ) { if( cnt == 0 ) first = q; cnt++; }
	 if( cnt )
	 {
	    cout << "grid " << g << " array " << fields[f] << " (asked for " << name << ") found " << cnt << " nans. First nan at linear index "
		 << first << endl;
	    clean = false;
	 }
      }
   return clean;
}

// --------------------------------------------------------------------------------------------- forcing
// The reference evaluates the source time functions on the device: under SW4_CUDA the function pointers of a
// GridPointSource are set by its __device__ init_dev() only (GridPointSource.C:78-80,826-833), on the device
// copies that EW::copy_point_sources_to_gpu builds (EW.C:6435-6466).  These two small kernels call the
// reference's own __device__ members (GridPointSource.o / time_functions_cu.o, linked as relocatable device
// code) the way init_forcing_dev / forcing_dev do (device-routines.C:8306-8343), but produce the amplitudes of
// the unique source points as a compact table instead of writing into a dense F array.
__global__ void b200_init_sources( GridPointSource* s, int n )
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if( i < n ) s[i].init_dev();
}
__global__ void b200_eval_forces( float_sw4 t, const GridPointSource* s, const int* ident, int nu, const int* slot, int tt,
				  double* f )
{
   const int r = blockIdx.x * blockDim.x + threadIdx.x;
   if( r >= nu ) return;
   float_sw4 sum[3] = { 0, 0, 0 }, v[3];
   for( int i = ident[r]; i < ident[r + 1]; i++ ) // EW::Force sums the sources sharing a grid point (EW.C:3092-3121)
   {
      if( tt ) s[i].getFxyztt( t, v );
      else s[i].getFxyz( t, v );
      sum[0] += v[0]; sum[1] += v[1]; sum[2] += v[2];
   }
   const int q = slot[r];
   f[3 * q] = sum[0]; f[3 * q + 1] = sum[1]; f[3 * q + 2] = sum[2];
}

void EW::init_point_sourcesCU()
{
   const int n = (int)m_point_sources.size();
   if( n == 0 ) return;
   b200_init_sources<<<( n + 127 ) / 128, 128, 0, (cudaStream_t)stream0()>>>( dev_point_sources, n );
   cudaError_t e = cudaGetLastError();
   if( e != cudaSuccess ) { fprintf( stderr, "sw4lite_b200: init_point_sourcesCU: %s\n", cudaGetErrorString( e ) ); exit( 1 ); }
   B200( sw4b200_sync_stream( 0 ) );
}

static int* g_d_slot = 0;
static double* g_d_fall = 0;
static int g_nunique = 0;

static void build_sources( EW* ew )
{
   if( g_src_built ) return;
   g_src.assign( ew->mNumberOfGrids, GridSources() );
   for( int g = 0; g < ew->mNumberOfGrids; g++ ) { g_src[g].n = 0; g_src[g].d_idx = 0; g_src[g].d_f = 0; g_src[g].h_f = 0; }
   vector<vector<long long> > idx( ew->mNumberOfGrids );
   vector<vector<int> > uniq( ew->mNumberOfGrids );
   const int nu = (int)ew->m_identsources.size() - 1;
   g_nunique = nu > 0 ? nu : 0;
   for( int r = 0; r < nu; r++ )
   {
      const int s0 = ew->m_identsources[r];
      const GridPointSource* p = ew->m_point_sources[s0];
      const int g = p->m_grid;
      const long long ni = ew->m_iEnd[g] - ew->m_iStart[g] + 1, nj = ew->m_jEnd[g] - ew->m_jStart[g] + 1;
      idx[g].push_back( ( p->m_i0 - ew->m_iStart[g] ) + ni * ( p->m_j0 - ew->m_jStart[g] ) + ni * nj * ( p->m_k0 - ew->m_kStart[g] ) );
      uniq[g].push_back( r );
      g_src[g].ijk.push_back( p->m_i0 ); g_src[g].ijk.push_back( p->m_j0 ); g_src[g].ijk.push_back( p->m_k0 );
   }
   if( nu > 0 )
   {
      // compact table of amplitudes, the unique points of grid 0 first, then grid 1, ...
      vector<int> slot( nu );
      g_d_fall = (double*)sw4b200_malloc( 3 * (size_t)nu * sizeof( double ) );
      g_d_slot = (int*)sw4b200_malloc( nu * sizeof( int ) );
      if( !g_d_fall || !g_d_slot ) ok( 1, "allocation of the source tables" );
      int start = 0;
      for( int g = 0; g < ew->mNumberOfGrids; g++ )
      {
	 GridSources& S = g_src[g];
	 S.n = (int)idx[g].size();
	 if( S.n == 0 ) continue;
	 for( int m = 0; m < S.n; m++ ) slot[uniq[g][m]] = start + m;
	 S.d_f = g_d_fall + 3 * (size_t)start;
	 S.d_idx = (long long*)sw4b200_malloc( S.n * sizeof( long long ) );
	 if( !S.d_idx ) ok( 1, "allocation of the source tables" );
	 B200( sw4b200_memcpy_h2d( S.d_idx, idx[g].data(), S.n * sizeof( long long ), 0 ) );
	 B200( sw4b200_sync_stream( 0 ) );
	 start += S.n;
      }
      B200( sw4b200_memcpy_h2d( g_d_slot, slot.data(), nu * sizeof( int ), 0 ) );
      B200( sw4b200_sync_stream( 0 ) );
   }
   g_src_built = true;
}

void EW::ForceCU( float_sw4 t, Sarray* dev_F, bool tt, int st )
{
   (void)dev_F; (void)st;
   if( !tt ) g_t_step = t;
   if( !g_src_built ) build_sources( this );
   if( g_nunique == 0 ) return;
   b200_eval_forces<<<( g_nunique + 127 ) / 128, 128, 0, (cudaStream_t)stream0()>>>( t, dev_point_sources, dev_identsources, g_nunique,
										   g_d_slot, tt ? 1 : 0, g_d_fall );
   cudaError_t e = cudaGetLastError();
   if( e != cudaSuccess ) { fprintf( stderr, "sw4lite_b200: ForceCU: %s\n", cudaGetErrorString( e ) ); exit( 1 ); }
}

static void inject( EW* ew, int g, double* up, const double* rho, double factor )
{
   const GridSources& S = g_src[g];
   if( !g_src_built || S.n == 0 ) return;
   const size_t npts = (size_t)( ew->m_iEnd[g] - ew->m_iStart[g] + 1 ) * ( ew->m_jEnd[g] - ew->m_jStart[g] + 1 ) *
		       ( ew->m_kEnd[g] - ew->m_kStart[g] + 1 );
   B200( sw4b200_add_point_forces( Sarray::m_corder, npts, up, rho, S.n, S.d_idx, S.d_f, factor, stream0() ) );
}

static double* lu_scratch( EW* ew, int g )
{
   if( (int)g_lu.size() < ew->mNumberOfGrids ) g_lu.resize( ew->mNumberOfGrids, (double*)0 );
   if( !g_lu[g] )
   {
      const size_t n = 3 * (size_t)( ew->m_iEnd[g] - ew->m_iStart[g] + 1 ) * ( ew->m_jEnd[g] - ew->m_jStart[g] + 1 ) *
		       ( ew->m_kEnd[g] - ew->m_kStart[g] + 1 );
      g_lu[g] = (double*)sw4b200_malloc( n * sizeof( double ) );
      if( !g_lu[g] ) ok( 1, "sw4b200_malloc" );
      B200( sw4b200_memset_zero( g_lu[g], n * sizeof( double ), stream0() ) );
   }
   return g_lu[g];
}

// --------------------------------------------------------------------------------------------- time step
#define BOUNDS( g ) m_iStart[g], m_iEnd[g], m_jStart[g], m_jEnd[g], m_kStart[g], m_kEnd[g]

// ---- grid-block time step.  EW::timesteploop (EW.C:2527-2842) calls, in this order: ForceCU(t), RHSPredCU_boundary,
// RHSPredCU_center, halo members, cartesian_bc_forcingCU + enforceBCCU, ForceCU(t, tt), evalDpDmInTimeCU, RHSCorrCU_boundary,
// addSuperGridDampingCU_upper_boundary, RHSCorrCU_center, addSuperGridDampingCU_center, halo members, cartesian_bc_forcingCU +
// enforceBCCU, sync_stream, (check points), extractRecordDataCU, cycleSolutionArrays.  Here a step is
//   RHSPredCU_center  -> sw4b200_grid_predictor_dev  (fused rhs4sg + predictor + stored acceleration + sparse forcing)
//   enforceBCCU       -> sw4b200_grid_enforce_bc (+ enforce_cart_topo)
//   RHSCorrCU_center  -> sw4b200_grid_corrector_dev  (fused rhs4sg + corrector on the stored acceleration + F_tt + damping)
//   enforceBCCU       -> boundary conditions; the reference's device arrays are refreshed only if the host will read them
// and the other members have nothing left to do.  cycleSolutionArrays is the reference's own: the grid blocks rotate at the
// start of the next step instead.
static void import_initial_data( EW* ew, vector<Sarray>& a_U, vector<Sarray>& a_Um )
{
   // U, Um as the host holds them at the start of the loop (zero, or a restart file's: EW.C:2403-2415)
   for( int g = 0; g < ew->mNumberOfGrids; g++ )
   {
      B200( sw4b200_grid_upload( g_grid[g], "U", a_U[g].c_ptr() ) );
      B200( sw4b200_grid_upload( g_grid[g], "Um", a_Um[g].c_ptr() ) );
   }
   g_imported = true;
}

// grid block -> the reference's (unpadded) device array, so that Sarray::copy_from_device sees the current values
static void export_field( int g, const char* name, Sarray& dst )
{
   const double* src = (const double*)sw4b200_grid_device_ptr( g_grid[g], name );
   const size_t pitch = (size_t)sw4b200_grid_row_pitch( g_grid[g] ), ni = dst.m_ni;
   const size_t rows = (size_t)dst.m_nc * dst.m_nj * dst.m_nk;
   cudaError_t e;
   if( Sarray::m_corder && pitch != ni )
      e = cudaMemcpy2DAsync( dst.dev_ptr(), ni * 8, src, pitch * 8, ni * 8, rows, cudaMemcpyDeviceToDevice, (cudaStream_t)stream0() );
   else
      e = cudaMemcpyAsync( dst.dev_ptr(), src, rows * ni * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream0() );
   if( e != cudaSuccess ) { fprintf( stderr, "sw4lite_b200: export of %s failed: %s\n", name, cudaGetErrorString( e ) ); exit( 1 ); }
}

void EW::RHSPredCU_boundary( vector<Sarray>&, vector<Sarray>&, vector<Sarray>&, vector<Sarray>&, vector<Sarray>&,
			     vector<Sarray>&, vector<Sarray>&, int ) {}
void EW::RHSPredCU_center( vector<Sarray>& a_Up, vector<Sarray>& a_U, vector<Sarray>& a_Um, vector<Sarray>& a_Mu,
			   vector<Sarray>& a_Lambda, vector<Sarray>& a_Rho, vector<Sarray>& a_F, int st )
{
   (void)a_Mu; (void)a_Lambda; (void)a_Rho; (void)a_F; (void)st;
   if( !g_imported ) import_initial_data( this, a_U, a_Um );
   else
      for( int g = 0; g < mNumberOfGrids; g++ ) B200( sw4b200_grid_cycle( g_grid[g] ) ); // the rotation of the previous step
   g_sUp.resize( mNumberOfGrids ); g_sU.resize( mNumberOfGrids );
   for( int g = 0; g < mNumberOfGrids; g++ )
   {
      g_sUp[g] = &a_Up[g]; g_sU[g] = &a_U[g];
      B200( sw4b200_grid_predictor_dev( g_grid[g], 0, g_src_built ? g_src[g].d_f : 0 ) );
   }
   g_corr_done = false;
}

// the acceleration is stored by the predictor pass (Cartesian grids) or formed inside the corrector phase (curvilinear grid)
void EW::evalDpDmInTimeCU( vector<Sarray>&, vector<Sarray>&, vector<Sarray>&, vector<Sarray>&, int ) {}

void EW::RHSCorrCU_boundary( vector<Sarray>&, vector<Sarray>&, vector<Sarray>&, vector<Sarray>&, vector<Sarray>&,
			     vector<Sarray>&, int ) {}
void EW::RHSCorrCU_center( vector<Sarray>& a_Up, vector<Sarray>& a_Uacc, vector<Sarray>& a_Mu, vector<Sarray>& a_Lambda,
			   vector<Sarray>& a_Rho, vector<Sarray>& a_F, int st )
{
   (void)a_Up; (void)a_Uacc; (void)a_Mu; (void)a_Lambda; (void)a_Rho; (void)a_F; (void)st;
   for( int g = 0; g < mNumberOfGrids; g++ )
      B200( sw4b200_grid_corrector_dev( g_grid[g], 0, g_src_built ? g_src[g].d_f : 0 ) ); // (the table holds F_tt now: ForceCU(t,tt))
   g_corr_done = true;
}

// supergrid damping is part of the corrector phase (box-decomposed: only where some damping coefficient is non-zero)
void EW::addSuperGridDampingCU_upper_boundary( vector<Sarray>&, vector<Sarray>&, vector<Sarray>&, vector<Sarray>&, int ) {}
void EW::addSuperGridDampingCU_center( vector<Sarray>&, vector<Sarray>&, vector<Sarray>&, vector<Sarray>&, int ) {}
// (the unfused operator on caller-owned arrays, for hosts that sequence the step themselves)
void EW::addSuperGridDampingCU( vector<Sarray>& a_Up, vector<Sarray>& a_U, vector<Sarray>& a_Um, vector<Sarray>& a_Rho, int st )
{
   (void)st;
   for( int g = 0; g < mNumberOfGrids; g++ )
   {
      if( g < mNumberOfCartesianGrids )
	 B200( sw4b200_addsgd( Sarray::m_corder, m_sg_damping_order, BOUNDS( g ), a_Up[g].dev_ptr(), a_U[g].dev_ptr(),
			       a_Um[g].dev_ptr(), a_Rho[g].dev_ptr(), dev_sg_dc_x[g], dev_sg_dc_y[g], dev_sg_dc_z[g],
			       dev_sg_str_x[g], dev_sg_str_y[g], dev_sg_str_z[g], dev_sg_corner_x[g], dev_sg_corner_y[g],
			       dev_sg_corner_z[g], m_supergrid_damping_coefficient, stream0() ) );
      else
	 B200( sw4b200_addsgdc( Sarray::m_corder, m_sg_damping_order, BOUNDS( g ), a_Up[g].dev_ptr(), a_U[g].dev_ptr(),
				a_Um[g].dev_ptr(), a_Rho[g].dev_ptr(), dev_sg_dc_x[g], dev_sg_dc_y[g], dev_sg_str_x[g],
				dev_sg_str_y[g], mJ.dev_ptr(), dev_sg_corner_x[g], dev_sg_corner_y[g],
				m_supergrid_damping_coefficient, stream0() ) );
   }
}

// boundary forcing: zeros except on Dirichlet sides of the point-source test (EW.C:3320-3344)
void EW::cartesian_bc_forcingCU( float_sw4 t, vector<float_sw4**>& a_BCForcing, vector<Source*>& a_sources, int st )
{
   (void)st;
   bool any = false;
   for( int g = 0; g < mNumberOfGrids && !any; g++ )
      for( int side = 0; side < 6; side++ ) any = any || ( m_point_source_test && m_bcType[g][side] == bDirichlet );
   if( !any ) return; // the grid blocks hold the zeros they were created with
   cartesian_bc_forcing( t, a_BCForcing, a_sources );
   for( int g = 0; g < mNumberOfGrids; g++ )
      for( int side = 0; side < 6; side++ )
	 if( m_bcType[g][side] == bDirichlet && a_BCForcing[g][side] )
	 {
	    char name[16];
	    snprintf( name, sizeof( name ), "bforce%d", side );
	    if( sw4b200_grid_array_size( g_grid[g], name ) == 3 * (size_t)m_NumberOfBCPoints[g][side] )
	       B200( sw4b200_grid_upload( g_grid[g], name, a_BCForcing[g][side] ) );
	 }
}

// replaces enforceBCCU (EW_cuda.C:2070) and adds what EW::enforceBC does for the curvilinear grid (EW.C:3477-3501)
void EW::enforceBCCU( vector<Sarray>& a_U, vector<Sarray>& a_Mu, vector<Sarray>& a_Lambda, float_sw4 t,
		      vector<float_sw4**>& a_BCForcing, int st )
{
   (void)a_U; (void)a_Mu; (void)a_Lambda; (void)a_BCForcing; (void)st;
   for( int g = 0; g < mNumberOfGrids; g++ ) B200( sw4b200_grid_enforce_bc( g_grid[g] ) );
   if( topographyExists() )
      B200( sw4b200_grid_enforce_cart_topo( g_grid[mNumberOfCartesianGrids - 1], g_grid[mNumberOfGrids - 1] ) );
   if( !g_corr_done ) return;
   // The step is complete.  The host reads the reference's device arrays in two places only: a check point about to be
   // written (U and Up, EW.C:2778-2791) and the error norm after the last step (the array that holds Up now is U after
   // cycleSolutionArrays, EW.C:2886-2888).  Refresh them from the grid blocks exactly then.
   const int cycle = (int)floor( ( g_t_step - mTstart ) / mDt + 0.5 ) + 1; // currentTimeStep of the reference's loop
   bool need_up = cycle >= mNumberOfTimeSteps, need_u = false;
   for( size_t c = 0; c < m_check_points.size(); c++ )
      if( CheckPointPeek::due( m_check_points[c], t, cycle, mDt ) ) need_up = need_u = true;
   if( m_checkfornan ) need_up = true;
   for( int g = 0; g < mNumberOfGrids; g++ )
   {
      if( need_up ) export_field( g, "Up", *g_sUp[g] );
      if( need_u ) export_field( g, "U", *g_sU[g] );
   }
}

// --------------------------------------------------------------------------------------------- unfused operators
void EW::evalRHSCU( vector<Sarray>& a_U, vector<Sarray>& a_Mu, vector<Sarray>& a_Lambda, vector<Sarray>& a_Lu, int st )
{
   (void)st;
   for( int g = 0; g < mNumberOfGrids; g++ )
   {
      if( g < mNumberOfCartesianGrids )
	 B200( sw4b200_rhs4sg( Sarray::m_corder, BOUNDS( g ), m_global_nz[g], m_onesided[g], a_Lu[g].dev_ptr(), a_U[g].dev_ptr(),
			       a_Mu[g].dev_ptr(), a_Lambda[g].dev_ptr(), mGridSize[g], dev_sg_str_x[g], dev_sg_str_y[g],
			       dev_sg_str_z[g], stream0() ) );
      else
	 B200( sw4b200_rhs4sgcurv( Sarray::m_corder, BOUNDS( g ), a_U[g].dev_ptr(), a_Mu[g].dev_ptr(), a_Lambda[g].dev_ptr(),
				   mMetric.dev_ptr(), mJ.dev_ptr(), a_Lu[g].dev_ptr(), m_onesided[g], dev_sg_str_x[g],
				   dev_sg_str_y[g], stream0() ) );
   }
}
void EW::evalPredictorCU( vector<Sarray>& a_Up, vector<Sarray>& a_U, vector<Sarray>& a_Um, vector<Sarray>& a_Rho,
			  vector<Sarray>& a_Lu, vector<Sarray>& a_F, int st )
{
   (void)st;
   for( int g = 0; g < mNumberOfGrids; g++ )
      B200( sw4b200_predfort( Sarray::m_corder, BOUNDS( g ), a_Up[g].dev_ptr(), a_U[g].dev_ptr(), a_Um[g].dev_ptr(), a_Lu[g].dev_ptr(),
			      a_F[g].dev_ptr(), a_Rho[g].dev_ptr(), mDt * mDt, stream0() ) );
}
void EW::evalCorrectorCU( vector<Sarray>& a_Up, vector<Sarray>& a_Rho, vector<Sarray>& a_Lu, vector<Sarray>& a_F, int st )
{
   (void)st;
   for( int g = 0; g < mNumberOfGrids; g++ )
      B200( sw4b200_corrfort( Sarray::m_corder, BOUNDS( g ), a_Up[g].dev_ptr(), a_Lu[g].dev_ptr(), a_F[g].dev_ptr(),
			      a_Rho[g].dev_ptr(), mDt * mDt * mDt * mDt, stream0() ) );
}

// --------------------------------------------------------------------------------------------- receivers
void EW::allocateTimeSeriesOnDeviceCU( int& nvals, int& ntloc, int*& i0dev, int*& j0dev, int*& k0dev, int*& g0dev,
				       int*& modedev, float_sw4**& urec_dev, float_sw4**& urec_host, float_sw4**& urec_hdev )
{
   // the index arrays of the reference's device kernel are not needed: the stations are addressed from the host
   i0dev = j0dev = k0dev = g0dev = modedev = 0;
   urec_dev = 0;
   ntloc = 0;
   for( size_t ts = 0; ts < m_GlobalTimeSeries.size(); ts++ )
      if( m_GlobalTimeSeries[ts]->myPoint() )
      {
	 nvals += m_GlobalTimeSeries[ts]->urec_size();
	 ntloc++;
      }
   urec_host = new float_sw4*[ntloc > 0 ? ntloc : 1];
   urec_hdev = new float_sw4*[ntloc > 0 ? ntloc : 1];
   if( ntloc == 0 ) return;
   float_sw4* devmem = (float_sw4*)sw4b200_malloc( 2 * 3 * ntloc * sizeof( double ) ); // Up and Um samples of every station
   float_sw4* hostmem = new float_sw4[nvals];
   if( !devmem ) ok( 1, "sw4b200_malloc" );
   size_t ptr = 0;
   int tsnr = 0;
   for( size_t ts = 0; ts < m_GlobalTimeSeries.size(); ts++ )
      if( m_GlobalTimeSeries[ts]->myPoint() )
      {
	 urec_hdev[tsnr] = devmem; // one device buffer for all stations
	 urec_host[tsnr] = &hostmem[ptr];
	 ptr += m_GlobalTimeSeries[ts]->urec_size();
	 tsnr++;
      }
}

void EW::extractRecordDataCU( int nt, int* mode, int* i0v, int* j0v, int* k0v, int* g0v, float_sw4** urec_dev, Sarray* dev_Um,
			      Sarray* dev_U, float_sw4 dt, float_sw4* h_dev, Sarray* dev_metric, Sarray* dev_j, int st,
			      int nvals, float_sw4* urec_hostmem, float_sw4* urec_devmem )
{
   (void)mode; (void)i0v; (void)j0v; (void)k0v; (void)g0v; (void)urec_dev; (void)dev_Um; (void)dev_U; (void)h_dev;
   (void)dev_metric; (void)dev_j; (void)st; (void)nvals;
   // station indices, grouped per grid (built once)
   static vector<vector<long long> > idx;
   static vector<long long*> d_idx;
   static vector<vector<int> > station; // station number (among myPoint stations) of every entry
   if( idx.empty() )
   {
      idx.resize( mNumberOfGrids ); d_idx.assign( mNumberOfGrids, (long long*)0 ); station.resize( mNumberOfGrids );
      int tsnr = 0;
      for( size_t ts = 0; ts < m_GlobalTimeSeries.size(); ts++ )
      {
	 TimeSeries* s = m_GlobalTimeSeries[ts];
	 if( !s->myPoint() ) continue;
	 if( s->getMode() != TimeSeries::Displacement && s->getMode() != TimeSeries::Velocity )
	 {
	    cerr << "sw4lite_b200: station " << ts << ": only displacement and velocity records are implemented on the device" << endl;
	    exit( 1 );
	 }
	 const int g = s->m_grid0;
	 const long long ni = sw4b200_grid_row_pitch( g_grid[g] ), nj = m_jEnd[g] - m_jStart[g] + 1; // (device rows may be padded)
	 idx[g].push_back( ( s->m_i0 - m_iStart[g] ) + ni * ( s->m_j0 - m_jStart[g] ) + ni * nj * ( s->m_k0 - m_kStart[g] ) );
	 station[g].push_back( tsnr++ );
      }
      for( int g = 0; g < mNumberOfGrids; g++ )
	 if( !idx[g].empty() )
	 {
	    d_idx[g] = (long long*)sw4b200_malloc( idx[g].size() * sizeof( long long ) );
	    B200( sw4b200_memcpy_h2d( d_idx[g], idx[g].data(), idx[g].size() * sizeof( long long ), 0 ) );
	    B200( sw4b200_sync_stream( 0 ) );
	 }
   }
   // gather Up (the new solution) and Um at the stations, one copy to the host
   vector<double> h( 2 * 3 * nt );
   size_t off = 0;
   vector<size_t> goff( mNumberOfGrids, 0 );
   for( int g = 0; g < mNumberOfGrids; g++ )
   {
      const int n = (int)idx[g].size();
      goff[g] = off;
      if( n == 0 ) continue;
      const size_t npts = (size_t)sw4b200_grid_row_pitch( g_grid[g] ) * ( m_jEnd[g] - m_jStart[g] + 1 ) * ( m_kEnd[g] - m_kStart[g] + 1 );
      B200( sw4b200_gather_points( Sarray::m_corder, npts, (const double*)sw4b200_grid_device_ptr( g_grid[g], "Up" ), n, d_idx[g],
				   urec_devmem + off, stream0() ) );
      B200( sw4b200_gather_points( Sarray::m_corder, npts, (const double*)sw4b200_grid_device_ptr( g_grid[g], "Um" ), n, d_idx[g],
				   urec_devmem + off + 3 * n, stream0() ) );
      off += 6 * n;
   }
   B200( sw4b200_memcpy_d2h( h.data(), urec_devmem, off * sizeof( double ), stream0() ) );
   B200( sw4b200_sync_stream( 0 ) );
   // hand the samples to the reference's TimeSeries objects in their own order (EW::extractRecordData, EW.C:5854)
   vector<float_sw4*> out( nt );
   {
      size_t ptr = 0;
      int tsnr = 0;
      for( size_t ts = 0; ts < m_GlobalTimeSeries.size(); ts++ )
	 if( m_GlobalTimeSeries[ts]->myPoint() )
	 {
	    out[tsnr++] = urec_hostmem + ptr;
	    ptr += m_GlobalTimeSeries[ts]->urec_size();
	 }
   }
   int tsnr = 0;
   vector<TimeSeries::receiverMode> modes;
   for( size_t ts = 0; ts < m_GlobalTimeSeries.size(); ts++ )
      if( m_GlobalTimeSeries[ts]->myPoint() ) modes.push_back( m_GlobalTimeSeries[ts]->getMode() );
   for( int g = 0; g < mNumberOfGrids; g++ )
   {
      const int n = (int)idx[g].size();
      for( int m = 0; m < n; m++ )
      {
	 const int s = station[g][m];
	 const double* up = &h[goff[g] + 3 * m];
	 const double* um = &h[goff[g] + 3 * n + 3 * m];
	 for( int c = 0; c < 3; c++ )
	    out[s][c] = modes[s] == TimeSeries::Displacement ? up[c] : ( up[c] - um[c] ) / ( 2 * dt );
      }
   }
   (void)tsnr;
}
