// C++ multi-GPU driver of the Cartesian hot path: one process per GPU, the grid decomposed into z-slabs, the halo exchange
// inside libsw4b200 (sw4b200_grid_exchange_begin / _end: NCCL straight from the field arrays), overlapped with the bulk rows.
//
// This is the time loop of EW::timesteploop (reference EW.C:2527-2842) for a synthetic half-space, written against the
// C-ABI only (include/sw4b200.h): what a C++ host that decomposes in z does per step, in the reference's order of
// operations -- Force, RHS+predictor, halo, boundary conditions, Force_tt, RHS+corrector, supergrid damping, halo, boundary
// conditions, cycle -- with the face rows computed first so that their transfer runs under the bulk rows (the
// boundary/centre split of RHSPredCU_boundary/_center, EW_cuda.C:1228-1410, applied to the z faces).  The reference's own
// host decomposes in x-y over MPI ranks (proc_decompose_2d, EW.C:2931); there is no MPI in this image, and NVSwitch makes
// the thin z faces the cheaper cut, so the multi-GPU host is this program (and its Python twin, sw4lite_b200/slabs.py).
//
// Set-up mirrors sw4lite_b200/setup.py (supergrid arrays SuperGrid.C:108-198, dt EW.C:5041-5066, windows EW.C:3347-3420,
// decomp1d EW.C:2963-2985).  Ranks meet through a file: rank 0 writes the 128-byte NCCL id (host/run_slabs.sh makes the
// directory and starts the ranks; torchrun's RANK / LOCAL_RANK / WORLD_SIZE are understood too).
//
//   slab_driver [--nx N --ny N --nzl N | --nz-total N] [--steps K --warmup W] [--h H] [--gp G]
// prints one JSON line on rank 0: Gpts/s of the whole job (device time, max over ranks) and a checksum of the wavefield.
#include "../include/sw4b200.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <unistd.h>

static void ok( int rc, const char* what )
{
   if( rc != 0 )
   {
      fprintf( stderr, "slab_driver: %s failed: %s\n", what, sw4b200_last_error() );
      exit( 1 );
   }
}
#define B200( call ) ok( ( call ), #call )

// ---- set-up (host side, never on the hot path)
static double psi0( double xi ) // SuperGrid::Psi0 (SuperGrid.C:170-192)
{
   if( xi <= 0 ) return 0;
   if( xi >= 1 ) return 1;
   return xi * xi * xi * xi * xi * xi * ( 462 - 1980 * xi + 3465 * xi * xi - 3080 * xi * xi * xi + 1386 * xi * xi * xi * xi - 252 * xi * xi * xi * xi * xi );
}
struct Sg1d { std::vector<double> dc, str, co; };
// dampingCoeff / stretching / cornerTaper at the coordinates x (SuperGrid.C:108-198), layers of `width` on the chosen sides
static Sg1d supergrid_1d( const std::vector<double>& x, bool left, bool right, double x0, double x1, double width )
{
   const double epsL = 1e-4, cmin = 0.33, tw = 0.5 * width;
   Sg1d s;
   for( double xv : x )
   {
      double psi = 0, damp = 0, lin = 0;
      const bool inl = left && xv < x0 + width;
      if( inl ) { psi = psi0( ( x0 + width - xv ) / width ); damp = psi0( ( x0 + width - xv ) / tw ); lin = ( x0 + width - xv ) / width; }
      if( right && xv > x1 - width && !inl ) { psi = psi0( ( xv - ( x1 - width ) ) / width ); damp = psi0( ( xv - ( x1 - width ) ) / tw ); lin = ( xv - ( x1 - width ) ) / width; }
      const double stretch = 1 - ( 1 - epsL ) * psi;
      s.dc.push_back( damp / stretch ); s.str.push_back( stretch ); s.co.push_back( 1.0 - ( 1.0 - cmin ) * lin );
   }
   return s;
}
static double c6smoothbump( double freq, double t ) { const double x = t * freq; return ( x < 0 || x > 1 ) ? 0.0 : 51480 * pow( x * ( 1 - x ), 7 ); }
static double c6smoothbump_tt( double freq, double t )
{
   const double x = t * freq;
   if( x < 0 || x > 1 ) return 0.0;
   return 51480 * freq * freq * 7 * ( 6 * ( 1 - 2 * x ) * ( 1 - 2 * x ) * pow( x * ( 1 - x ), 5 ) - 2 * pow( x * ( 1 - x ), 6 ) );
}
// EW::decomp1d (EW.C:2963-2985) and the planes a slab owns
static void decomp1d( int nglobal, int myid, int nproc, int& s, int& e )
{
   const int olap = 4;
   int nlocal = ( nglobal + ( nproc - 1 ) * olap ) / nproc;
   const int deficit = ( nglobal + ( nproc - 1 ) * olap ) % nproc;
   if( myid < deficit ) { s = myid * ( nlocal - olap ) + myid + 1; nlocal++; }
   else s = myid * ( nlocal - olap ) + deficit + 1;
   e = s + nlocal - 1;
}

static int envi( const char* n, int d ) { const char* e = getenv( n ); return e ? atoi( e ) : d; }

int main( int argc, char** argv )
{
   int nx = 2048, ny = 2048, nzl = 128, nztot = 0, steps = 10, warmup = 3, gp = 30;
   double h = 10.0;
   for( int a = 1; a + 1 < argc; a += 2 )
   {
      const std::string k = argv[a];
      if( k == "--nx" ) nx = atoi( argv[a + 1] );
      else if( k == "--ny" ) ny = atoi( argv[a + 1] );
      else if( k == "--nzl" ) nzl = atoi( argv[a + 1] );
      else if( k == "--nz-total" ) nztot = atoi( argv[a + 1] );
      else if( k == "--steps" ) steps = atoi( argv[a + 1] );
      else if( k == "--warmup" ) warmup = atoi( argv[a + 1] );
      else if( k == "--gp" ) gp = atoi( argv[a + 1] );
      else if( k == "--h" ) h = atof( argv[a + 1] );
      else { fprintf( stderr, "slab_driver: unknown option %s\n", argv[a] ); return 2; }
   }
   const int rank = envi( "RANK", 0 ), world = envi( "WORLD_SIZE", 1 ), local = envi( "LOCAL_RANK", rank );
   const int nz = nztot > 0 ? nztot : nzl * world;
   B200( sw4b200_init( local ) );

   // ---- communicator: rank 0 publishes the NCCL id in a file
   unsigned char id[128] = { 0 };
   if( world > 1 )
   {
      const char* f = getenv( "SW4B200_ID_FILE" );
      if( !f ) { fprintf( stderr, "slab_driver: SW4B200_ID_FILE is not set (use host/run_slabs.sh)\n" ); return 2; }
      const std::string path = f, tmp = path + ".tmp";
      if( rank == 0 )
      {
	 B200( sw4b200_comm_unique_id( id ) );
	 FILE* fp = fopen( tmp.c_str(), "wb" );
	 if( !fp || fwrite( id, 1, 128, fp ) != 128 ) { fprintf( stderr, "slab_driver: cannot write %s\n", tmp.c_str() ); return 1; }
	 fclose( fp );
	 rename( tmp.c_str(), path.c_str() );
      }
      else
      {
	 FILE* fp = 0;
	 for( int t = 0; t < 1200 && !( fp = fopen( path.c_str(), "rb" ) ); t++ ) std::this_thread::sleep_for( std::chrono::milliseconds( 50 ) );
	 if( !fp || fread( id, 1, 128, fp ) != 128 ) { fprintf( stderr, "slab_driver: rank %d found no NCCL id in %s\n", rank, path.c_str() ); return 1; }
	 fclose( fp );
      }
   }
   {
      // (NCCL prints a version banner on stdout: keep stdout for the one result line)
      fflush( stdout );
      const int saved = dup( 1 );
      dup2( 2, 1 );
      B200( sw4b200_comm_init( rank, world, id ) );
      fflush( stdout );
      dup2( saved, 1 );
      close( saved );
   }

   // ---- the problem: half-space, free surface on top, supergrid layers on the other five sides, one layer interface
   const double vp[2] = { 4000, 6000 }, vs[2] = { 2000, 3464 }, rho[2] = { 2600, 2700 }, ztop = 0.6 * nz * h, cfl = 1.3, beta = 0.02;
   const int ni = nx + 4, nj = ny + 4, nkg = nz + 4;
   std::vector<double> xs( ni ), ys( nj ), zs( nkg );
   for( int i = 0; i < ni; i++ ) xs[i] = ( i - 2 ) * h;
   for( int j = 0; j < nj; j++ ) ys[j] = ( j - 2 ) * h;
   for( int k = 0; k < nkg; k++ ) zs[k] = ( k - 2 ) * h;
   const double width = gp * h;
   const Sg1d sx = supergrid_1d( xs, true, true, 0.0, ( nx - 1 ) * h, width ), sy = supergrid_1d( ys, true, true, 0.0, ( ny - 1 ) * h, width ),
	      sz = supergrid_1d( zs, false, true, 0.0, ( nz - 1 ) * h, width );
   std::vector<double> muk( nkg ), lak( nkg ), rhk( nkg );
   double cmax = 0;
   for( int k = 0; k < nkg; k++ )
   {
      const int m = zs[k] >= ztop ? 1 : 0;
      rhk[k] = rho[m]; muk[k] = rho[m] * vs[m] * vs[m]; lak[k] = rho[m] * vp[m] * vp[m] - 2 * muk[k];
      cmax = fmax( cmax, ( 4 * muk[k] + lak[k] ) / rhk[k] );
   }
   const double dt = cfl * h / sqrt( cmax ); // EW::computeDT (EW.C:5041-5066)

   // ---- this rank's slab
   int s, e;
   decomp1d( nz, rank, world, s, e );
   const bool halo_lo = rank > 0, halo_hi = rank < world - 1;
   const int k0 = s + ( halo_lo ? 2 : 0 ), k1 = e - ( halo_hi ? 2 : 0 ); // owned planes
   sw4b200_grid_desc d;
   memset( &d, 0, sizeof( d ) );
   d.corder = 1;
   d.ifirst = -1; d.ilast = nx + 2; d.jfirst = -1; d.jlast = ny + 2; d.kfirst = k0 - 2; d.klast = k1 + 2;
   d.nx = nx; d.ny = ny; d.nz = nz; d.h = h; d.dt = dt;
   const int bSuperGrid = 2, bStressFree = 0, bProcessor = 7;
   for( int sd = 0; sd < 6; sd++ ) d.bctype[sd] = bSuperGrid;
   d.bctype[4] = halo_lo ? bProcessor : bStressFree;
   if( halo_hi ) d.bctype[5] = bProcessor;
   d.onesided[4] = halo_lo ? 0 : 1;
   const int bounds[6] = { d.ifirst, d.ilast, d.jfirst, d.jlast, d.kfirst, d.klast };
   for( int sd = 0; sd < 6; sd++ ) // m_BndryWindow (EW.C:3347-3420)
   {
      int w[6] = { 999, -999, 999, -999, 999, -999 };
      if( d.bctype[sd] == bStressFree || d.bctype[sd] == bSuperGrid )
      {
	 for( int q = 0; q < 6; q++ ) w[q] = bounds[q];
	 const int lo = 2 * ( sd / 2 );
	 if( d.bctype[sd] == bStressFree ) w[lo] = w[lo + 1] = ( sd % 2 == 0 ) ? bounds[lo] + 2 : bounds[lo + 1] - 2;
	 else if( sd % 2 == 0 ) w[lo + 1] = w[lo] + 1;
	 else w[lo] = w[lo + 1] - 1;
      }
      for( int q = 0; q < 6; q++ ) d.wind[6 * sd + q] = w[q];
   }
   d.sg_order = 4; d.beta = beta; d.halo_lo = halo_lo; d.halo_hi = halo_hi;
   sw4b200_grid* G = sw4b200_grid_create( &d );
   if( !G ) ok( 1, "sw4b200_grid_create" );
   const int nk = d.klast - d.kfirst + 1, koff = d.kfirst - ( -1 );
   B200( sw4b200_grid_upload( G, "strx", sx.str.data() ) ); B200( sw4b200_grid_upload( G, "dcx", sx.dc.data() ) ); B200( sw4b200_grid_upload( G, "cox", sx.co.data() ) );
   B200( sw4b200_grid_upload( G, "stry", sy.str.data() ) ); B200( sw4b200_grid_upload( G, "dcy", sy.dc.data() ) ); B200( sw4b200_grid_upload( G, "coy", sy.co.data() ) );
   B200( sw4b200_grid_upload( G, "strz", sz.str.data() + koff ) ); B200( sw4b200_grid_upload( G, "dcz", sz.dc.data() + koff ) );
   B200( sw4b200_grid_upload( G, "coz", sz.co.data() + koff ) );
   B200( sw4b200_grid_fill_profile( G, "mu", muk.data() + koff ) ); B200( sw4b200_grid_fill_profile( G, "lambda", lak.data() + koff ) );
   B200( sw4b200_grid_fill_profile( G, "rho", rhk.data() + koff ) );
   if( world > 1 ) B200( sw4b200_grid_set_neighbours( G, halo_lo ? rank - 1 : -1, halo_hi ? rank + 1 : -1 ) );

   // ---- sources: a 6 x 6 x 6 cloud of point forces (one moment-tensor source, GridPointSource.C), owned by the slab holding them
   const int ci = nx / 2, cj = ny / 2, ck = std::max( 8, std::min( 32, nz - 8 ) ); // (the same point whatever the number of slabs)
   std::vector<int> ijk;
   std::vector<double> amp;
   unsigned long long lcg = 12345;
   for( int di = -3; di < 3; di++ )
      for( int dj = -3; dj < 3; dj++ )
	 for( int dk = -3; dk < 3; dk++ )
	 {
	    double a3[3];
	    for( int c = 0; c < 3; c++ ) { lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL; a3[c] = ( ( lcg >> 11 ) / 9007199254740992.0 * 2 - 1 ) * 1e12; }
	    const int k = ck + dk;
	    if( k < k0 || k > k1 ) continue;
	    ijk.push_back( ci + di ); ijk.push_back( cj + dj ); ijk.push_back( k );
	    for( int c = 0; c < 3; c++ ) amp.push_back( a3[c] );
	 }
   const int nsrc = (int)ijk.size() / 3;
   if( nsrc ) B200( sw4b200_grid_set_source_points( G, nsrc, ijk.data() ) );
   // a smooth initial wavefield defined by global indices, so that all slabs agree on their halo planes
   {
      const size_t npts = (size_t)ni * nj * nk;
      std::vector<double> u( 3 * npts ), um( 3 * npts );
      std::vector<double> fi( ni ), fj( nj );
      for( int c = 0; c < 3; c++ )
      {
	 for( int ph = 0; ph < 2; ph++ )
	 {
	    const double phs = ph ? 0.013 : 0.0;
	    for( int i = 0; i < ni; i++ ) fi[i] = sin( 0.11 * i + 0.7 * c + phs );
	    for( int j = 0; j < nj; j++ ) fj[j] = cos( 0.07 * j + 0.3 * c );
	    std::vector<double>& dst = ph ? um : u;
	    for( int k = 0; k < nk; k++ )
	    {
	       const double fk = 1e-3 * sin( 0.05 * ( k + d.kfirst ) + c + phs );
	       double* p = dst.data() + c * npts + (size_t)k * ni * nj;
	       for( int j = 0; j < nj; j++ )
		  for( int i = 0; i < ni; i++ ) p[(size_t)j * ni + i] = fk * fj[j] * fi[i];
	    }
	 }
      }
      B200( sw4b200_grid_upload( G, "U", u.data() ) );
      B200( sw4b200_grid_upload( G, "Um", um.data() ) );
   }

   // ---- the time loop
   std::vector<double> f( 3 * ( nsrc ? nsrc : 1 ) ), ftt( f.size() );
   auto step = [&]( int n ) {
      const double t = n * dt, g = c6smoothbump( 2.0, t ), gtt = c6smoothbump_tt( 2.0, t );
      for( int q = 0; q < 3 * nsrc; q++ ) { f[q] = amp[q] * g; ftt[q] = amp[q] * gtt; }
      const double* pf = nsrc ? f.data() : 0;
      const double* pftt = nsrc ? ftt.data() : 0;
      B200( sw4b200_grid_predictor_part( G, 1, pf ) );	       // face rows (+ ghost shell)
      B200( sw4b200_grid_exchange_begin( G, 1 ) );	       // their planes travel ...
      B200( sw4b200_grid_predictor_part( G, 2, pf ) );	       // ... while the bulk rows compute
      B200( sw4b200_grid_exchange_end( G ) );
      B200( sw4b200_grid_enforce_bc( G ) );
      B200( sw4b200_grid_corrector_part( G, 1, pftt ) );
      B200( sw4b200_grid_exchange_begin( G, 0 ) );
      B200( sw4b200_grid_corrector_part( G, 2, pftt ) );
      B200( sw4b200_grid_exchange_end( G ) );
      B200( sw4b200_grid_enforce_bc( G ) );
      B200( sw4b200_grid_cycle( G ) );
   };
   auto barrier = [&]() {
      double one = 1;
      B200( sw4b200_sync_device() );
      B200( sw4b200_comm_allreduce( &one, 1, 1 ) );
   };
   for( int n = 0; n < warmup; n++ ) step( n );
   barrier();
   const int l0 = sw4b200_kernel_launch_count();
   B200( sw4b200_timer_start() );
   for( int n = warmup; n < warmup + steps; n++ ) step( n );
   double ms = 0;
   B200( sw4b200_timer_stop_ms( &ms ) );
   const int launches = sw4b200_kernel_launch_count() - l0;
   barrier();
   B200( sw4b200_comm_allreduce( &ms, 1, 0 ) ); // max over ranks

   // ---- checksum of the owned part of the final wavefield (sum of squares, max): comparable between decompositions
   double sums[2] = { 0, 0 };
   {
      const size_t npts = (size_t)ni * nj * nk;
      std::vector<double> u( 3 * npts );
      B200( sw4b200_grid_download( G, "U", u.data() ) );
      for( int c = 0; c < 3; c++ )
	 for( int k = 2; k < nk - 2; k++ )
	    for( int j = 2; j < nj - 2; j++ )
	       for( int i = 2; i < ni - 2; i++ )
	       {
		  const double v = u[c * npts + ( (size_t)k * nj + j ) * ni + i];
		  sums[0] += v * v;
		  sums[1] = fmax( sums[1], fabs( v ) );
	       }
   }
   double mx = sums[1];
   B200( sw4b200_comm_allreduce( &sums[0], 1, 1 ) );
   B200( sw4b200_comm_allreduce( &mx, 1, 0 ) );
   if( rank == 0 )
   {
      const double gpts = (double)nx * ny * nz * steps / ( ms * 1e-3 ) / 1e9;
      printf( "{\"impl\": \"ours-cxx-slabs\", \"metric\": \"grid-point updates/sec per timestep\", \"value\": %.6f, \"unit\": \"Gpts/s\", \"n_gpus\": %d, "
	      "\"steps\": %d, \"warmup\": %d, \"ms_per_step\": %.6f, \"higher_is_better\": true, \"scaling\": \"%s\", \"dtype\": \"f64\", \"data\": \"synthetic\", "
	      "\"config\": {\"workload\": \"synthetic Cartesian half-space %dx%dx%d, z-slabs over %d GPU(s), free surface + supergrid gp=%d, "
	      "%d-point source; C++ driver host/slab_driver.C on the C-ABI, halo exchange inside libsw4b200 (NCCL)\", \"grid\": [%d, %d, %d]}, "
	      "\"gpu_launches\": %d, \"dt\": %.17g, \"checksum\": {\"sum_sq\": %.17g, \"max_abs\": %.17g}}\n",
	      gpts, world, steps, warmup, ms / steps, nztot > 0 ? "strong" : "weak", nx, ny, nz, world, gp, 216, nx, ny, nz, launches, dt, sums[0], mx );
   }
   B200( sw4b200_grid_destroy( G ) );
   B200( sw4b200_comm_finalize() );
   B200( sw4b200_finalize() );
   return 0;
}
