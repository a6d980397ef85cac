// TEST INFRASTRUCTURE ONLY (oracle/).  Not part of the product; nothing here is shipped or timed
// as "ours".  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load the library this file is linked into.
//
// C-ABI shim over the UNMODIFIED reference sources (compiled where they lie under
// /root/reference/src by oracle/build_ref.py; outputs only into oracle/_ref/).  It exposes
//   (1) the reference CPU kernels of the hot path, one extern "C" entry per kernel, and
//   (2) a handle on a fully set-up reference `EW` object (parser + setupRun run by the
//       reference itself) that can be stepped one reference CPU time step at a time.
//
// (2) works because oracle/build_ref.py weakens the reference's own EW::timesteploop symbol
// in the copy of EW.o that goes into libsw4ref.so, so the definition below wins at link time:
// the reference constructor (EW.C:162-224) then returns after setupRun, and ref_ew_step()
// sequences exactly the calls of the reference's CPU branch (EW.C:2527-2878).  The pure
// reference executable oracle/_ref/sw4lite_ref keeps the original loop and is used to check
// that this sequencing reproduces it (tests/test_oracle_ref.py).
#include <mpi.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "EW.h"
#include "GridPointSource.h"
#include "Source.h"
#include "TimeSeries.h"
#include "EWCuda.h"

// free-function kernels of the reference (C++ linkage), rhs4sg.C:38, rhs4sg_rev.C:44,
// rhs4sgcurv.C:34, rhs4sgcurv_rev.C:34
void rhs4sg( int ifirst, int ilast, int jfirst, int jlast, int kfirst, int klast,
	     int nk, int* onesided, float_sw4* a_acof, float_sw4* a_bope, float_sw4* a_ghcof,
	     float_sw4* a_lu, float_sw4* a_u, float_sw4* a_mu, float_sw4* a_lambda,
	     float_sw4 h, float_sw4* a_strx, float_sw4* a_stry, float_sw4* a_strz );
void rhs4sg_rev( int ifirst, int ilast, int jfirst, int jlast, int kfirst, int klast,
		 int nk, int* onesided, float_sw4* a_acof, float_sw4* a_bope, float_sw4* a_ghcof,
		 float_sw4* a_lu, float_sw4* a_u, float_sw4* a_mu, float_sw4* a_lambda,
		 float_sw4 h, float_sw4* a_strx, float_sw4* a_stry, float_sw4* a_strz );
void rhs4sgcurv( int ifirst, int ilast, int jfirst, int jlast, int kfirst, int klast,
		 float_sw4* a_u, float_sw4* a_mu, float_sw4* a_lambda, float_sw4* a_met,
		 float_sw4* a_jac, float_sw4* a_lu, int* onesided, float_sw4* acof,
		 float_sw4* bope, float_sw4* ghcof, float_sw4* a_strx, float_sw4* a_stry );
void rhs4sgcurv_rev( int ifirst, int ilast, int jfirst, int jlast, int kfirst, int klast,
		     float_sw4* a_u, float_sw4* a_mu, float_sw4* a_lambda, float_sw4* a_met,
		     float_sw4* a_jac, float_sw4* a_lu, int* onesided, float_sw4* acof,
		     float_sw4* bope, float_sw4* ghcof, float_sw4* a_strx, float_sw4* a_stry );

// Replaces the (weakened) reference definition inside libsw4ref.so only; see header comment.
void EW::timesteploop( vector<Sarray>& U, vector<Sarray>& Um )
{
   (void)U; (void)Um;
}

namespace {
// An EW whose constructor never ran: the kernel members below only read m_corder.
EW* kernel_ew( int corder )
{
   static EW* e = 0;
   if( e == 0 )
      e = static_cast<EW*>( calloc( 1, sizeof(EW) ) );
   e->m_corder = (corder != 0);
   Sarray::m_corder = (corder != 0);
   return e;
}

struct RefRun
{
   EW* ew;
   std::vector<Sarray> F, Lu, Uacc, Up;
   double t;
   int step;
   std::string cwd0;
};
}

extern "C" {

//----------------------------------------------------------------------- kernels
void ref_get_stencil_coefficients( double* acof, double* ghcof, double* bope, double* sbop )
{
   kernel_ew(1)->GetStencilCoefficients( acof, ghcof, bope, sbop );
}

void ref_rhs4sg( int corder, int ifirst, int ilast, int jfirst, int jlast, int kfirst, int klast,
		 int nk, int* onesided, double* acof, double* bope, double* ghcof,
		 double* lu, double* u, double* mu, double* la, double h,
		 double* strx, double* stry, double* strz )
{
   if( corder )
      rhs4sg_rev( ifirst, ilast, jfirst, jlast, kfirst, klast, nk, onesided, acof, bope, ghcof,
		  lu, u, mu, la, h, strx, stry, strz );
   else
      rhs4sg( ifirst, ilast, jfirst, jlast, kfirst, klast, nk, onesided, acof, bope, ghcof,
	      lu, u, mu, la, h, strx, stry, strz );
}

void ref_rhs4sgcurv( int corder, int ifirst, int ilast, int jfirst, int jlast, int kfirst, int klast,
		     double* u, double* mu, double* la, double* met, double* jac, double* lu,
		     int* onesided, double* acof, double* bope, double* ghcof,
		     double* strx, double* stry )
{
   if( corder )
      rhs4sgcurv_rev( ifirst, ilast, jfirst, jlast, kfirst, klast, u, mu, la, met, jac, lu,
		      onesided, acof, bope, ghcof, strx, stry );
   else
      rhs4sgcurv( ifirst, ilast, jfirst, jlast, kfirst, klast, u, mu, la, met, jac, lu,
		  onesided, acof, bope, ghcof, strx, stry );
}

void ref_predfort( int corder, int ib, int ie, int jb, int je, int kb, int ke, double* up,
		   double* u, double* um, double* lu, double* fo, double* rho, double dt2 )
{
   kernel_ew(corder)->predfort( ib, ie, jb, je, kb, ke, up, u, um, lu, fo, rho, dt2 );
}

void ref_corrfort( int corder, int ib, int ie, int jb, int je, int kb, int ke, double* up,
		   double* lu, double* fo, double* rho, double dt4 )
{
   kernel_ew(corder)->corrfort( ib, ie, jb, je, kb, ke, up, lu, fo, rho, dt4 );
}

void ref_dpdmtfort( int corder, int ib, int ie, int jb, int je, int kb, int ke, double* up,
		    double* u, double* um, double* u2, double dt2i )
{
   kernel_ew(corder)->dpdmtfort( ib, ie, jb, je, kb, ke, up, u, um, u2, dt2i );
}

void ref_addsgd( int corder, int order, int ifirst, int ilast, int jfirst, int jlast, int kfirst, int klast,
		 double* up, double* u, double* um, double* rho,
		 double* dcx, double* dcy, double* dcz, double* strx, double* stry, double* strz,
		 double* cox, double* coy, double* coz, double beta )
{
   EW* e = kernel_ew(corder);
   if( order == 4 && corder )
      e->addsgd4fort_indrev( ifirst, ilast, jfirst, jlast, kfirst, klast, up, u, um, rho,
			     dcx, dcy, dcz, strx, stry, strz, cox, coy, coz, beta );
   else if( order == 4 )
      e->addsgd4fort( ifirst, ilast, jfirst, jlast, kfirst, klast, up, u, um, rho,
		      dcx, dcy, dcz, strx, stry, strz, cox, coy, coz, beta );
   else if( order == 6 && corder )
      e->addsgd6fort_indrev( ifirst, ilast, jfirst, jlast, kfirst, klast, up, u, um, rho,
			     dcx, dcy, dcz, strx, stry, strz, cox, coy, coz, beta );
   else
      e->addsgd6fort( ifirst, ilast, jfirst, jlast, kfirst, klast, up, u, um, rho,
		      dcx, dcy, dcz, strx, stry, strz, cox, coy, coz, beta );
}

void ref_addsgdc( int corder, int order, int ifirst, int ilast, int jfirst, int jlast, int kfirst, int klast,
		  double* up, double* u, double* um, double* rho,
		  double* dcx, double* dcy, double* strx, double* stry, double* jac,
		  double* cox, double* coy, double beta )
{
   EW* e = kernel_ew(corder);
   if( order == 4 && corder )
      e->addsgd4cfort_indrev( ifirst, ilast, jfirst, jlast, kfirst, klast, up, u, um, rho,
			      dcx, dcy, strx, stry, jac, cox, coy, beta );
   else if( order == 4 )
      e->addsgd4cfort( ifirst, ilast, jfirst, jlast, kfirst, klast, up, u, um, rho,
		       dcx, dcy, strx, stry, jac, cox, coy, beta );
   else if( order == 6 && corder )
      e->addsgd6cfort_indrev( ifirst, ilast, jfirst, jlast, kfirst, klast, up, u, um, rho,
			      dcx, dcy, strx, stry, jac, cox, coy, beta );
   else
      e->addsgd6cfort( ifirst, ilast, jfirst, jlast, kfirst, klast, up, u, um, rho,
		       dcx, dcy, strx, stry, jac, cox, coy, beta );
}

void ref_bcfortsg( int corder, int ib, int ie, int jb, int je, int kb, int ke, int* wind,
		   int nx, int ny, int nz, double* u, double h, int* bccnd, double* sbop,
		   double* mu, double* la, double t,
		   double* bforce1, double* bforce2, double* bforce3,
		   double* bforce4, double* bforce5, double* bforce6,
		   double* strx, double* stry )
{
   EW* e = kernel_ew(corder);
   boundaryConditionType bc[6];
   for( int s=0 ; s < 6 ; s++ )
      bc[s] = static_cast<boundaryConditionType>(bccnd[s]);
   if( corder )
      e->bcfortsg_indrev( ib, ie, jb, je, kb, ke, wind, nx, ny, nz, u, h, bc, sbop, mu, la, t,
			  bforce1, bforce2, bforce3, bforce4, bforce5, bforce6, 0, 0, 0, strx, stry );
   else
      e->bcfortsg( ib, ie, jb, je, kb, ke, wind, nx, ny, nz, u, h, bc, sbop, mu, la, t,
		   bforce1, bforce2, bforce3, bforce4, bforce5, bforce6, 0, 0, 0, strx, stry );
}

void ref_freesurfcurvisg( int corder, int ib, int ie, int jb, int je, int kb, int ke, int nz, int side,
			  double* u, double* mu, double* la, double* met, double* sbop,
			  double* forcing, double* strx, double* stry )
{
   EW* e = kernel_ew(corder);
   if( corder )
      e->freesurfcurvisg_rev( ib, ie, jb, je, kb, ke, nz, side, u, mu, la, met, sbop, forcing, strx, stry );
   else
      e->freesurfcurvisg( ib, ie, jb, je, kb, ke, nz, side, u, mu, la, met, sbop, forcing, strx, stry );
}

//----------------------------------------------------------------------- EW handle
// Runs the reference parser + setupRun on `infile` (in directory `workdir`, where the
// reference creates its output directory) and allocates the time-loop arrays exactly as
// EW.C:2352-2427 does.  Returns an opaque handle.
void* ref_ew_create( const char* infile, const char* workdir )
{
   RefRun* r = new RefRun;
   char buf[4096];
   r->cwd0 = getcwd( buf, sizeof(buf) ) ? buf : ".";
   if( workdir && workdir[0] )
      if( chdir( workdir ) != 0 ) { delete r; return 0; }
   r->ew = new EW( std::string(infile) );
   EW* ew = r->ew;
   int ng = ew->mNumberOfGrids;
   r->F.resize(ng); r->Lu.resize(ng); r->Uacc.resize(ng); r->Up.resize(ng);
   ew->mU.resize(ng); ew->mUm.resize(ng);
   for( int g=0 ; g < ng ; g++ )
   {
      int ifirst = ew->m_iStart[g], ilast = ew->m_iEnd[g];
      int jfirst = ew->m_jStart[g], jlast = ew->m_jEnd[g];
      int kfirst = ew->m_kStart[g], klast = ew->m_kEnd[g];
      r->F[g].define(3,ifirst,ilast,jfirst,jlast,kfirst,klast);
      r->Lu[g].define(3,ifirst,ilast,jfirst,jlast,kfirst,klast);
      r->Uacc[g].define(3,ifirst,ilast,jfirst,jlast,kfirst,klast);
      r->Up[g].define(3,ifirst,ilast,jfirst,jlast,kfirst,klast);
      ew->mU[g].define(3,ifirst,ilast,jfirst,jlast,kfirst,klast);
      ew->mUm[g].define(3,ifirst,ilast,jfirst,jlast,kfirst,klast);
      ew->mU[g].set_value(0.0);  ew->mUm[g].set_value(0.0);
      r->F[g].set_value(0.0);    r->Up[g].set_value(0.0);
      r->Uacc[g].set_value(0.0); r->Lu[g].set_value(0.0);
   }
   ew->BCForcing.resize(ng);
   for( int g=0 ; g < ng ; g++ )
   {
      ew->BCForcing[g] = new float_sw4*[6];
      for( int side=0 ; side < 6 ; side++ )
      {
	 ew->BCForcing[g][side] = NULL;
	 if( ew->m_bcType[g][side] == bStressFree || ew->m_bcType[g][side] == bDirichlet ||
	     ew->m_bcType[g][side] == bSuperGrid )
	 {
	    size_t n = 3*static_cast<size_t>(ew->m_NumberOfBCPoints[g][side]);
	    ew->BCForcing[g][side] = new float_sw4[n];
	    for( size_t q=0 ; q < n ; q++ ) ew->BCForcing[g][side][q] = 0;
	 }
      }
   }
   for( size_t ts=0 ; ts < ew->m_GlobalTimeSeries.size() ; ts++ )
      ew->m_GlobalTimeSeries[ts]->allocateRecordingArrays( ew->mNumberOfTimeSteps+1, ew->mTstart, ew->mDt );
   // initial data on receiver records, EW.C:2485-2502
   std::vector<float_sw4> uRec;
   for( size_t ts=0 ; ts < ew->m_GlobalTimeSeries.size() ; ts++ )
   {
      TimeSeries* s = ew->m_GlobalTimeSeries[ts];
      if( s->getMode() != TimeSeries::Velocity && s->myPoint() )
      {
	 ew->extractRecordData( s->getMode(), s->m_i0, s->m_j0, s->m_k0, s->m_grid0, uRec, ew->mUm, ew->mU );
	 s->recordData( uRec );
      }
   }
   r->t = ew->mTstart;
   r->step = 0;
   if( workdir && workdir[0] )
      if( chdir( r->cwd0.c_str() ) != 0 ) {}
   return r;
}

// scalar queries: what = "ngrids","ncart","nsteps","corder","topo","nsources","nrec","sgorder","usesg"
int ref_ew_int( void* h, const char* what )
{
   EW* ew = static_cast<RefRun*>(h)->ew;
   std::string w(what);
   if( w == "ngrids" ) return ew->mNumberOfGrids;
   if( w == "ncart" ) return ew->mNumberOfCartesianGrids;
   if( w == "nsteps" ) return ew->mNumberOfTimeSteps;
   if( w == "corder" ) return ew->m_corder ? 1 : 0;
   if( w == "topo" ) return ew->m_topography_exists ? 1 : 0;
   if( w == "npointsources" ) return (int)ew->m_point_sources.size();
   if( w == "nunique" ) return (int)ew->m_identsources.size()-1;
   if( w == "nrec" ) return (int)ew->m_GlobalTimeSeries.size();
   if( w == "sgorder" ) return ew->m_sg_damping_order;
   if( w == "usesg" ) return ew->m_use_supergrid ? 1 : 0;
   if( w == "step" ) return static_cast<RefRun*>(h)->step;
   if( w == "pointsourcetest" ) return ew->m_point_source_test ? 1 : 0;
   return -999999;
}

double ref_ew_double( void* h, const char* what )
{
   RefRun* r = static_cast<RefRun*>(h);
   EW* ew = r->ew;
   std::string w(what);
   if( w == "dt" ) return ew->mDt;
   if( w == "t" ) return r->t;
   if( w == "tstart" ) return ew->mTstart;
   if( w == "beta" ) return ew->m_supergrid_damping_coefficient;
   return -1e300;
}

// per-grid integers: out = {ifirst,ilast,jfirst,jlast,kfirst,klast, nx,ny,nz(global), onesided[6], bctype[6], wind[36], nbcpts[6]}
void ref_ew_grid_ints( void* h, int g, int* out )
{
   EW* ew = static_cast<RefRun*>(h)->ew;
   int p = 0;
   out[p++] = ew->m_iStart[g]; out[p++] = ew->m_iEnd[g];
   out[p++] = ew->m_jStart[g]; out[p++] = ew->m_jEnd[g];
   out[p++] = ew->m_kStart[g]; out[p++] = ew->m_kEnd[g];
   out[p++] = ew->m_global_nx[g]; out[p++] = ew->m_global_ny[g]; out[p++] = ew->m_global_nz[g];
   for( int s=0 ; s < 6 ; s++ ) out[p++] = ew->m_onesided[g][s];
   for( int s=0 ; s < 6 ; s++ ) out[p++] = static_cast<int>(ew->m_bcType[g][s]);
   for( int s=0 ; s < 36 ; s++ ) out[p++] = ew->m_BndryWindow[g][s];
   for( int s=0 ; s < 6 ; s++ ) out[p++] = ew->m_NumberOfBCPoints[g][s];
}

double ref_ew_grid_h( void* h, int g ) { return static_cast<RefRun*>(h)->ew->mGridSize[g]; }
double ref_ew_grid_zmin( void* h, int g ) { return static_cast<RefRun*>(h)->ew->m_zmin[g]; }

// pointer to a host array owned by the reference object (no copy).
double* ref_ew_array( void* h, const char* name, int g )
{
   RefRun* r = static_cast<RefRun*>(h);
   EW* ew = r->ew;
   std::string w(name);
   if( w == "U" ) return ew->mU[g].c_ptr();
   if( w == "Um" ) return ew->mUm[g].c_ptr();
   if( w == "Up" ) return r->Up[g].c_ptr();
   if( w == "F" ) return r->F[g].c_ptr();
   if( w == "Lu" ) return r->Lu[g].c_ptr();
   if( w == "Uacc" ) return r->Uacc[g].c_ptr();
   if( w == "mu" ) return ew->mMu[g].c_ptr();
   if( w == "lambda" ) return ew->mLambda[g].c_ptr();
   if( w == "rho" ) return ew->mRho[g].c_ptr();
   if( w == "metric" ) return ew->mMetric.c_ptr();
   if( w == "jac" ) return ew->mJ.c_ptr();
   if( w == "X" ) return ew->mX.c_ptr();
   if( w == "Y" ) return ew->mY.c_ptr();
   if( w == "Z" ) return ew->mZ.c_ptr();
   if( w == "strx" ) return ew->m_sg_str_x[g];
   if( w == "stry" ) return ew->m_sg_str_y[g];
   if( w == "strz" ) return ew->m_sg_str_z[g];
   if( w == "dcx" ) return ew->m_sg_dc_x[g];
   if( w == "dcy" ) return ew->m_sg_dc_y[g];
   if( w == "dcz" ) return ew->m_sg_dc_z[g];
   if( w == "cox" ) return ew->m_sg_corner_x[g];
   if( w == "coy" ) return ew->m_sg_corner_y[g];
   if( w == "coz" ) return ew->m_sg_corner_z[g];
   if( w == "acof" ) return ew->m_acof;
   if( w == "bope" ) return ew->m_bope;
   if( w == "ghcof" ) return ew->m_ghcof;
   if( w == "sbop" ) return ew->m_sbop;
   if( w.size() == 7 && w.substr(0,6) == "bforce" )
      return ew->BCForcing[g][w[6]-'0'];
   return 0;
}

// grid-point sources in the reference's sorted order: idx[4*s..] = {grid,i,j,k}; ident = m_identsources
void ref_ew_point_sources( void* h, int* idx, double* fxyz_notime )
{
   EW* ew = static_cast<RefRun*>(h)->ew;
   for( size_t s=0 ; s < ew->m_point_sources.size() ; s++ )
   {
      GridPointSource* p = ew->m_point_sources[s];
      idx[4*s] = p->m_grid; idx[4*s+1] = p->m_i0; idx[4*s+2] = p->m_j0; idx[4*s+3] = p->m_k0;
      if( fxyz_notime ) p->getFxyz_notime( fxyz_notime + 3*s );
   }
}
void ref_ew_identsources( void* h, int* ident )
{
   EW* ew = static_cast<RefRun*>(h)->ew;
   for( size_t s=0 ; s < ew->m_identsources.size() ; s++ ) ident[s] = ew->m_identsources[s];
}
// forcing of every grid-point source at time t (tt: second time derivative), fxyz[3*s..]
void ref_ew_eval_forces( void* h, double t, int tt, double* fxyz )
{
   EW* ew = static_cast<RefRun*>(h)->ew;
   for( size_t s=0 ; s < ew->m_point_sources.size() ; s++ )
   {
      if( tt ) ew->m_point_sources[s]->getFxyztt( t, fxyz+3*s );
      else     ew->m_point_sources[s]->getFxyz( t, fxyz+3*s );
   }
}

// receivers: out[4*r..] = {grid,i,j,k}, mode[r]
void ref_ew_receivers( void* h, int* out, int* mode )
{
   EW* ew = static_cast<RefRun*>(h)->ew;
   for( size_t ts=0 ; ts < ew->m_GlobalTimeSeries.size() ; ts++ )
   {
      TimeSeries* s = ew->m_GlobalTimeSeries[ts];
      out[4*ts] = s->m_grid0; out[4*ts+1] = s->m_i0; out[4*ts+2] = s->m_j0; out[4*ts+3] = s->m_k0;
      mode[ts] = static_cast<int>(s->getMode());
   }
}

// boundary forcing at time t into the reference's BCForcing arrays (EW.C:3320)
void ref_ew_bc_forcing( void* h, double t )
{
   EW* ew = static_cast<RefRun*>(h)->ew;
   ew->cartesian_bc_forcing( t, ew->BCForcing, ew->m_globalUniqueSources );
}

// phases of one reference CPU step, in the order of EW.C:2527-2878.  `upto` lets a test stop
// after a given phase to compare intermediate fields:
//  1 Force(t)  2 evalRHS(U)  3 predictor  4 halo  5 bc  6 Force_tt  7 dpdmt  8 evalRHS(Uacc)
//  9 corrector 10 supergrid damping 11 halo 12 bc 13 (t+=dt, receivers, cycle)
void ref_ew_step_phases( void* h, int from, int upto )
{
   RefRun* r = static_cast<RefRun*>(h);
   EW* ew = r->ew;
   std::vector<Sarray>& U = ew->mU;
   std::vector<Sarray>& Um = ew->mUm;
   double t = r->t;
   for( int ph=from ; ph <= upto ; ph++ )
   {
      switch( ph )
      {
      case 1: ew->Force( t, r->F, ew->m_point_sources, false ); break;
      case 2: ew->evalRHS( U, ew->mMu, ew->mLambda, r->Lu ); break;
      case 3: ew->evalPredictor( r->Up, U, Um, ew->mRho, r->Lu, r->F ); break;
      case 4: for( int g=0 ; g < ew->mNumberOfGrids ; g++ ) ew->communicate_array( r->Up[g], g ); break;
      case 5: ew->cartesian_bc_forcing( t+ew->mDt, ew->BCForcing, ew->m_globalUniqueSources );
	 ew->enforceBC( r->Up, ew->mMu, ew->mLambda, t+ew->mDt, ew->BCForcing ); break;
      case 6: ew->Force( t, r->F, ew->m_point_sources, true ); break;
      case 7: ew->evalDpDmInTime( r->Up, U, Um, r->Uacc ); break;
      case 8: ew->evalRHS( r->Uacc, ew->mMu, ew->mLambda, r->Lu ); break;
      case 9: ew->evalCorrector( r->Up, ew->mRho, r->Lu, r->F ); break;
      case 10: if( ew->m_use_supergrid ) ew->addSuperGridDamping( r->Up, U, Um, ew->mRho ); break;
      case 11: for( int g=0 ; g < ew->mNumberOfGrids ; g++ ) ew->communicate_array( r->Up[g], g ); break;
      case 12: ew->cartesian_bc_forcing( t+ew->mDt, ew->BCForcing, ew->m_globalUniqueSources );
	 ew->enforceBC( r->Up, ew->mMu, ew->mLambda, t+ew->mDt, ew->BCForcing ); break;
      case 13:
      {
	 r->t += ew->mDt;
	 r->step++;
	 std::vector<float_sw4> uRec;
	 for( size_t ts=0 ; ts < ew->m_GlobalTimeSeries.size() ; ts++ )
	 {
	    TimeSeries* s = ew->m_GlobalTimeSeries[ts];
	    if( s->myPoint() )
	    {
	       ew->extractRecordData( s->getMode(), s->m_i0, s->m_j0, s->m_k0, s->m_grid0, uRec, Um, r->Up );
	       s->recordData( uRec );
	    }
	 }
	 Sarray *d1=0, *d2=0, *d3=0;
	 ew->cycleSolutionArrays( Um, U, r->Up, d1, d2, d3 );
	 break;
      }
      default: break;
      }
   }
}

void ref_ew_step( void* h ) { ref_ew_step_phases( h, 1, 13 ); }

// rotate Um<-U<-Up<-Um host pointers and advance time without computing (used when a test
// has overwritten Up with a field computed elsewhere)
void ref_ew_cycle( void* h ) { ref_ew_step_phases( h, 13, 13 ); }

// point-source test error norms of field `u` (layout of the reference's U[g]) against the
// reference's exact solution at time t: out = {Linf, L2, norm of solution}  (EW.C:2884-2903)
int ref_ew_pointsource_error( void* h, double t, double** u_per_grid, double* out )
{
   RefRun* r = static_cast<RefRun*>(h);
   EW* ew = r->ew;
   if( !ew->m_point_source_test ) return 1;
   int ng = ew->mNumberOfGrids;
   std::vector<Sarray> Uex(ng), Ucmp(ng);
   for( int g=0 ; g < ng ; g++ )
   {
      Uex[g].define(3,ew->m_iStart[g],ew->m_iEnd[g],ew->m_jStart[g],ew->m_jEnd[g],ew->m_kStart[g],ew->m_kEnd[g]);
      Ucmp[g].define(3,ew->m_iStart[g],ew->m_iEnd[g],ew->m_jStart[g],ew->m_jEnd[g],ew->m_kStart[g],ew->m_kEnd[g]);
      Uex[g].set_value(0.0);
      memcpy( Ucmp[g].c_ptr(), u_per_grid[g], sizeof(double)*Ucmp[g].m_npts ); // m_npts counts all components
   }
   ew->exactSol( t, Uex, ew->m_globalUniqueSources );
   float_sw4 errInf=0, errL2=0, solInf=0;
   ew->normOfDifference( Uex, Ucmp, errInf, errL2, solInf, ew->m_globalUniqueSources );
   out[0] = errInf; out[1] = errL2; out[2] = solInf;
   return 0;
}

// recorded station data: component q of receiver r, n values starting at sample 0
int ref_ew_write_receivers( void* h, const char* workdir )
{
   RefRun* r = static_cast<RefRun*>(h);
   EW* ew = r->ew;
   if( workdir && workdir[0] )
      if( chdir( workdir ) != 0 ) return 1;
   for( size_t ts=0 ; ts < ew->m_GlobalTimeSeries.size() ; ts++ )
      ew->m_GlobalTimeSeries[ts]->writeFile();
   if( workdir && workdir[0] )
      if( chdir( r->cwd0.c_str() ) != 0 ) return 2;
   return 0;
}

int ref_num_threads()
{
   int n = 1;
#ifdef _OPENMP
#pragma omp parallel
   {
#pragma omp master
      n = omp_get_num_threads();
   }
#endif
   return n;
}

// the OpenMP team size of the reference kernels (torchrun exports OMP_NUM_THREADS=1 to its workers: the reference arm of the
// benchmark sets the host's core count explicitly)
void ref_set_num_threads( int n )
{
#ifdef _OPENMP
   if( n > 0 ) omp_set_num_threads( n );
#endif
}

} // extern "C"
