// Supergrid damping, 4th order, SoA layout: 2.5-D z-marching kernel over one box of the damping
// layers.  Reference arithmetic: EW::addsgd4fort_indrev (ew-cfromfort.C:952-1056):
//   up -= beta/rho * [ strx coy coz D_x + stry cox coz D_y + strz cox coy D_z ],
//   D_d = delta_d^2( rho dc_d delta_d^2 (u - um) )
// The difference d = u - um of the current plane is staged in shared memory once per point (ring of
// width 2 filled by helper loads); the z direction comes from a register column.
#include "common.cuh"

namespace sw4b200 {
namespace {

constexpr int SG_TX = 32, SG_TY = 8, SG_PX = SG_TX + 4, SG_PY = SG_TY + 4;

__device__ __forceinline__ double sg_term( double dm2, double dm1, double d0, double dp1, double dp2, double rm, double r0,
					   double rp, double cm, double c0, double cp )
{
   const double e0 = rm * cm * ( d0 - 2 * dm1 + dm2 );
   const double e1 = r0 * c0 * ( dp1 - 2 * d0 + dm1 );
   const double e2 = rp * cp * ( dp2 - 2 * dp1 + d0 );
   return e2 - 2 * e1 + e0;
}

__global__ void __launch_bounds__( SG_TX* SG_TY ) k_addsgd4_fast( Block b, Int6 box, int kchunk, double* __restrict__ up,
								   const double* __restrict__ u, const double* __restrict__ um,
								   const double* __restrict__ rho, const double* __restrict__ dcx,
								   const double* __restrict__ dcy, const double* __restrict__ dcz,
								   const double* __restrict__ strx, const double* __restrict__ stry,
								   const double* __restrict__ strz, const double* __restrict__ cox,
								   const double* __restrict__ coy, const double* __restrict__ coz,
								   double beta )
{
   __shared__ double sd[3][SG_PY][SG_PX];
   __shared__ double sr[SG_PY][SG_PX];
   const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * SG_TX + tx;
   const int li0 = box.v[0] + blockIdx.x * SG_TX, lj0 = box.v[2] + blockIdx.y * SG_TY;
   const int ka = box.v[4] + blockIdx.z * kchunk;
   const int kb = ka + kchunk - 1 < box.v[5] ? ka + kchunk - 1 : box.v[5];
   if( ka > kb ) return;
   const int li = li0 + tx, lj = lj0 + ty;
   const bool inarr = li < b.ni && lj < b.nj;		 // loads
   const bool act = li <= box.v[1] && lj <= box.v[3];	 // update
   const long long own = (long long)lj * b.ni + li;
   // ring point handled by this thread (tid < 176): the 2-wide frame around the tile
   int hsx = -1, hsy = -1;
   if( tid < 2 * SG_PX ) { hsy = tid / SG_PX; hsx = tid % SG_PX; }
   else if( tid < 4 * SG_PX ) { hsy = SG_TY + 2 + ( tid - 2 * SG_PX ) / SG_PX; hsx = ( tid - 2 * SG_PX ) % SG_PX; }
   else if( tid < 4 * SG_PX + 4 * SG_TY )
   {
      const int t = tid - 4 * SG_PX;
      hsy = 2 + t / 4;
      hsx = ( t & 3 ) < 2 ? ( t & 3 ) : SG_TX + ( t & 3 );
   }
   const int hli = li0 - 2 + hsx, hlj = lj0 - 2 + hsy;
   const bool hval = hsx >= 0 && hli >= 0 && hlj >= 0 && hli < b.ni && hlj < b.nj;
   const long long hoff = hval ? (long long)hlj * b.ni + hli : 0;

   double d[3][5], rh[3];
#pragma unroll
   for( int c = 0; c < 3; c++ )
#pragma unroll
      for( int m = 0; m < 5; m++ ) d[c][m] = 0;
   rh[0] = rh[1] = rh[2] = 0;
   // prologue: planes ka-2 .. ka+1 of d, rho(ka-1), rho(ka)
   if( inarr )
   {
#pragma unroll
      for( int m = 1; m < 5; m++ )
      {
	 const long long q = b.nij * ( ka - 3 + m ) + own;
#pragma unroll
	 for( int c = 0; c < 3; c++ ) d[c][m] = u[c * b.npts + q] - um[c * b.npts + q];
      }
      rh[1] = rho[b.nij * ( ka - 1 ) + own];
      rh[2] = rho[b.nij * ka + own];
   }
   const double cxm = dcx[li < b.ni ? li - 1 : 0], cx0 = dcx[li < b.ni ? li : 0], cxp = dcx[li + 1 < b.ni ? li + 1 : 0];
   const double cym = dcy[lj < b.nj ? lj - 1 : 0], cy0 = dcy[lj < b.nj ? lj : 0], cyp = dcy[lj + 1 < b.nj ? lj + 1 : 0];
   const double sxi = strx[li < b.ni ? li : 0], syj = stry[lj < b.nj ? lj : 0];
   const double cxi = cox[li < b.ni ? li : 0], cyj = coy[lj < b.nj ? lj : 0];

   for( int k = ka; k <= kb; k++ )
   {
      // rotate the columns, fetch plane k+2 of d and rho(k+1)
#pragma unroll
      for( int c = 0; c < 3; c++ )
#pragma unroll
	 for( int m = 0; m < 4; m++ ) d[c][m] = d[c][m + 1];
      rh[0] = rh[1]; rh[1] = rh[2];
      double upv[3] = { 0, 0, 0 };
      if( inarr )
      {
	 const long long q2 = b.nij * ( k + 2 ) + own;
#pragma unroll
	 for( int c = 0; c < 3; c++ ) d[c][4] = u[c * b.npts + q2] - um[c * b.npts + q2];
	 rh[2] = rho[b.nij * ( k + 1 ) + own];
      }
      if( act )
      {
	 const long long q = b.nij * k + own;
#pragma unroll
	 for( int c = 0; c < 3; c++ ) upv[c] = up[c * b.npts + q];
      }
      sd[0][ty + 2][tx + 2] = d[0][2]; sd[1][ty + 2][tx + 2] = d[1][2]; sd[2][ty + 2][tx + 2] = d[2][2];
      sr[ty + 2][tx + 2] = rh[1];
      if( hsx >= 0 )
      {
	 double h0 = 0, h1 = 0, h2 = 0, hr = 0;
	 if( hval )
	 {
	    const long long q = b.nij * k + hoff;
	    h0 = u[q] - um[q];
	    h1 = u[b.npts + q] - um[b.npts + q];
	    h2 = u[2 * b.npts + q] - um[2 * b.npts + q];
	    hr = rho[q];
	 }
	 sd[0][hsy][hsx] = h0; sd[1][hsy][hsx] = h1; sd[2][hsy][hsx] = h2; sr[hsy][hsx] = hr;
      }
      __syncthreads();
      if( act )
      {
	 const double czm = dcz[k - 1], cz0 = dcz[k], czp = dcz[k + 1];
	 const double szk = strz[k], czk = coz[k];
	 const double prex = sxi * cyj * czk, prey = syj * cxi * czk, prez = szk * cxi * cyj;
	 const double birho = beta / rh[1];
	 const int X = tx + 2, Y = ty + 2;
	 const double rxm = sr[Y][X - 1], rxp = sr[Y][X + 1], rym = sr[Y - 1][X], ryp = sr[Y + 1][X];
#pragma unroll
	 for( int c = 0; c < 3; c++ )
	 {
	    const double s = prex * sg_term( sd[c][Y][X - 2], sd[c][Y][X - 1], d[c][2], sd[c][Y][X + 1], sd[c][Y][X + 2], rxm,
					     rh[1], rxp, cxm, cx0, cxp ) +
			     prey * sg_term( sd[c][Y - 2][X], sd[c][Y - 1][X], d[c][2], sd[c][Y + 1][X], sd[c][Y + 2][X], rym,
					     rh[1], ryp, cym, cy0, cyp ) +
			     prez * sg_term( d[c][0], d[c][1], d[c][2], d[c][3], d[c][4], rh[0], rh[1], rh[2], czm, cz0, czp );
	    up[c * b.npts + b.nij * k + own] = upv[c] - birho * s;
	 }
      }
      __syncthreads();
   }
}

// Boxes in which only dcz is non-zero (the bottom layer away from the x/y layers: most of the damping points of a
// production grid).  The x and y terms of addsgd4fort_indrev are exactly zero there, so the update reduces to the z term:
// no neighbours in the plane, hence no shared memory and no barriers -- a pure streaming march down the z columns
// (104 B per point).  Same arithmetic as k_addsgd4_fast for its z term, so the two kernels agree bit for bit where
// both apply.
__global__ void __launch_bounds__( 256 ) k_addsgd4_zonly( Block b, Int6 box, int kchunk, double* __restrict__ up, const double* __restrict__ u,
							    const double* __restrict__ um, const double* __restrict__ rho,
							    const double* __restrict__ dcz, const double* __restrict__ strz,
							    const double* __restrict__ cox, const double* __restrict__ coy, double beta )
{
   const int li = box.v[0] + blockIdx.x * 32 + threadIdx.x, lj = box.v[2] + blockIdx.y * 8 + threadIdx.y;
   const int ka = box.v[4] + blockIdx.z * kchunk;
   const int kb = ka + kchunk - 1 < box.v[5] ? ka + kchunk - 1 : box.v[5];
   if( ka > kb || li > box.v[1] || lj > box.v[3] ) return;
   const long long own = (long long)lj * b.ni + li;
   const double cxi = cox[li], cyj = coy[lj];
   double d[3][5], rh[3];
#pragma unroll
   for( int c = 0; c < 3; c++ )
   {
      d[c][0] = 0;
#pragma unroll
      for( int m = 1; m < 5; m++ )
      {
	 const long long q = b.nij * ( ka - 3 + m ) + own;
	 d[c][m] = u[c * b.npts + q] - um[c * b.npts + q];
      }
   }
   rh[0] = 0; rh[1] = rho[b.nij * ( ka - 1 ) + own]; rh[2] = rho[b.nij * ka + own];
   double upn[3]; // up of the next plane, fetched one plane ahead
#pragma unroll
   for( int c = 0; c < 3; c++ ) upn[c] = up[c * b.npts + b.nij * ka + own];
   for( int k = ka; k <= kb; k++ )
   {
#pragma unroll
      for( int c = 0; c < 3; c++ )
#pragma unroll
	 for( int m = 0; m < 4; m++ ) d[c][m] = d[c][m + 1];
      rh[0] = rh[1]; rh[1] = rh[2];
      const long long q = b.nij * k + own, q2 = q + 2 * b.nij;
      const double upv[3] = { upn[0], upn[1], upn[2] };
#pragma unroll
      for( int c = 0; c < 3; c++ ) d[c][4] = u[c * b.npts + q2] - um[c * b.npts + q2];
      rh[2] = rho[q + b.nij];
      if( k < kb )
      {
#pragma unroll
	 for( int c = 0; c < 3; c++ ) upn[c] = up[c * b.npts + q + b.nij];
      }
      const double prez = strz[k] * cxi * cyj; // as k_addsgd4_fast forms it
      const double birho = beta / rh[1];
      const double czm = dcz[k - 1], cz0 = dcz[k], czp = dcz[k + 1];
#pragma unroll
      for( int c = 0; c < 3; c++ )
      {
	 const double s = prez * sg_term( d[c][0], d[c][1], d[c][2], d[c][3], d[c][4], rh[0], rh[1], rh[2], czm, cz0, czp );
	 up[c * b.npts + q] = upv[c] - birho * s;
      }
   }
}

} // namespace

int launch_addsgd4_zonly( const Block& b, const Int6& box, double* up, const double* u, const double* um, const double* rho,
			  const double* dcz, const double* strz, const double* cox, const double* coy, double beta, cudaStream_t st )
{
   const int nx = box.v[1] - box.v[0] + 1, ny = box.v[3] - box.v[2] + 1, nz = box.v[5] - box.v[4] + 1;
   if( beta == 0 || nx <= 0 || ny <= 0 || nz <= 0 ) return 0;
   ProfScope prof( "addsgd", st );
   ProfScope prof2( "addsgd_zonly", st );
   const long long tiles = (long long)( ( nx + 31 ) / 32 ) * ( ( ny + 7 ) / 8 );
   long long nch = ( 148LL * 32 + tiles - 1 ) / tiles;
   if( nch < 1 ) nch = 1;
   int kchunk = (int)( ( nz + nch - 1 ) / nch );
   if( kchunk < 16 ) kchunk = 16; // (every chunk re-reads 4 planes of u, um)
   if( kchunk > nz ) kchunk = nz;
   dim3 bs( 32, 8, 1 );
   dim3 gs( ( nx + 31 ) / 32, ( ny + 7 ) / 8, ( nz + kchunk - 1 ) / kchunk );
   k_addsgd4_zonly<<<gs, bs, 0, st>>>( b, box, kchunk, up, u, um, rho, dcz, strz, cox, coy, beta );
   count_launch();
   return check_launch( "k_addsgd4_zonly" );
}

int launch_addsgd4_fast( const Block& b, const Int6& box, double* up, const double* u, const double* um,
			 const double* rho, const double* dcx, const double* dcy, const double* dcz, const double* strx,
			 const double* stry, const double* strz, const double* cox, const double* coy, const double* coz,
			 double beta, cudaStream_t st )
{
   const int nx = box.v[1] - box.v[0] + 1, ny = box.v[3] - box.v[2] + 1, nz = box.v[5] - box.v[4] + 1;
   if( beta == 0 || nx <= 0 || ny <= 0 || nz <= 0 ) return 0;
   ProfScope prof( "addsgd", st );
   const long long tiles = (long long)( ( nx + SG_TX - 1 ) / SG_TX ) * ( ( ny + SG_TY - 1 ) / SG_TY );
   long long nch = ( 148LL * 16 + tiles - 1 ) / tiles;
   if( nch < 1 ) nch = 1;
   int kchunk = (int)( ( nz + nch - 1 ) / nch );
   if( kchunk < 8 ) kchunk = 8;
   if( kchunk > nz ) kchunk = nz;
   dim3 bs( SG_TX, SG_TY, 1 );
   dim3 gs( ( nx + SG_TX - 1 ) / SG_TX, ( ny + SG_TY - 1 ) / SG_TY, ( nz + kchunk - 1 ) / kchunk );
   k_addsgd4_fast<<<gs, bs, 0, st>>>( b, box, kchunk, up, u, um, rho, dcx, dcy, dcz, strx, stry, strz, cox, coy, coz, beta );
   count_launch();
   return check_launch( "k_addsgd4_fast" );
}

} // namespace sw4b200
