// TEST INFRASTRUCTURE ONLY: runs the kernels of sw4lite_b200/csrc/rhs4sg_fast2.cu and rhs4sg_fast4.cu on the CPU through
// tests/emu/cuda_emu.h.  extern "C" entry for ctypes (tests/test_emu_fast.py).
#include "cuda_emu.h"
namespace emu {
thread_local dim3 t_threadIdx;
dim3 g_blockIdx, g_blockDim, g_gridDim;
double* g_smem = 0;
std::barrier<>* g_barrier = 0;
std::atomic<int> g_vote_false( 0 );
int g_force_general = 0, g_nostr_ctas = 0;
}
extern "C" int emu_nostr_ctas() { const int n = emu::g_nostr_ctas; emu::g_nostr_ctas = 0; return n; }
extern "C" void emu_force_general( int on ) { emu::g_force_general = on; }
#include "../../sw4lite_b200/csrc/rhs4sg_fast2.cu"
#include "../../sw4lite_b200/csrc/rhs4sg_fast4.cu"

using namespace sw4b200;

// what launch_fast4 does: the Lu epilogue in one launch of all tiles (KIND 0); the fused passes as a launch of the tiles with
// stretching (KIND 2) and one of the plain tiles (KIND 1) -- or, with the test hook emu_force_general, as ONE launch of all
// tiles with the general march, for the bit comparison of the two marches
#define EMU_FAST4_K( EPI_, SPLIT_, DER_, KIND_ ) \
   emu::launch( gs, bs, C4::SMEM_DOUBLES, [&]() { k_rhs_fast4<TY4, EPI_, SPLIT_, DER_, KIND_>( a, maps ); } )
#define EMU_FAST4( SPLIT_ )                                                                             \
   if( epi == EPI_LU ) EMU_FAST4_K( EPI_LU, SPLIT_, false, 0 );                                          \
   else if( epi == EPI_PRED )                                                                            \
   {                                                                                                     \
      if( emu::g_force_general ) EMU_FAST4_K( EPI_PRED, SPLIT_, true, 0 );                               \
      else { EMU_FAST4_K( EPI_PRED, SPLIT_, true, 2 ); EMU_FAST4_K( EPI_PRED, SPLIT_, true, 1 ); }       \
   }                                                                                                     \
   else                                                                                                  \
   {                                                                                                     \
      if( emu::g_force_general ) EMU_FAST4_K( EPI_CORR, SPLIT_, true, 0 );                               \
      else { EMU_FAST4_K( EPI_CORR, SPLIT_, true, 2 ); EMU_FAST4_K( EPI_CORR, SPLIT_, true, 1 ); }       \
   }

// gen 4: rhs4sg_fast4.cu (the product configuration; the SPLIT variant when pitch > ni), gen 2: rhs4sg_fast2.cu.
// pitch = row pitch of the arrays in doubles (ni, or ni+1 for a block whose odd rows are padded to an even pitch)
extern "C" int emu_rhs_fast( int gen, int epi, int ifirst, int ilast, int jfirst, int jlast, int kfirst, int klast, int pitch, int klo,
			     int khi, int kchunk, const double* u, const double* mu, const double* la, const double* strx,
			     const double* stry, const double* strz, double cof, double* out, double* out2,
			     const double* um, const double* rho, const double* fo, double fac )
{
   FastArgs a;
   a.b = make_block( 1, ifirst, ilast, jfirst, jlast, kfirst, klast, pitch > ilast - ifirst + 1 );
   if( a.b.ni != pitch ) return 1;
   a.klo = klo; a.khi = khi; a.kchunk = kchunk;
   const long long n = a.b.npts;
   for( int c = 0; c < 3; c++ )
   {
      a.u[c] = u + c * n;
      a.out[c] = out + c * n;
      a.out2[c] = out2 ? out2 + c * n : 0;
      a.um[c] = um ? um + c * n : 0;
      a.fo[c] = fo ? fo + c * n : 0;
   }
   a.la2 = a.rhoi = 0;
   a.mu = mu; a.la = la; a.strx = strx; a.stry = stry; a.strz = strz; a.cof6 = cof / 6; a.cof144 = cof / 144; a.rho = rho; a.fac = fac;
   if( gen == 4 )
   {
      constexpr int TY4 = 16;
      typedef fast4::Cfg<TY4> C4;
      dim3 bs( C4::NT, 1, 1 );
      dim3 gs( ( a.b.nil - 4 + C4::TX - 1 ) / C4::TX, ( a.b.nj - 4 + TY4 - 1 ) / TY4, ( khi - klo + 1 + kchunk - 1 ) / kchunk );
      FastMaps maps; // (the emulated TMA tile load reads through the array base)
      for( int c = 0; c < 3; c++ ) { maps.u[c].base = a.u[c]; maps.um[c].base = a.um[c]; }
      // the fused passes read the derived arrays of a grid block (2 mu + lambda, 1 / rho: launch_derive_materials)
      std::vector<double> la2, rhoi;
      if( epi != EPI_LU )
      {
	 la2.resize( n ); rhoi.resize( n );
	 for( long long q = 0; q < n; q++ ) { la2[q] = 2 * mu[q] + la[q]; rhoi[q] = 1.0 / rho[q]; }
	 a.la2 = la2.data(); a.rhoi = rhoi.data();
      }
      maps.mu.base = a.mu; maps.la.base = epi != EPI_LU ? a.la2 : a.la; maps.rho.base = epi != EPI_LU ? a.rhoi : a.rho;
      if( a.b.nil == a.b.ni )
      {
	 EMU_FAST4( 0 )
      }
      else
      {
	 EMU_FAST4( 1 )
      }
      return 0;
   }
   constexpr int TY = 8;
   typedef fast2::Cfg<TY> C2;
   dim3 bs( C2::TX, TY, 1 );
   dim3 gs( ( a.b.nil - 4 + C2::TX - 1 ) / C2::TX, ( a.b.nj - 4 + TY - 1 ) / TY, ( khi - klo + 1 + kchunk - 1 ) / kchunk );
   if( epi == EPI_LU ) emu::launch( gs, bs, C2::SMEM_DOUBLES, [&]() { k_rhs_fast2<TY, EPI_LU>( a ); } );
   else if( epi == EPI_PRED ) emu::launch( gs, bs, C2::SMEM_DOUBLES, [&]() { k_rhs_fast2<TY, EPI_PRED>( a ); } );
   else emu::launch( gs, bs, C2::SMEM_DOUBLES, [&]() { k_rhs_fast2<TY, EPI_CORR>( a ); } );
   return 0;
}
