// Curvilinear-grid operators (top grid under topography).  Placeholder launchers: implemented in
// a later milestone; they fail loudly instead of falling back to anything.
#include "common.cuh"
namespace sw4b200 {
int launch_rhs4sgcurv( const Block&, const double*, const double*, const double*, const double*, const double*,
		       double*, int, const double*, const double*, cudaStream_t )
{ return set_error( "rhs4sgcurv: not implemented yet" ); }
int launch_addsgdc( int, const Block&, double*, const double*, const double*, const double*, const double*,
		    const double*, const double*, const double*, const double*, const double*, const double*,
		    double, cudaStream_t )
{ return set_error( "addsgdc: not implemented yet" ); }
int launch_freesurfcurvisg( const Block&, int, int, double*, const double*, const double*, const double*,
			    const double*, const double*, const double*, cudaStream_t )
{ return set_error( "freesurfcurvisg: not implemented yet" ); }
int launch_enforce_cart_topo( int, double*, const Block&, double*, const Block&, cudaStream_t )
{ return set_error( "enforce_cart_topo: not implemented yet" ); }
}
