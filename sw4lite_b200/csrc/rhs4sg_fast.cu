// Throughput path for the SoA Cartesian interior (placeholder: routes to the general kernel).
#include "common.cuh"
namespace sw4b200 {
int launch_rhs_fast( RhsMode mode, const RhsArgs& a, cudaStream_t st ) { return launch_rhs_v1( mode, a, st ); }
}
