// Single translation unit of libsw4b200.so (constant-memory tables are shared by all kernels).
#include "cart_v1.cu"
#include "rhs4sg_fast2.cu"
#include "rhs4sg_fast4.cu"
#include "addsgd_fast.cu"
#include "curvilinear.cu"
#include "peaks.cu"
#include "exchange.cu"
#include "api.cu"
