/*
 * sw4b200.h -- C-ABI of the B200-native (sm_100a) implementation of SW4's explicit elastic
 * time-stepping hot path.  Plain pointers and sizes only; no C++/torch types.
 *
 * This is the drop-in boundary for the GPU entry points that sw4lite's EW::timesteploop
 * branches to under m_cuobj->has_gpu() (reference EW.C:2536-2842), which bottom out in the
 * launchers of reference src/device-routines.h:290-390 and the CUDA members of src/EW_cuda.C.
 * Each entry point below names the reference interface it replaces (file:line in
 * /root/reference/src).  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions (same as the reference CPU kernels rhs4sg.C / ew-cfromfort.C):
 *  - index bounds are inclusive, global, Fortran style and include the ghost points:
 *    ifirst..ilast, jfirst..jlast, kfirst..klast (interior = [first+2, last-2]);
 *    `nk` / `nz` is the GLOBAL number of interior points in k of the grid.
 *  - all arrays are fp64.  Field layout is selected by `corder` like Sarray::m_corder
 *    (Sarray.C:753-778): corder=1 -> (i,j,k,c), component stride npts ("SoA");
 *    corder=0 -> (c,i,j,k), component fastest ("AoS").  Scalar fields are i-fastest.
 *  - pointers named d_* are DEVICE pointers; h_* are HOST pointers.  Small index/flag arrays
 *    (onesided[6], wind[36], bccnd[6]) are always host pointers and are passed by value to kernels.
 *  - `stream` is a cudaStream_t passed as void* (NULL = the library's stream 0).
 *  - every function returns 0 on success, non-zero on failure (sw4b200_last_error() describes
 *    it); nothing here falls back to a CPU path.
 */
#ifndef SW4B200_H
#define SW4B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- lifecycle
 * replaces EW::find_cuda_device (EW_cuda.C:560), EWCuda::EWCuda/initialize_gpu/sync_stream/
 * sync_device/reset_gpu (EWCuda.h:39-53), setupgpu (EWCuda.C:94). */
int sw4b200_init( int device );            /* select device, create the library's 4 streams   */
int sw4b200_finalize( void );
int sw4b200_device_count( void );
const char* sw4b200_last_error( void );
const char* sw4b200_version( void );
void* sw4b200_stream( int st );            /* cudaStream_t of library stream st (0..3)         */
int sw4b200_sync_stream( int st );
int sw4b200_sync_device( void );
int sw4b200_kernel_launch_count( void );   /* number of kernels launched by this library so far */
/* per-kernel device time, measured with CUDA events on the launching stream (the timing hooks of
 * EW::timesteploop's time_measure[], EW.C:2529-2873, at kernel granularity).  Names: "rhs_fast_pred",
 * "rhs_fast_corr", "rhs_fast_lu", "rhs_v1", "addsgd", "shell", "bc".  "tiles_plain" / "tiles_general": no time, `launches` =
 * thread blocks of the interior kernel since the last reset that ran the march without / with the stretching factors
 * (tiles on which strx = stry = 1 / the others). */
int sw4b200_profile_enable( int on );
int sw4b200_profile_reset( void );
int sw4b200_profile_read( const char* kernel, double* ms_total, long long* launches );
/* run-time switches between equivalent kernels, for A/B measurements and tests (all default to 1):
 *   "sgd_zonly"  damping boxes in which only dcz is non-zero take the streaming z-only kernel (bit-identical to the
 *                general kernel there; the reference has one kernel, addsgd4fort, ew-cfromfort.C:748) */
int sw4b200_set_option( const char* name, int value );

/* ---------------------------------------------------------------- memory
 * replaces Sarray::allocate_on_device / copy_to_device / copy_from_device / page_lock
 * (Sarray.C:818-947) for raw buffers. */
void* sw4b200_malloc( size_t bytes );
int sw4b200_free( void* d_ptr );
void* sw4b200_malloc_host( size_t bytes ); /* pinned host memory */
int sw4b200_free_host( void* h_ptr );
int sw4b200_memcpy_h2d( void* d_dst, const void* h_src, size_t bytes, void* stream );
int sw4b200_memcpy_d2h( void* h_dst, const void* d_src, size_t bytes, void* stream );
int sw4b200_memcpy_d2d( void* d_dst, const void* d_src, size_t bytes, void* stream );
int sw4b200_memset_zero( void* d_ptr, size_t bytes, void* stream );

/* ---------------------------------------------------------------- SBP coefficients
 * sw4b200_get_stencil_coefficients: EW::GetStencilCoefficients (ew-cfromfort.C:1487-1944), host.
 * sw4b200_copy_stencilcoefficients: copy_stencilcoefficients1 (device-routines.h:3,
 *   device-routines.C:70-94) -> constant memory.  Called implicitly with the built-in tables by
 *   sw4b200_init; call it only to override them. */
int sw4b200_get_stencil_coefficients( double* h_acof /*384*/, double* h_ghcof /*6*/,
                                      double* h_bope /*48*/, double* h_sbop /*5*/ );
int sw4b200_copy_stencilcoefficients( const double* h_acof, const double* h_ghcof,
                                      const double* h_bope, const double* h_sbop );

/* ---------------------------------------------------------------- unfused operators
 * One entry per reference CPU kernel, same argument meaning, device pointers. */

/* rhs4sg / rhs4sg_rev (rhs4sg.C:38, rhs4sg_rev.C:44); GPU twins evalRHSCU (EW_cuda.C:57).
 * lu = (1/h^2) L(u) on the interior; closure rows k=1..6 if onesided[4], nk-5..nk if onesided[5]. */
int sw4b200_rhs4sg( int corder, int ifirst, int ilast, int jfirst, int jlast, int kfirst, int klast,
                    int nk, const int* h_onesided, double* d_lu, const double* d_u,
                    const double* d_mu, const double* d_lambda, double h,
                    const double* d_strx, const double* d_stry, const double* d_strz, void* stream );

/* EW::predfort (ew-cfromfort.C:70); evalPredictorCU (EW_cuda.C:205): up = 2u-um+dt2/rho*(lu+fo) */
int sw4b200_predfort( int corder, int ib, int ie, int jb, int je, int kb, int ke, double* d_up,
                      const double* d_u, const double* d_um, const double* d_lu, const double* d_fo,
                      const double* d_rho, double dt2, void* stream );
/* EW::corrfort (ew-cfromfort.C:40); evalCorrectorCU (EW_cuda.C:236): up += dt4/(12 rho)*(lu+fo) */
int sw4b200_corrfort( int corder, int ib, int ie, int jb, int je, int kb, int ke, double* d_up,
                      const double* d_lu, const double* d_fo, const double* d_rho, double dt4,
                      void* stream );
/* EW::dpdmtfort (ew-cfromfort.C:116); dpdmt_dev (device-routines.C:186): u2 = dt2i*(up-2u+um) */
int sw4b200_dpdmtfort( int ib, int ie, int jb, int je, int kb, int ke, const double* d_up,
                       const double* d_u, const double* d_um, double* d_u2, double dt2i, void* stream );

/* EW::addsgd4fort(_indrev) / addsgd6fort(_indrev) (ew-cfromfort.C:748,952 / 852,1059);
 * addsgd4_gpu (device-routines.h:360), addSuperGridDampingCU (EW_cuda.C:291).  order = 4 or 6. */
int sw4b200_addsgd( int corder, int order, int ifirst, int ilast, int jfirst, int jlast, int kfirst, int klast,
                    double* d_up, const double* d_u, const double* d_um, const double* d_rho,
                    const double* d_dcx, const double* d_dcy, const double* d_dcz,
                    const double* d_strx, const double* d_stry, const double* d_strz,
                    const double* d_cox, const double* d_coy, const double* d_coz, double beta, void* stream );

/* EW::bcfortsg(_indrev) (ew-cfromfort.C:205,480); bcfortsg_gpu (device-routines.h:384), enforceBCCU
 * (EW_cuda.C:2070).  bccnd values are boundaryConditionType (sw4.h): 0 StressFree, 1 Dirichlet,
 * 2 SuperGrid, 3 Periodic; others are ignored.  Unlike the reference CUDA kernel, the whole
 * window of every side is filled (as the CPU kernel does).  d_bforce[s] may be NULL for sides
 * without forcing. */
int sw4b200_bcfortsg( int corder, int ib, int ie, int jb, int je, int kb, int ke, const int* h_wind /*36*/,
                      int nx, int ny, int nz, double* d_u, double h, const int* h_bccnd /*6*/,
                      const double* d_mu, const double* d_lambda,
                      const double* const* h_d_bforce /*6 device pointers in a host array*/,
                      const double* d_strx, const double* d_stry, void* stream );

/* curvilinear grid (top grid under topography) */
/* rhs4sgcurv / rhs4sgcurv_rev (rhs4sgcurv.C:34, rhs4sgcurv_rev.C:34) */
int sw4b200_rhs4sgcurv( int corder, int ifirst, int ilast, int jfirst, int jlast, int kfirst, int klast,
                        const double* d_u, const double* d_mu, const double* d_lambda, const double* d_met,
                        const double* d_jac, double* d_lu, const int* h_onesided,
                        const double* d_strx, const double* d_stry, void* stream );
/* EW::addsgd4cfort(_indrev) / addsgd6cfort(_indrev) (ew-cfromfort.C:1160,1323 / 1241,1405) */
int sw4b200_addsgdc( int corder, int order, int ifirst, int ilast, int jfirst, int jlast, int kfirst, int klast,
                     double* d_up, const double* d_u, const double* d_um, const double* d_rho,
                     const double* d_dcx, const double* d_dcy, const double* d_strx, const double* d_stry,
                     const double* d_jac, const double* d_cox, const double* d_coy, double beta, void* stream );
/* EW::freesurfcurvisg(_rev) (curvilinear-c.C:465,621) */
int sw4b200_freesurfcurvisg( int corder, int ib, int ie, int jb, int je, int kb, int ke, int nz, int side,
                             double* d_u, const double* d_mu, const double* d_lambda, const double* d_met,
                             const double* d_forcing, const double* d_strx, const double* d_stry, void* stream );
/* EW::enforceCartTopo (EW.C:3504-3531): interface injection between the top Cartesian grid
 * (bounds b*) and the curvilinear grid above it (bounds c*) */
int sw4b200_enforce_cart_topo( int corder, double* d_ucart, int ib, int ie, int jb, int je, int kb, int ke,
                               double* d_ucurv, int ckb, int cke, void* stream );

/* ---------------------------------------------------------------- fused step kernels
 * replace rhs4_pred_gpu / rhs4_X/Y_pred_gpu / rhs4_lowk/highk_pred_gpu (device-routines.h:290-358)
 * and their callers RHSPredCU_center/_boundary (EW_cuda.C:1228,1266): ONE call covers the whole
 * block -- interior rows, SBP closure rows and the 2-point shell where L(u)=0.
 *   up = 2u - um + dt^2/rho * ( L(u)/h^2 + fo )          d_fo may be NULL (treated as 0)   */
int sw4b200_rhs4_pred( int corder, int ifirst, int ilast, int jfirst, int jlast, int kfirst, int klast,
                       int nk, const int* h_onesided, double* d_up, const double* d_u, const double* d_um,
                       const double* d_mu, const double* d_lambda, const double* d_rho, const double* d_fo,
                       const double* d_strx, const double* d_stry, const double* d_strz,
                       double h, double dt, void* stream );
/* replace dpdmt_dev + rhs4_corr_gpu family + addsgd4_gpu family (device-routines.h:7,297-382;
 * RHSCorrCU_*, addSuperGridDampingCU_* EW_cuda.C:1325-1510):
 *   uacc = (up - 2u + um)/dt^2 ;  up += dt^4/(12 rho) * ( L(uacc)/h^2 + fo ) ;
 *   up -= supergrid damping(u-um) of order sg_order (0 = none).   d_up_out may equal d_up only when
 *   the library says so: the corrector reads up with a 2-point halo, so it writes to d_up_out. */
int sw4b200_rhs4_corr( int corder, int ifirst, int ilast, int jfirst, int jlast, int kfirst, int klast,
                       int nk, const int* h_onesided, double* d_up_out, const double* d_up,
                       const double* d_u, const double* d_um,
                       const double* d_mu, const double* d_lambda, const double* d_rho, const double* d_fo,
                       const double* d_strx, const double* d_stry, const double* d_strz,
                       const double* d_dcx, const double* d_dcy, const double* d_dcz,
                       const double* d_cox, const double* d_coy, const double* d_coz,
                       double beta, int sg_order, double h, double dt, void* stream );

/* replaces rhs4_corr_gpu (device-routines.h:297) as the reference's time loop calls it (RHSCorrCU_center, EW_cuda.C:1325):
 * the corrector with the acceleration already formed by evalDpDmInTimeCU, in place on up:
 *   up += dt^4/(12 rho) * ( L(uacc)/h^2 + fo )              d_fo may be NULL */
int sw4b200_rhs4_corr_acc( int corder, int ifirst, int ilast, int jfirst, int jlast, int kfirst, int klast,
                           int nk, const int* h_onesided, double* d_up, const double* d_uacc, const double* d_mu,
                           const double* d_lambda, const double* d_rho, const double* d_fo, const double* d_strx,
                           const double* d_stry, const double* d_strz, double h, double dt, void* stream );

/* ---------------------------------------------------------------- sparse forcing / receivers
 * replaces EW::ForceCU (EW_cuda.C:709) + forcing_dev (device-routines.C:8306): adds
 * factor/rho(p) * f to up at n unique source points; d_pidx = flat point index (i-fastest),
 * d_f = 3*n values (point-major).  Equivalent to a dense F that is zero elsewhere. */
int sw4b200_add_point_forces( int corder, size_t npts, double* d_up, const double* d_rho, int n,
                              const long long* d_pidx, const double* d_f, double factor, void* stream );
/* replaces extractRecordDataCU (EW_cuda.C:2180) for displacement receivers: gathers the 3
 * components at n points into d_out[3*n] */
int sw4b200_gather_points( int corder, size_t npts, const double* d_u, int n, const long long* d_pidx,
                           double* d_out, void* stream );

/* ---------------------------------------------------------------- host-buffer entry points
 * Same operators on HOST arrays: the library stages through its own device buffers (H2D, kernel,
 * D2H on stream 0, synchronous on return).  This is the call a host code without any CUDA of
 * its own makes, and the one bench.py times for the end-to-end number of the kernel benchmark. */
int sw4b200_rhs4sg_host( int corder, int ifirst, int ilast, int jfirst, int jlast, int kfirst, int klast,
                         int nk, const int* h_onesided, double* h_lu, const double* h_u,
                         const double* h_mu, const double* h_lambda, double h,
                         const double* h_strx, const double* h_stry, const double* h_strz );

/* ---------------------------------------------------------------- grid-block solver object
 * The device-resident state of one grid block (one z-slab of one grid on one GPU) and the
 * sequence of one time step on it: the part of EW::timesteploop (EW.C:2527-2878) between
 * "fields are on the device" and "cycle the arrays".  Replaces the per-grid bodies of
 * RHSPredCU_*, RHSCorrCU_*, addSuperGridDampingCU_*, enforceBCCU, ForceCU, cycleSolutionArrays.
 */
typedef struct sw4b200_grid sw4b200_grid;

typedef struct
{
   int corder;                 /* Sarray::m_corder                                              */
   int ifirst, ilast, jfirst, jlast, kfirst, klast; /* block bounds incl. ghost/halo points     */
   int nx, ny, nz;             /* global interior sizes (m_global_nx/ny/nz)                      */
   double h;                   /* grid size                                                      */
   double dt;
   int onesided[6];            /* m_onesided[g]                                                  */
   int bctype[6];              /* m_bcType[g] as ints                                            */
   int wind[36];               /* m_BndryWindow[g]                                               */
   int sg_order;               /* 0 (no supergrid damping), 4 or 6                               */
   double beta;                /* m_supergrid_damping_coefficient                                */
   int curvilinear;            /* 1: rhs4sgcurv/addsgd4c/freesurfcurvisg path (needs met, jac)   */
   int halo_lo, halo_hi;       /* 1: the low/high-k 2 planes are halos owned by a neighbour slab  */
} sw4b200_grid_desc;

sw4b200_grid* sw4b200_grid_create( const sw4b200_grid_desc* desc );
int sw4b200_grid_destroy( sw4b200_grid* g );
/* names: "U","Um","Up","Uacc" (3*npts), "mu","lambda","rho","jac" (npts), "metric" (4*npts),
 * "strx","dcx","cox" (ni), "stry","dcy","coy" (nj), "strz","dcz","coz" (nk),
 * "bforce0".."bforce5" (3*points of the side window).  Host arrays are always in the reference's
 * Sarray layout (Sarray.C:753-778) with npts = ni*nj*nk.  On the device the rows of an (i,j,k,c) block with odd
 * ni are padded to an even pitch (sw4b200_grid_row_pitch doubles; 16-byte aligned rows for the TMA-staged
 * kernels): upload/download convert, code that uses sw4b200_grid_device_ptr must honour the pitch.
 * Materials are time-invariant: the block keeps 2 mu + lambda and 1 / rho next to them for its fused passes and refreshes
 * the two after sw4b200_grid_upload / _fill_profile / _device_ptr of "mu", "lambda" or "rho" -- code that writes the materials
 * through a device pointer AFTER the first time step must ask for the pointer again (or upload) afterwards. */
int sw4b200_grid_upload( sw4b200_grid* g, const char* name, const double* h_src );
int sw4b200_grid_download( sw4b200_grid* g, const char* name, double* h_dst );
void* sw4b200_grid_device_ptr( sw4b200_grid* g, const char* name );
size_t sw4b200_grid_array_size( sw4b200_grid* g, const char* name ); /* number of doubles (host view) */
int sw4b200_grid_row_pitch( sw4b200_grid* g );                       /* device row pitch in doubles (ni or ni+1) */
/* unique source points of this block: (i,j,k) global indices, n points */
int sw4b200_grid_set_source_points( sw4b200_grid* g, int n, const int* h_ijk /*3*n*/ );
/* receivers (displacement) */
int sw4b200_grid_set_receiver_points( sw4b200_grid* g, int n, const int* h_ijk /*3*n*/ );

/* step phases; h_f / h_ftt = 3*n source values at time t (F and its 2nd time derivative) or NULL */
int sw4b200_grid_predictor( sw4b200_grid* g, const double* h_f );      /* EW.C:2537-2584           */
/* EW::enforceCartTopo (EW.C:3504-3531) on the new solution (Up) of the top Cartesian block and of the
 * curvilinear block above it; call after sw4b200_grid_enforce_bc of both, as EW::enforceBC does (EW.C:3500) */
int sw4b200_grid_enforce_cart_topo( sw4b200_grid* gcart, sw4b200_grid* gcurv );
int sw4b200_grid_enforce_bc( sw4b200_grid* g );                        /* EW.C:2622-2631 / 2748-2757 on Up, using the uploaded bforce arrays */
int sw4b200_grid_corrector( sw4b200_grid* g, const double* h_ftt );    /* EW.C:2644-2717           */
int sw4b200_grid_cycle( sw4b200_grid* g );                             /* EW.C:3060-3082           */
int sw4b200_grid_record( sw4b200_grid* g, double* h_out /*3*nrec*/ );  /* receivers from Up, EW.C:2802-2835 */
/* all of the above for a single block with no neighbours: one full time step */
int sw4b200_grid_step( sw4b200_grid* g, const double* h_f, const double* h_ftt, double* h_rec );
/* phase-split variants for z-slab runs, where the caller exchanges halo planes while the bulk of
 * the block is still being computed (the boundary/centre split of RHSPredCU_boundary/_center,
 * EW_cuda.C:1228-1410, applied to z faces).  part 0 = whole block, 1 = only the two interior
 * planes next to every halo face (what a neighbour slab needs) plus the ghost shell, 2 = the rest
 * (+ source injection, + supergrid damping for the corrector). */
int sw4b200_grid_predictor_part( sw4b200_grid* g, int part, const double* h_f );
int sw4b200_grid_corrector_part( sw4b200_grid* g, int part, const double* h_ftt );
/* device-resident stepping: the source amplitudes of steps [0,nsteps) (nsteps x 3*nsrc values, F and
 * F_tt) are uploaded once; sw4b200_grid_run advances nsteps time steps without any host
 * synchronisation and keeps the receiver samples on the device until they are fetched
 * (replaces the per-step blocking copy of extractRecordDataCU, EW_cuda.C:2198). */
int sw4b200_grid_set_source_series( sw4b200_grid* g, int nsteps, const double* h_f, const double* h_ftt );
int sw4b200_grid_run( sw4b200_grid* g, int first_step, int nsteps );
int sw4b200_grid_fetch_records( sw4b200_grid* g, int first_step, int nsteps, double* h_out /* nsteps*3*nrec */ );
/* the same device-resident record for drivers that sequence the phases themselves (z-slab runs): samples the receivers from
 * the new solution of step `step` (call after the second sw4b200_grid_enforce_bc, before sw4b200_grid_cycle) */
int sw4b200_grid_record_resident( sw4b200_grid* g, int step );
/* layered media: fill scalar field `name` ("mu","lambda","rho") with h_kvalues[k-kfirst] on plane k
 * (what MaterialBlock produces for depth-only blocks, MaterialBlock.C) without a host copy of the field */
int sw4b200_grid_fill_profile( sw4b200_grid* g, const char* name, const double* h_kvalues /* nk */ );
/* run this block's kernels on library stream st (0..3) */
int sw4b200_grid_set_stream( sw4b200_grid* g, int st );
/* z-slab halo planes: copy the 2 interior planes of Up next to the low (side=0) / high (side=1) face
 * into d_dst (3*2*ni*nj doubles, component-major), or from d_src into the 2 halo planes.  with_acc=1
 * (the exchange after the predictor) appends the same planes of the stored acceleration, so that a
 * slab run stays bit-identical to the undivided run; sw4b200_grid_halo_doubles gives the buffer size. */
int sw4b200_grid_pack_halo( sw4b200_grid* g, int side, int with_acc, double* d_dst, void* stream );
int sw4b200_grid_unpack_halo( sw4b200_grid* g, int side, int with_acc, const double* d_src, void* stream );
int sw4b200_grid_halo_doubles( sw4b200_grid* g, int with_acc );
int sw4b200_grid_sync( sw4b200_grid* g );
/* predictor / corrector phases with the source amplitudes already on the device (3*nsrc doubles, order of
 * sw4b200_grid_set_source_points; what ForceCU + forcing_dev produce, EW_cuda.C:709, device-routines.C:8306), or NULL */
int sw4b200_grid_predictor_dev( sw4b200_grid* g, int part, const double* d_f );
int sw4b200_grid_corrector_dev( sw4b200_grid* g, int part, const double* d_ftt );

/* ---------------------------------------------------------------- halo exchange between z-slabs (one process per GPU)
 * replaces EW::communicate_array (EW.C:3247-3317) / communicate_arrayCU_X/_Y + pack/unpack_HaloArrayCU_* (EW_cuda.C:1515-1997)
 * and setup_device_communication_array (:2002): the two face planes of the new solution go straight from the field arrays
 * into the neighbour's halo planes with grouped ncclSend/ncclRecv on the library's communication stream (no pack kernels,
 * no staging buffers, no host).  Rendezvous: rank 0 obtains 128 bytes with sw4b200_comm_unique_id and hands them to the
 * other ranks by any means (a file, torch.distributed, MPI_Bcast); every rank then calls sw4b200_comm_init. */
int sw4b200_comm_unique_id( void* out128 );
int sw4b200_comm_init( int rank, int nranks, const void* id128 );
int sw4b200_comm_finalize( void );
/* host values reduced over the ranks, in place (what the reference does with MPI_Allreduce for dt and the error norms,
 * EW.C:5134, 4606-4608): op 0 = max, 1 = sum, 2 = min.  Synchronises the device. */
int sw4b200_comm_allreduce( double* h_values, int n, int op );
/* device timer (CUDA events on library stream 0) */
int sw4b200_timer_start( void );
int sw4b200_timer_stop_ms( double* ms );
int sw4b200_grid_set_neighbours( sw4b200_grid* g, int rank_lo, int rank_hi );   /* -1: no neighbour on that face */
/* transport chosen by sw4b200_grid_set_neighbours: 1 = the copy engines push the planes into the neighbour's arrays, mapped
 * through CUDA IPC (no SM involved; needs peer access between the GPUs), 0 = NCCL send / receive, -1 = no neighbours.
 * sw4b200_set_option( "exchange_p2p", 0 ) before set_neighbours forces NCCL. */
int sw4b200_grid_exchange_transport( sw4b200_grid* g );
int sw4b200_grid_exchange_begin( sw4b200_grid* g, int with_acc );              /* after the face rows (part 1)   */
int sw4b200_grid_exchange_end( sw4b200_grid* g );                              /* before the boundary conditions  */

/* ---------------------------------------------------------------- measurement aid
 * fp64 FMA throughput of the device (a register-only chain kernel): the co-bound of the stencil kernels next to the HBM
 * bandwidth ("to be measured"); TFLOP/s at 2 flop per FMA and FMA lane-operations per second. */
int sw4b200_measure_fp64_peak( double* tflops, double* fma_per_s );

#ifdef __cplusplus
}
#endif
#endif
