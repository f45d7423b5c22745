// vofx: host-callable launchers of the kernels, one translation unit per group so that nvcc compiles them in parallel
// (the geometry code is heavily inlined).  The C-ABI implementation (vofx_capi.cu) only sees these prototypes.
#pragma once

#include <cuda_runtime.h>

#include "vofx_types.cuh"

namespace vofx
{
  namespace launch
  {
#if defined( __CUDACC__ )
    // kernel<<< grid, block, 0, s >>>( args... ) with the programmatic-serialisation attribute (see pdl_prologue, vofx_math.cuh):
    // for the kernels of the loop pass, every one of which starts with pdl_prologue().  Only with VOFX_PDL=1 (measured slower).
    bool pdl_enabled ();
    template< class... P, class... A >
    inline void pdl ( void ( *kernel )( P... ), dim3 grid, dim3 block, cudaStream_t s, A &&... args )
    {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = grid;
      cfg.blockDim = block;
      cfg.dynamicSmemBytes = 0;
      cfg.stream = s;
      cudaLaunchAttribute attr[ 1 ];
      attr[ 0 ].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[ 0 ].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = pdl_enabled() ? 1 : 0;
      cudaLaunchKernelEx( &cfg, kernel, static_cast< P >( args )... );
    }
#endif

    // every translation unit: load its kernels now (see preload_one in k_dense.cu)
    void preload_dense ();
    void preload_band2d ();
    void preload_recon3d ();
    void preload_flux3d ();
    void preload_swartz3d ();
    // k_diag.cu
    size_t diag_scratch_bytes ();
    double *diagnostics ( cudaStream_t s, long long n, const double *u, const double *exact, double volume, void *scratch );
    void circle_init ( int blocks, cudaStream_t s, Grid g, double cx, double cy, double radius, double *u );
    double *l1error_circle ( cudaStream_t s, Grid g, const unsigned char *flags, const double *hs, double cx, double cy, double radius, void *scratch );
    // k_dense.cu
    void flag_rows ( int dim, int blocks, cudaStream_t s, Grid g, const double *u, double eps, unsigned char *flags, int *mixedList, int *slot,
                     int capacity, StepState *state, int compact, int rowsPerBlock, int rowBegin, int rowEnd );
    void flag_compact ( int dim, int blocks, cudaStream_t s, Grid g, const double *u, double eps, unsigned char *flags, int *mixedList, int *slot,
                        int capacity, StepState *state, int compact );
    void compact_flags ( int dim, int blocks, cudaStream_t s, Grid g, const unsigned char *flags, int *mixedList, int *slot, int capacity, StepState *state );
    void update ( int dim, int blocks, cudaStream_t s, Grid g, const unsigned char *flags, const int *mixedList, const int *slot, const StepState *state,
                  int capacity, const FluxRecord *records, double *target, int accumulate, StepState *stateRW, int *changedIdx, double *changedVal,
                  int changedCapacity, unsigned char *marks, int segsPerRow, const CommPeers *peers );
    // distributed step over peer-mapped memory
    void comm_wait_halo ( cudaStream_t s, const CommPeers &P, StepState *state );
    void comm_flux_done ( cudaStream_t s, const CommPeers &P, StepState *state );
    void finalize_dist ( cudaStream_t s, const CommPeers &P, StepState *state, double cfl );
    void halo_push ( int blocks, cudaStream_t s, const CommPeers &P, const double *u, long long layerCells, int own0, int own1 );
    // incremental flagging: k_flag_segments on the segment list the previous pass's side branch (k_mark_band + k_scan_marks) left
    void reflag ( int dim, int blocks, cudaStream_t s, Grid g, const double *u, double eps, unsigned char *flags, int *mixedList, int *slot, int capacity,
                  StepState *state, const int *segList, int segsPerRow );
    void mark_scan ( int dim, int blocks, cudaStream_t s, Grid g, const int *mixedList, int capacity, StepState *state, unsigned char *marks, int *segList,
                     int segsPerRow, long long nSegments, int layer0, int layer1 );
    void curvature_local ( int dim, int blocks, cudaStream_t s, Grid g, const double *u, const double *hs, const int *mixedList, const StepState *state,
                           int capacity, double *curv1, unsigned char *ok1 );
    void curvature_average ( int dim, int blocks, cudaStream_t s, Grid g, const unsigned char *flags, const int *mixedList, const int *slot,
                             const StepState *state, int capacity, const double *curv1, const unsigned char *ok1, double *kappa, unsigned char *valid );
    void finalize ( cudaStream_t s, StepState *state, double cfl );
    void reset_counters ( cudaStream_t s, StepState *state );
    void axpy ( int blocks, cudaStream_t s, long long n, double *u, double a, const double *x );
    void scatter_values ( int blocks, cudaStream_t s, long long n, const int *idx, const double *val, double *u );
    void init ( int dim, int blocks, cudaStream_t s, Grid g, Indicator I, double *u );
    void test_face_velocity ( cudaStream_t s, Grid g, const Velocity *velocity, double time, long long cell, int face, double *out );

    // k_interface.cu
    int scan_blocks ( long long n );
    void scan ( int dim, cudaStream_t s, Grid g, long long n, const unsigned char *flags, const int *counts, int *blockSums, long long *total, int *out, int phase );
    void interface_cut ( int dim, int blocks, cudaStream_t s, Grid g, const double *hs, const int *cells, long long nMixed, int *nverts, double *verts );
    void interface_gather ( int dim, int blocks, cudaStream_t s, long long nMixed, const int *nverts, const int *offsets, const double *verts, double *csr );

    // k_band2d.cu
    void youngs2 ( int blocks, cudaStream_t s, Grid g, const double *u, const int *mixedList, const StepState *state, int capacity, double *hs );
    void hf2 ( int blocks, cudaStream_t s, Grid g, const double *u, const int *mixedList, const StepState *state, int capacity, double *hs );
    void hf_fused2 ( int blocks, cudaStream_t s, Grid g, const double *u, const int *mixedList, const StepState *state, int capacity, double *hs );
    void clear_listed ( int blocks, cudaStream_t s, const int *list, const int *count, double *target );
    void copy_list ( int blocks, cudaStream_t s, const int *list, const StepState *state, int capacity, int *out, int *outCount );
    void swartz_curvature2 ( int blocks, cudaStream_t s, Grid g, const double *hs, const unsigned char *flags, const int *mixedList, const int *slot,
                             const StepState *state, int capacity, double *curv1, double *kappa );
    void swartz2 ( int blocks, cudaStream_t s, Grid g, const double *u, const unsigned char *flags, const int *mixedList, const StepState *state,
                   int capacity, double *hs, int maxIterations );
    void flux2 ( int blocks, cudaStream_t s, Grid g, const Velocity *velocity, const double *hs, const unsigned char *flags, const int *mixedList,
                 StepState *state, int capacity, FluxRecord *records, const double *timeOverride, const double *dtOverride );
    void test_locate2 ( int blocks, long long count, const double *lo, const double *h, const double *normal, const double *fraction, double *out );
    void test_flux2 ( int blocks, long long count, const double *lo, const double *h, const int *face, const double *v, const double *hs, double *out );
    void test_interface2 ( int blocks, long long count, const double *lo, const double *h, const double *hs, int *nverts, double *centroid, double *measure );

    // k_recon3d.cu, k_swartz3d.cu, k_flux3d.cu
    void youngs3 ( int blocks, cudaStream_t s, Grid g, const double *u, const int *mixedList, const StepState *state, int capacity, double *hs );
    void cost_order ( int blocks, cudaStream_t s, const int *mixedList, StepState *state, int capacity, const unsigned char *costClass, int *order, int parts = 1 );
    void youngs_coop3 ( int lanes, int blocks, cudaStream_t s, Grid g, const double *u, const int *mixedList, StepState *state, int capacity, double *hs,
                        const void *sweepTable, int *serialCells, void *rootTasks, int heightFunction, const int *order, unsigned char *costClass, BandPart bp = BandPart() );
    void locate_root3 ( int blocks, cudaStream_t s, StepState *state, int capacity, const void *rootTasks, double *hs, BandPart bp = BandPart() );
    size_t root_task_bytes ();
    void locate_serial3 ( cudaStream_t s, Grid g, const double *u, const StepState *state, double *hs, const int *serialCells, int capacity, BandPart bp = BandPart() );
    int make_sweep_table ( void *host, size_t bytes );   // returns the number of cases (96) or -1
    size_t sweep_table_bytes ();
    void hf3 ( int blocks, cudaStream_t s, Grid g, const double *u, const int *mixedList, const StepState *state, int capacity, double *hs );
    void swartz3 ( int blocks, cudaStream_t s, Grid g, const double *u, const unsigned char *flags, const int *mixedList, const StepState *state,
                   int capacity, double *hs, int maxIterations );
    void swartz_coop3 ( int lanes, int blocks, cudaStream_t s, Grid g, const double *u, const unsigned char *flags, const int *mixedList, const StepState *state,
                        int capacity, double *hs, int maxIterations, const void *sweepTable );
    size_t swz_cell_bytes ();
    size_t swz_task_bytes ();
    void swartz_batched3 ( int lanes, int sms, cudaStream_t s, Grid g, const double *u, const unsigned char *flags, const int *mixedList, StepState *state,
                           int capacity, double *hs, int maxIterations, const void *sweepTable, void *cells, void *tasks, int *counter,
                           int cellCapacity, int taskCapacity, int *order = nullptr, int *sort = nullptr, int *cellList = nullptr, int *listCount = nullptr );
    void flux3 ( int blocks, cudaStream_t s, Grid g, const Velocity *velocity, const double *hs, const unsigned char *flags, const int *mixedList,
                 StepState *state, int capacity, FluxRecord *records, const double *timeOverride, const double *dtOverride );
    void flux_cell3 ( int lanes, int blocks, cudaStream_t s, Grid g, const Velocity *velocity, const Velocity &velocityByValue, const double *hs, const unsigned char *flags, const int *mixedList,
                      StepState *state, int capacity, FluxRecord *records, int *tasks, int taskCapacity, const void *clipTable,
                      const double *timeOverride, const double *dtOverride, BandPart bp = BandPart(), int withSerialClips = 1 );
    void flux_clip_serial3 ( cudaStream_t s, Grid g, const Velocity *velocity, const double *hs, const unsigned char *flags, const int *mixedList,
                             StepState *state, FluxRecord *records, int *tasks, int taskCapacity, const double *timeOverride, const double *dtOverride );
    size_t clip_table_bytes ();
    void make_clip_table ( void *host );   // clip_table_bytes(): 256 + 6561 entries of 64 bytes (coop::ClipCase)
    void test_locate3 ( int blocks, long long count, const double *lo, const double *h, const double *normal, const double *fraction, double *out );
    void test_flux3 ( int blocks, long long count, const double *lo, const double *h, const int *face, const double *v, const double *hs, double *out );
    float fp64_rate ( int mode, int blocks, int iters, int reps, double *scratch );
    void test_arith ( long long n, unsigned long long seed, unsigned long long *mismatch );
    void test_interface3 ( int blocks, long long count, const double *lo, const double *h, const double *hs, int *nverts, double *centroid, double *measure );
  } // namespace launch
} // namespace vofx
