// vofx: launchers of the dense / bookkeeping kernels (flags + compaction, gather update, curvature, init).
#include "vofx_launch.h"
#include "vofx_kernels.cuh"

#include <cstdlib>

namespace vofx
{
  namespace launch
  {
    bool pdl_enabled ()
    {
      // measured on B200 (512^3, graph replay): 0.592 ms per pass with programmatic dependent launches, 0.565 ms without --
      // the early-launched successor's blocks take SM slots from the tail of the running kernel.  Off unless VOFX_PDL=1.
      static const bool on = std::getenv( "VOFX_PDL" ) != nullptr;
      return on;
    }

    void flag_rows ( int dim, int blocks, cudaStream_t s, Grid g, const double *u, double eps, unsigned char *flags, int *mixedList, int *slot,
                     int capacity, StepState *state, int compact, int rowsPerBlock, int rowBegin, int rowEnd )
    {
      if( dim == 2 )
        pdl( k_flag_rows< 2 >, blocks, 128, s, g, u, eps, flags, mixedList, slot, capacity, state, compact, rowsPerBlock, rowBegin, rowEnd );
      else
        pdl( k_flag_rows< 3 >, blocks, 128, s, g, u, eps, flags, mixedList, slot, capacity, state, compact, rowsPerBlock, rowBegin, rowEnd );
    }
    void flag_compact ( int dim, int blocks, cudaStream_t s, Grid g, const double *u, double eps, unsigned char *flags, int *mixedList, int *slot,
                        int capacity, StepState *state, int compact )
    {
      if( dim == 2 )
        k_flag_compact< 2 ><<< blocks, 256, 0, s >>>( g, u, eps, flags, mixedList, slot, capacity, state, compact );
      else
        k_flag_compact< 3 ><<< blocks, 256, 0, s >>>( g, u, eps, flags, mixedList, slot, capacity, state, compact );
    }
    void compact_flags ( int dim, int blocks, cudaStream_t s, Grid g, const unsigned char *flags, int *mixedList, int *slot, int capacity, StepState *state )
    {
      if( dim == 2 )
        k_compact_flags< 2 ><<< blocks, 256, 0, s >>>( g, flags, mixedList, slot, capacity, state );
      else
        k_compact_flags< 3 ><<< blocks, 256, 0, s >>>( g, flags, mixedList, slot, capacity, state );
    }
    void update ( int dim, int blocks, cudaStream_t s, Grid g, const unsigned char *flags, const int *mixedList, const int *slot, const StepState *state,
                  int capacity, const FluxRecord *records, double *target, int accumulate, StepState *stateRW, int *changedIdx, double *changedVal,
                  int changedCapacity, unsigned char *marks, int segsPerRow, const CommPeers *peers )
    {
      CommPeers P;
      if( peers )
        P = *peers;
      else
      {
        P.rank = 0;
        P.world = 1;
        P.uLo = P.uHi = nullptr;
        P.offLo = P.offHi = 0;
        P.pushLo1 = P.pushHi0 = 0;
      }
      if( dim == 2 )
        pdl( k_update< 2 >, blocks, 128, s, g, flags, mixedList, slot, state, capacity, records, target, accumulate, stateRW, changedIdx, changedVal, changedCapacity, marks, segsPerRow, P );
      else
        pdl( k_update< 3 >, blocks, 128, s, g, flags, mixedList, slot, state, capacity, records, target, accumulate, stateRW, changedIdx, changedVal, changedCapacity, marks, segsPerRow, P );
    }
    void comm_wait_halo ( cudaStream_t s, const CommPeers &P, StepState *state ) { k_comm_wait_halo<<< 1, 32, 0, s >>>( P, state ); }
    void comm_flux_done ( cudaStream_t s, const CommPeers &P, StepState *state ) { k_comm_flux_done<<< 1, 32, 0, s >>>( P, state ); }
    void finalize_dist ( cudaStream_t s, const CommPeers &P, StepState *state, double cfl ) { k_finalize_dist<<< 1, 32, 0, s >>>( P, state, cfl ); }
    void halo_push ( int blocks, cudaStream_t s, const CommPeers &P, const double *u, long long layerCells, int own0, int own1 )
    { k_halo_push<<< blocks, 256, 0, s >>>( P, u, layerCells, own0, own1 ); }
    // incremental flagging: the marked segments were compacted by the previous pass's side branch (mark_scan)
    void reflag ( int dim, int blocks, cudaStream_t s, Grid g, const double *u, double eps, unsigned char *flags, int *mixedList, int *slot, int capacity,
                  StepState *state, const int *segList, int segsPerRow )
    {
      if( dim == 2 )
        pdl( k_flag_segments< 2 >, blocks, 256, s, g, u, eps, flags, mixedList, slot, capacity, state, segList, segsPerRow );
      else
        pdl( k_flag_segments< 3 >, blocks, 256, s, g, u, eps, flags, mixedList, slot, capacity, state, segList, segsPerRow );
    }
    // side branch of a pass: marks from the band in flight, then their compaction into the segment list of the next pass
    void mark_scan ( int dim, int blocks, cudaStream_t s, Grid g, const int *mixedList, int capacity, StepState *state, unsigned char *marks, int *segList,
                     int segsPerRow, long long nSegments, int layer0, int layer1 )
    {
      const long long chunks = ( nSegments + 15 ) / 16;
      long long scanBlocks = ( chunks + 255 ) / 256;
      if( scanBlocks > blocks )
        scanBlocks = blocks;
      if( dim == 2 )
      {
        k_mark_band< 2 ><<< blocks, 256, 0, s >>>( g, mixedList, state, capacity, marks, segsPerRow );
        k_scan_marks< 2 ><<< (int) scanBlocks, 256, 0, s >>>( g, marks, segsPerRow, nSegments, layer0, layer1, segList, state );
      }
      else
      {
        k_mark_band< 3 ><<< blocks, 256, 0, s >>>( g, mixedList, state, capacity, marks, segsPerRow );
        k_scan_marks< 3 ><<< (int) scanBlocks, 256, 0, s >>>( g, marks, segsPerRow, nSegments, layer0, layer1, segList, state );
      }
    }
    void curvature_local ( int dim, int blocks, cudaStream_t s, Grid g, const double *u, const double *hs, const int *mixedList, const StepState *state,
                           int capacity, double *curv1, unsigned char *ok1 )
    {
      if( dim == 2 )
        k_curvature_local< 2 ><<< blocks, 128, 0, s >>>( g, u, hs, mixedList, state, capacity, curv1, ok1 );
      else
        k_curvature_local< 3 ><<< blocks, 128, 0, s >>>( g, u, hs, mixedList, state, capacity, curv1, ok1 );
    }
    void curvature_average ( int dim, int blocks, cudaStream_t s, Grid g, const unsigned char *flags, const int *mixedList, const int *slot,
                             const StepState *state, int capacity, const double *curv1, const unsigned char *ok1, double *kappa, unsigned char *valid )
    {
      if( dim == 2 )
        k_curvature_average< 2 ><<< blocks, 128, 0, s >>>( g, flags, mixedList, slot, state, capacity, curv1, ok1, kappa, valid );
      else
        k_curvature_average< 3 ><<< blocks, 128, 0, s >>>( g, flags, mixedList, slot, state, capacity, curv1, ok1, kappa, valid );
    }
    void cost_order ( int blocks, cudaStream_t s, const int *mixedList, StepState *state, int capacity, const unsigned char *costClass, int *order, int parts )
    {
      pdl( k_cost_hist, blocks, 256, s, mixedList, state, capacity, costClass, parts );
      pdl( k_cost_order, blocks, 256, s, mixedList, state, capacity, costClass, order, parts );
    }
    void finalize ( cudaStream_t s, StepState *state, double cfl ) { pdl( k_finalize, 1, 1, s, state, cfl ); }
    void reset_counters ( cudaStream_t s, StepState *state ) { k_reset_counters<<< 1, 1, 0, s >>>( state ); }
    void axpy ( int blocks, cudaStream_t s, long long n, double *u, double a, const double *x ) { k_axpy<<< blocks, 256, 0, s >>>( n, u, a, x ); }
    void scatter_values ( int blocks, cudaStream_t s, long long n, const int *idx, const double *val, double *u ) { k_scatter_values<<< blocks, 256, 0, s >>>( n, idx, val, u ); }
    void init ( int dim, int blocks, cudaStream_t s, Grid g, Indicator I, double *u )
    {
      if( dim == 2 )
        k_init< 2 ><<< blocks, 128, 0, s >>>( g, I, u );
      else
        k_init< 3 ><<< blocks, 128, 0, s >>>( g, I, u );
    }
    void test_face_velocity ( cudaStream_t s, Grid g, const Velocity *velocity, double time, long long cell, int face, double *out )
    {
      k_test_face_velocity<<< 1, 1, 0, s >>>( g, velocity, time, cell, face, out );
    }
    // forces the (lazily loaded) kernels of this translation unit into the context: a first launch loads its kernel, which
    // may wait for the device -- fatal while a kernel of another context spins on a signal (vofx_comm_connect*)
    template< class K >
    static void preload_one ( K kernel )
    {
      cudaFuncAttributes a;
      cudaFuncGetAttributes( &a, kernel );
    }
    void preload_dense ()
    {
      preload_one( k_flag_rows< 2 > );
      preload_one( k_flag_rows< 3 > );
      preload_one( k_flag_compact< 2 > );
      preload_one( k_flag_compact< 3 > );
      preload_one( k_compact_flags< 2 > );
      preload_one( k_compact_flags< 3 > );
      preload_one( k_update< 2 > );
      preload_one( k_update< 3 > );
      preload_one( k_mark_band< 2 > );
      preload_one( k_mark_band< 3 > );
      preload_one( k_scan_marks< 2 > );
      preload_one( k_scan_marks< 3 > );
      preload_one( k_flag_segments< 2 > );
      preload_one( k_flag_segments< 3 > );
      preload_one( k_curvature_local< 2 > );
      preload_one( k_curvature_local< 3 > );
      preload_one( k_curvature_average< 2 > );
      preload_one( k_curvature_average< 3 > );
      preload_one( k_cost_hist );
      preload_one( k_cost_order );
      preload_one( k_finalize );
      preload_one( k_reset_counters );
      preload_one( k_finalize_dist );
      preload_one( k_comm_wait_halo );
      preload_one( k_comm_flux_done );
      preload_one( k_halo_push );
      preload_one( k_axpy );
      preload_one( k_scatter_values );
      preload_one( k_init< 2 > );
      preload_one( k_init< 3 > );
    }
  } // namespace launch
} // namespace vofx
