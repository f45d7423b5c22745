// vofx: launchers of band kernels (see vofx_launch.h); generated layout, one group of kernels per translation unit.
#include <cstdlib>

#include "vofx_launch.h"
#include "vofx_kernels.cuh"

namespace vofx
{
  namespace launch
  {
    void swartz3 ( int blocks, cudaStream_t s, Grid g, const double *u, const unsigned char *flags, const int *mixedList, const StepState *state,
                   int capacity, double *hs, int maxIterations )
    { k_swartz< 3 ><<< blocks, 64, 0, s >>>( g, u, flags, mixedList, state, capacity, hs, maxIterations ); }
    void swartz_coop3 ( int lanes, int blocks, cudaStream_t s, Grid g, const double *u, const unsigned char *flags, const int *mixedList, const StepState *state,
                        int capacity, double *hs, int maxIterations, const void *sweepTable )
    {
      if( lanes == 4 )
        k_swartz_coop3< 4 ><<< blocks, 32, 0, s >>>( g, u, flags, mixedList, state, capacity, hs, maxIterations, (const coop::SweepTable *) sweepTable );
      else
        k_swartz_coop3< 8 ><<< blocks, 64, 0, s >>>( g, u, flags, mixedList, state, capacity, hs, maxIterations, (const coop::SweepTable *) sweepTable );
    }
    size_t swz_cell_bytes () { return sizeof( SwzCell ); }
    size_t swz_task_bytes () { return sizeof( SwzTask ); }
    // all iterations of the batched Swartz reconstruction (4 kernels each) after the setup; `counter` is zeroed by the caller
    void swartz_batched3 ( int lanes, int sms, cudaStream_t s, Grid g, const double *u, const unsigned char *flags, const int *mixedList, StepState *state,
                           int capacity, double *hs, int maxIterations, const void *sweepTable, void *cells, void *tasks, int *counter,
                           int cellCapacity, int taskCapacity, int *order, int *sort, int *cellList, int *listCount )
    {
      const coop::SweepTable *T = (const coop::SweepTable *) sweepTable;
      SwzCell *C = (SwzCell *) cells;
      SwzTask *K = (SwzTask *) tasks;
      k_swz_setup< 3 ><<< sms * 8, 128, 0, s >>>( g, flags, mixedList, state, capacity, hs, C, K, counter, cellCapacity, taskCapacity, cellList, listCount );
      // grid-stride kernels over the (cell, neighbour) tasks: grids well beyond the resident blocks balance the waves
      // (256^3, ms per step: 16 blocks/SM 9.75, 40 8.99, 80 8.72)
      static const int locateBlocks = std::getenv( "VOFX_SWZ_BLOCKS_PER_SM" ) ? std::atoi( std::getenv( "VOFX_SWZ_BLOCKS_PER_SM" ) ) : 80;
      static const int otherBlocks = std::getenv( "VOFX_SWZ_OTHER_BLOCKS_PER_SM" ) ? std::atoi( std::getenv( "VOFX_SWZ_OTHER_BLOCKS_PER_SM" ) ) : 48;
      for( int it = 0; it < maxIterations; ++it )
      {
        // the active tasks, sorted by the bracket of their last solve: order[ 0 .. sort[ 32 ] )
        const int *nActive = counter;
        const int *ord = nullptr;
        if( order && sort )
        {
          cudaMemsetAsync( sort, 0, 64 * sizeof( int ), s );
          k_swz_hist<<< sms * 2, 256, 0, s >>>( C, K, counter, sort );
          k_swz_order<<< sms * 2, 256, 0, s >>>( C, K, counter, sort, order );
          nActive = sort + 32;
          ord = order;
        }
        if( lanes == 8 )
          k_swz_locate< 8 ><<< sms * 8, 128, 0, s >>>( g, u, C, K, nActive, T, ord );
        else
          k_swz_locate< 4 ><<< sms * locateBlocks, 64, 0, s >>>( g, u, C, K, nActive, T, ord );
        k_swz_centroid< 3 ><<< sms * otherBlocks, 64, 0, s >>>( g, u, C, K, nActive, ord );
        if( lanes == 8 )
          k_swz_normal< 8 ><<< sms * 8, 64, 0, s >>>( g, u, C, K, state, capacity, T, cellList, listCount, it, cellCapacity );
        else
          k_swz_normal< 4 ><<< sms * otherBlocks, 32, 0, s >>>( g, u, C, K, state, capacity, T, cellList, listCount, it, cellCapacity );
        k_swz_finish< 3 ><<< sms * 8, 64, 0, s >>>( g, u, C, state, capacity, hs, maxIterations, cellList, listCount, it, cellCapacity );
      }
    }
    // forces the (lazily loaded) kernels of this translation unit into the context: a first launch loads its kernel, which
    // may wait for the device -- fatal while a kernel of another context spins on a signal (vofx_comm_connect*)
    template< class K >
    static void preload_one ( K kernel )
    {
      cudaFuncAttributes a;
      cudaFuncGetAttributes( &a, kernel );
    }
    void preload_swartz3d ()
    {
      preload_one( k_swartz< 3 > );
      preload_one( k_swartz_coop3< 4 > );
      preload_one( k_swartz_coop3< 8 > );
      preload_one( k_swz_setup< 3 > );
      preload_one( k_swz_locate< 4 > );
      preload_one( k_swz_locate< 8 > );
      preload_one( k_swz_centroid< 3 > );
      preload_one( k_swz_normal< 4 > );
      preload_one( k_swz_normal< 8 > );
      preload_one( k_swz_finish< 3 > );
    }
  } // namespace launch
} // namespace vofx
