// vofx: launchers of band kernels (see vofx_launch.h); generated layout, one group of kernels per translation unit.
#include "vofx_launch.h"
#include "vofx_kernels.cuh"

namespace vofx
{
  namespace launch
  {
    void youngs3 ( int blocks, cudaStream_t s, Grid g, const double *u, const int *mixedList, const StepState *state, int capacity, double *hs )
    { k_youngs< 3 ><<< blocks, 128, 0, s >>>( g, u, mixedList, state, capacity, hs ); }
    void youngs_coop3 ( int lanes, int blocks, cudaStream_t s, Grid g, const double *u, const int *mixedList, StepState *state, int capacity, double *hs,
                        const void *sweepTable, int *serialCells, void *rootTasks, int heightFunction, const int *order, unsigned char *costClass, BandPart bp )
    {
      const coop::SweepTable *T = (const coop::SweepTable *) sweepTable;
      coop::RootTask *R = (coop::RootTask *) rootTasks;
      if( lanes == 4 && !heightFunction )
        pdl( k_youngs_coop3< 4, false >, blocks * 2, 64, s, g, u, mixedList, state, capacity, hs, T, serialCells, R, order, costClass, bp );
      else if( lanes == 4 )
        pdl( k_youngs_coop3< 4, true >, blocks * 2, 64, s, g, u, mixedList, state, capacity, hs, T, serialCells, R, order, costClass, bp );
      else if( !heightFunction )
        pdl( k_youngs_coop3< 8, false >, blocks, 128, s, g, u, mixedList, state, capacity, hs, T, serialCells, R, order, costClass, bp );
      else
        pdl( k_youngs_coop3< 8, true >, blocks, 128, s, g, u, mixedList, state, capacity, hs, T, serialCells, R, order, costClass, bp );
    }
    void locate_root3 ( int blocks, cudaStream_t s, StepState *state, int capacity, const void *rootTasks, double *hs, BandPart bp )
    { pdl( k_locate_root3< 3 >, blocks, 128, s, state, capacity, (const coop::RootTask *) rootTasks, hs, bp ); }
    size_t root_task_bytes () { return sizeof( coop::RootTask ); }
    void locate_serial3 ( cudaStream_t s, Grid g, const double *u, const StepState *state, double *hs, const int *serialCells, int capacity, BandPart bp )
    { pdl( k_locate_serial3< 3 >, 148, 64, s, g, u, state, hs, serialCells, capacity, bp ); }
    int make_sweep_table ( void *host, size_t bytes )
    {
      if( bytes < sizeof( coop::SweepTable ) )
        return -1;
      return coop::make_sweep_table( *(coop::SweepTable *) host );
    }
    size_t sweep_table_bytes () { return sizeof( coop::SweepTable ); }
    void hf3 ( int blocks, cudaStream_t s, Grid g, const double *u, const int *mixedList, const StepState *state, int capacity, double *hs )
    { k_hf< 3 ><<< blocks, 128, 0, s >>>( g, u, mixedList, state, capacity, hs ); }
    // forces the (lazily loaded) kernels of this translation unit into the context: a first launch loads its kernel, which
    // may wait for the device -- fatal while a kernel of another context spins on a signal (vofx_comm_connect*)
    template< class K >
    static void preload_one ( K kernel )
    {
      cudaFuncAttributes a;
      cudaFuncGetAttributes( &a, kernel );
    }
    void preload_recon3d ()
    {
      preload_one( k_youngs_coop3< 4, false > );
      preload_one( k_youngs_coop3< 4, true > );
      preload_one( k_youngs_coop3< 8, false > );
      preload_one( k_youngs_coop3< 8, true > );
      preload_one( k_locate_root3< 3 > );
      preload_one( k_locate_serial3< 3 > );
    }
  } // namespace launch
} // namespace vofx

#if defined( VOFX_COUNT_WALKS )
extern "C" int vofx_debug_counters ( unsigned long long *out, int reset )
{
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol( out, vofx::coop::g_dbg, 8 * sizeof( unsigned long long ) );
  if( reset )
  {
    unsigned long long zero[ 8 ] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    cudaMemcpyToSymbol( vofx::coop::g_dbg, zero, sizeof( zero ) );
  }
  return 0;
}
#endif
