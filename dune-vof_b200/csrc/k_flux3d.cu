// vofx: launchers of band kernels (see vofx_launch.h); generated layout, one group of kernels per translation unit.
#include "vofx_launch.h"
#include "vofx_kernels.cuh"

namespace vofx
{
  namespace launch
  {
    void flux3 ( int blocks, cudaStream_t s, Grid g, const Velocity *velocity, const double *hs, const unsigned char *flags, const int *mixedList,
                 StepState *state, int capacity, FluxRecord *records, const double *timeOverride, const double *dtOverride )
    { k_flux< 3 ><<< blocks, 128, 0, s >>>( g, velocity, hs, flags, mixedList, state, capacity, records, timeOverride, dtOverride ); }
    void flux_cell3 ( int lanes, int blocks, cudaStream_t s, Grid g, const Velocity *velocity, const Velocity &velocityByValue, const double *hs, const unsigned char *flags, const int *mixedList,
                      StepState *state, int capacity, FluxRecord *records, int *tasks, int taskCapacity, const void *clipTable,
                      const double *timeOverride, const double *dtOverride, BandPart bp, int withSerialClips )
    {
      if( lanes == 4 )
        pdl( k_flux_cell3< 3, 4 >, blocks * 2, 64, s, g, velocityByValue, hs, flags, mixedList, state, capacity, records, tasks, taskCapacity, (const coop::ClipCase *) clipTable, timeOverride, dtOverride, bp );
      else
        pdl( k_flux_cell3< 3, 8 >, blocks, 128, s, g, velocityByValue, hs, flags, mixedList, state, capacity, records, tasks, taskCapacity, (const coop::ClipCase *) clipTable, timeOverride, dtOverride, bp );
      if( withSerialClips )
        flux_clip_serial3( s, g, velocity, hs, flags, mixedList, state, records, tasks, taskCapacity, timeOverride, dtOverride );
    }
    void flux_clip_serial3 ( cudaStream_t s, Grid g, const Velocity *velocity, const double *hs, const unsigned char *flags, const int *mixedList,
                             StepState *state, FluxRecord *records, int *tasks, int taskCapacity, const double *timeOverride, const double *dtOverride )
    { pdl( k_flux_clip_serial3< 3 >, 148, 64, s, g, velocity, hs, flags, mixedList, state, records, tasks, taskCapacity, timeOverride, dtOverride ); }
    size_t clip_table_bytes () { return sizeof( coop::ClipCase ) * ( 256 + coop::kClipBoundaryCases ); }
    void make_clip_table ( void *host )
    {
      coop::make_clip_table( (coop::ClipCase *) host );
      coop::make_clip_boundary_table( (coop::ClipCase *) host + 256 );
    }
    // forces the (lazily loaded) kernels of this translation unit into the context: a first launch loads its kernel, which
    // may wait for the device -- fatal while a kernel of another context spins on a signal (vofx_comm_connect*)
    template< class K >
    static void preload_one ( K kernel )
    {
      cudaFuncAttributes a;
      cudaFuncGetAttributes( &a, kernel );
    }
    void preload_flux3d ()
    {
      preload_one( k_flux_cell3< 3, 4 > );
      preload_one( k_flux_cell3< 3, 8 > );
      preload_one( k_flux_clip_serial3< 3 > );
    }
  } // namespace launch
} // namespace vofx
