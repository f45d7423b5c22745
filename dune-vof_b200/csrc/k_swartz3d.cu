// vofx: launchers of band kernels (see vofx_launch.h); generated layout, one group of kernels per translation unit.
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
  } // namespace launch
} // namespace vofx
