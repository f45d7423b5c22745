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
  } // namespace launch
} // namespace vofx
