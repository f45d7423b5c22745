// vofx: launchers of band kernels (see vofx_launch.h); generated layout, one group of kernels per translation unit.
#include "vofx_launch.h"
#include "vofx_kernels.cuh"

namespace vofx
{
  namespace launch
  {
    void youngs3 ( int blocks, cudaStream_t s, Grid g, const double *u, const int *mixedList, const StepState *state, int capacity, double *hs )
    { k_youngs< 3 ><<< blocks, 128, 0, s >>>( g, u, mixedList, state, capacity, hs ); }
    void hf3 ( int blocks, cudaStream_t s, Grid g, const double *u, const int *mixedList, const StepState *state, int capacity, double *hs )
    { k_hf< 3 ><<< blocks, 128, 0, s >>>( g, u, mixedList, state, capacity, hs ); }
  } // namespace launch
} // namespace vofx
