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
  } // namespace launch
} // namespace vofx
