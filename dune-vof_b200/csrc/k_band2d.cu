// vofx: launchers of band kernels (see vofx_launch.h); generated layout, one group of kernels per translation unit.
#include "vofx_launch.h"
#include "vofx_kernels.cuh"

namespace vofx
{
  namespace launch
  {
    void youngs2 ( int blocks, cudaStream_t s, Grid g, const double *u, const int *mixedList, const StepState *state, int capacity, double *hs )
    { k_youngs< 2 ><<< blocks, 128, 0, s >>>( g, u, mixedList, state, capacity, hs ); }
    void hf2 ( int blocks, cudaStream_t s, Grid g, const double *u, const int *mixedList, const StepState *state, int capacity, double *hs )
    { k_hf< 2 ><<< blocks, 128, 0, s >>>( g, u, mixedList, state, capacity, hs ); }
    void hf_fused2 ( int blocks, cudaStream_t s, Grid g, const double *u, const int *mixedList, const StepState *state, int capacity, double *hs )
    { k_hf_fused< 2 ><<< blocks, 128, 0, s >>>( g, u, mixedList, state, capacity, hs ); }
    void clear_listed ( int blocks, cudaStream_t s, const int *list, const int *count, double *target )
    { k_clear_listed<<< blocks, 256, 0, s >>>( list, count, target ); }
    void copy_list ( int blocks, cudaStream_t s, const int *list, const StepState *state, int capacity, int *out, int *outCount )
    { k_copy_list<<< blocks, 256, 0, s >>>( list, state, capacity, out, outCount ); }
    void swartz_curvature2 ( int blocks, cudaStream_t s, Grid g, const double *hs, const unsigned char *flags, const int *mixedList, const int *slot,
                             const StepState *state, int capacity, double *curv1, double *kappa )
    {
      k_swartz_curv_local<<< blocks, 128, 0, s >>>( g, hs, flags, mixedList, state, capacity, curv1 );
      k_swartz_curv_average<<< blocks, 128, 0, s >>>( g, flags, mixedList, slot, state, capacity, curv1, kappa );
    }
    void swartz2 ( int blocks, cudaStream_t s, Grid g, const double *u, const unsigned char *flags, const int *mixedList, const StepState *state,
                   int capacity, double *hs, int maxIterations )
    { k_swartz< 2 ><<< blocks, 64, 0, s >>>( g, u, flags, mixedList, state, capacity, hs, maxIterations ); }
    void flux2 ( int blocks, cudaStream_t s, Grid g, const Velocity *velocity, const double *hs, const unsigned char *flags, const int *mixedList,
                 StepState *state, int capacity, FluxRecord *records, const double *timeOverride, const double *dtOverride )
    { k_flux< 2 ><<< blocks, 128, 0, s >>>( g, velocity, hs, flags, mixedList, state, capacity, records, timeOverride, dtOverride ); }
    void test_locate2 ( int blocks, long long count, const double *lo, const double *h, const double *normal, const double *fraction, double *out )
    { k_test_locate< 2 ><<< blocks, 64 >>>( count, lo, h, normal, fraction, out ); }
    void test_flux2 ( int blocks, long long count, const double *lo, const double *h, const int *face, const double *v, const double *hs, double *out )
    { k_test_flux< 2 ><<< blocks, 64 >>>( count, lo, h, face, v, hs, out ); }
    void test_interface2 ( int blocks, long long count, const double *lo, const double *h, const double *hs, int *nverts, double *centroid, double *measure )
    { k_test_interface< 2 ><<< blocks, 64 >>>( count, lo, h, hs, nverts, centroid, measure ); }
    // forces the (lazily loaded) kernels of this translation unit into the context: a first launch loads its kernel, which
    // may wait for the device -- fatal while a kernel of another context spins on a signal (vofx_comm_connect*)
    template< class K >
    static void preload_one ( K kernel )
    {
      cudaFuncAttributes a;
      cudaFuncGetAttributes( &a, kernel );
    }
    void preload_band2d ()
    {
      preload_one( k_youngs< 2 > );
      preload_one( k_hf< 2 > );
      preload_one( k_hf_fused< 2 > );
      preload_one( k_clear_listed );
      preload_one( k_copy_list );
      preload_one( k_swartz_curv_local );
      preload_one( k_swartz_curv_average );
      preload_one( k_swartz< 2 > );
      preload_one( k_flux< 2 > );
    }
  } // namespace launch
} // namespace vofx
