// vofx: on-device diagnostics of the colour function over the OWNED cells of a context:
// total mass, minimum, maximum and the cell-wise L1 error of test/errors.hh:138-151 ( sum |uh - uhExact| * volume ),
// so that a driver's conservation / EOC reporting does not need a full-field device-to-host copy per step.
// One streaming pass (8 B/cell, 16 B with an exact solution), HBM-bound.  The summation tree is FIXED (1184 blocks of 256
// threads, warp shuffles in a fixed pattern, one final block): results are reproducible run to run and independent of
// the SM schedule; they differ from the reference's sequential sum by round-off only (tests: 1e-12 relative).
#include "vofx_launch.h"
#include "vofx_geom2d.cuh"

namespace vofx
{

  namespace
  {
    constexpr int kDiagBlocks = 148 * 8, kDiagThreads = 256;

    struct Diag
    {
      double sum, lo, hi, l1;
    };

    __device__ __forceinline__ Diag combine ( const Diag &a, const Diag &b )
    {
      return { a.sum + b.sum, fmin( a.lo, b.lo ), fmax( a.hi, b.hi ), a.l1 + b.l1 };
    }

    __device__ __forceinline__ Diag block_reduce ( Diag v )
    {
      __shared__ Diag warpResult[ kDiagThreads / 32 ];
#pragma unroll
      for( int o = 16; o > 0; o >>= 1 )
      {
        Diag w;
        w.sum = __shfl_down_sync( 0xffffffffu, v.sum, o );
        w.lo = __shfl_down_sync( 0xffffffffu, v.lo, o );
        w.hi = __shfl_down_sync( 0xffffffffu, v.hi, o );
        w.l1 = __shfl_down_sync( 0xffffffffu, v.l1, o );
        v = combine( v, w );
      }
      if( ( threadIdx.x & 31 ) == 0 )
        warpResult[ threadIdx.x >> 5 ] = v;
      __syncthreads();
      if( threadIdx.x == 0 )
        for( int w = 1; w < kDiagThreads / 32; ++w )
          v = combine( v, warpResult[ w ] );
      return v;
    }

    __global__ void __launch_bounds__( kDiagThreads ) k_diag_partial ( long long n, const double *__restrict__ u, const double *__restrict__ exact,
                                                                        Diag *__restrict__ partial )
    {
      Diag v = { 0.0, DBL_MAX, -DBL_MAX, 0.0 };
      // pairs of cells (the owned range starts at a multiple of the row length: 16-byte aligned for even rows)
      const bool vec = ( ( reinterpret_cast< uintptr_t >( u ) | ( exact ? reinterpret_cast< uintptr_t >( exact ) : 0 ) ) & 15 ) == 0;
      const long long pairs = vec ? n / 2 : 0;
      for( long long p = blockIdx.x * (long long) blockDim.x + threadIdx.x; p < pairs; p += (long long) gridDim.x * blockDim.x )
      {
        const double2 a = __ldg( reinterpret_cast< const double2 * >( u ) + p );
        v.sum += a.x;
        v.sum += a.y;
        v.lo = fmin( v.lo, fmin( a.x, a.y ) );
        v.hi = fmax( v.hi, fmax( a.x, a.y ) );
        if( exact )
        {
          const double2 e = __ldg( reinterpret_cast< const double2 * >( exact ) + p );
          v.l1 += fabs( a.x - e.x );
          v.l1 += fabs( a.y - e.y );
        }
      }
      for( long long i = 2 * pairs + blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x )
      {
        const double a = u[ i ];
        v.sum += a;
        v.lo = fmin( v.lo, a );
        v.hi = fmax( v.hi, a );
        if( exact )
          v.l1 += fabs( a - exact[ i ] );
      }
      v = block_reduce( v );
      if( threadIdx.x == 0 )
        partial[ blockIdx.x ] = v;
    }

    // out: [0] sum of the cell values, [1] min, [2] max, [3] sum |u - exact| * volume, [4] mass = sum u * volume
    __global__ void __launch_bounds__( kDiagThreads ) k_diag_final ( const Diag *__restrict__ partial, double volume, double *__restrict__ out )
    {
      Diag v = { 0.0, DBL_MAX, -DBL_MAX, 0.0 };
      for( int b = threadIdx.x; b < kDiagBlocks; b += kDiagThreads )
        v = combine( v, partial[ b ] );
      v = block_reduce( v );
      if( threadIdx.x == 0 )
      {
        out[ 0 ] = v.sum;
        out[ 1 ] = v.lo;
        out[ 2 ] = v.hi;
        out[ 3 ] = v.l1 * volume;
        out[ 4 ] = v.sum * volume;
      }
    }
    // circleInterpolation( center, radius, uh ), interpolation/circleinterpolation.hh:140-153: exact area fraction of the disc per cell
    __global__ void __launch_bounds__( 128 ) k_circle_init ( Grid g, double cx, double cy, double radius, double *__restrict__ u )
    {
      for( long long c = blockIdx.x * (long long) blockDim.x + threadIdx.x; c < g.cells; c += (long long) gridDim.x * blockDim.x )
      {
        int ijk[ 3 ];
        cell_ijk( g, c, ijk );
        if( !in_domain( g, ijk ) )
        {
          u[ c ] = 0.0;
          continue;
        }
        Poly2 cell;
        make_cell( cell_lower( g, 0, ijk[ 0 ] ), cell_lower( g, 1, ijk[ 1 ] ), g.h[ 0 ], g.h[ 1 ], cell );
        u[ c ] = circle_polygon_volume( cell, cx, cy, radius ) / ( g.h[ 0 ] * g.h[ 1 ] );
      }
    }

    // l1error against the exact circle over the OWNED cells: per-cell terms (l1error_circle_cell), fixed summation tree
    __global__ void __launch_bounds__( kDiagThreads ) k_l1error_circle_partial ( Grid g, const unsigned char *__restrict__ flags, const double *__restrict__ hs,
                                                                                  double cx, double cy, double radius, Diag *__restrict__ partial )
    {
      Diag v = { 0.0, DBL_MAX, -DBL_MAX, 0.0 };
      const long long layer = g.cells / g.ns[ 1 ];
      const long long first = layer * g.own0, n = layer * ( g.own1 - g.own0 );
      for( long long t = blockIdx.x * (long long) blockDim.x + threadIdx.x; t < n; t += (long long) gridDim.x * blockDim.x )
      {
        const long long c = first + t;
        int ijk[ 3 ];
        cell_ijk( g, c, ijk );
        const double lox = cell_lower( g, 0, ijk[ 0 ] ), loy = cell_lower( g, 1, ijk[ 1 ] );
        // the reference clears the reconstructions every pass (modifiedyoungs.hh:65): only mixed cells carry one.  The resident
        // loop never clears its array, so a stale entry of a cell that is no longer mixed must not be read.
        const unsigned char f = flags[ c ];
        const bool mixed = ( f == 1 || f == 2 );
        const Hs2 h = { mixed ? hs[ 3 * c ] : 0.0, mixed ? hs[ 3 * c + 1 ] : 0.0, mixed ? hs[ 3 * c + 2 ] : 0.0 };
        v.l1 += l1error_circle_cell( lox, loy, g.h[ 0 ], g.h[ 1 ], h, f == 3, cx, cy, radius );
      }
      v = block_reduce( v );
      if( threadIdx.x == 0 )
        partial[ blockIdx.x ] = v;
    }

    __global__ void __launch_bounds__( kDiagThreads ) k_l1error_final ( const Diag *__restrict__ partial, double *__restrict__ out )
    {
      Diag v = { 0.0, DBL_MAX, -DBL_MAX, 0.0 };
      for( int b = threadIdx.x; b < kDiagBlocks; b += kDiagThreads )
        v = combine( v, partial[ b ] );
      v = block_reduce( v );
      if( threadIdx.x == 0 )
        out[ 0 ] = v.l1;
    }
  } // namespace

  namespace launch
  {
    void circle_init ( int blocks, cudaStream_t s, Grid g, double cx, double cy, double radius, double *u )
    { k_circle_init<<< blocks, 128, 0, s >>>( g, cx, cy, radius, u ); }

    double *l1error_circle ( cudaStream_t s, Grid g, const unsigned char *flags, const double *hs, double cx, double cy, double radius, void *scratch )
    {
      Diag *partial = static_cast< Diag * >( scratch );
      double *out = reinterpret_cast< double * >( partial + kDiagBlocks );
      k_l1error_circle_partial<<< kDiagBlocks, kDiagThreads, 0, s >>>( g, flags, hs, cx, cy, radius, partial );
      k_l1error_final<<< 1, kDiagThreads, 0, s >>>( partial, out );
      return out;
    }

    size_t diag_scratch_bytes () { return sizeof( Diag ) * kDiagBlocks + 8 * sizeof( double ); }

    // scratch: diag_scratch_bytes() of device memory; the five results are written behind the block partials
    double *diagnostics ( cudaStream_t s, long long n, const double *u, const double *exact, double volume, void *scratch )
    {
      Diag *partial = static_cast< Diag * >( scratch );
      double *out = reinterpret_cast< double * >( partial + kDiagBlocks );
      k_diag_partial<<< kDiagBlocks, kDiagThreads, 0, s >>>( n, u, exact, partial );
      k_diag_final<<< 1, kDiagThreads, 0, s >>>( partial, volume, out );
      return out;
    }
  } // namespace launch
} // namespace vofx
