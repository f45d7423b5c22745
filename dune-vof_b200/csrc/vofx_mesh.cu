// vofx on unstructured grids given as flattened arrays (include/vofx.h: vofx_mesh_desc) -- triangles and quadrilaterals in 2D,
// tetrahedra and hexahedra in 3D: kernels and the C ABI.  The arithmetic restates, operation for operation (vofx_math.cuh, -fmad=false),
//   F1  flagging.hh:35-70
//   R1  reconstruction/modifiedyoungs.hh:63-134 over the CSR vertex-neighbour stencil (stencil/vertexneighborsstencil.hh:52-94)
//   G1  locateHalfSpace( Polygon ), geometry/algorithm.hh:68-111        (vofx_geom2d.cuh)
//   E1  evolution/evolution.hh:58-145, G7 2d/upwindpolygon.hh:24-30
//   3D: makePolyhedron( simplex / cube ), 3d/polyhedron.hh:366-404; G2 locateHalfSpace( Polyhedron ), geometry/algorithm.hh:119-236;
//       G7 upwindPolygon3d for triangular and quadrilateral intersections, 3d/upwindpolygon.hh:24-74; G5 (vofx_geom3d.cuh: topology
//       policies topo::Tet / topo::Prism / topo::Cube)
// Layout: SoA device copies of the descriptor arrays; flags 1 B/cell; half-spaces (n, d) 24 B/cell; per-intersection flux
// records (flux, v.N, direction) written by the mixed cells and GATHERED by every cell in the reference's accumulation
// order (ascending index of the contributing mixed cell, its faces in order), so that no atomics touch the update and
// the sums carry the reference's bits.
#include "../../include/vofx.h"

#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include <cuda_runtime.h>

#include "vofx_geom2d.cuh"
#include "vofx_geom3d.cuh"

namespace vofx
{
  int set_error ( int status, const std::string &msg );   // vofx_capi.cu
}

using namespace vofx;

namespace
{

#define VOFX_MESH_CUDA( call )                                                                                  \
  do                                                                                                            \
  {                                                                                                             \
    const cudaError_t vofx_err_ = ( call );                                                                     \
    if( vofx_err_ != cudaSuccess )                                                                              \
      return set_error( VOFX_ERR_CUDA, std::string( #call ) + ": " + cudaGetErrorString( vofx_err_ ) );         \
  } while( 0 )

  struct MeshDev
  {
    int dim;
    int nCells, nFaces;
    const int *faceCornerOffsets;    // 3D: corners of intersection s are faceCorners[ 3 * faceCornerOffsets[ s ] ... ]
    const int *cornerOffsets;
    const double *corners, *centers, *volumes;
    const int *faceOffsets, *faceNeighbor, *backSlot;
    const double *faceCorners, *faceCenters, *faceUnit, *faceInt, *faceVolumes;
    const int *stencilOffsets, *stencil;
  };

  // device-resident loop state
  struct MeshState
  {
    double time, dt;
    unsigned long long dtEstBits;   // min over the mixed cells of volume / sumFluxes, as the bits of a non-negative double
    int mixedCount, pad;
  };

  __device__ __forceinline__ bool is_mixed ( unsigned char f ) { return f == 1 || f == 2; }   // flagset.hh:71-72,78

  // flagging.hh:35-70; mixed cells are appended to the list (its order does not enter any result)
  __global__ void __launch_bounds__( 256 ) k_mesh_flag ( MeshDev M, const double *__restrict__ u, double eps, unsigned char *__restrict__ flags,
                                                          int *__restrict__ mixedList, MeshState *state )
  {
    for( int c = blockIdx.x * blockDim.x + threadIdx.x; c < M.nCells; c += gridDim.x * blockDim.x )
    {
      const double colorEn = u[ c ];
      unsigned char flag;
      if( colorEn < eps )
        flag = 0;
      else if( colorEn <= ( 1 - eps ) )
        flag = 1;
      else if( colorEn != colorEn )
        flag = 99;
      else
      {
        flag = 3;
        for( int s = M.faceOffsets[ c ]; s < M.faceOffsets[ c + 1 ]; ++s )
        {
          const int nb = M.faceNeighbor[ s ];
          if( nb >= 0 && u[ nb ] < eps )
          {
            flag = 2;
            break;
          }
        }
      }
      flags[ c ] = flag;
      if( mixedList && is_mixed( flag ) )
        mixedList[ atomicAdd( &state->mixedCount, 1 ) ] = c;
    }
  }

  __global__ void __launch_bounds__( 256 ) k_mesh_list_mixed ( MeshDev M, const unsigned char *__restrict__ flags, int *__restrict__ mixedList,
                                                                MeshState *state )
  {
    for( int c = blockIdx.x * blockDim.x + threadIdx.x; c < M.nCells; c += gridDim.x * blockDim.x )
      if( is_mixed( flags[ c ] ) )
        mixedList[ atomicAdd( &state->mixedCount, 1 ) ] = c;
  }

  // one least-squares contribution, modifiedyoungs.hh:104-108 / :117-123
  __device__ __forceinline__ void ls_accumulate2 ( double dx, double dy, double colorDiff, double AtA[ 2 ][ 2 ], double *Atb )
  {
    double n2 = 0.0;
    n2 += dx * dx;
    n2 += dy * dy;
    const double weight = 1.0 / n2;
    const double d[ 2 ] = { dx * weight, dy * weight };
#pragma unroll
    for( int i = 0; i < 2; ++i )
#pragma unroll
      for( int j = 0; j < 2; ++j )
      {
        double m = 0.0;          // outerProduct: axpy onto a zero matrix, geometry/utility.hh:161-169
        m += d[ i ] * d[ j ];
        AtA[ i ][ j ] += m;
      }
    const double w = weight * colorDiff;
    Atb[ 0 ] += w * d[ 0 ];
    Atb[ 1 ] += w * d[ 1 ];
  }

  // makePolygon( geometry ), 2d/polygon.hh:230-241
  __device__ __forceinline__ void cell_polygon ( const MeshDev &M, int c, Poly2 &p )
  {
    const int b = M.cornerOffsets[ c ], n = M.cornerOffsets[ c + 1 ] - b;
    p.n = n;
    for( int k = 0; k < n; ++k )
    {
      const int src = ( n == 4 && k >= 2 ) ? 5 - k : k;   // quadrilateral: corners 0,1,3,2
      p.x[ k ] = M.corners[ 2 * ( b + src ) ];
      p.y[ k ] = M.corners[ 2 * ( b + src ) + 1 ];
    }
  }

  // applyLocal, modifiedyoungs.hh:90-134: one thread per mixed cell; the half-spaces were cleared before
  __global__ void __launch_bounds__( 128 ) k_mesh_youngs ( MeshDev M, const double *__restrict__ u, const int *__restrict__ mixedList,
                                                            const MeshState *state, double *__restrict__ hs )
  {
    const int count = state->mixedCount;
    for( int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x )
    {
      const int c = mixedList[ idx ];
      const double cx = M.centers[ 2 * c ], cy = M.centers[ 2 * c + 1 ];
      const double colorEn = u[ c ];
      double AtA[ 2 ][ 2 ] = { { 0.0, 0.0 }, { 0.0, 0.0 } }, Atb[ 2 ] = { 0.0, 0.0 };
      for( int s = M.stencilOffsets[ c ]; s < M.stencilOffsets[ c + 1 ]; ++s )
      {
        const int nb = M.stencil[ s ];
        const double colorNb = clamp01( u[ nb ] );
        ls_accumulate2( M.centers[ 2 * nb ] - cx, M.centers[ 2 * nb + 1 ] - cy, colorNb - colorEn, AtA, Atb );
      }
      // boundary intersections get a ghost with colour 0.5 at twice the distance of the intersection centre, :112-124
      for( int s = M.faceOffsets[ c ]; s < M.faceOffsets[ c + 1 ]; ++s )
      {
        if( M.faceNeighbor[ s ] >= 0 )
          continue;
        double dx = M.faceCenters[ 2 * s ] - cx, dy = M.faceCenters[ 2 * s + 1 ] - cy;
        dx *= 2.0;
        dy *= 2.0;
        ls_accumulate2( dx, dy, 0.5 - colorEn, AtA, Atb );
      }
      // FieldMatrix< double, 2, 2 >::solve (dune-common closed form, DESIGN.md "parity unpinned")
      double detinv = AtA[ 0 ][ 0 ] * AtA[ 1 ][ 1 ] - AtA[ 0 ][ 1 ] * AtA[ 1 ][ 0 ];
      detinv = 1.0 / detinv;
      double nx = detinv * ( AtA[ 1 ][ 1 ] * Atb[ 0 ] - AtA[ 0 ][ 1 ] * Atb[ 1 ] );
      double ny = detinv * ( AtA[ 0 ][ 0 ] * Atb[ 1 ] - AtA[ 1 ][ 0 ] * Atb[ 0 ] );
      const double n2 = dot2( nx, ny, nx, ny );
      if( n2 < kEps )
        continue;                       // the reconstruction stays the zero half-space, :128-129
      const double nrm = sqrt( n2 );
      nx = nx / nrm;
      ny = ny / nrm;
      Poly2 p;
      cell_polygon( M, c, p );
      const double dist = locate( p, nx, ny, colorEn );
      hs[ 3 * c ] = nx;
      hs[ 3 * c + 1 ] = ny;
      hs[ 3 * c + 2 ] = dist;
    }
  }

  // ---- 3D -----------------------------------------------------------------------------------------------------------------

  __device__ __forceinline__ void ls_accumulate3 ( double *d, double colorDiff, double AtA[ 3 ][ 3 ], double *Atb )
  {
    double n2 = 0.0;
#pragma unroll
    for( int k = 0; k < 3; ++k )
      n2 += d[ k ] * d[ k ];
    const double weight = 1.0 / n2;
#pragma unroll
    for( int k = 0; k < 3; ++k )
      d[ k ] *= weight;
#pragma unroll
    for( int i = 0; i < 3; ++i )
#pragma unroll
      for( int j = 0; j < 3; ++j )
      {
        double m = 0.0;          // outerProduct: axpy onto a zero matrix, geometry/utility.hh:161-169
        m += d[ i ] * d[ j ];
        AtA[ i ][ j ] += m;
      }
    const double w = weight * colorDiff;
#pragma unroll
    for( int k = 0; k < 3; ++k )
      Atb[ k ] += w * d[ k ];
  }

  // FieldMatrix< double, 3, 3 >::solve of dune-common 2.6 (closed form; DESIGN.md "parity unpinned"), as the oracle's solve_small
  __device__ __forceinline__ void solve3_plain ( const double M[ 3 ][ 3 ], const double *b, double *x )
  {
    const double t4 = M[ 0 ][ 0 ] * M[ 1 ][ 1 ];
    const double t6 = M[ 0 ][ 0 ] * M[ 1 ][ 2 ];
    const double t8 = M[ 0 ][ 1 ] * M[ 1 ][ 0 ];
    const double t10 = M[ 0 ][ 2 ] * M[ 1 ][ 0 ];
    const double t12 = M[ 0 ][ 1 ] * M[ 2 ][ 0 ];
    const double t14 = M[ 0 ][ 2 ] * M[ 2 ][ 0 ];
    const double d = ( t4 * M[ 2 ][ 2 ] - t6 * M[ 2 ][ 1 ] - t8 * M[ 2 ][ 2 ] + t10 * M[ 2 ][ 1 ] + t12 * M[ 1 ][ 2 ] - t14 * M[ 1 ][ 1 ] );
    x[ 0 ] = ( b[ 0 ] * M[ 1 ][ 1 ] * M[ 2 ][ 2 ] - b[ 0 ] * M[ 2 ][ 1 ] * M[ 1 ][ 2 ]
               - b[ 1 ] * M[ 0 ][ 1 ] * M[ 2 ][ 2 ] + b[ 1 ] * M[ 2 ][ 1 ] * M[ 0 ][ 2 ]
               + b[ 2 ] * M[ 0 ][ 1 ] * M[ 1 ][ 2 ] - b[ 2 ] * M[ 1 ][ 1 ] * M[ 0 ][ 2 ] ) / d;
    x[ 1 ] = ( M[ 0 ][ 0 ] * b[ 1 ] * M[ 2 ][ 2 ] - M[ 0 ][ 0 ] * b[ 2 ] * M[ 1 ][ 2 ]
               - M[ 1 ][ 0 ] * b[ 0 ] * M[ 2 ][ 2 ] + M[ 1 ][ 0 ] * b[ 2 ] * M[ 0 ][ 2 ]
               + M[ 2 ][ 0 ] * b[ 0 ] * M[ 1 ][ 2 ] - M[ 2 ][ 0 ] * b[ 1 ] * M[ 0 ][ 2 ] ) / d;
    x[ 2 ] = ( M[ 0 ][ 0 ] * M[ 1 ][ 1 ] * b[ 2 ] - M[ 0 ][ 0 ] * M[ 2 ][ 1 ] * b[ 1 ]
               - M[ 1 ][ 0 ] * M[ 0 ][ 1 ] * b[ 2 ] + M[ 1 ][ 0 ] * M[ 2 ][ 1 ] * b[ 0 ]
               + M[ 2 ][ 0 ] * M[ 0 ][ 1 ] * b[ 1 ] - M[ 2 ][ 0 ] * M[ 1 ][ 1 ] * b[ 0 ] ) / d;
  }

  // applyLocal, modifiedyoungs.hh:90-134, on tetrahedra and hexahedra: one thread per mixed cell
  __global__ void __launch_bounds__( 64 ) k_mesh_youngs3 ( MeshDev M, const double *__restrict__ u, const int *__restrict__ mixedList,
                                                            const MeshState *state, double *__restrict__ hs )
  {
    const int count = state->mixedCount;
    for( int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x )
    {
      const int c = mixedList[ idx ];
      const double center[ 3 ] = { M.centers[ 3 * c ], M.centers[ 3 * c + 1 ], M.centers[ 3 * c + 2 ] };
      const double colorEn = u[ c ];
      double AtA[ 3 ][ 3 ] = { { 0.0, 0.0, 0.0 }, { 0.0, 0.0, 0.0 }, { 0.0, 0.0, 0.0 } }, Atb[ 3 ] = { 0.0, 0.0, 0.0 };
      for( int s = M.stencilOffsets[ c ]; s < M.stencilOffsets[ c + 1 ]; ++s )
      {
        const int nb = M.stencil[ s ];
        const double colorNb = clamp01( u[ nb ] );
        double d[ 3 ];
#pragma unroll
        for( int k = 0; k < 3; ++k )
          d[ k ] = M.centers[ 3 * nb + k ] - center[ k ];
        ls_accumulate3( d, colorNb - colorEn, AtA, Atb );
      }
      for( int s = M.faceOffsets[ c ]; s < M.faceOffsets[ c + 1 ]; ++s )
      {
        if( M.faceNeighbor[ s ] >= 0 )
          continue;
        double d[ 3 ];
#pragma unroll
        for( int k = 0; k < 3; ++k )
        {
          d[ k ] = M.faceCenters[ 3 * s + k ] - center[ k ];
          d[ k ] *= 2.0;
        }
        ls_accumulate3( d, 0.5 - colorEn, AtA, Atb );
      }
      double normal[ 3 ] = { 0.0, 0.0, 0.0 };
      solve3_plain( AtA, Atb, normal );
      const double n2 = dot3( normal, normal );
      if( n2 < kEps )
        continue;                       // the reconstruction stays the zero half-space, :128-129
      const double nrm = sqrt( n2 );
#pragma unroll
      for( int k = 0; k < 3; ++k )
        normal[ k ] = normal[ k ] / nrm;
      const int b = M.cornerOffsets[ c ], nc = M.cornerOffsets[ c + 1 ] - b;
      const double dist = locate_corners( M.corners + 3 * size_t( b ), nc, normal, colorEn );
      hs[ 4 * c ] = normal[ 0 ];
      hs[ 4 * c + 1 ] = normal[ 1 ];
      hs[ 4 * c + 2 ] = normal[ 2 ];
      hs[ 4 * c + 3 ] = dist;
    }
  }

  // the per-intersection part of Evolution::applyLocal (evolution.hh:100-141) in 3D
  __global__ void __launch_bounds__( 64 ) k_mesh_flux3 ( MeshDev M, const double *__restrict__ hs, const double *__restrict__ faceVelocity,
                                                          const double *dtPtr, const int *__restrict__ mixedList, MeshState *state,
                                                          double *__restrict__ recFlux, double *__restrict__ recVN, signed char *__restrict__ recDir )
  {
    const int count = state->mixedCount;
    const double dt = *dtPtr;
    for( int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x )
    {
      const int c = mixedList[ idx ];
      const Hs3 rec = { { hs[ 4 * c ], hs[ 4 * c + 1 ], hs[ 4 * c + 2 ] }, hs[ 4 * c + 3 ] };
      double sumFluxes = 0.0;
      for( int s = M.faceOffsets[ c ]; s < M.faceOffsets[ c + 1 ]; ++s )
      {
        if( M.faceNeighbor[ s ] < 0 )
          continue;
        const double *n = M.faceUnit + 3 * size_t( s );
        double v[ 3 ] = { faceVelocity[ 3 * size_t( s ) ], faceVelocity[ 3 * size_t( s ) + 1 ], faceVelocity[ 3 * size_t( s ) + 2 ] };
        sumFluxes += M.faceVolumes[ s ] * fabs( dot3( n, v ) );
#pragma unroll
        for( int k = 0; k < 3; ++k )
          v[ k ] *= dt;
        const double vn = dot3( v, n );
        double flux = 0.0;
        if( vn > 0 )
        {
          const int fb = M.faceCornerOffsets[ s ], fnc = M.faceCornerOffsets[ s + 1 ] - fb;
          flux = flux_corners( M.faceCorners + 3 * size_t( fb ), fnc, v, rec );
        }
        recFlux[ s ] = flux;
        recVN[ s ] = dot3( v, M.faceInt + 3 * size_t( s ) );
        recDir[ s ] = ( vn > 0 ) ? 1 : ( ( vn < 0 ) ? -1 : 0 );
      }
      const double est = M.volumes[ c ] / sumFluxes;
      if( est >= 0.0 )
        atomicMin( &state->dtEstBits, (unsigned long long) __double_as_longlong( est ) );
    }
  }

  // interface( entity, reconstructions ) of a cell located with a given normal: centroid of the cut segment
  // ( locateHalfSpace + intersect( polygon, boundary ) + Line::centroid, 2d/polygon.hh:183-188 )
  __device__ void mesh_interface_centroid ( const MeshDev &M, int c, double nx, double ny, double color, double &cx, double &cy )
  {
    Poly2 p;
    cell_polygon( M, c, p );
    const Hs2 H = { nx, ny, locate( p, nx, ny, color ) };
    double lx[ 2 ], ly[ 2 ];
    plane_cut( p, H, lx, ly );
    cx = ( lx[ 0 ] + lx[ 1 ] ) * 0.5;
    cy = ( ly[ 0 ] + ly[ 1 ] ) * 0.5;
  }

  // ModifiedSwartzReconstruction::applyLocal in 2D, reconstruction/modifiedswartz.hh:102-153: one thread per mixed cell,
  // on top of the Young's half-spaces already in hs (a cell reads and writes only its own entry)
  __global__ void __launch_bounds__( 64 ) k_mesh_swartz ( MeshDev M, const double *__restrict__ u, const unsigned char *__restrict__ flags,
                                                           const int *__restrict__ mixedList, const MeshState *state, double *__restrict__ hs,
                                                           int maxIterations )
  {
    const int count = state->mixedCount;
    for( int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x )
    {
      const int c = mixedList[ idx ];
      const double colorEn = u[ c ];
      double nx = hs[ 3 * c ], ny = hs[ 3 * c + 1 ];
      int iterations = 0;
      double residuum = 0.0;
      do
      {
        const double ox = nx, oy = ny;
        nx = ny = 0.0;
        double ex, ey;
        mesh_interface_centroid( M, c, ox, oy, colorEn, ex, ey );
        for( int s = M.stencilOffsets[ c ]; s < M.stencilOffsets[ c + 1 ]; ++s )
        {
          const int nb = M.stencil[ s ];
          if( !is_mixed( flags[ nb ] ) )
            continue;
          double bx, by;
          mesh_interface_centroid( M, nb, ox, oy, u[ nb ], bx, by );
          const double dx = bx - ex, dy = by - ey;
          double cx = -dy, cy = dx;                     // generalizedCrossProduct, geometry/utility.hh:82-85
          if( dot2( cx, cy, ox, oy ) < 0 )
          {
            cx *= -1.0;
            cy *= -1.0;
          }
          const double nrm = norm2( cx, cy );
          cx /= nrm;
          cy /= nrm;
          nx += 1.0 * cx;
          ny += 1.0 * cy;
        }
        const double nrm = norm2( nx, ny );
        if( nrm < 0.5 )
          break;                                        // "return": keep the last reconstruction, :140-141
        nx /= nrm;
        ny /= nrm;
        Poly2 p;
        cell_polygon( M, c, p );
        hs[ 3 * c ] = nx;
        hs[ 3 * c + 1 ] = ny;
        hs[ 3 * c + 2 ] = locate( p, nx, ny, colorEn );
        residuum = 1.0 - dot2( nx, ny, ox, oy );
        ++iterations;
      }
      while( residuum > 1e-12 && iterations < maxIterations );
    }
  }

  // truncVolume( upwindPolygon2d( iGeometry, v ), hs ), 2d/upwindpolygon.hh:24-30 + geometry/upwindpolygon.hh:58-62
  __device__ double mesh_flux ( const double *fc, double vx, double vy, const Hs2 &hs )
  {
    double ax = fc[ 0 ], ay = fc[ 1 ], bx = fc[ 2 ], by = fc[ 3 ];
    const double dx = bx - ax, dy = by - ay;
    if( !( dot2( -dy, dx, vx, vy ) < 0.0 ) )   // generalizedCrossProduct( c1 - c0 ) * v, geometry/utility.hh:82-85
    {
      double t = ax; ax = bx; bx = t;
      t = ay; ay = by; by = t;
    }
    Poly2 up, cl;
    up.n = 4;
    up.x[ 0 ] = ax; up.y[ 0 ] = ay;
    up.x[ 1 ] = bx; up.y[ 1 ] = by;
    up.x[ 2 ] = bx - vx; up.y[ 2 ] = by - vy;
    up.x[ 3 ] = ax - vx; up.y[ 3 ] = ay - vy;
    poly_clip( up, hs, cl );
    return poly_volume( cl );
  }

  // the per-intersection part of Evolution::applyLocal (evolution.hh:100-141) for every mixed cell: flux and v.N of each
  // of its intersections, the direction of v there, and the cell's time-step estimate volume / sumFluxes
  __global__ void __launch_bounds__( 128 ) k_mesh_flux ( MeshDev M, const double *__restrict__ hs, const double *__restrict__ faceVelocity,
                                                          const double *dtPtr, const int *__restrict__ mixedList, MeshState *state,
                                                          double *__restrict__ recFlux, double *__restrict__ recVN, signed char *__restrict__ recDir )
  {
    const int count = state->mixedCount;
    const double dt = *dtPtr;
    for( int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x )
    {
      const int c = mixedList[ idx ];
      const Hs2 rec = { hs[ 3 * c ], hs[ 3 * c + 1 ], hs[ 3 * c + 2 ] };
      double sumFluxes = 0.0;
      for( int s = M.faceOffsets[ c ]; s < M.faceOffsets[ c + 1 ]; ++s )
      {
        if( M.faceNeighbor[ s ] < 0 )
          continue;
        const double nx = M.faceUnit[ 2 * s ], ny = M.faceUnit[ 2 * s + 1 ];
        double vx = faceVelocity[ 2 * s ], vy = faceVelocity[ 2 * s + 1 ];
        sumFluxes += M.faceVolumes[ s ] * fabs( dot2( nx, ny, vx, vy ) );
        vx *= dt;
        vy *= dt;
        const double vn = dot2( vx, vy, nx, ny );
        double flux = 0.0;
        if( vn > 0 )
          flux = mesh_flux( M.faceCorners + 4 * s, vx, vy, rec );
        recFlux[ s ] = flux;
        recVN[ s ] = dot2( vx, vy, M.faceInt[ 2 * s ], M.faceInt[ 2 * s + 1 ] );
        recDir[ s ] = ( vn > 0 ) ? 1 : ( ( vn < 0 ) ? -1 : 0 );
      }
      const double est = M.volumes[ c ] / sumFluxes;
      if( est >= 0.0 )                  // not NaN; +inf never wins against the initial DBL_MAX
        atomicMin( &state->dtEstBits, (unsigned long long) __double_as_longlong( est ) );
    }
  }

  // the accumulation of Evolution::operator() (evolution.hh:64-71, :117-141) as a gather: the entity loop visits the
  // mixed cells in ascending index and each adds to itself and to the outside cells of its intersections; cell c therefore
  // receives, in this order, the contributions of its mixed neighbours below c, its own (if mixed), its mixed neighbours above
  template< bool FUSED >
  __global__ void __launch_bounds__( 256 ) k_mesh_update ( MeshDev M, const unsigned char *__restrict__ flags, const double *__restrict__ recFlux,
                                                            const double *__restrict__ recVN, const signed char *__restrict__ recDir,
                                                            double *__restrict__ target )
  {
    for( int c = blockIdx.x * blockDim.x + threadIdx.x; c < M.nCells; c += gridDim.x * blockDim.x )
    {
      const unsigned char flag = flags[ c ];
      const int begin = M.faceOffsets[ c ], end = M.faceOffsets[ c + 1 ];
      // contributors ( cell, slot in that cell ) sorted lexicographically; own entry ( c, -1 )
      int key1[ 7 ], key2[ 7 ], n = 0;     // the cell itself and at most six intersections (hexahedra)
      if( is_mixed( flag ) )
      {
        key1[ n ] = c;
        key2[ n++ ] = -1;
      }
      for( int s = begin; s < end && n < 7; ++s )
      {
        const int nb = M.faceNeighbor[ s ];
        if( nb < 0 || !is_mixed( flags[ nb ] ) )
          continue;
        const int back = M.backSlot[ s ];
        int k = n++;
        while( k > 0 && ( key1[ k - 1 ] > nb || ( key1[ k - 1 ] == nb && key2[ k - 1 ] > back ) ) )
        {
          key1[ k ] = key1[ k - 1 ];
          key2[ k ] = key2[ k - 1 ];
          --k;
        }
        key1[ k ] = nb;
        key2[ k ] = back;
      }
      if( n == 0 )
      {
        if( !FUSED )
          target[ c ] = 0.0;             // update.clear(), :62
        continue;
      }
      const double volume = M.volumes[ c ];
      double upd = 0.0;
      for( int k = 0; k < n; ++k )
      {
        if( key2[ k ] < 0 )
        {
          // own intersections in order, :100-141
          for( int s = begin; s < end; ++s )
          {
            const int nb = M.faceNeighbor[ s ];
            if( nb < 0 )
              continue;
            if( recDir[ s ] > 0 )
              upd -= recFlux[ s ] / volume;
            if( flags[ nb ] == 3 && recDir[ s ] < 0 )
              upd -= recVN[ s ] / volume;
          }
        }
        else
        {
          const int b = key2[ k ];
          if( recDir[ b ] > 0 )
          {
            if( flag == 3 )
              upd -= ( recVN[ b ] - recFlux[ b ] ) / volume;
            if( flag == 0 )
              upd += recFlux[ b ] / volume;
            if( is_mixed( flag ) )
              upd += recFlux[ b ] / volume;
          }
        }
      }
      if( FUSED )
        target[ c ] += upd;              // uh.axpy( 1.0, update ), colorfunction.hh:31-37
      else
        target[ c ] = upd;
    }
  }

  __global__ void k_mesh_begin ( MeshState *state )
  {
    state->mixedCount = 0;
    state->dtEstBits = (unsigned long long) __double_as_longlong( DBL_MAX );
  }

  // algorithm.hh:89-91: time += dt (if dt > 0), dt = cfl * dtEst
  __global__ void k_mesh_finish ( MeshState *state, double cfl )
  {
    if( state->dt > 0.0 )
      state->time += state->dt;
    state->dt = __longlong_as_double( (long long) state->dtEstBits ) * cfl;
  }

  __global__ void k_mesh_set_dt ( MeshState *state, double time, double dt )
  {
    state->time = time;
    state->dt = dt;
  }

  template< class T >
  cudaError_t upload ( T *&dst, const T *src, size_t count )
  {
    cudaError_t e = cudaMalloc( &dst, ( count ? count : 1 ) * sizeof( T ) );
    if( e != cudaSuccess )
      return e;
    return cudaMemcpy( dst, src, count * sizeof( T ), cudaMemcpyHostToDevice );
  }

} // namespace

struct vofx_mesh
{
  int device = 0;
  cudaStream_t stream = nullptr, ownStream = nullptr;
  MeshDev dev {};
  std::vector< void * > owned;       // device allocations behind dev
  MeshState *state = nullptr;
  int *mixedList = nullptr;
  double *recFlux = nullptr, *recVN = nullptr;
  signed char *recDir = nullptr;
  // device-resident loop + staging of the host variants
  double *u = nullptr, *hs = nullptr, *upd = nullptr, *vel = nullptr;
  unsigned char *flags = nullptr;
  bool colorSet = false, velocitySet = false;
  long long launches = 0;
  int blocksCells = 1;
  int reconstruction = VOFX_RECON_YOUNGS, maxIterations = 10;   // of vofx_mesh_run (vofx_mesh_set_reconstruction)
  // the captured loop pass (vofx_mesh_run)
  cudaGraphExec_t passGraph = nullptr;
  double graphEps = 0.0, graphCfl = 0.0;
  cudaStream_t graphStream = nullptr;
  int launchesPerPass = 0;
};

namespace
{

  int bandBlocks ( const vofx_mesh *m ) { return 148 * 2; }

  int launchFlag ( vofx_mesh *m, const double *u, double eps, unsigned char *flags, bool list )
  {
    k_mesh_begin<<< 1, 1, 0, m->stream >>>( m->state );
    k_mesh_flag<<< m->blocksCells, 256, 0, m->stream >>>( m->dev, u, eps, flags, list ? m->mixedList : nullptr, m->state );
    m->launches += 2;
    VOFX_MESH_CUDA( cudaGetLastError() );
    return VOFX_OK;
  }

  int launchYoungs ( vofx_mesh *m, const double *u, double *hs )
  {
    VOFX_MESH_CUDA( cudaMemsetAsync( hs, 0, sizeof( double ) * ( m->dev.dim + 1 ) * size_t( m->dev.nCells ), m->stream ) );   // reconstructions.clear(), :65
    if( m->dev.dim == 3 )
      k_mesh_youngs3<<< bandBlocks( m ) * 2, 64, 0, m->stream >>>( m->dev, u, m->mixedList, m->state, hs );
    else
      k_mesh_youngs<<< bandBlocks( m ), 128, 0, m->stream >>>( m->dev, u, m->mixedList, m->state, hs );
    m->launches += 1;
    VOFX_MESH_CUDA( cudaGetLastError() );
    return VOFX_OK;
  }

  int launchSwartz ( vofx_mesh *m, const double *u, const unsigned char *flags, double *hs, int maxIterations )
  {
    if( m->dev.dim != 2 )
      return set_error( VOFX_ERR_ARGUMENT, "the Swartz reconstruction on unstructured grids is built for 2D meshes only" );
    k_mesh_swartz<<< bandBlocks( m ), 64, 0, m->stream >>>( m->dev, u, flags, m->mixedList, m->state, hs, maxIterations );
    m->launches += 1;
    VOFX_MESH_CUDA( cudaGetLastError() );
    return VOFX_OK;
  }

  template< bool FUSED >
  int launchEvolve ( vofx_mesh *m, const double *hs, const unsigned char *flags, const double *vel, const double *dtPtr, double *target )
  {
    if( m->dev.dim == 3 )
      k_mesh_flux3<<< bandBlocks( m ) * 2, 64, 0, m->stream >>>( m->dev, hs, vel, dtPtr, m->mixedList, m->state, m->recFlux, m->recVN, m->recDir );
    else
      k_mesh_flux<<< bandBlocks( m ), 128, 0, m->stream >>>( m->dev, hs, vel, dtPtr, m->mixedList, m->state, m->recFlux, m->recVN, m->recDir );
    k_mesh_update< FUSED ><<< m->blocksCells, 256, 0, m->stream >>>( m->dev, flags, m->recFlux, m->recVN, m->recDir, target );
    m->launches += 2;
    VOFX_MESH_CUDA( cudaGetLastError() );
    return VOFX_OK;
  }

  // the mixed list for operator-level calls that receive the flags from the caller
  int listMixed ( vofx_mesh *m, const unsigned char *flags )
  {
    k_mesh_begin<<< 1, 1, 0, m->stream >>>( m->state );
    k_mesh_list_mixed<<< m->blocksCells, 256, 0, m->stream >>>( m->dev, flags, m->mixedList, m->state );
    m->launches += 2;
    VOFX_MESH_CUDA( cudaGetLastError() );
    return VOFX_OK;
  }

  int checkMesh ( const vofx_mesh *m )
  {
    if( !m )
      return set_error( VOFX_ERR_ARGUMENT, "null mesh" );
    VOFX_MESH_CUDA( cudaSetDevice( m->device ) );
    return VOFX_OK;
  }

} // namespace

extern "C"
{

  int vofx_mesh_create ( const vofx_mesh_desc *d, int device, vofx_mesh **out )
  {
    if( !d || !out )
      return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_create: null argument" );
    *out = nullptr;
    if( d->dim != 2 && d->dim != 3 )
      return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_create: dim must be 2 (triangles, quadrilaterals) or 3 (tetrahedra, hexahedra)" );
    const int dim = d->dim;
    if( dim == 3 && !d->face_corner_offsets )
      return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_create: a 3D mesh needs face_corner_offsets" );
    if( d->n_cells <= 0 || d->n_cells > 0x7fffffffLL )
      return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_create: n_cells out of range" );
    if( !d->corner_offsets || !d->corners || !d->centers || !d->volumes || !d->face_offsets || !d->face_neighbor || !d->face_corners
        || !d->face_centers || !d->face_unit_normals || !d->face_int_normals || !d->face_volumes || !d->stencil_offsets || !d->stencil )
      return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_create: null array in the descriptor" );
    const int n = int( d->n_cells );
    if( d->corner_offsets[ 0 ] != 0 || d->face_offsets[ 0 ] != 0 || d->stencil_offsets[ 0 ] != 0 )
      return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_create: CSR offsets must start at 0" );
    for( int c = 0; c < n; ++c )
    {
      const int nc = d->corner_offsets[ c + 1 ] - d->corner_offsets[ c ];
      if( dim == 2 ? ( nc != 3 && nc != 4 ) : ( nc != 4 && nc != 8 ) )
        return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_create: Invalid GeometryType (cells must be triangles or quadrilaterals in 2D, tetrahedra or hexahedra in 3D)" );
      const int nf = d->face_offsets[ c + 1 ] - d->face_offsets[ c ];
      if( nf < 0 || nf > ( dim == 2 ? 4 : 6 ) )
        return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_create: a cell has too many intersections" );
      for( int s = d->face_offsets[ c ]; s < d->face_offsets[ c + 1 ]; ++s )
        if( d->face_neighbor[ s ] >= n || d->face_neighbor[ s ] == c )
          return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_create: face_neighbor out of range" );
      if( d->stencil_offsets[ c + 1 ] < d->stencil_offsets[ c ] )
        return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_create: stencil_offsets not monotone" );
      for( int s = d->stencil_offsets[ c ]; s < d->stencil_offsets[ c + 1 ]; ++s )
      {
        const int nb = d->stencil[ s ];
        if( nb < 0 || nb >= n || nb == c || ( s > d->stencil_offsets[ c ] && d->stencil[ s - 1 ] >= nb ) )
          return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_create: stencil must be ascending, without the cell itself" );
      }
    }
    const int nFaces = d->face_offsets[ n ], nCorners = d->corner_offsets[ n ], nStencil = d->stencil_offsets[ n ];
    int nFaceCorners = 2 * nFaces;
    if( dim == 3 )
    {
      if( d->face_corner_offsets[ 0 ] != 0 )
        return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_create: CSR offsets must start at 0" );
      for( int s = 0; s < nFaces; ++s )
      {
        const int fnc = d->face_corner_offsets[ s + 1 ] - d->face_corner_offsets[ s ];
        if( fnc != 3 && fnc != 4 )
          return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_create: Invalid GeometryType (intersections must be triangles or quadrilaterals)" );
      }
      nFaceCorners = d->face_corner_offsets[ nFaces ];
    }

    // for every intersection the slot of the same intersection seen from the outside cell
    std::vector< int > back( size_t( nFaces ), -1 );
    for( int c = 0; c < n; ++c )
      for( int s = d->face_offsets[ c ]; s < d->face_offsets[ c + 1 ]; ++s )
      {
        const int nb = d->face_neighbor[ s ];
        if( nb < 0 )
          continue;
        int first = -1, match = -1;
        for( int t = d->face_offsets[ nb ]; t < d->face_offsets[ nb + 1 ]; ++t )
        {
          if( d->face_neighbor[ t ] != c )
            continue;
          if( first < 0 )
            first = t;
          if( dim == 3 )
            continue;                    // two cells of a conforming 3D mesh share at most one face
          const double *a = d->face_corners + 4 * size_t( s ), *b = d->face_corners + 4 * size_t( t );
          const bool same = a[ 0 ] == b[ 0 ] && a[ 1 ] == b[ 1 ] && a[ 2 ] == b[ 2 ] && a[ 3 ] == b[ 3 ];
          const bool swapped = a[ 0 ] == b[ 2 ] && a[ 1 ] == b[ 3 ] && a[ 2 ] == b[ 0 ] && a[ 3 ] == b[ 1 ];
          if( match < 0 && ( same || swapped ) )
            match = t;
        }
        if( first < 0 )
          return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_create: intersections are not symmetric (outside cell does not see the inside cell)" );
        back[ size_t( s ) ] = ( match >= 0 ) ? match : first;
      }

    if( device >= 0 )
      VOFX_MESH_CUDA( cudaSetDevice( device ) );
    int current = 0;
    VOFX_MESH_CUDA( cudaGetDevice( &current ) );

    vofx_mesh *m = new vofx_mesh;
    m->device = current;
    MeshDev &D = m->dev;
    D.dim = dim;
    D.nCells = n;
    D.nFaces = nFaces;
    D.faceCornerOffsets = nullptr;
    cudaError_t e = cudaSuccess;
    auto up = [ & ] ( auto *&dst, const auto *src, size_t count ) {
      if( e != cudaSuccess )
        return;
      std::remove_const_t< std::remove_reference_t< decltype( *src ) > > *p = nullptr;
      e = upload( p, src, count );
      if( p )
        m->owned.push_back( p );
      dst = p;
    };
    up( D.cornerOffsets, d->corner_offsets, size_t( n ) + 1 );
    up( D.corners, d->corners, size_t( dim ) * size_t( nCorners ) );
    up( D.centers, d->centers, size_t( dim ) * size_t( n ) );
    up( D.volumes, d->volumes, size_t( n ) );
    up( D.faceOffsets, d->face_offsets, size_t( n ) + 1 );
    up( D.faceNeighbor, d->face_neighbor, size_t( nFaces ) );
    up( D.backSlot, back.data(), size_t( nFaces ) );
    if( dim == 3 )
      up( D.faceCornerOffsets, d->face_corner_offsets, size_t( nFaces ) + 1 );
    up( D.faceCorners, d->face_corners, size_t( dim ) * size_t( nFaceCorners ) );
    up( D.faceCenters, d->face_centers, size_t( dim ) * size_t( nFaces ) );
    up( D.faceUnit, d->face_unit_normals, size_t( dim ) * size_t( nFaces ) );
    up( D.faceInt, d->face_int_normals, size_t( dim ) * size_t( nFaces ) );
    up( D.faceVolumes, d->face_volumes, size_t( nFaces ) );
    up( D.stencilOffsets, d->stencil_offsets, size_t( n ) + 1 );
    up( D.stencil, d->stencil, size_t( nStencil ) );
    auto alloc = [ & ] ( auto *&p, size_t count ) {
      if( e == cudaSuccess )
        e = cudaMalloc( &p, ( count ? count : 1 ) * sizeof( *p ) );
    };
    alloc( m->state, 1 );
    alloc( m->mixedList, size_t( n ) );
    alloc( m->recFlux, size_t( nFaces ) );
    alloc( m->recVN, size_t( nFaces ) );
    alloc( m->recDir, size_t( nFaces ) );
    alloc( m->u, size_t( n ) );
    alloc( m->hs, size_t( dim + 1 ) * size_t( n ) );
    alloc( m->upd, size_t( n ) );
    alloc( m->vel, size_t( dim ) * size_t( nFaces ) );
    alloc( m->flags, size_t( n ) );
    if( e == cudaSuccess )
      e = cudaStreamCreateWithFlags( &m->ownStream, cudaStreamNonBlocking );
    m->stream = m->ownStream;
    if( e == cudaSuccess )
      e = cudaMemset( m->state, 0, sizeof( MeshState ) );
    if( e != cudaSuccess )
    {
      const std::string msg = std::string( "vofx_mesh_create: " ) + cudaGetErrorString( e );
      vofx_mesh_destroy( m );
      return set_error( VOFX_ERR_CUDA, msg );
    }
    // grids in multiples of the SM count (148), grid-stride loops
    const long long need = ( (long long) n + 255 ) / 256;
    m->blocksCells = int( need < 148 ? need : ( ( need + 147 ) / 148 < 16 ? ( need + 147 ) / 148 : 16 ) * 148 );
    if( m->blocksCells < 1 )
      m->blocksCells = 1;
    *out = m;
    return VOFX_OK;
  }

  void vofx_mesh_destroy ( vofx_mesh *m )
  {
    if( !m )
      return;
    cudaSetDevice( m->device );
    if( m->stream )
      cudaStreamSynchronize( m->stream );
    if( m->passGraph )
      cudaGraphExecDestroy( m->passGraph );
    if( m->ownStream )
      cudaStreamDestroy( m->ownStream );
    for( void *p : m->owned )
      cudaFree( p );
    cudaFree( m->state );
    cudaFree( m->mixedList );
    cudaFree( m->recFlux );
    cudaFree( m->recVN );
    cudaFree( m->recDir );
    cudaFree( m->u );
    cudaFree( m->hs );
    cudaFree( m->upd );
    cudaFree( m->vel );
    cudaFree( m->flags );
    delete m;
  }

  int vofx_mesh_set_stream ( vofx_mesh *m, void *stream )
  {
    if( int rc = checkMesh( m ) )
      return rc;
    VOFX_MESH_CUDA( cudaStreamSynchronize( m->stream ) );
    m->stream = stream ? static_cast< cudaStream_t >( stream ) : m->ownStream;
    return VOFX_OK;
  }

  int64_t vofx_mesh_cells ( const vofx_mesh *m ) { return m ? m->dev.nCells : 0; }
  int64_t vofx_mesh_faces ( const vofx_mesh *m ) { return m ? m->dev.nFaces : 0; }
  int64_t vofx_mesh_launch_count ( const vofx_mesh *m ) { return m ? m->launches : 0; }

  int vofx_mesh_flag ( vofx_mesh *m, const double *u, double eps, uint8_t *flags )
  {
    if( int rc = checkMesh( m ) )
      return rc;
    if( !u || !flags )
      return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_flag: null argument" );
    return launchFlag( m, u, eps, flags, false );
  }

  int vofx_mesh_reconstruct ( vofx_mesh *m, const double *u, const uint8_t *flags, double *hs )
  {
    if( int rc = checkMesh( m ) )
      return rc;
    if( !u || !flags || !hs )
      return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_reconstruct: null argument" );
    if( int rc = listMixed( m, flags ) )
      return rc;
    return launchYoungs( m, u, hs );
  }

  int vofx_mesh_reconstruct_swartz ( vofx_mesh *m, const double *u, const uint8_t *flags, double *hs, int maxIterations )
  {
    if( int rc = vofx_mesh_reconstruct( m, u, flags, hs ) )      // initializer()( color, reconstructions, flags ), modifiedswartz.hh:72
      return rc;
    if( maxIterations < 0 )
      return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_reconstruct_swartz: max_iterations < 0" );
    return launchSwartz( m, u, flags, hs, maxIterations );
  }

  int vofx_mesh_set_reconstruction ( vofx_mesh *m, int kind, int maxIterations )
  {
    if( int rc = checkMesh( m ) )
      return rc;
    if( ( kind != VOFX_RECON_YOUNGS && kind != VOFX_RECON_SWARTZ ) || maxIterations < 0 )
      return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_set_reconstruction: VOFX_RECON_YOUNGS or VOFX_RECON_SWARTZ (height functions are Cartesian-only)" );
    if( m->passGraph && ( kind != m->reconstruction || maxIterations != m->maxIterations ) )
    {
      cudaGraphExecDestroy( m->passGraph );
      m->passGraph = nullptr;
    }
    m->reconstruction = kind;
    m->maxIterations = maxIterations;
    return VOFX_OK;
  }

  int vofx_mesh_evolve ( vofx_mesh *m, const double *hs, const uint8_t *flags, const double *vel, double dt, double *update, double *dtEst )
  {
    if( int rc = checkMesh( m ) )
      return rc;
    if( !hs || !flags || !vel || !update || !dtEst )
      return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_evolve: null argument" );
    if( int rc = listMixed( m, flags ) )
      return rc;
    k_mesh_set_dt<<< 1, 1, 0, m->stream >>>( m->state, 0.0, dt );
    m->launches += 1;
    if( int rc = launchEvolve< false >( m, hs, flags, vel, &m->state->dt, update ) )
      return rc;
    MeshState host;
    VOFX_MESH_CUDA( cudaMemcpyAsync( &host, m->state, sizeof( MeshState ), cudaMemcpyDeviceToHost, m->stream ) );
    VOFX_MESH_CUDA( cudaStreamSynchronize( m->stream ) );
    std::memcpy( dtEst, &host.dtEstBits, sizeof( double ) );
    return VOFX_OK;
  }

  int vofx_mesh_flag_host ( vofx_mesh *m, const double *u, double eps, uint8_t *flags )
  {
    if( int rc = checkMesh( m ) )
      return rc;
    if( !u || !flags )
      return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_flag_host: null argument" );
    const size_t n = size_t( m->dev.nCells );
    VOFX_MESH_CUDA( cudaMemcpyAsync( m->u, u, n * sizeof( double ), cudaMemcpyHostToDevice, m->stream ) );
    if( int rc = launchFlag( m, m->u, eps, m->flags, false ) )
      return rc;
    VOFX_MESH_CUDA( cudaMemcpyAsync( flags, m->flags, n, cudaMemcpyDeviceToHost, m->stream ) );
    VOFX_MESH_CUDA( cudaStreamSynchronize( m->stream ) );
    m->colorSet = false;
    return VOFX_OK;
  }

  int vofx_mesh_reconstruct_host ( vofx_mesh *m, const double *u, const uint8_t *flags, double *hs )
  {
    if( int rc = checkMesh( m ) )
      return rc;
    if( !u || !flags || !hs )
      return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_reconstruct_host: null argument" );
    const size_t n = size_t( m->dev.nCells );
    VOFX_MESH_CUDA( cudaMemcpyAsync( m->u, u, n * sizeof( double ), cudaMemcpyHostToDevice, m->stream ) );
    VOFX_MESH_CUDA( cudaMemcpyAsync( m->flags, flags, n, cudaMemcpyHostToDevice, m->stream ) );
    if( int rc = vofx_mesh_reconstruct( m, m->u, m->flags, m->hs ) )
      return rc;
    VOFX_MESH_CUDA( cudaMemcpyAsync( hs, m->hs, size_t( m->dev.dim + 1 ) * n * sizeof( double ), cudaMemcpyDeviceToHost, m->stream ) );
    VOFX_MESH_CUDA( cudaStreamSynchronize( m->stream ) );
    m->colorSet = false;
    return VOFX_OK;
  }

  int vofx_mesh_reconstruct_swartz_host ( vofx_mesh *m, const double *u, const uint8_t *flags, double *hs, int maxIterations )
  {
    if( int rc = checkMesh( m ) )
      return rc;
    if( !u || !flags || !hs )
      return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_reconstruct_swartz_host: null argument" );
    const size_t n = size_t( m->dev.nCells );
    VOFX_MESH_CUDA( cudaMemcpyAsync( m->u, u, n * sizeof( double ), cudaMemcpyHostToDevice, m->stream ) );
    VOFX_MESH_CUDA( cudaMemcpyAsync( m->flags, flags, n, cudaMemcpyHostToDevice, m->stream ) );
    if( int rc = vofx_mesh_reconstruct_swartz( m, m->u, m->flags, m->hs, maxIterations ) )
      return rc;
    VOFX_MESH_CUDA( cudaMemcpyAsync( hs, m->hs, size_t( m->dev.dim + 1 ) * n * sizeof( double ), cudaMemcpyDeviceToHost, m->stream ) );
    VOFX_MESH_CUDA( cudaStreamSynchronize( m->stream ) );
    m->colorSet = false;
    return VOFX_OK;
  }

  int vofx_mesh_evolve_host ( vofx_mesh *m, const double *hs, const uint8_t *flags, const double *vel, double dt, double *update, double *dtEst )
  {
    if( int rc = checkMesh( m ) )
      return rc;
    if( !hs || !flags || !vel || !update || !dtEst )
      return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_evolve_host: null argument" );
    const size_t n = size_t( m->dev.nCells ), nf = size_t( m->dev.nFaces );
    VOFX_MESH_CUDA( cudaMemcpyAsync( m->hs, hs, size_t( m->dev.dim + 1 ) * n * sizeof( double ), cudaMemcpyHostToDevice, m->stream ) );
    VOFX_MESH_CUDA( cudaMemcpyAsync( m->flags, flags, n, cudaMemcpyHostToDevice, m->stream ) );
    VOFX_MESH_CUDA( cudaMemcpyAsync( m->vel, vel, size_t( m->dev.dim ) * nf * sizeof( double ), cudaMemcpyHostToDevice, m->stream ) );
    m->velocitySet = false;
    if( int rc = vofx_mesh_evolve( m, m->hs, m->flags, m->vel, dt, m->upd, dtEst ) )
      return rc;
    VOFX_MESH_CUDA( cudaMemcpyAsync( update, m->upd, n * sizeof( double ), cudaMemcpyDeviceToHost, m->stream ) );
    VOFX_MESH_CUDA( cudaStreamSynchronize( m->stream ) );
    return VOFX_OK;
  }

  int vofx_mesh_color_upload ( vofx_mesh *m, const double *u )
  {
    if( int rc = checkMesh( m ) )
      return rc;
    if( !u )
      return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_color_upload: null argument" );
    VOFX_MESH_CUDA( cudaMemcpyAsync( m->u, u, size_t( m->dev.nCells ) * sizeof( double ), cudaMemcpyHostToDevice, m->stream ) );
    VOFX_MESH_CUDA( cudaStreamSynchronize( m->stream ) );
    m->colorSet = true;
    return VOFX_OK;
  }

  int vofx_mesh_velocity_upload ( vofx_mesh *m, const double *vel )
  {
    if( int rc = checkMesh( m ) )
      return rc;
    if( !vel )
      return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_velocity_upload: null argument" );
    VOFX_MESH_CUDA( cudaMemcpyAsync( m->vel, vel, size_t( m->dev.dim ) * size_t( m->dev.nFaces ) * sizeof( double ), cudaMemcpyHostToDevice, m->stream ) );
    VOFX_MESH_CUDA( cudaStreamSynchronize( m->stream ) );
    m->velocitySet = true;
    return VOFX_OK;
  }

  int vofx_mesh_run ( vofx_mesh *m, double eps, double cfl, int passes, double *timeIO, double *dtIO )
  {
    if( int rc = checkMesh( m ) )
      return rc;
    if( !timeIO || !dtIO || passes < 0 )
      return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_run: bad argument" );
    if( !m->colorSet || !m->velocitySet )
      return set_error( VOFX_ERR_STATE, "vofx_mesh_run: call vofx_mesh_color_upload and vofx_mesh_velocity_upload first" );
    k_mesh_set_dt<<< 1, 1, 0, m->stream >>>( m->state, *timeIO, *dtIO );
    m->launches += 1;
    // one loop pass reads only device-resident state: it is captured once into a CUDA graph and replayed
    // (7 small launches per pass are launch-latency bound when issued one by one; VOFX_NO_GRAPH=1 issues them eagerly)
    static const bool noGraph = std::getenv( "VOFX_NO_GRAPH" ) != nullptr;
    if( !noGraph && passes > 1 && ( !m->passGraph || m->graphEps != eps || m->graphCfl != cfl || m->graphStream != m->stream ) )
    {
      if( m->passGraph )
      {
        cudaGraphExecDestroy( m->passGraph );
        m->passGraph = nullptr;
      }
      const long long before = m->launches;
      cudaGraph_t graph = nullptr;
      VOFX_MESH_CUDA( cudaStreamBeginCapture( m->stream, cudaStreamCaptureModeThreadLocal ) );
      int rc = launchFlag( m, m->u, eps, m->flags, true );
      if( !rc )
        rc = launchYoungs( m, m->u, m->hs );
      if( !rc && m->reconstruction == VOFX_RECON_SWARTZ )
        rc = launchSwartz( m, m->u, m->flags, m->hs, m->maxIterations );
      if( !rc )
        rc = launchEvolve< true >( m, m->hs, m->flags, m->vel, &m->state->dt, m->u );
      k_mesh_finish<<< 1, 1, 0, m->stream >>>( m->state, cfl );
      const cudaError_t endErr = cudaStreamEndCapture( m->stream, &graph );
      m->launchesPerPass = int( m->launches - before ) + 1;
      m->launches = before;
      if( rc )
        return rc;
      VOFX_MESH_CUDA( endErr );
      VOFX_MESH_CUDA( cudaGraphInstantiate( &m->passGraph, graph, 0 ) );
      cudaGraphDestroy( graph );
      m->graphEps = eps;
      m->graphCfl = cfl;
      m->graphStream = m->stream;
    }
    for( int pass = 0; pass < passes; ++pass )
    {
      if( m->passGraph && !noGraph && passes > 1 )
      {
        VOFX_MESH_CUDA( cudaGraphLaunch( m->passGraph, m->stream ) );
        m->launches += m->launchesPerPass;
        continue;
      }
      if( int rc = launchFlag( m, m->u, eps, m->flags, true ) )
        return rc;
      if( int rc = launchYoungs( m, m->u, m->hs ) )
        return rc;
      if( m->reconstruction == VOFX_RECON_SWARTZ )
        if( int rc = launchSwartz( m, m->u, m->flags, m->hs, m->maxIterations ) )
          return rc;
      if( int rc = launchEvolve< true >( m, m->hs, m->flags, m->vel, &m->state->dt, m->u ) )
        return rc;
      k_mesh_finish<<< 1, 1, 0, m->stream >>>( m->state, cfl );
      m->launches += 1;
    }
    MeshState host;
    VOFX_MESH_CUDA( cudaMemcpyAsync( &host, m->state, sizeof( MeshState ), cudaMemcpyDeviceToHost, m->stream ) );
    VOFX_MESH_CUDA( cudaStreamSynchronize( m->stream ) );
    *timeIO = host.time;
    *dtIO = host.dt;
    return VOFX_OK;
  }

  int vofx_mesh_color_download ( vofx_mesh *m, double *u, uint8_t *flags )
  {
    if( int rc = checkMesh( m ) )
      return rc;
    if( !u )
      return set_error( VOFX_ERR_ARGUMENT, "vofx_mesh_color_download: null argument" );
    if( !m->colorSet )
      return set_error( VOFX_ERR_STATE, "vofx_mesh_color_download: no device-resident colour function" );
    VOFX_MESH_CUDA( cudaMemcpyAsync( u, m->u, size_t( m->dev.nCells ) * sizeof( double ), cudaMemcpyDeviceToHost, m->stream ) );
    if( flags )
      VOFX_MESH_CUDA( cudaMemcpyAsync( flags, m->flags, size_t( m->dev.nCells ), cudaMemcpyDeviceToHost, m->stream ) );
    VOFX_MESH_CUDA( cudaStreamSynchronize( m->stream ) );
    return VOFX_OK;
  }

} // extern "C"
