// vofx: CUDA kernels of the per-timestep VoF hot path (sm_100a).
//
// Step structure (DESIGN.md "Kernels"):
//   k_flag_compact   dense, HBM-bound: u -> flags (1 B/cell) + compacted list of mixed cells         (F1)
//   k_youngs         band: least-squares normal + plane-constant solve per mixed cell                 (R1, G1/G2)
//   k_hf / k_swartz  band: height-function / iterative Swartz normals on top of Young's                (R2, R3)
//   k_flux           band: one lane per (mixed cell, face): face velocity, swept-box clip, flux record (E1, G5, G7)
//   k_update         band: owner-computes gather of the flux records in the reference's summation order,
//                    u += du in place (the dense update.clear()/axpy passes of the reference vanish)   (E1, U1)
//   k_finalize       one thread: time += dt, dt = cfl * dtEst, reset counters                          (T1)
// No tensor cores: nothing here is a contraction.  Arithmetic order follows the reference (vofx_math.cuh).
#pragma once

#include <cuda_runtime.h>

#include "vofx_geom2d.cuh"
#include "vofx_geom3d.cuh"
#include "vofx_coop3d.cuh"
#include "vofx_grid.cuh"
#include "vofx_types.cuh"

namespace vofx
{



  // ------------------------------------------------------------------------------------------------------------
  // F1: flags + compaction.  Each thread owns 4 consecutive cells (two 16-byte loads, one 4-byte flag store).
  // ------------------------------------------------------------------------------------------------------------

  template< int DIM >
  __device__ __forceinline__ unsigned char flag_of ( const Grid &g, const double *__restrict__ u, double eps, double colorEn, const int *ijk )
  {
    // flagging.hh:46-64
    if( colorEn < eps )
      return 0;
    if( colorEn <= ( 1 - eps ) )
      return 1;
    if( colorEn != colorEn )
      return 99;
    unsigned char flag = 3;
#pragma unroll
    for( int face = 0; face < 2 * DIM; ++face )
    {
      int nb[ 3 ] = { ijk[ 0 ], ijk[ 1 ], ijk[ 2 ] };
      nb[ face >> 1 ] += ( face & 1 ) ? 1 : -1;
      if( cell_exists( g, nb ) && __ldg( u + cell_index( g, nb ) ) < eps )
      {
        flag = 2;
        break;
      }
    }
    return flag;
  }

  template< int DIM >
  __global__ void __launch_bounds__( 256 ) k_flag_compact ( Grid g, const double *__restrict__ u, double eps, unsigned char *__restrict__ flags,
                                                             int *__restrict__ mixedList, int *__restrict__ slot, int capacity, StepState *state, int compact )
  {
    const long long nQuads = ( g.cells + 3 ) >> 2;
    const int lane = threadIdx.x & 31;
    for( long long q = blockIdx.x * (long long) blockDim.x + threadIdx.x; q < ( ( nQuads + 31 ) & ~31LL ); q += (long long) gridDim.x * blockDim.x )
    {
      const long long c0 = q << 2;
      double uu[ 4 ] = { 0.0, 0.0, 0.0, 0.0 };
      unsigned char ff[ 4 ] = { 0, 0, 0, 0 };
      int nMixed = 0;
      unsigned mixedBits = 0;
      const bool full4 = ( q < nQuads ) && ( c0 + 3 < g.cells );
      if( full4 )
      {
        const double2 a = __ldg( reinterpret_cast< const double2 * >( u + c0 ) );
        const double2 b = __ldg( reinterpret_cast< const double2 * >( u + c0 + 2 ) );
        uu[ 0 ] = a.x; uu[ 1 ] = a.y; uu[ 2 ] = b.x; uu[ 3 ] = b.y;
      }
      else if( q < nQuads )
        for( int k = 0; k < 4; ++k )
          if( c0 + k < g.cells )
            uu[ k ] = u[ c0 + k ];

      if( q < nQuads )
      {
        int ijk[ 3 ];
        cell_ijk( g, c0, ijk );
#pragma unroll
        for( int k = 0; k < 4; ++k )
        {
          if( c0 + k < g.cells )
          {
            ff[ k ] = flag_of< DIM >( g, u, eps, uu[ k ], ijk );
            const int la = ijk[ DIM - 1 ];
            if( is_mixed( ff[ k ] ) && la >= g.list0 && la < g.list1 && in_domain( g, ijk ) )
            {
              mixedBits |= 1u << k;
              nMixed++;
            }
          }
          // advance the multi-index with carry
          if( ++ijk[ 0 ] == g.ns[ 0 ] )
          {
            ijk[ 0 ] = 0;
            if( ++ijk[ 1 ] == g.ns[ 1 ] )
            {
              ijk[ 1 ] = 0;
              ++ijk[ 2 ];
            }
          }
        }
        if( full4 )
          *reinterpret_cast< uchar4 * >( flags + c0 ) = make_uchar4( ff[ 0 ], ff[ 1 ], ff[ 2 ], ff[ 3 ] );
        else
          for( int k = 0; k < 4; ++k )
            if( c0 + k < g.cells )
              flags[ c0 + k ] = ff[ k ];
      }

      if( compact )
      {
        // warp-aggregated append: one atomic per warp that holds any mixed cell
        int incl = nMixed;
#pragma unroll
        for( int o = 1; o < 32; o <<= 1 )
        {
          const int t = __shfl_up_sync( 0xffffffffu, incl, o );
          if( lane >= o )
            incl += t;
        }
        const int total = __shfl_sync( 0xffffffffu, incl, 31 );
        if( total > 0 )
        {
          int base = 0;
          if( lane == 31 )
            base = atomicAdd( &state->mixedCount, total );
          base = __shfl_sync( 0xffffffffu, base, 31 );
          int pos = base + incl - nMixed;
          for( int k = 0; k < 4; ++k )
            if( mixedBits & ( 1u << k ) )
            {
              if( pos < capacity )
              {
                mixedList[ pos ] = int( c0 + k );
                slot[ c0 + k ] = pos;
              }
              else
                state->overflow = 1;
              pos++;
            }
        }
      }
    }
  }

  // Fast path of F1 for row lengths that are a multiple of 4 (every production grid): one block walks `rowsPerBlock`
  // consecutive x-rows, a thread owns 4 consecutive cells of a row (two 16-byte loads, one 4-byte flag store), all
  // index arithmetic is 32-bit and row-uniform, and the warp-aggregated append only runs in warps that hold a mixed
  // cell (about 1 in 100 on the BASELINE configs).  ~11 instructions per cell: the kernel is HBM-bound, not issue-bound.
  template< int DIM >
  __global__ void __launch_bounds__( 128 ) k_flag_rows ( Grid g, const double *__restrict__ u, double eps, unsigned char *__restrict__ flags,
                                                          int *__restrict__ mixedList, int *__restrict__ slot, int capacity, StepState *state,
                                                          int compact, int rowsPerBlock, int rowBegin, int rowEnd )
  {
    pdl_prologue();
    const int nx = g.ns[ 0 ];
    const int ny = g.ns[ 1 ];
    const int nRows = rowEnd;            // rows [ rowBegin, rowEnd ) of the ns[1] * ns[2] rows of the storage box
    const int quadsPerRow = nx >> 2;
    const int quadsRounded = ( quadsPerRow + 31 ) & ~31;
    const int lane = threadIdx.x & 31;
    const double oneMinusEps = 1 - eps;
    const int planeStride = nx * ny;

    for( int r = 0; r < rowsPerBlock; ++r )
    {
      const int row = rowBegin + blockIdx.x * rowsPerBlock + r;
      if( row >= nRows )
        break;
      const int j = ( DIM == 3 ) ? row % ny : row;
      const int k = ( DIM == 3 ) ? row / ny : 0;
      const int la = ( DIM == 3 ) ? k : j;                       // index along the decomposed (last) axis
      const int gla = la + g.g0[ DIM - 1 ];
      const bool inList = la >= g.list0 && la < g.list1;
      // neighbour rows that exist (inside the global domain and inside this context's storage)
      const bool lastLo = la > 0 && gla > 0;
      const bool lastHi = la + 1 < g.ns[ DIM - 1 ] && gla + 1 < g.ng[ DIM - 1 ];
      const bool yLo = ( DIM == 3 ) ? ( j > 0 ) : lastLo;
      const bool yHi = ( DIM == 3 ) ? ( j + 1 < ny ) : lastHi;
      const bool zLo = ( DIM == 3 ) && lastLo;
      const bool zHi = ( DIM == 3 ) && lastHi;
      const int rowBase = row * nx;

      for( int q = threadIdx.x; q < quadsRounded; q += blockDim.x )
      {
        int nMixed = 0;
        unsigned mixedBits = 0;
        const int i0 = q << 2;
        const int c0 = rowBase + i0;
        if( q < quadsPerRow )
        {
          const double2 a = __ldg( reinterpret_cast< const double2 * >( u + c0 ) );
          const double2 b = __ldg( reinterpret_cast< const double2 * >( u + c0 + 2 ) );
          const double uu[ 4 ] = { a.x, a.y, b.x, b.y };
          unsigned ff[ 4 ];
          bool anyFull = false;
#pragma unroll
          for( int t = 0; t < 4; ++t )
          {
            // flagging.hh:46-55
            ff[ t ] = ( uu[ t ] < eps ) ? 0u : ( ( uu[ t ] <= oneMinusEps ) ? 1u : ( ( uu[ t ] != uu[ t ] ) ? 99u : 3u ) );
            anyFull |= ( ff[ t ] == 3u );
          }
          if( anyFull )
          {
            // full cells with an empty face neighbour become mixedfull, flagging.hh:56-63.  The quads of the neighbour rows
            // are fetched with the same 16-byte loads as the row itself (inside the fluid every cell of the quad is full:
            // per-cell 8-byte loads would touch every sector four times), the x neighbours come from the registers
            bool emptyRow[ 4 ] = { false, false, false, false };
            auto scanRow = [ & ] ( int offset ) {
              const double2 p = __ldg( reinterpret_cast< const double2 * >( u + c0 + offset ) );
              const double2 q2 = __ldg( reinterpret_cast< const double2 * >( u + c0 + offset + 2 ) );
              emptyRow[ 0 ] |= p.x < eps;
              emptyRow[ 1 ] |= p.y < eps;
              emptyRow[ 2 ] |= q2.x < eps;
              emptyRow[ 3 ] |= q2.y < eps;
            };
            if( yLo )
              scanRow( -nx );
            if( yHi )
              scanRow( nx );
            if( DIM == 3 )
            {
              if( zLo )
                scanRow( -planeStride );
              if( zHi )
                scanRow( planeStride );
            }
#pragma unroll
            for( int t = 0; t < 4; ++t )
              if( ff[ t ] == 3u )
              {
                const int i = i0 + t;
                const int c = c0 + t;
                bool emptyNb = emptyRow[ t ];
                if( t > 0 )
                  emptyNb |= uu[ t > 0 ? t - 1 : 0 ] < eps;
                else if( i > 0 )
                  emptyNb |= __ldg( u + c - 1 ) < eps;
                if( t < 3 )
                  emptyNb |= uu[ t < 3 ? t + 1 : 3 ] < eps;
                else if( i + 1 < nx )
                  emptyNb |= __ldg( u + c + 1 ) < eps;
                if( emptyNb )
                  ff[ t ] = 2u;
              }
          }
          *reinterpret_cast< unsigned * >( flags + c0 ) = ff[ 0 ] | ( ff[ 1 ] << 8 ) | ( ff[ 2 ] << 16 ) | ( ff[ 3 ] << 24 );
          if( inList )
          {
#pragma unroll
            for( int t = 0; t < 4; ++t )
              if( ff[ t ] == 1u || ff[ t ] == 2u )
              {
                mixedBits |= 1u << t;
                nMixed++;
              }
          }
        }

        if( compact && __any_sync( 0xffffffffu, nMixed > 0 ) )
        {
          int incl = nMixed;
#pragma unroll
          for( int o = 1; o < 32; o <<= 1 )
          {
            const int t = __shfl_up_sync( 0xffffffffu, incl, o );
            if( lane >= o )
              incl += t;
          }
          const int total = __shfl_sync( 0xffffffffu, incl, 31 );
          int base = 0;
          if( lane == 31 )
            base = atomicAdd( &state->mixedCount, total );
          base = __shfl_sync( 0xffffffffu, base, 31 );
          int pos = base + incl - nMixed;
          for( int t = 0; t < 4; ++t )
            if( mixedBits & ( 1u << t ) )
            {
              if( pos < capacity )
              {
                mixedList[ pos ] = c0 + t;
                slot[ c0 + t ] = pos;
              }
              else
                state->overflow = 1;
              pos++;
            }
        }
      }
    }
  }

  // mixed list from given flags (operator-level API: the caller supplies the flags)
  template< int DIM >
  __global__ void __launch_bounds__( 256 ) k_compact_flags ( Grid g, const unsigned char *__restrict__ flags, int *__restrict__ mixedList,
                                                              int *__restrict__ slot, int capacity, StepState *state )
  {
    const int lane = threadIdx.x & 31;
    const long long nRound = ( g.cells + 31 ) & ~31LL;
    for( long long c = blockIdx.x * (long long) blockDim.x + threadIdx.x; c < nRound; c += (long long) gridDim.x * blockDim.x )
    {
      bool m = false;
      if( c < g.cells && is_mixed( flags[ c ] ) )
      {
        int ijk[ 3 ];
        cell_ijk( g, c, ijk );
        const int la = ijk[ DIM - 1 ];
        m = la >= g.list0 && la < g.list1 && in_domain( g, ijk );
      }
      const unsigned ballot = __ballot_sync( 0xffffffffu, m );
      if( ballot )
      {
        int base = 0;
        if( lane == 0 )
          base = atomicAdd( &state->mixedCount, __popc( ballot ) );
        base = __shfl_sync( 0xffffffffu, base, 0 );
        if( m )
        {
          const int pos = base + __popc( ballot & ( ( 1u << lane ) - 1 ) );
          if( pos < capacity )
          {
            mixedList[ pos ] = int( c );
            slot[ c ] = pos;
          }
          else
            state->overflow = 1;
        }
      }
    }
  }

  // number of entries of the band list; 0 once the list has overflowed: the band kernels, the update and the time
  // advance of that pass become no-ops and the host sees VOFX_ERR_STATE at its next read of the loop state
  __device__ __forceinline__ int mixed_count ( const StepState *state, int capacity )
  {
    const int n = state->mixedCount;
    if( state->overflow )
      return 0;
    return n < capacity ? n : capacity;
  }

  // ------------------------------------------------------------------------------------------------------------
  // F1, incremental form for a colour function that only the loop itself changes (resident loop, graph replay).
  // The update of the previous pass wrote u only at the band cells and their face neighbours; a flag depends on u
  // of the cell and of its face neighbours (flagging.hh:46-63).  So only cells within two face steps of a cell of
  // the previous band can change their flag, and only they can be mixed now.  Rows are cut into aligned segments of 32
  // cells; k_mark_band marks (one byte per segment, plain idempotent stores: no atomics) every segment that holds such
  // a cell, k_flag_segments re-flags the marked segments with one warp per segment (coalesced 256-byte loads of the
  // segment and of its four neighbour rows) and rebuilds the band list from them.  All other flags stay what the last
  // dense pass / last re-flag wrote.  Layers outside [ layer0, layer1 ) are left to the dense kernel (halo layers,
  // whose colour values arrive from another rank).
  // ------------------------------------------------------------------------------------------------------------
  // marks of one band cell: the rows ( j + dj, k + dk ), |dj| + |dk| <= 2, hold cells within two face steps of it at
  // i - w .. i + w, w = 2 - |dj| - |dk|.  Called by the lanes of the cell's group in k_update (row r by lane r % nLanes).
  template< int DIM >
  __device__ __forceinline__ void mark_band_cell ( const Grid &g, const int *ijk, int lane, int nLanes, unsigned char *__restrict__ marks, int segsPerRow )
  {
    constexpr int NROWS = ( DIM == 3 ) ? 13 : 5;
    for( int r = lane; r < NROWS; r += nLanes )
    {
      int dj = 0, dk = 0;
      if( DIM == 3 )
      {
        // 0: (0,0); 1-4: (+-1,0),(0,+-1); 5-8: (+-2,0),(0,+-2); 9-12: (+-1,+-1)
        if( r >= 1 && r <= 8 )
        {
          const int q = ( r - 1 ) & 3, len = ( r <= 4 ) ? 1 : 2;
          dj = ( q == 0 ) ? -len : ( q == 1 ? len : 0 );
          dk = ( q == 2 ) ? -len : ( q == 3 ? len : 0 );
        }
        else if( r >= 9 )
        {
          dj = ( ( r - 9 ) & 1 ) ? 1 : -1;
          dk = ( ( r - 9 ) & 2 ) ? 1 : -1;
        }
      }
      else
        dj = r - 2;
      const int j = ijk[ 1 ] + dj, k = ijk[ 2 ] + dk;
      if( j < 0 || j >= g.ns[ 1 ] || k < 0 || k >= g.ns[ 2 ] )
        continue;
      const int w = 2 - ( dj < 0 ? -dj : dj ) - ( dk < 0 ? -dk : dk );
      const int i0 = ( ijk[ 0 ] - w > 0 ) ? ijk[ 0 ] - w : 0;
      const int i1 = ( ijk[ 0 ] + w < g.ns[ 0 ] - 1 ) ? ijk[ 0 ] + w : g.ns[ 0 ] - 1;
      const long long row = j + (long long) g.ns[ 1 ] * k;
      marks[ row * segsPerRow + ( i0 >> 5 ) ] = 1;
      if( ( i1 >> 5 ) != ( i0 >> 5 ) )
        marks[ row * segsPerRow + ( i1 >> 5 ) ] = 1;
    }
  }

  // the band of the pass in flight marks the segments the NEXT pass has to re-flag.  Runs on a side branch of the pass (it only
  // reads the band list), next to the reconstruction and the fluxes, and is followed there by k_scan_marks.
  template< int DIM >
  __global__ void __launch_bounds__( 256 ) k_mark_band ( Grid g, const int *__restrict__ mixedList, StepState *state, int capacity,
                                                          unsigned char *__restrict__ marks, int segsPerRow )
  {
    constexpr int LANES = 8;
    if( blockIdx.x == 0 && threadIdx.x == 0 )
      state->segCount = 0;                      // k_scan_marks (next on this branch) counts from zero
    const int count = mixed_count( state, capacity );
    const int lane = threadIdx.x % LANES;
    const int groupsPerBlock = blockDim.x / LANES;
    for( int idx = blockIdx.x * groupsPerBlock + threadIdx.x / LANES; idx < count; idx += gridDim.x * groupsPerBlock )
    {
      int ijk[ 3 ];
      cell_ijk( g, mixedList[ idx ], ijk );
      mark_band_cell< DIM >( g, ijk, lane, LANES, marks, segsPerRow );
    }
  }

  // compaction of the marks into a list of segment ids (and reset of the marks): a warp that worked through its own marked
  // segments one after the other was bound by a handful of warps whose 32 marks all lay on the sphere.  A thread owns 16
  // marks (one 16-byte load; the array is padded to a multiple of 16 zero bytes).
  template< int DIM >
  __global__ void __launch_bounds__( 256 ) k_scan_marks ( Grid g, unsigned char *__restrict__ marks, int segsPerRow, long long nSegments,
                                                           int layer0, int layer1, int *__restrict__ segList, StepState *state )
  {
    pdl_prologue();
    __shared__ int warpCount[ 8 ], blockBase;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long nChunks = ( nSegments + 15 ) >> 4;
    const long long rounded = ( nChunks + 255 ) & ~255LL;
    const int ny = g.ns[ 1 ];
    for( long long first = blockIdx.x * 256LL; first < rounded; first += gridDim.x * 256LL )
    {
      const long long chunk = first + threadIdx.x;
      unsigned bits = 0;                 // bit b: segment 16 * chunk + b is marked and inside the layer range
      if( chunk < nChunks )
      {
        const uint4 m = *reinterpret_cast< const uint4 * >( marks + 16 * chunk );
        if( m.x | m.y | m.z | m.w )
        {
          *reinterpret_cast< uint4 * >( marks + 16 * chunk ) = make_uint4( 0u, 0u, 0u, 0u );
          const unsigned words[ 4 ] = { m.x, m.y, m.z, m.w };
#pragma unroll
          for( int b = 0; b < 16; ++b )
            if( ( words[ b >> 2 ] >> ( 8 * ( b & 3 ) ) ) & 0xffu )
            {
              const long long row = ( 16 * chunk + b ) / segsPerRow;
              const int la = ( DIM == 3 ) ? int( row / ny ) : int( row );
              if( la >= layer0 && la < layer1 )
                bits |= 1u << b;
            }
        }
      }
      const int mine = __popc( bits );
      if( !__syncthreads_or( mine ) )
        continue;
      // exclusive prefix of `mine` over the block: warp scan, then one global atomic per block iteration
      int incl = mine;
#pragma unroll
      for( int o = 1; o < 32; o <<= 1 )
      {
        const int t = __shfl_up_sync( 0xffffffffu, incl, o );
        if( lane >= o )
          incl += t;
      }
      if( lane == 31 )
        warpCount[ warp ] = incl;
      __syncthreads();
      if( threadIdx.x == 0 )
      {
        int n = 0;
        for( int w = 0; w < 8; ++w )
        {
          const int c = warpCount[ w ];
          warpCount[ w ] = n;
          n += c;
        }
        blockBase = atomicAdd( &state->segCount, n );
      }
      __syncthreads();
      int pos = blockBase + warpCount[ warp ] + incl - mine;
      while( bits )
      {
        const int b = __ffs( bits ) - 1;
        bits &= bits - 1;
        segList[ pos++ ] = int( 16 * chunk + b );
      }
      __syncthreads();
    }
  }

#ifndef VOFX_FLAGSEG_BLOCKS
#define VOFX_FLAGSEG_BLOCKS 4     // resident blocks per SM the register allocation aims for (92 registers allowed only 2)
#endif
  template< int DIM >
  __global__ void __launch_bounds__( 256, VOFX_FLAGSEG_BLOCKS ) k_flag_segments ( Grid g, const double *__restrict__ u, double eps, unsigned char *__restrict__ flags,
                                                              int *__restrict__ mixedList, int *__restrict__ slot, int capacity, StepState *state,
                                                              const int *__restrict__ segList, int segsPerRow )
  {
    pdl_prologue();
    constexpr int kBlockList = 2048;
    constexpr int U = 4;               // segments in flight per warp: the kernel is bound by the latency of one round trip per segment
    __shared__ int sList[ kBlockList ];
    __shared__ int sCount, sBase;
    if( threadIdx.x == 0 )
      sCount = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = int( ( blockIdx.x * (long long) blockDim.x + threadIdx.x ) >> 5 );
    const int nWarps = int( ( gridDim.x * (long long) blockDim.x ) >> 5 );
    const int nx = g.ns[ 0 ], ny = g.ns[ 1 ];
    const long long planeStride = (long long) nx * ny;
    const double oneMinusEps = 1 - eps;
    const double big = 1.0;            // "not empty": a missing neighbour never makes a cell mixedfull
    const int nSegs = state->segCount;
    for( int q0 = warp * U; q0 < nSegs; q0 += nWarps * U )
    {
      long long c[ U ];
      int la[ U ];
      bool valid[ U ];
      double uc[ U ], uyl[ U ], uyh[ U ], uzl[ U ], uzh[ U ], uxl[ U ], uxh[ U ];
#pragma unroll
      for( int r = 0; r < U; ++r )
      {
        const bool have = q0 + r < nSegs;
        const int seg = have ? segList[ q0 + r ] : 0;
        const int row = seg / segsPerRow;
        const int s = seg - row * segsPerRow;
        const int j = ( DIM == 3 ) ? row % ny : row;
        const int k = ( DIM == 3 ) ? row / ny : 0;
        la[ r ] = ( DIM == 3 ) ? k : j;
        const int gla = la[ r ] + g.g0[ DIM - 1 ];
        const bool lastLo = la[ r ] > 0 && gla > 0;
        const bool lastHi = la[ r ] + 1 < g.ns[ DIM - 1 ] && gla + 1 < g.ng[ DIM - 1 ];
        const bool yLo = ( DIM == 3 ) ? ( j > 0 ) : lastLo;
        const bool yHi = ( DIM == 3 ) ? ( j + 1 < ny ) : lastHi;
        const bool zLo = ( DIM == 3 ) && lastLo;
        const bool zHi = ( DIM == 3 ) && lastHi;
        const int i = ( s << 5 ) + lane;
        valid[ r ] = have && i < nx;
        c[ r ] = (long long) row * nx + i;
        // the loads of the U segments are independent: one round trip for all of them
        uc[ r ] = valid[ r ] ? __ldg( u + c[ r ] ) : big;
        uxl[ r ] = ( valid[ r ] && lane == 0 && i > 0 ) ? __ldg( u + c[ r ] - 1 ) : big;
        uxh[ r ] = ( valid[ r ] && lane == 31 && i + 1 < nx ) ? __ldg( u + c[ r ] + 1 ) : big;
        // the neighbour rows are read unconditionally: loading them only for segments that hold a full cell costs a second
        // round trip, and the kernel is bound by round trips, not by bytes (measured: 24.6 vs 31.3 us)
        uyl[ r ] = ( valid[ r ] && yLo ) ? __ldg( u + c[ r ] - nx ) : big;
        uyh[ r ] = ( valid[ r ] && yHi ) ? __ldg( u + c[ r ] + nx ) : big;
        uzl[ r ] = ( valid[ r ] && zLo ) ? __ldg( u + c[ r ] - planeStride ) : big;
        uzh[ r ] = ( valid[ r ] && zHi ) ? __ldg( u + c[ r ] + planeStride ) : big;
      }
#pragma unroll
      for( int r = 0; r < U; ++r )
      {
        const double fromLo = __shfl_up_sync( 0xffffffffu, uc[ r ], 1 );
        const double fromHi = __shfl_down_sync( 0xffffffffu, uc[ r ], 1 );
        if( lane > 0 )
          uxl[ r ] = fromLo;
        if( lane < 31 )
          uxh[ r ] = fromHi;      // an invalid upper neighbour (beyond the row) carries `big`
        // flagging.hh:46-63
        unsigned f = ( uc[ r ] < eps ) ? 0u : ( ( uc[ r ] <= oneMinusEps ) ? 1u : ( ( uc[ r ] != uc[ r ] ) ? 99u : 3u ) );
        if( f == 3u && ( uxl[ r ] < eps || uxh[ r ] < eps || uyl[ r ] < eps || uyh[ r ] < eps || uzl[ r ] < eps || uzh[ r ] < eps ) )
          f = 2u;
        if( valid[ r ] )
          flags[ c[ r ] ] = (unsigned char) f;
        const bool mixedHere = valid[ r ] && ( f == 1u || f == 2u ) && la[ r ] >= g.list0 && la[ r ] < g.list1;
        const unsigned ballot = __ballot_sync( 0xffffffffu, mixedHere );
        if( ballot )
        {
          // the block collects its mixed cells in shared memory and appends them with ONE global atomic at the end: a
          // counter shared by every segment of the band (3e4 atomics with a return value on one address) serialises
          int pos0 = 0;
          if( lane == 0 )
            pos0 = atomicAdd( &sCount, __popc( ballot ) );
          pos0 = __shfl_sync( 0xffffffffu, pos0, 0 );
          const int p = pos0 + __popc( ballot & ( ( 1u << lane ) - 1 ) );
          if( mixedHere && p < kBlockList )
            sList[ p ] = int( c[ r ] );
          const bool spill = mixedHere && p >= kBlockList;
          const unsigned spillBallot = __ballot_sync( 0xffffffffu, spill );
          if( spillBallot )
          {
            int g0 = 0;
            const int leader = __ffs( spillBallot ) - 1;
            if( lane == leader )
              g0 = atomicAdd( &state->mixedCount, __popc( spillBallot ) );
            g0 = __shfl_sync( 0xffffffffu, g0, leader );
            if( spill )
            {
              const int pos = g0 + __popc( spillBallot & ( ( 1u << lane ) - 1 ) );
              if( pos < capacity )
              {
                mixedList[ pos ] = int( c[ r ] );
                slot[ c[ r ] ] = pos;
              }
              else
                state->overflow = 1;
            }
          }
        }
      }
    }
    __syncthreads();
    const int nMine = ( sCount < kBlockList ) ? sCount : kBlockList;
    if( threadIdx.x == 0 )
      sBase = nMine ? atomicAdd( &state->mixedCount, nMine ) : 0;
    __syncthreads();
    for( int t = threadIdx.x; t < nMine; t += blockDim.x )
    {
      const int pos = sBase + t;
      const int cc = sList[ t ];
      if( pos < capacity )
      {
        mixedList[ pos ] = cc;
        slot[ cc ] = pos;
      }
      else
        state->overflow = 1;
    }
  }

  // ------------------------------------------------------------------------------------------------------------
  // plane-constant solve for the cell ijk with a given unit normal: locateHalfSpace( makePolytope( geometry ), n, u )
  // ------------------------------------------------------------------------------------------------------------

  template< int DIM >
  __device__ __forceinline__ double locate_cell ( const Grid &g, const int *ijk, const double *normal, double fraction )
  {
    if( DIM == 2 )
    {
      Poly2 p;
      make_cell( cell_lower( g, 0, ijk[ 0 ] ), cell_lower( g, 1, ijk[ 1 ] ), g.h[ 0 ], g.h[ 1 ], p );
      return locate( p, normal[ 0 ], normal[ 1 ], fraction );
    }
    const double lo[ 3 ] = { cell_lower( g, 0, ijk[ 0 ] ), cell_lower( g, 1, ijk[ 1 ] ), cell_lower( g, 2, ijk[ 2 ] ) };
    Hex c;
    make_cell( lo, g.h, c );
    return locate( c, normal, fraction );
  }

  // ------------------------------------------------------------------------------------------------------------
  // R1: modified Young's, reconstruction/modifiedyoungs.hh:90-134.  One thread per mixed cell.
  // ------------------------------------------------------------------------------------------------------------

  template< int DIM >
  __device__ __forceinline__ void ls_accumulate ( double *d, double colorDiff, double AtA[ 3 ][ 3 ], double *Atb )
  {
    double n2 = 0.0;
#pragma unroll
    for( int k = 0; k < DIM; ++k )
      n2 += d[ k ] * d[ k ];
    const double weight = 1.0 / n2;
#pragma unroll
    for( int k = 0; k < DIM; ++k )
      d[ k ] *= weight;
#pragma unroll
    for( int i = 0; i < DIM; ++i )
#pragma unroll
      for( int j = 0; j < DIM; ++j )
      {
        double m = 0.0;          // outerProduct: axpy onto a zero matrix, geometry/utility.hh:161-169
        m += d[ i ] * d[ j ];
        AtA[ i ][ j ] += m;
      }
    const double w = weight * colorDiff;
#pragma unroll
    for( int k = 0; k < DIM; ++k )
      Atb[ k ] += w * d[ k ];
  }

  // FieldMatrix::solve closed forms (dune-common 2.6 densematrix.hh; third party, see DESIGN.md "parity unpinned")
  template< int DIM >
  __device__ __forceinline__ void solve_small ( const double M[ 3 ][ 3 ], const double *b, double *x )
  {
    if( DIM == 2 )
    {
      double detinv = M[ 0 ][ 0 ] * M[ 1 ][ 1 ] - M[ 0 ][ 1 ] * M[ 1 ][ 0 ];
      detinv = 1.0 / detinv;
      x[ 0 ] = detinv * ( M[ 1 ][ 1 ] * b[ 0 ] - M[ 0 ][ 1 ] * b[ 1 ] );
      x[ 1 ] = detinv * ( M[ 0 ][ 0 ] * b[ 1 ] - M[ 1 ][ 0 ] * b[ 0 ] );
      x[ 2 ] = 0.0;
      return;
    }
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

  // returns false if the normal degenerates (the reconstruction stays the zero half-space, :128-129)
  template< int DIM >
  __device__ bool youngs_normal ( const Grid &g, const double *__restrict__ u, const int *ijk, double colorEn, double *normal )
  {
    double center[ 3 ] = { 0.0, 0.0, 0.0 }, lower[ 3 ] = { 0.0, 0.0, 0.0 };
#pragma unroll
    for( int d = 0; d < DIM; ++d )
    {
      lower[ d ] = cell_lower( g, d, ijk[ d ] );
      center[ d ] = lower[ d ] + 0.5 * g.h[ d ];
    }
    double AtA[ 3 ][ 3 ] = { { 0.0, 0.0, 0.0 }, { 0.0, 0.0, 0.0 }, { 0.0, 0.0, 0.0 } };
    double Atb[ 3 ] = { 0.0, 0.0, 0.0 };

    // vertex-neighbour stencil in ascending cell index, self excluded (stencil/vertexneighborsstencil.hh:52-94)
    for( int dz = ( DIM == 3 ? -1 : 0 ); dz <= ( DIM == 3 ? 1 : 0 ); ++dz )
      for( int dy = -1; dy <= 1; ++dy )
        for( int dx = -1; dx <= 1; ++dx )
        {
          if( dx == 0 && dy == 0 && dz == 0 )
            continue;
          const int nb[ 3 ] = { ijk[ 0 ] + dx, ijk[ 1 ] + dy, ijk[ 2 ] + dz };
          if( !cell_exists( g, nb ) )
            continue;
          const double colorNb = clamp01( __ldg( u + cell_index( g, nb ) ) );
          double d[ 3 ];
#pragma unroll
          for( int k = 0; k < DIM; ++k )
            d[ k ] = ( cell_lower( g, k, nb[ k ] ) + 0.5 * g.h[ k ] ) - center[ k ];
          ls_accumulate< DIM >( d, colorNb - colorEn, AtA, Atb );
        }

    // domain-boundary faces get a ghost with colour 0.5 at twice the face-centre distance, :112-124
    for( int face = 0; face < 2 * DIM; ++face )
    {
      int nb[ 3 ] = { ijk[ 0 ], ijk[ 1 ], ijk[ 2 ] };
      nb[ face >> 1 ] += ( face & 1 ) ? 1 : -1;
      if( in_domain( g, nb ) )
        continue;
      double d[ 3 ];
#pragma unroll
      for( int k = 0; k < DIM; ++k )
      {
        double fc = lower[ k ];
        if( k == ( face >> 1 ) )
        {
          if( face & 1 )
            fc += g.h[ k ];
        }
        else
          fc += 0.5 * g.h[ k ];
        d[ k ] = fc - center[ k ];
        d[ k ] *= 2.0;
      }
      ls_accumulate< DIM >( d, 0.5 - colorEn, AtA, Atb );
    }

    normal[ 0 ] = normal[ 1 ] = normal[ 2 ] = 0.0;
    solve_small< DIM >( AtA, Atb, normal );
    double n2 = 0.0;
#pragma unroll
    for( int k = 0; k < DIM; ++k )
      n2 += normal[ k ] * normal[ k ];
    if( n2 < kEps )
      return false;
    const double nrm = sqrt( n2 );
#pragma unroll
    for( int k = 0; k < DIM; ++k )
      normal[ k ] = vdiv( normal[ k ], nrm );
    return true;
  }

  template< int DIM >
  __global__ void __launch_bounds__( 128 ) k_youngs ( Grid g, const double *__restrict__ u, const int *__restrict__ mixedList,
                                                       const StepState *state, int capacity, double *__restrict__ hs )
  {
    const int count = mixed_count( state, capacity );
    for( int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x )
    {
      const long long c = mixedList[ idx ];
      int ijk[ 3 ];
      cell_ijk( g, c, ijk );
      const double colorEn = u[ c ];
      double normal[ 3 ];
      double *out = hs + c * ( DIM + 1 );
      if( !youngs_normal< DIM >( g, u, ijk, colorEn, normal ) )
      {
#pragma unroll
        for( int k = 0; k <= DIM; ++k )
          out[ k ] = 0.0;
        continue;
      }
      const double dist = locate_cell< DIM >( g, ijk, normal, colorEn );
#pragma unroll
      for( int k = 0; k < DIM; ++k )
        out[ k ] = normal[ k ];
      out[ DIM ] = dist;
    }
  }

  // ------------------------------------------------------------------------------------------------------------
  // S2 + R2: Cartesian height functions, stencil/heightfunctionstencil.hh:22-107, reconstruction/heightfunction.hh:101-259
  // ------------------------------------------------------------------------------------------------------------

  // getOrientation, heightfunction.hh:126-144
  template< int DIM >
  __device__ __forceinline__ void hf_orientation ( const double *normal, int &dir, int &sign )
  {
    dir = 0;
    double max = DBL_MIN;
#pragma unroll
    for( int i = 0; i < DIM; ++i )
      if( fabs( normal[ i ] ) > max )
      {
        dir = i;
        max = fabs( normal[ i ] );
      }
    if( fabs( max - 1.0 / sqrt( 2.0 ) ) < 1e-3 )
      dir = DIM - 1;
    sign = ( normal[ dir ] > 0 ) ? -1 : 1;
  }

  // getMultiIndex in cell-index units, heightfunctionstencil.hh:69-100 (the reference works on SPGrid ids 2*i+1)
  template< int DIM >
  __device__ __forceinline__ void hf_offset ( int i, int j, int c, int t, int *m )
  {
    m[ 0 ] = m[ 1 ] = m[ 2 ] = 0;
    if( DIM == 2 )
    {
      m[ 1 - i ] += ( c - 1 ) * ( i == 0 ? -1 : 1 );
      m[ i ] += t;
      m[ 0 ] *= j;
      m[ 1 ] *= j;
    }
    else
    {
      m[ i ] += j * t;
      m[ ( i + 1 ) % 3 ] += ( ( c % 3 ) - 1 ) * -j;
      m[ ( i + 2 ) % 3 ] += ( ( c / 3 ) - 1 );
    }
  }

  // the stencil cell (c,t) or -1 if it is outside (valid(), heightfunctionstencil.hh:63-66)
  template< int DIM >
  __device__ __forceinline__ long long hf_cell ( const Grid &g, const int *ijk, int i, int j, int c, int t )
  {
    int m[ 3 ];
    hf_offset< DIM >( i, j, c, t, m );
    const int nb[ 3 ] = { ijk[ 0 ] + m[ 0 ], ijk[ 1 ] + m[ 1 ], ijk[ 2 ] + m[ 2 ] };
    return cell_exists( g, nb ) ? cell_index( g, nb ) : -1;
  }

  // one column of getHeightValues, heightfunction.hh:147-196
  template< int DIM >
  __device__ double hf_column ( const Grid &g, const double *__restrict__ u, const int *ijk, int i, int j, int tup, int c )
  {
    const double TOL = 1e-10;
    double height = 0.0;
    long long cell = hf_cell< DIM >( g, ijk, i, j, c, 0 );
    if( cell < 0 )
      return height;
    const double u0 = clamp01( __ldg( u + cell ) );
    height += u0;
    double lastU = u0;
    for( int t = 1; t <= tup; ++t )
    {
      cell = hf_cell< DIM >( g, ijk, i, j, c, t );
      if( cell < 0 )
        break;
      const double uu = clamp01( __ldg( u + cell ) );
      if( uu > lastU - TOL && !( uu > 1.0 - TOL ) )
        break;
      height += uu;
      lastU = uu;
    }
    lastU = u0;
    for( int t = -1; t >= -tup; --t )
    {
      cell = hf_cell< DIM >( g, ijk, i, j, c, t );
      if( cell < 0 )
        break;
      double uu = clamp01( __ldg( u + cell ) );
      if( uu < lastU + TOL && !( uu < TOL ) )
        uu = 1.0;
      height += uu;
      lastU = uu;
    }
    return height;
  }

  // getHeightValues, heightfunction.hh:147-196
  template< int DIM >
  __device__ void hf_heights ( const Grid &g, const double *__restrict__ u, const int *ijk, int i, int j, int tup, double *heights )
  {
    constexpr int noc = ( DIM == 2 ) ? 3 : 9;
    for( int c = 0; c < noc; ++c )
      heights[ c ] = hf_column< DIM >( g, u, ijk, i, j, tup, c );
  }

  // computeNormal, heightfunction.hh:200-226 (2D) / :229-259 (3D)
  template< int DIM >
  __device__ __forceinline__ void hf_normal ( const double *heights, int i, int j, double *n )
  {
    if( DIM == 2 )
    {
      double Hx;
      if( fabs( heights[ 0 ] ) < kEps )
        Hx = ( heights[ 2 ] - heights[ 1 ] );
      else if( fabs( heights[ 2 ] ) < kEps )
        Hx = ( heights[ 1 ] - heights[ 0 ] );
      else
        Hx = ( heights[ 2 ] - heights[ 0 ] ) / 2.0;
      n[ 0 ] = Hx;
      n[ 1 ] = -1.0;
      n[ 2 ] = 0.0;
      const double nrm = norm2( n[ 0 ], n[ 1 ] );
      n[ 0 ] /= nrm;
      n[ 1 ] /= nrm;
      if( ( i == 0 && j == 1 ) || ( i == 1 && j == -1 ) )
      {
        n[ 0 ] *= -1.0;
        n[ 1 ] *= -1.0;
      }
      if( i == 0 )
      {
        const double a = -n[ 1 ], b = n[ 0 ];
        n[ 0 ] = a;
        n[ 1 ] = b;
      }
      return;
    }
    double Hx, Hy;
    if( heights[ 5 ] == 0.0 )
      Hx = ( heights[ 4 ] - heights[ 3 ] );
    else if( heights[ 3 ] == 0.0 )
      Hx = ( heights[ 5 ] - heights[ 4 ] );
    else
      Hx = ( heights[ 5 ] - heights[ 3 ] ) / 2.0;
    if( heights[ 7 ] == 0.0 )
      Hy = ( heights[ 4 ] - heights[ 1 ] );
    else if( heights[ 1 ] == 0.0 )
      Hy = ( heights[ 7 ] - heights[ 4 ] );
    else
      Hy = ( heights[ 7 ] - heights[ 1 ] ) / 2.0;
    n[ 0 ] = n[ 1 ] = n[ 2 ] = 0.0;
    n[ i ] = -j;
    n[ ( i + 1 ) % 3 ] = Hx * -j;
    n[ ( i + 2 ) % 3 ] = Hy;
    const double nrm = norm3( n );
#pragma unroll
    for( int k = 0; k < 3; ++k )
      n[ k ] /= nrm;
  }

  // HeightFunctionReconstruction::applyLocal on top of the Young's planes already in hs, heightfunction.hh:101-123
  template< int DIM >
  __global__ void __launch_bounds__( 128 ) k_hf ( Grid g, const double *__restrict__ u, const int *__restrict__ mixedList,
                                                   const StepState *state, int capacity, double *__restrict__ hs )
  {
    const int count = mixed_count( state, capacity );
    for( int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x )
    {
      const long long c = mixedList[ idx ];
      int ijk[ 3 ];
      cell_ijk( g, c, ijk );
      double *rec = hs + c * ( DIM + 1 );
      double youngs[ 3 ] = { rec[ 0 ], rec[ 1 ], DIM == 3 ? rec[ 2 ] : 0.0 };
      int dir, sign;
      hf_orientation< DIM >( youngs, dir, sign );
      double heights[ 9 ], newNormal[ 3 ];
      hf_heights< DIM >( g, u, ijk, dir, sign, 3, heights );
      hf_normal< DIM >( heights, dir, sign, newNormal );
      const double dist = locate_cell< DIM >( g, ijk, newNormal, u[ c ] );
#pragma unroll
      for( int k = 0; k < DIM; ++k )
        rec[ k ] = newNormal[ k ];
      rec[ DIM ] = dist;
    }
  }

  // per band part (BandPart, vofx_types.cuh): its counter of serially located cells and its region of the serialCells array
  __device__ __forceinline__ int *serial_locate_counter ( StepState *state, int part )
  {
    return part == 0 ? &state->serialLocate : &state->serialLocateMore[ part - 1 ];
  }
  __host__ __device__ __forceinline__ int serial_locate_offset ( int capacity, const BandPart &bp )
  {
    return bp.part * ( capacity / bp.parts + 1 );      // a part holds at most ceil( capacity / parts ) cells
  }

  // R1 in 3D for the B200: L lanes per mixed cell (vofx_coop3d.cuh).  The stencil contributions, the faces of the cell
  // volume, the direction vectors and the per-bracket terms of the plane-constant solve are spread over the lanes;
  // the sweep topology comes from a table of the 96 generic height orders.  Cells with tied node heights (axis-aligned
  // normals) are queued for k_locate_serial3.
#ifndef VOFX_YOUNGS_THREADS_PER_SM
#define VOFX_YOUNGS_THREADS_PER_SM 640   // resident threads per SM the register allocation aims for
#endif
  template< int L, bool HF >
  __global__ void __launch_bounds__( 16 * L, VOFX_YOUNGS_THREADS_PER_SM / ( 16 * L ) ) k_youngs_coop3 ( Grid g, const double *__restrict__ u, const int *__restrict__ mixedList,
                                                             StepState *state, int capacity, double *__restrict__ hs,
                                                             const coop::SweepTable *__restrict__ table, int *__restrict__ serialCells,
                                                             coop::RootTask *__restrict__ rootTasks, const int *__restrict__ order,
                                                             unsigned char *__restrict__ costClass, BandPart bp )
  {
    pdl_prologue();
    constexpr int GROUPS = 16;          // 16 groups per block: 128 threads for L = 8, 64 for L = 4
    __shared__ coop::LocateScratch scratch[ GROUPS ];
    const int count = mixed_count( state, capacity );
    int *serialCounter = serial_locate_counter( state, bp.part );
    serialCells += serial_locate_offset( capacity, bp );
    const int lane = threadIdx.x % L;
    const int group = threadIdx.x / L;
    coop::Group G;
    G.lane = lane;
    G.mask = ( ( L == 32 ) ? 0xffffffffu : ( ( 1u << L ) - 1u ) ) << ( ( threadIdx.x & 31 ) - lane );
    coop::LocateScratch &S = scratch[ group ];
    // part p owns the band-list entries p, p + parts, ...; in the order array they follow the entries of the parts before it,
    // sorted by last step's cost class (k_cost_order): cells of one class share a warp
    const int nMine = ( count - bp.part + bp.parts - 1 ) / bp.parts;
    int orderBase = 0;
    for( int q = 0; q < bp.part; ++q )
      orderBase += ( count - q + bp.parts - 1 ) / bp.parts;
    for( int mine = blockIdx.x * GROUPS + group; mine < nMine; mine += gridDim.x * GROUPS )
    {
      const int idx = order ? order[ orderBase + mine ] : mine * bp.parts + bp.part;   // position in the band list
      const int c = mixedList[ idx ];
      int ijk[ 3 ];
      cell_ijk( g, c, ijk );
      const double colorEn = u[ c ];
      double *out = hs + (long long) c * 4;
      double normal[ 3 ];
      const bool haveYoungs = coop::youngs_normal_coop< L >( G, S, g, u, ijk, colorEn, normal );
      if( !HF && !haveYoungs )
      {
        if( lane == 0 )
        {
          out[ 0 ] = out[ 1 ] = out[ 2 ] = out[ 3 ] = 0.0;
          rootTasks[ idx ].pending = 0;
        }
        G.sync();
        continue;
      }
      if( HF )
      {
        // HeightFunctionReconstruction::applyLocal, heightfunction.hh:101-123: of the Young's reconstruction only the normal
        // of the cell is read, so its plane-constant solve is skipped; the nine columns are summed by different lanes
        if( !haveYoungs )
          normal[ 0 ] = normal[ 1 ] = normal[ 2 ] = 0.0;       // the zero half-space, modifiedyoungs.hh:128-129
        int dir, sign;
        hf_orientation< 3 >( normal, dir, sign );
        G.sync();                                              // S.sums (the least-squares sums) has been read by every lane
        for( int col = lane; col < 9; col += L )
          S.sums[ col ] = hf_column< 3 >( g, u, ijk, dir, sign, 3, col );
        G.sync();
        double heights[ 9 ];
#pragma unroll
        for( int col = 0; col < 9; ++col )
          heights[ col ] = S.sums[ col ];
        hf_normal< 3 >( heights, dir, sign, normal );
        G.sync();                                              // before locate_coop reuses the scratch
      }
      const double lo[ 3 ] = { cell_lower( g, 0, ijk[ 0 ] ), cell_lower( g, 1, ijk[ 1 ] ), cell_lower( g, 2, ijk[ 2 ] ) };
      double dist = 0.0;
      coop::RootTask root;
      root.pending = 0;
      int cost = 0;
      const bool located = coop::locate_coop< L >( G, S, *table, lo, g.h, normal, colorEn, dist, &root, &cost );
      if( lane == 0 )
      {
        if( costClass )
          costClass[ c ] = (unsigned char) cost;
        out[ 0 ] = normal[ 0 ];
        out[ 1 ] = normal[ 1 ];
        out[ 2 ] = normal[ 2 ];
        // the root tasks sit at the cell's position in the band list: no counter shared by 1e5 cells
        if( !located )
          serialCells[ atomicAdd( serialCounter, 1 ) ] = c;
        else if( root.pending )
        {
          root.cell = c;
          rootTasks[ idx ] = root;
        }
        else
          out[ 3 ] = dist;
        if( !located || !root.pending )
          rootTasks[ idx ].pending = 0;
      }
      G.sync();
    }
  }

  // Brent's method for the cells whose bracket k_youngs_coop3 has found: one thread per cell
  template< int DIM >
  __global__ void __launch_bounds__( 128 ) k_locate_root3 ( StepState *state, int capacity, const coop::RootTask *__restrict__ rootTasks, double *__restrict__ hs,
                                                            BandPart bp )
  {
    pdl_prologue();
    const int n = mixed_count( state, capacity );
    int mine = 0;
    for( int t = ( blockIdx.x * blockDim.x + threadIdx.x ) * bp.parts + bp.part; t < n; t += gridDim.x * blockDim.x * bp.parts )
    {
      if( !rootTasks[ t ].pending )
        continue;
      const coop::RootTask r = rootTasks[ t ];
      hs[ (long long) r.cell * 4 + 3 ] = r.base + brent( r.V, 0.0, r.h, 1e-14 );
      mine++;
    }
    // diagnostic counter (vofx_step_counters): one atomic per warp
#pragma unroll
    for( int o = 16; o > 0; o >>= 1 )
      mine += __shfl_xor_sync( 0xffffffffu, mine, o );
    if( ( threadIdx.x & 31 ) == 0 && mine > 0 )
      atomicAdd( &state->rootCount, mine );
  }

  // plane-constant solve by the serial walk for the cells queued by k_youngs_coop3 (normal already in hs)
  template< int DIM >
  __global__ void __launch_bounds__( 64 ) k_locate_serial3 ( Grid g, const double *__restrict__ u, const StepState *state, double *__restrict__ hs,
                                                              const int *__restrict__ serialCells, int capacity, BandPart bp )
  {
    pdl_prologue();
    const int n = *serial_locate_counter( const_cast< StepState * >( state ), bp.part );
    serialCells += serial_locate_offset( capacity, bp );
    for( int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x )
    {
      const int c = serialCells[ t ];
      int ijk[ 3 ];
      cell_ijk( g, c, ijk );
      double *rec = hs + (long long) c * 4;
      const double normal[ 3 ] = { rec[ 0 ], rec[ 1 ], rec[ 2 ] };
      rec[ 3 ] = locate_cell< 3 >( g, ijk, normal, u[ c ] );
    }
  }

  // HeightFunctionReconstruction in one kernel (2D): the reference first runs ModifiedYoungs on every mixed cell
  // (heightfunction.hh:73) and then overwrites each reconstruction (:122); of the Young's result only the NORMAL of the
  // cell itself is read (getOrientation, :104), so its plane-constant solve is dead work and the two band passes fuse.
  template< int DIM >
  __global__ void __launch_bounds__( 128 ) k_hf_fused ( Grid g, const double *__restrict__ u, const int *__restrict__ mixedList,
                                                         const StepState *state, int capacity, double *__restrict__ hs )
  {
    const int count = mixed_count( state, capacity );
    for( int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x )
    {
      const long long c = mixedList[ idx ];
      int ijk[ 3 ];
      cell_ijk( g, c, ijk );
      const double colorEn = u[ c ];
      double youngs[ 3 ];
      if( !youngs_normal< DIM >( g, u, ijk, colorEn, youngs ) )
        youngs[ 0 ] = youngs[ 1 ] = youngs[ 2 ] = 0.0;     // the zero half-space, modifiedyoungs.hh:128-129
      int dir, sign;
      hf_orientation< DIM >( youngs, dir, sign );
      double heights[ 9 ], newNormal[ 3 ];
      hf_heights< DIM >( g, u, ijk, dir, sign, 3, heights );
      hf_normal< DIM >( heights, dir, sign, newNormal );
      const double dist = locate_cell< DIM >( g, ijk, newNormal, colorEn );
      double *rec = hs + c * ( DIM + 1 );
#pragma unroll
      for( int k = 0; k < DIM; ++k )
        rec[ k ] = newNormal[ k ];
      rec[ DIM ] = dist;
    }
  }

  // curvature[ c ] = 0 for the cells that received a value in the previous step (instead of a dense clear)
  static __global__ void __launch_bounds__( 256 ) k_clear_listed ( const int *__restrict__ list, const int *__restrict__ count, double *__restrict__ target )
  {
    const int n = *count;
    for( int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x )
      target[ list[ i ] ] = 0.0;
  }

  static __global__ void __launch_bounds__( 256 ) k_copy_list ( const int *__restrict__ list, const StepState *state, int capacity, int *__restrict__ out, int *__restrict__ outCount )
  {
    const int n = mixed_count( state, capacity );
    for( int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x )
      out[ i ] = list[ i ];
    if( blockIdx.x == 0 && threadIdx.x == 0 )
      *outCount = n;
  }

  // ------------------------------------------------------------------------------------------------------------
  // R3: modified Swartz, reconstruction/modifiedswartz.hh:102-243.  One thread per mixed cell; the neighbours'
  // interface centroids depend only on (neighbour, oldNormal) and are computed once per iteration.
  // ------------------------------------------------------------------------------------------------------------

  template< int DIM >
  __device__ void interface_centroid ( const Grid &g, const int *ijk, const double *normal, double color, double *centroid )
  {
    const double dist = locate_cell< DIM >( g, ijk, normal, color );
    if( DIM == 2 )
    {
      Poly2 p;
      make_cell( cell_lower( g, 0, ijk[ 0 ] ), cell_lower( g, 1, ijk[ 1 ] ), g.h[ 0 ], g.h[ 1 ], p );
      const Hs2 H = { normal[ 0 ], normal[ 1 ], dist };
      double lx[ 2 ], ly[ 2 ];
      plane_cut( p, H, lx, ly );
      centroid[ 0 ] = ( lx[ 0 ] + lx[ 1 ] ) * 0.5;
      centroid[ 1 ] = ( ly[ 0 ] + ly[ 1 ] ) * 0.5;
      centroid[ 2 ] = 0.0;
      return;
    }
    const double lo[ 3 ] = { cell_lower( g, 0, ijk[ 0 ] ), cell_lower( g, 1, ijk[ 1 ] ), cell_lower( g, 2, ijk[ 2 ] ) };
    Hex cell;
    make_cell( lo, g.h, cell );
    const Hs3 H = { { normal[ 0 ], normal[ 1 ], normal[ 2 ] }, dist };
    Face3 f;
    plane_cut( cell, H, f );
    face_centroid( f, centroid );
  }

  template< int DIM >
  __global__ void __launch_bounds__( 64 ) k_swartz ( Grid g, const double *__restrict__ u, const unsigned char *__restrict__ flags,
                                                      const int *__restrict__ mixedList, const StepState *state, int capacity,
                                                      double *__restrict__ hs, int maxIterations )
  {
    constexpr int maxNb = ( DIM == 3 ) ? 26 : 8;
    const int count = mixed_count( state, capacity );
    for( int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x )
    {
      const long long c = mixedList[ idx ];
      int ijk[ 3 ];
      cell_ijk( g, c, ijk );
      double *rec = hs + c * ( DIM + 1 );
      const double colorEn = u[ c ];

      // mixed vertex neighbours, ascending index; packed offsets (dx+1) | (dy+1)<<2 | (dz+1)<<4
      unsigned char nbOff[ maxNb ];
      int nNb = 0;
      for( int dz = ( DIM == 3 ? -1 : 0 ); dz <= ( DIM == 3 ? 1 : 0 ); ++dz )
        for( int dy = -1; dy <= 1; ++dy )
          for( int dx = -1; dx <= 1; ++dx )
          {
            if( dx == 0 && dy == 0 && dz == 0 )
              continue;
            const int nb[ 3 ] = { ijk[ 0 ] + dx, ijk[ 1 ] + dy, ijk[ 2 ] + dz };
            if( !cell_exists( g, nb ) )
              continue;
            if( !is_mixed( flags[ cell_index( g, nb ) ] ) )
              continue;
            nbOff[ nNb++ ] = (unsigned char) ( ( dx + 1 ) | ( ( dy + 1 ) << 2 ) | ( ( dz + 1 ) << 4 ) );
          }

      double normal[ 3 ] = { rec[ 0 ], rec[ 1 ], DIM == 3 ? rec[ 2 ] : 0.0 };
      int iterations = 0;
      double residuum = 0.0;
      do
      {
        const double oldNormal[ 3 ] = { normal[ 0 ], normal[ 1 ], normal[ 2 ] };
        normal[ 0 ] = normal[ 1 ] = normal[ 2 ] = 0.0;

        double cEn[ 3 ];
        interface_centroid< DIM >( g, ijk, oldNormal, colorEn, cEn );

        double cNb[ maxNb ][ 3 ];
        for( int a = 0; a < nNb; ++a )
        {
          const int nb[ 3 ] = { ijk[ 0 ] + ( nbOff[ a ] & 3 ) - 1, ijk[ 1 ] + ( ( nbOff[ a ] >> 2 ) & 3 ) - 1, ijk[ 2 ] + ( ( nbOff[ a ] >> 4 ) & 3 ) - 1 };
          interface_centroid< DIM >( g, nb, oldNormal, u[ cell_index( g, nb ) ], cNb[ a ] );
        }

        if( DIM == 2 )
        {
          for( int a = 0; a < nNb; ++a )
          {
            const double dx = cNb[ a ][ 0 ] - cEn[ 0 ], dy = cNb[ a ][ 1 ] - cEn[ 1 ];
            double cx = -dy, cy = dx;
            if( dot2( cx, cy, oldNormal[ 0 ], oldNormal[ 1 ] ) < 0 )
            {
              cx *= -1.0;
              cy *= -1.0;
            }
            const double nrm = norm2( cx, cy );
            cx /= nrm;
            cy /= nrm;
            normal[ 0 ] += 1.0 * cx;
            normal[ 1 ] += 1.0 * cy;
          }
        }
        else
        {
          for( int a = 0; a < nNb; ++a )
            for( int b = 0; b < nNb; ++b )
            {
              if( a == b )
                continue;
              double d1[ 3 ], d2[ 3 ], cn[ 3 ];
#pragma unroll
              for( int k = 0; k < 3; ++k )
              {
                d1[ k ] = cNb[ a ][ k ] - cEn[ k ];
                d2[ k ] = cNb[ b ][ k ] - cEn[ k ];
              }
              cross3( d1, d2, cn );
              if( norm3( cn ) < 1e-12 )
                continue;
              if( dot3( cn, oldNormal ) < 0 )
              {
#pragma unroll
                for( int k = 0; k < 3; ++k )
                  cn[ k ] *= -1.0;
              }
              const double nrm = norm3( cn );
#pragma unroll
              for( int k = 0; k < 3; ++k )
              {
                cn[ k ] /= nrm;
                normal[ k ] += 1.0 * cn[ k ];
              }
            }
        }

        double n2 = 0.0;
#pragma unroll
        for( int k = 0; k < DIM; ++k )
          n2 += normal[ k ] * normal[ k ];
        const double nrm = sqrt( n2 );
        if( nrm < 0.5 )
          break;   // "return": keep the last reconstruction, modifiedswartz.hh:140-141,231-232
#pragma unroll
        for( int k = 0; k < DIM; ++k )
          normal[ k ] /= nrm;

        const double dist = locate_cell< DIM >( g, ijk, normal, colorEn );
#pragma unroll
        for( int k = 0; k < DIM; ++k )
          rec[ k ] = normal[ k ];
        rec[ DIM ] = dist;

        double dotp = 0.0;
#pragma unroll
        for( int k = 0; k < DIM; ++k )
          dotp += normal[ k ] * oldNormal[ k ];
        residuum = 1.0 - dotp;
        ++iterations;
      }
      while( residuum > 1e-12 && iterations < maxIterations );
    }
  }

  // R3 in 3D for the B200: L lanes per mixed cell (coop::swartz_cell_coop).  The serial k_swartz< 3 > spends 0.26 s per
  // step on 3e4 mixed cells at 256^3 (one thread walks ~90 plane-constant solves per cell).
  template< int L >
  __global__ void __launch_bounds__( 8 * L ) k_swartz_coop3 ( Grid g, const double *__restrict__ u, const unsigned char *__restrict__ flags,
                                                               const int *__restrict__ mixedList, const StepState *state, int capacity,
                                                               double *__restrict__ hs, int maxIterations, const coop::SweepTable *__restrict__ table )
  {
    constexpr int GROUPS = 8;
    __shared__ coop::SwartzScratch< L > scratch[ GROUPS ];
    const int count = mixed_count( state, capacity );
    const int lane = threadIdx.x % L;
    const int group = threadIdx.x / L;
    coop::Group G;
    G.lane = lane;
    G.mask = ( ( L == 32 ) ? 0xffffffffu : ( ( 1u << L ) - 1u ) ) << ( ( threadIdx.x & 31 ) - lane );
    coop::SwartzScratch< L > &S = scratch[ group ];
    for( int idx = blockIdx.x * GROUPS + group; idx < count; idx += gridDim.x * GROUPS )
    {
      const int c = mixedList[ idx ];
      int ijk[ 3 ];
      cell_ijk( g, c, ijk );
      double *out = hs + (long long) c * 4;
      double rec[ 4 ] = { out[ 0 ], out[ 1 ], out[ 2 ], out[ 3 ] };
      coop::swartz_cell_coop< L >( G, S, *table, g, u, flags, ijk, maxIterations, rec );
      if( lane == 0 )
      {
        out[ 0 ] = rec[ 0 ];
        out[ 1 ] = rec[ 1 ];
        out[ 2 ] = rec[ 2 ];
        out[ 3 ] = rec[ 3 ];
      }
      G.sync();
    }
  }

  // ------------------------------------------------------------------------------------------------------------
  // R3 in 3D, batched over the whole band (modifiedswartz.hh:167-243).  One iteration of every still-active cell is four
  // small kernels with compact code and full grids instead of one long per-cell program:
  //   k_swz_locate   group of L lanes per (cell, neighbour) task: bracket of the plane constant with the cell's normal
  //   k_swz_centroid one thread per task: root search, interface polygon, centroid
  //   k_swz_normal   group per cell: the n(n-1) cross products row by row, new normal, bracket for the cell itself
  //   k_swz_finish   one thread per cell: root search, write the reconstruction, convergence test
  // The iteration count is data dependent (<= maxIterations, default 10): all iterations are enqueued and the cells
  // that have converged are skipped on the device (no host round trip).
  // ------------------------------------------------------------------------------------------------------------

  struct SwzCell
  {
    double normal[ 3 ], oldNormal[ 3 ];
    double dist;
    coop::RootTask root;              // bracket of the cell's own plane constant (k_swz_normal -> k_swz_finish)
    int taskBase, nNb;
    int active, iterations, state;    // state of `root`: 0 dist known, 1 root pending, 2 serial walk, 3 keep the last reconstruction
    int cell;
    unsigned char nbOff[ 28 ];
  };

  struct SwzTask
  {
    coop::RootTask root;
    double dist;
    double cen[ 3 ];
    int owner;                        // index into the mixed list / SwzCell array
    int which;                        // 0: the cell itself, a > 0: its (a-1)-th mixed neighbour
    int state;
    int cls;                          // bracket the last plane-constant solve of this task ended in (7: walked on lane 0): sort key of the next iteration
  };

  template< int DIM_ >
  __global__ void __launch_bounds__( 128 ) k_swz_setup ( Grid g, const unsigned char *__restrict__ flags, const int *__restrict__ mixedList,
                                                          StepState *state, int capacity, const double *__restrict__ hs,
                                                          SwzCell *__restrict__ cells, SwzTask *__restrict__ tasks, int *taskCounter,
                                                          int cellCapacity, int taskCapacity, int *__restrict__ cellList, int *listCount )
  {
    const int count = mixed_count( state, capacity );
    // the cells still iterating, as a list per iteration (k_swz_finish fills the next one): list 0 = every cell of the band
    if( cellList && blockIdx.x == 0 && threadIdx.x == 0 )
    {
      listCount[ 0 ] = ( count > cellCapacity ) ? 0 : count;
      listCount[ 1 ] = 0;
    }
    // the buffers are sized by the host from the band of an earlier pass (no read-back per pass, the pass stays
    // graph-capturable): a band that outgrew them turns the pass into a no-op and the host sees the overflow at its next
    // read of the loop state (and enlarges the buffers)
    if( count > cellCapacity )
    {
      if( blockIdx.x == 0 && threadIdx.x == 0 )
        state->overflow = 2;
      return;
    }
    for( int p = blockIdx.x * blockDim.x + threadIdx.x; p < count; p += gridDim.x * blockDim.x )
    {
      const int c = mixedList[ p ];
      int ijk[ 3 ];
      cell_ijk( g, c, ijk );
      SwzCell &C = cells[ p ];
      if( cellList )
        cellList[ p ] = p;
      int n = 0;
      for( int s = 0; s < 26; ++s )
      {
        const int m = ( s < 13 ) ? s : s + 1;
        const int nb[ 3 ] = { ijk[ 0 ] + ( m % 3 ) - 1, ijk[ 1 ] + ( ( m / 3 ) % 3 ) - 1, ijk[ 2 ] + ( m / 9 ) - 1 };
        if( cell_exists( g, nb ) && is_mixed( flags[ cell_index( g, nb ) ] ) )
          C.nbOff[ n++ ] = (unsigned char) ( ( m % 3 ) | ( ( ( m / 3 ) % 3 ) << 2 ) | ( ( m / 9 ) << 4 ) );
      }
      const double *rec = hs + (long long) c * 4;
      for( int k = 0; k < 3; ++k )
        C.normal[ k ] = rec[ k ];
      C.dist = rec[ 3 ];
      C.nNb = n;
      C.cell = c;
      C.active = 1;
      C.iterations = 0;
      C.state = 0;
      const int base = atomicAdd( taskCounter, n + 1 );
      C.taskBase = base;
      if( base + n + 1 > taskCapacity )
      {
        state->overflow = 2;
        C.nNb = 0;
        C.active = 0;
        continue;
      }
      for( int a = 0; a <= n; ++a )
      {
        tasks[ base + a ].owner = p;
        tasks[ base + a ].which = a;
        tasks[ base + a ].cls = 0;
      }
    }
  }

  __device__ __forceinline__ void swz_task_cell ( const SwzCell &C, const Grid &g, int which, int *cell )
  {
    cell_ijk( g, C.cell, cell );
    if( which > 0 )
    {
      const int off = C.nbOff[ which - 1 ];
      cell[ 0 ] += ( off & 3 ) - 1;
      cell[ 1 ] += ( ( off >> 2 ) & 3 ) - 1;
      cell[ 2 ] += ( ( off >> 4 ) & 3 ) - 1;
    }
  }

  // Every iteration the tasks of the still-active cells are counting-sorted by the bracket their last solve ended in (the
  // same idea as k_cost_order: the eight tasks of a warp of k_swz_locate leave the bracket loop together, and the tasks of
  // converged cells no longer occupy lanes of k_swz_locate / k_swz_centroid).  sort[ 0 .. 15 ] histogram, [ 16 .. 31 ] fill
  // counters, [ 32 ] number of active tasks; bucket 8 = inactive.
  __device__ __forceinline__ int swz_task_bucket ( const SwzCell *cells, const SwzTask &T ) { return cells[ T.owner ].active ? ( T.cls & 7 ) : 8; }

  static __global__ void __launch_bounds__( 256 ) k_swz_hist ( const SwzCell *__restrict__ cells, const SwzTask *__restrict__ tasks, const int *taskCounter, int *sort )
  {
    __shared__ int hist[ 16 ];
    if( threadIdx.x < 16 )
      hist[ threadIdx.x ] = 0;
    __syncthreads();
    const int nTasks = *taskCounter;
    for( int t = blockIdx.x * 256 + threadIdx.x; t < nTasks; t += gridDim.x * 256 )
      atomicAdd( &hist[ swz_task_bucket( cells, tasks[ t ] ) ], 1 );
    __syncthreads();
    if( threadIdx.x < 16 && hist[ threadIdx.x ] )
      atomicAdd( &sort[ threadIdx.x ], hist[ threadIdx.x ] );
  }

  static __global__ void __launch_bounds__( 256 ) k_swz_order ( const SwzCell *__restrict__ cells, const SwzTask *__restrict__ tasks, const int *taskCounter, int *sort,
                                                                 int *__restrict__ order )
  {
    __shared__ int warpCount[ 16 ][ 8 ], tileBase[ 16 ], start[ 16 ];
    const int nTasks = *taskCounter;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if( threadIdx.x == 0 )
    {
      int at = 0;
      for( int c = 7; c >= 0; --c )      // expensive first; the inactive tasks behind all active ones
      {
        start[ c ] = at;
        at += sort[ c ];
      }
      start[ 8 ] = at;
      if( blockIdx.x == 0 )
        sort[ 32 ] = at;                 // active tasks: the grid-stride loops of this iteration's kernels end here
    }
    __syncthreads();
    const int rounded = ( nTasks + 255 ) & ~255;
    for( int first = blockIdx.x * 256; first < rounded; first += gridDim.x * 256 )
    {
      const int t = first + threadIdx.x;
      const bool valid = t < nTasks;
      const int bucket = valid ? swz_task_bucket( cells, tasks[ t ] ) : 16 + warp;
      if( threadIdx.x < 128 )
        warpCount[ threadIdx.x >> 3 ][ threadIdx.x & 7 ] = 0;
      __syncthreads();
      const unsigned peers = __match_any_sync( 0xffffffffu, bucket );
      const int below = __popc( peers & ( ( 1u << lane ) - 1u ) );
      if( valid && below == 0 )
        warpCount[ bucket ][ warp ] = __popc( peers );
      __syncthreads();
      if( threadIdx.x < 9 )
      {
        const int b = threadIdx.x;
        int n = 0;
        for( int w = 0; w < 8; ++w )
        {
          const int a = warpCount[ b ][ w ];
          warpCount[ b ][ w ] = n;
          n += a;
        }
        tileBase[ b ] = n ? atomicAdd( &sort[ 16 + b ], n ) : 0;
      }
      __syncthreads();
      if( valid )
      {
        const int pos = start[ bucket ] + tileBase[ bucket ] + warpCount[ bucket ][ warp ] + below;
        if( pos < nTasks )
          order[ pos ] = t;
      }
      __syncthreads();
    }
  }

  template< int L >
  __global__ void __launch_bounds__( 16 * L, VOFX_YOUNGS_THREADS_PER_SM / ( 16 * L ) ) k_swz_locate ( Grid g, const double *__restrict__ u, const SwzCell *__restrict__ cells,
                                                                                                        SwzTask *__restrict__ tasks, const int *taskCounter,
                                                                                                        const coop::SweepTable *__restrict__ table, const int *__restrict__ order )
  {
    constexpr int GROUPS = 16;
    __shared__ coop::LocateScratch scratch[ GROUPS ];
    const int nTasks = *taskCounter;     // with `order`: the tasks of the still-active cells, sorted by class (k_swz_order)
    const int lane = threadIdx.x % L;
    const int group = threadIdx.x / L;
    coop::Group G;
    G.lane = lane;
    G.mask = ( ( 1u << L ) - 1u ) << ( ( threadIdx.x & 31 ) - lane );
    coop::LocateScratch &S = scratch[ group ];
    for( int q = blockIdx.x * GROUPS + group; q < nTasks; q += gridDim.x * GROUPS )
    {
      SwzTask &T = tasks[ order ? order[ q ] : q ];
      const SwzCell &C = cells[ T.owner ];
      if( !C.active )
        continue;
      int cell[ 3 ];
      swz_task_cell( C, g, T.which, cell );
      const double lo[ 3 ] = { cell_lower( g, 0, cell[ 0 ] ), cell_lower( g, 1, cell[ 1 ] ), cell_lower( g, 2, cell[ 2 ] ) };
      const double normal[ 3 ] = { C.normal[ 0 ], C.normal[ 1 ], C.normal[ 2 ] };
      const double color = u[ cell_index( g, cell ) ];
      coop::RootTask root;
      root.pending = 0;
      double dist = 0.0;
      int cost = 0;
      const bool ok = coop::locate_coop< L >( G, S, *table, lo, g.h, normal, color, dist, &root, &cost );
      if( lane == 0 )
      {
        T.state = !ok ? 2 : ( root.pending ? 1 : 0 );
        T.dist = dist;
        T.root = root;
        T.cls = cost;
      }
      G.sync();
    }
  }

  template< int DIM_ >
  __global__ void __launch_bounds__( 64 ) k_swz_centroid ( Grid g, const double *__restrict__ u, const SwzCell *__restrict__ cells, SwzTask *__restrict__ tasks,
                                                            const int *taskCounter, const int *__restrict__ order )
  {
    const int nTasks = *taskCounter;
    for( int q = blockIdx.x * blockDim.x + threadIdx.x; q < nTasks; q += gridDim.x * blockDim.x )
    {
      SwzTask &T = tasks[ order ? order[ q ] : q ];
      const SwzCell &C = cells[ T.owner ];
      if( !C.active )
        continue;
      int cell[ 3 ];
      swz_task_cell( C, g, T.which, cell );
      const double lo[ 3 ] = { cell_lower( g, 0, cell[ 0 ] ), cell_lower( g, 1, cell[ 1 ] ), cell_lower( g, 2, cell[ 2 ] ) };
      const double normal[ 3 ] = { C.normal[ 0 ], C.normal[ 1 ], C.normal[ 2 ] };
      double dist = T.dist;
      if( T.state == 1 )
        dist = coop::root_of( T.root );
      else if( T.state == 2 )
        dist = coop::locate_serial( lo, g.h, normal, u[ cell_index( g, cell ) ] );
      double c[ 3 ];
      coop::interface_centroid_serial( lo, g.h, normal, dist, c );
      T.cen[ 0 ] = c[ 0 ];
      T.cen[ 1 ] = c[ 1 ];
      T.cen[ 2 ] = c[ 2 ];
    }
  }

  // the cross products of the new normal and the plane-constant solve that follows never overlap in time (a group barrier
  // separates them): one overlay, 1.3 kB instead of 2.5 kB per cell -- the kernel's occupancy is set by its shared memory
  struct SwzNormalScratch
  {
    union
    {
      coop::LocateScratch loc;
      struct
      {
        double row[ coop::kMaxSwartzNb ][ 3 ];
        double cen[ coop::kMaxSwartzNb + 1 ][ 3 ];
        double nsum[ 3 ];
        unsigned char rowValid[ 32 ];
      };
    };
  };

  template< int L >
  __global__ void __launch_bounds__( 8 * L ) k_swz_normal ( Grid g, const double *__restrict__ u, SwzCell *__restrict__ cells, const SwzTask *__restrict__ tasks,
                                                             const StepState *state, int capacity, const coop::SweepTable *__restrict__ table,
                                                             const int *__restrict__ cellList, int *listCount, int iteration, int cellCapacity )
  {
    constexpr int GROUPS = 8;
    __shared__ SwzNormalScratch scratch[ GROUPS ];
    // cellList: the cells still iterating (list iteration & 1, written by the previous iteration's k_swz_finish); the other
    // list is emptied here for this iteration's k_swz_finish
    const int *list = cellList ? cellList + ( iteration & 1 ) * cellCapacity : nullptr;
    const int count = list ? listCount[ iteration & 1 ] : mixed_count( state, capacity );
    if( list && blockIdx.x == 0 && threadIdx.x == 0 )
      listCount[ ( iteration & 1 ) ^ 1 ] = 0;
    const int lane = threadIdx.x % L;
    const int group = threadIdx.x / L;
    coop::Group G;
    G.lane = lane;
    G.mask = ( ( 1u << L ) - 1u ) << ( ( threadIdx.x & 31 ) - lane );
    SwzNormalScratch &S = scratch[ group ];
    for( int q = blockIdx.x * GROUPS + group; q < count; q += gridDim.x * GROUPS )
    {
      SwzCell &C = cells[ list ? list[ q ] : q ];
      if( !C.active )
        continue;
      const int nNb = C.nNb;
      const double oldNormal[ 3 ] = { C.normal[ 0 ], C.normal[ 1 ], C.normal[ 2 ] };
      for( int a = lane; a <= nNb; a += L )
        for( int k = 0; k < 3; ++k )
          S.cen[ a ][ k ] = tasks[ C.taskBase + a ].cen[ k ];
      if( lane < 3 )
        S.nsum[ lane ] = 0.0;
      G.sync();
      // normal = sum over ordered pairs ( a, b ), a != b, of the normalised cross products, modifiedswartz.hh:196-229
      for( int a = 0; a < nNb; ++a )
      {
        for( int b = lane; b < nNb; b += L )
        {
          bool valid = false;
          double cn[ 3 ] = { 0.0, 0.0, 0.0 };
          if( b != a )
          {
            double d1[ 3 ], d2[ 3 ];
            for( int k = 0; k < 3; ++k )
            {
              d1[ k ] = S.cen[ a + 1 ][ k ] - S.cen[ 0 ][ k ];
              d2[ k ] = S.cen[ b + 1 ][ k ] - S.cen[ 0 ][ k ];
            }
            cross3( d1, d2, cn );
            const double nrm = norm3( cn );            // | -cn | has the same bits: the reference's second two_norm() is this value
            if( !( nrm < 1e-12 ) )
            {
              valid = true;
              if( dot3( cn, oldNormal ) < 0 )
                for( int k = 0; k < 3; ++k )
                  cn[ k ] *= -1.0;
              vdiv3( cn[ 0 ], cn[ 1 ], cn[ 2 ], nrm );
            }
          }
          S.rowValid[ b ] = valid ? 1 : 0;
          for( int k = 0; k < 3; ++k )
            S.row[ b ][ k ] = cn[ k ];
        }
        G.sync();
        if( lane < 3 )
        {
          double acc = S.nsum[ lane ];
          for( int b = 0; b < nNb; ++b )
            if( S.rowValid[ b ] )
              acc += 1.0 * S.row[ b ][ lane ];
          S.nsum[ lane ] = acc;
        }
        G.sync();
      }
      double normal[ 3 ] = { S.nsum[ 0 ], S.nsum[ 1 ], S.nsum[ 2 ] };
      double n2 = 0.0;
      for( int k = 0; k < 3; ++k )
        n2 += normal[ k ] * normal[ k ];
      const double nrm = sqrt( n2 );
      G.sync();
      if( nrm < 0.5 )
      {
        // "return": keep the last reconstruction, modifiedswartz.hh:231-232
        if( lane == 0 )
          C.state = 3;
        continue;
      }
      for( int k = 0; k < 3; ++k )
        normal[ k ] = vdiv( normal[ k ], nrm );
      int ijk[ 3 ];
      cell_ijk( g, C.cell, ijk );
      const double lo[ 3 ] = { cell_lower( g, 0, ijk[ 0 ] ), cell_lower( g, 1, ijk[ 1 ] ), cell_lower( g, 2, ijk[ 2 ] ) };
      coop::RootTask root;
      root.pending = 0;
      double dist = 0.0;
      const bool ok = coop::locate_coop< L >( G, S.loc, *table, lo, g.h, normal, u[ C.cell ], dist, &root );
      if( lane == 0 )
      {
        for( int k = 0; k < 3; ++k )
        {
          C.oldNormal[ k ] = oldNormal[ k ];
          C.normal[ k ] = normal[ k ];
        }
        C.dist = dist;
        C.root = root;
        C.state = !ok ? 2 : ( root.pending ? 1 : 0 );
      }
      G.sync();
    }
  }

  template< int DIM_ >
  __global__ void __launch_bounds__( 64 ) k_swz_finish ( Grid g, const double *__restrict__ u, SwzCell *__restrict__ cells, const StepState *state, int capacity,
                                                          double *__restrict__ hs, int maxIterations, int *__restrict__ cellList, int *listCount, int iteration,
                                                          int cellCapacity )
  {
    const int *list = cellList ? cellList + ( iteration & 1 ) * cellCapacity : nullptr;
    int *next = cellList ? cellList + ( ( iteration & 1 ) ^ 1 ) * cellCapacity : nullptr;
    const int count = list ? listCount[ iteration & 1 ] : mixed_count( state, capacity );
    for( int q = blockIdx.x * blockDim.x + threadIdx.x; q < count; q += gridDim.x * blockDim.x )
    {
      const int p = list ? list[ q ] : q;
      SwzCell &C = cells[ p ];
      if( !C.active )
        continue;
      if( C.state == 3 )
      {
        C.active = 0;
        continue;
      }
      double dist = C.dist;
      if( C.state == 1 )
        dist = coop::root_of( C.root );
      else if( C.state == 2 )
      {
        int ijk[ 3 ];
        cell_ijk( g, C.cell, ijk );
        const double lo[ 3 ] = { cell_lower( g, 0, ijk[ 0 ] ), cell_lower( g, 1, ijk[ 1 ] ), cell_lower( g, 2, ijk[ 2 ] ) };
        dist = coop::locate_serial( lo, g.h, C.normal, u[ C.cell ] );
      }
      double *rec = hs + (long long) C.cell * 4;
      rec[ 0 ] = C.normal[ 0 ];
      rec[ 1 ] = C.normal[ 1 ];
      rec[ 2 ] = C.normal[ 2 ];
      rec[ 3 ] = dist;
      double dotp = 0.0;
      for( int k = 0; k < 3; ++k )
        dotp += C.normal[ k ] * C.oldNormal[ k ];
      const double residuum = 1.0 - dotp;
      const int it = C.iterations + 1;
      C.iterations = it;
      C.active = ( residuum > 1e-12 && it < maxIterations ) ? 1 : 0;
      if( next && C.active )
        next[ atomicAdd( &listCount[ ( iteration & 1 ) ^ 1 ], 1 ) ] = p;
    }
  }

  // ------------------------------------------------------------------------------------------------------------
  // E1: geometric fluxes, evolution/evolution.hh:90-174.  LANES lanes per mixed cell, one per face.
  // ------------------------------------------------------------------------------------------------------------

  __device__ __forceinline__ void atomic_min_positive ( double *addr, double value )
  {
    // std::min( dtEst, value ) for value >= 0 (NaN never wins, as in std::min): the bit patterns of non-negative
    // doubles are ordered like unsigned integers
    if( value == value )
      atomicMin( reinterpret_cast< unsigned long long * >( addr ), (unsigned long long) __double_as_longlong( value ) );
  }

  template< int DIM >
  __global__ void __launch_bounds__( 128 ) k_flux ( Grid g, const Velocity *__restrict__ velocity, const double *__restrict__ hs,
                                                     const unsigned char *__restrict__ flags, const int *__restrict__ mixedList,
                                                     StepState *state, int capacity, FluxRecord *__restrict__ records,
                                                     const double *timeOverride, const double *dtOverride )
  {
    constexpr int LANES = ( DIM == 3 ) ? 8 : 4;
    constexpr int NFACES = 2 * DIM;
    const int count = mixed_count( state, capacity );
    const double time = timeOverride ? *timeOverride : state->time;
    const double dt = dtOverride ? *dtOverride : state->dt;
    const double volume = cell_volume( g );
    const int groupsPerBlock = blockDim.x / LANES;
    const int lane = threadIdx.x % LANES;
    const int laneInWarp = threadIdx.x & 31;
    const int groupBase = laneInWarp - lane;
    // all groups of a warp iterate the same number of times so that the shuffles below see full warps
    const int nGroupsRounded = ( ( count + 32 / LANES - 1 ) / ( 32 / LANES ) ) * ( 32 / LANES );
    for( int idx = blockIdx.x * groupsPerBlock + threadIdx.x / LANES; idx < nGroupsRounded; idx += gridDim.x * groupsPerBlock )
    {
      const bool active = idx < count;
      double sf = 0.0, selfc = 0.0, nbc = 0.0;
      bool hasNb = false, hasSelf = false, hasNbc = false;
      long long c = 0;
      if( active && lane < NFACES )
      {
        c = mixedList[ idx ];
        int ijk[ 3 ];
        cell_ijk( g, c, ijk );
        const int face = lane;
        int nb[ 3 ] = { ijk[ 0 ], ijk[ 1 ], ijk[ 2 ] };
        nb[ face >> 1 ] += ( face & 1 ) ? 1 : -1;
        // a face has a neighbour iff the neighbour is inside the global domain (intersection.neighbor(), :102-103)
        if( in_domain( g, nb ) )
        {
          hasNb = true;
          const unsigned char flagNb = cell_exists( g, nb ) ? flags[ cell_index( g, nb ) ] : (unsigned char) 99;
          double outerNormal[ 3 ] = { 0.0, 0.0, 0.0 };
          outerNormal[ face >> 1 ] = ( face & 1 ) ? 1.0 : -1.0;
          double v[ 3 ];
          face_velocity( g, *velocity, ijk, face, time, v );
          const double faceVol = face_measure_of( g, face );
          double nv = 0.0;
#pragma unroll
          for( int k = 0; k < DIM; ++k )
            nv += outerNormal[ k ] * v[ k ];
          sf = faceVol * fabs( nv );
#pragma unroll
          for( int k = 0; k < DIM; ++k )
            v[ k ] *= dt;
          double vn = 0.0, vN = 0.0;
#pragma unroll
          for( int k = 0; k < DIM; ++k )
          {
            vn += v[ k ] * outerNormal[ k ];
            vN += v[ k ] * ( outerNormal[ k ] * faceVol );   // v * integrationOuterNormal
          }
          if( vn > 0 )
          {
            double flux;
            const double *rec = hs + c * ( DIM + 1 );
            if( DIM == 2 )
            {
              const Hs2 H = { rec[ 0 ], rec[ 1 ], rec[ 2 ] };
              flux = vofx::flux( cell_lower( g, 0, ijk[ 0 ] ), cell_lower( g, 1, ijk[ 1 ] ), g.h[ 0 ], g.h[ 1 ], face, v[ 0 ], v[ 1 ], H );
            }
            else
            {
              const Hs3 H = { { rec[ 0 ], rec[ 1 ], rec[ 2 ] }, rec[ 3 ] };
              const double lo[ 3 ] = { cell_lower( g, 0, ijk[ 0 ] ), cell_lower( g, 1, ijk[ 1 ] ), cell_lower( g, 2, ijk[ 2 ] ) };
              flux = vofx::flux( lo, g.h, face, v, H );
            }
            hasSelf = true;
            selfc = -( flux / volume );                          // update[ entity ] -= flux / volume
            if( flagNb == 3 )
            {
              hasNbc = true;
              nbc = -( ( vN - flux ) / volume );                 // update[ neighbor ] -= ( v*N - flux ) / volume
            }
            else if( flagNb == 0 || is_mixed( flagNb ) )
            {
              hasNbc = true;
              nbc = flux / volume;                               // update[ neighbor ] += flux / volume
            }
          }
          if( flagNb == 3 && vn < 0 )
          {
            hasSelf = true;
            selfc = -( vN / volume );                            // inflow from a full neighbour
          }
        }
      }

      // time-step estimate: volume / sum of |F| |n.v| over the faces in face order, :113,144
      double sumFluxes = 0.0;
#pragma unroll
      for( int f = 0; f < NFACES; ++f )
      {
        const double sff = __shfl_sync( 0xffffffffu, sf, groupBase + f );
        const int has = __shfl_sync( 0xffffffffu, (int) hasNb, groupBase + f );
        if( has )
          sumFluxes += sff;
      }
      const unsigned selfBits = __ballot_sync( 0xffffffffu, hasSelf ) >> groupBase;
      const unsigned nbBits = __ballot_sync( 0xffffffffu, hasNbc ) >> groupBase;

      if( active )
      {
        FluxRecord *r = records + idx;
        if( lane < NFACES )
        {
          r->self[ lane ] = selfc;
          r->nb[ lane ] = nbc;
        }
        if( lane == 0 )
        {
          r->selfMask = (unsigned char) ( selfBits & ( ( 1u << NFACES ) - 1 ) );
          r->nbMask = (unsigned char) ( nbBits & ( ( 1u << NFACES ) - 1 ) );
          // only cells this rank owns enter the estimate (every cell is owned by exactly one rank)
          int ijk[ 3 ];
          cell_ijk( g, c, ijk );
          const int la = ijk[ DIM - 1 ];
          if( la >= g.own0 && la < g.own1 )
            atomic_min_positive( &state->dtEst, volume / sumFluxes );
        }
      }
    }
  }

  // ------------------------------------------------------------------------------------------------------------
  // E1 in 3D for the B200 (k_flux_cell3 below).  A first version compacted the outflow faces into a task list
  // (k_flux_prepare3 + k_flux_clip3: 0.043 + 0.46 ms at 512^3); the cell-major kernel evaluates nothing twice and
  // passes no task list through memory (0.36 ms).
  // ------------------------------------------------------------------------------------------------------------

  struct FaceMotion
  {
    double v[ 3 ];     // dt * velocity at the face centre
    double sf;         // |F| |n.v| with the unscaled velocity (time-step estimate)
    double vn, vN;     // v . n and v . integrationOuterNormal
  };

  template< int DIM >
  __device__ __forceinline__ void face_motion_rest ( const Grid &g, int face, double dt, FaceMotion &m );

  // with the time factors of the step already evaluated (velocity_time_factors)
  template< int DIM >
  __device__ __forceinline__ void face_motion_tf ( const Grid &g, const Velocity &velocity, const int *ijk, int face, const double *tf, double dt, FaceMotion &m )
  {
    face_velocity_tf( g, velocity, ijk, face, tf, m.v );
    face_motion_rest< DIM >( g, face, dt, m );
  }

  template< int DIM >
  __device__ __forceinline__ void face_motion ( const Grid &g, const Velocity &velocity, const int *ijk, int face, double time, double dt, FaceMotion &m )
  {
    face_velocity( g, velocity, ijk, face, time, m.v );
    face_motion_rest< DIM >( g, face, dt, m );
  }

  template< int DIM >
  __device__ __forceinline__ void face_motion_rest ( const Grid &g, int face, double dt, FaceMotion &m )
  {
    double outerNormal[ 3 ] = { 0.0, 0.0, 0.0 };
    outerNormal[ face >> 1 ] = ( face & 1 ) ? 1.0 : -1.0;
    const double faceVol = face_measure_of( g, face );
    double nv = 0.0;
#pragma unroll
    for( int k = 0; k < DIM; ++k )
      nv += outerNormal[ k ] * m.v[ k ];
    m.sf = faceVol * fabs( nv );
#pragma unroll
    for( int k = 0; k < DIM; ++k )
      m.v[ k ] *= dt;
    m.vn = 0.0;
    m.vN = 0.0;
#pragma unroll
    for( int k = 0; k < DIM; ++k )
    {
      m.vn += m.v[ k ] * outerNormal[ k ];
      m.vN += m.v[ k ] * ( outerNormal[ k ] * faceVol );   // v * integrationOuterNormal
    }
  }

  // E1 in 3D, cell-major: a group of 8 lanes owns one mixed cell.  Lanes 0..5 evaluate the motion of the six faces
  // (velocity, time-step estimate, flux-free contributions); the outflow faces are then clipped one after the other by
  // all 8 lanes (vofx_coop3d.cuh).  Nothing is evaluated twice and no task list goes through memory.
  struct FaceScratch
  {
    double v[ 6 ][ 3 ];
    double vN[ 6 ], sf[ 6 ], selfc[ 6 ];
    unsigned char flagNb[ 6 ], hasNb[ 6 ], clip[ 6 ], hasSelf[ 6 ], hasNbc[ 6 ], shiftedFirst[ 6 ];
    unsigned char pad[ 12 ];
    double bankPad[ 1 ];
  };
  static_assert( sizeof( FaceScratch ) % 8 == 0 && ( sizeof( FaceScratch ) / 8 ) % 2 == 1, "FaceScratch stride must be an odd number of doubles" );

#ifndef VOFX_FLUX_BLOCKS_PER_SM
#define VOFX_FLUX_BLOCKS_PER_SM 4
#endif
  template< int DIM_, int L >
  __global__ void __launch_bounds__( 16 * L, VOFX_FLUX_BLOCKS_PER_SM * 8 / L ) k_flux_cell3 ( Grid g, const Velocity velocity, const double *__restrict__ hs,
                                                              const unsigned char *__restrict__ flags, const int *__restrict__ mixedList,
                                                              StepState *state, int capacity, FluxRecord *__restrict__ records,
                                                              int *__restrict__ tasks, int taskCapacity, const coop::ClipCase *__restrict__ clipTable,
                                                              const double *timeOverride, const double *dtOverride, BandPart bp )
  {
    pdl_prologue();
    constexpr int DIM = 3, LANES = L, NFACES = 6;
    constexpr int GROUPS = 16;          // 128 threads for 8 lanes per cell, 64 for 4
    __shared__ coop::ClipCase table[ 256 ];
    __shared__ coop::ClipScratch scratch[ GROUPS ];
    __shared__ FaceScratch faceScratch[ GROUPS ];
    __shared__ double blockMin[ ( 16 * L + 31 ) / 32 ];
    {
      const uint4 *src = reinterpret_cast< const uint4 * >( clipTable );
      uint4 *dst = reinterpret_cast< uint4 * >( table );
      for( int i = threadIdx.x; i < int( 256 * sizeof( coop::ClipCase ) / sizeof( uint4 ) ); i += blockDim.x )
        dst[ i ] = src[ i ];
    }
    __syncthreads();

    const int count = mixed_count( state, capacity );
    const double time = timeOverride ? *timeOverride : state->time;
    const double dt = dtOverride ? *dtOverride : state->dt;
    const double volume = cell_volume( g );
    double tf[ 3 ];
    velocity_time_factors( velocity, time, tf );
    const int lane = threadIdx.x % LANES;
    const int group = threadIdx.x / LANES;
    const unsigned groupMask = ( ( 1u << LANES ) - 1u ) << ( ( threadIdx.x & 31 ) - lane );
    coop::ClipScratch &S = scratch[ group ];
    FaceScratch &F = faceScratch[ group ];
    double dtMin = DBL_MAX;

    for( int pos = ( blockIdx.x * GROUPS + group ) * bp.parts + bp.part; pos < count; pos += gridDim.x * GROUPS * bp.parts )
    {
      // a band part takes the band-list entries its plane-constant kernel took ( part, part + parts, ... ) in list order:
      // neighbours in the list are neighbours in space with the same outflow faces, which keeps the warp's four cells in step
      const int idx = pos;
      const int c = mixedList[ idx ];
      int ijk[ 3 ];
      cell_ijk( g, c, ijk );
      FluxRecord *r = records + idx;

#pragma unroll 1
      for( int face = lane; face < NFACES; face += LANES )
      {
        bool hasNb = false, hasSelf = false, hasNbc = false, needClip = false;
        double sf = 0.0, selfc = 0.0;
        unsigned char flagNb = 99;
        int nb[ 3 ] = { ijk[ 0 ], ijk[ 1 ], ijk[ 2 ] };
        nb[ face >> 1 ] += ( face & 1 ) ? 1 : -1;
        // a face has a neighbour iff the neighbour is inside the global domain (intersection.neighbor(), :102-103)
        if( in_domain( g, nb ) )
        {
          hasNb = true;
          flagNb = cell_exists( g, nb ) ? flags[ cell_index( g, nb ) ] : (unsigned char) 99;
          FaceMotion m;
          face_motion_tf< DIM >( g, velocity, ijk, face, tf, dt, m );
          sf = m.sf;
          F.v[ face ][ 0 ] = m.v[ 0 ];
          F.v[ face ][ 1 ] = m.v[ 1 ];
          F.v[ face ][ 2 ] = m.v[ 2 ];
          F.vN[ face ] = m.vN;
          if( m.vn > 0 )
          {
            needClip = true;
            hasSelf = true;
            hasNbc = ( flagNb == 3 ) || ( flagNb == 0 ) || is_mixed( flagNb );
            const double lo[ 3 ] = { cell_lower( g, 0, ijk[ 0 ] ), cell_lower( g, 1, ijk[ 1 ] ), cell_lower( g, 2, ijk[ 2 ] ) };
            F.shiftedFirst[ face ] = coop::upwind_shifted_first( lo, g.h, face, m.v ) ? 1 : 0;
          }
          if( flagNb == 3 && m.vn < 0 )
          {
            hasSelf = true;
            selfc = -( m.vN / volume );                          // inflow from a full neighbour
          }
        }
        F.sf[ face ] = sf;
        F.selfc[ face ] = selfc;
        F.flagNb[ face ] = flagNb;
        F.hasNb[ face ] = hasNb;
        F.clip[ face ] = needClip;
        F.hasSelf[ face ] = hasSelf;
        F.hasNbc[ face ] = hasNbc;
        if( !needClip )
        {
          r->self[ face ] = selfc;
          r->nb[ face ] = 0.0;
        }
      }
      __syncwarp( groupMask );

      if( lane == 0 )
      {
        // time-step estimate: volume / sum of |F| |n.v| over the faces in face order, :113,144
        double sumFluxes = 0.0;
        unsigned selfBits = 0, nbBits = 0;
#pragma unroll
        for( int f = 0; f < NFACES; ++f )
        {
          if( F.hasNb[ f ] )
            sumFluxes += F.sf[ f ];
          selfBits |= F.hasSelf[ f ] ? ( 1u << f ) : 0u;
          nbBits |= F.hasNbc[ f ] ? ( 1u << f ) : 0u;
        }
        r->selfMask = (unsigned char) selfBits;
        r->nbMask = (unsigned char) nbBits;
        // only cells this rank owns enter the estimate (every cell is owned by exactly one rank)
        const int la = ijk[ DIM - 1 ];
        if( la >= g.own0 && la < g.own1 )
        {
          const double est = volume / sumFluxes;
          if( est == est && est < dtMin )                        // std::min( dtEst, est ): NaN never wins
            dtMin = est;
        }
      }

      const double *rec = hs + (long long) c * ( DIM + 1 );
      const Hs3 H = { { rec[ 0 ], rec[ 1 ], rec[ 2 ] }, rec[ 3 ] };
      const bool valid = hs_valid( H );
      const double lo[ 3 ] = { cell_lower( g, 0, ijk[ 0 ] ), cell_lower( g, 1, ijk[ 1 ] ), cell_lower( g, 2, ijk[ 2 ] ) };
#pragma unroll 1
      for( int face = 0; face < NFACES; ++face )
      {
        if( !F.clip[ face ] )
          continue;
        double flux = fabs( 0.0 / 3.0 );
        if( valid )
        {
          const double v[ 3 ] = { F.v[ face ][ 0 ], F.v[ face ][ 1 ], F.v[ face ][ 2 ] };
#pragma unroll
          for( int node = lane; node < 8; node += LANES )
          {
            double xn[ 3 ];
            coop::upwind_node_oriented( node, lo, g.h, face, v, F.shiftedFirst[ face ] != 0, xn );
            coop::clip_phase_classify( node, S, xn, H );
          }
          __syncwarp( groupMask );
          const coop::ClipCase *cs;
          int status = 0;
#pragma unroll
          for( int j = lane; j < 8; j += LANES )
            status = coop::clip_phase_cut( j, S, H, table, cs, clipTable + 256 );   // patterns with nodes on the plane: global table
          __syncwarp( groupMask );
          if( status == 0 )
          {
#pragma unroll
            for( int f = lane; f < 8; f += LANES )
              coop::clip_phase_faces( f, S, *cs );
            __syncwarp( groupMask );
            coop::clip_phase_edges< LANES >( lane, S, *cs );
            __syncwarp( groupMask );
#pragma unroll
            for( int f = lane; f < 8; f += LANES )
              coop::clip_phase_face_sums( f, S, *cs );
            __syncwarp( groupMask );
            flux = coop::clip_phase_sum( S, *cs );
          }
          __syncwarp( groupMask );     // the scratch is reused by the next clip
          if( status == 2 )
          {
            // a node on the plane / irregular mask: rare; queued for k_flux_clip_serial3
            if( lane == 0 )
              tasks[ taskCapacity - 1 - atomicAdd( &state->serialCount, 1 ) ] = idx * 8 + face;
            continue;
          }
        }
        if( lane == 0 )
        {
          const unsigned char flagNb = F.flagNb[ face ];
          r->self[ face ] = -( flux / volume );                  // update[ entity ] -= flux / volume
          double nbc = 0.0;
          if( flagNb == 3 )
            nbc = -( ( F.vN[ face ] - flux ) / volume );         // update[ neighbor ] -= ( v*N - flux ) / volume
          else if( flagNb == 0 || is_mixed( flagNb ) )
            nbc = flux / volume;                                 // update[ neighbor ] += flux / volume
          r->nb[ face ] = nbc;
        }
      }
      __syncwarp( groupMask );         // F is rewritten by the next cell
    }

    // one atomic per block
#pragma unroll
    for( int o = 16; o > 0; o >>= 1 )
    {
      const double other = __shfl_xor_sync( 0xffffffffu, dtMin, o );
      dtMin = ( other < dtMin ) ? other : dtMin;
    }
    if( ( threadIdx.x & 31 ) == 0 )
      blockMin[ threadIdx.x >> 5 ] = dtMin;
    __syncthreads();
    if( threadIdx.x == 0 )
    {
      for( int w = 1; w < ( 16 * L ) / 32; ++w )
        dtMin = ( blockMin[ w ] < dtMin ) ? blockMin[ w ] : dtMin;
      if( dtMin < DBL_MAX )
        atomic_min_positive( &state->dtEst, dtMin );
    }
  }

  // the clips k_flux_clip3 could not take from its table (a node of the swept box exactly on the plane): serial walk
  template< int DIM_ >
  __global__ void __launch_bounds__( 64 ) k_flux_clip_serial3 ( Grid g, const Velocity *__restrict__ velocity, const double *__restrict__ hs,
                                                                        const unsigned char *__restrict__ flags, const int *__restrict__ mixedList,
                                                                        const StepState *state, FluxRecord *__restrict__ records, const int *__restrict__ tasks,
                                                                        int taskCapacity, const double *timeOverride, const double *dtOverride )
  {
    pdl_prologue();
    constexpr int DIM = 3;
    const int n = state->serialCount;
    const double time = timeOverride ? *timeOverride : state->time;
    const double dt = dtOverride ? *dtOverride : state->dt;
    const double volume = cell_volume( g );
    for( int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x )
    {
      const int task = tasks[ taskCapacity - 1 - t ];
      const int idx = task >> 3, face = task & 7;
      const int c = mixedList[ idx ];
      int ijk[ 3 ];
      cell_ijk( g, c, ijk );
      FaceMotion m;
      face_motion< DIM >( g, *velocity, ijk, face, time, dt, m );
      const double *rec = hs + (long long) c * ( DIM + 1 );
      const Hs3 H = { { rec[ 0 ], rec[ 1 ], rec[ 2 ] }, rec[ 3 ] };
      const double lo[ 3 ] = { cell_lower( g, 0, ijk[ 0 ] ), cell_lower( g, 1, ijk[ 1 ] ), cell_lower( g, 2, ijk[ 2 ] ) };
      const double flux = vofx::flux( lo, g.h, face, m.v, H );
      int nb[ 3 ] = { ijk[ 0 ], ijk[ 1 ], ijk[ 2 ] };
      nb[ face >> 1 ] += ( face & 1 ) ? 1 : -1;
      const unsigned char flagNb = cell_exists( g, nb ) ? flags[ cell_index( g, nb ) ] : (unsigned char) 99;
      FluxRecord *r = records + idx;
      r->self[ face ] = -( flux / volume );
      double nbc = 0.0;
      if( flagNb == 3 )
        nbc = -( ( m.vN - flux ) / volume );
      else if( flagNb == 0 || is_mixed( flagNb ) )
        nbc = flux / volume;
      r->nb[ face ] = nbc;
    }
  }

  // ------------------------------------------------------------------------------------------------------------
  // E1 + U1: owner-computes gather.  LANES lanes per mixed cell: lane 0 = the cell itself,
  // lanes 1..2*DIM = its face neighbours; a non-mixed neighbour X is handled by its first mixed face neighbour in
  // the order z-,y-,x-,x+,y+,z+.  The contributions are summed in the order the reference's scatter loop produces:
  // mixed neighbours with smaller index, the cell's own faces in face order, mixed neighbours with larger index.
  // ------------------------------------------------------------------------------------------------------------

  template< int DIM >
  __device__ __forceinline__ int gather_face ( int k )
  {
    // k-th entry of the order z-,y-,x-,x+,y+,z+ as a face number of X (2D: y-,x-,x+,y+)
    if( DIM == 3 )
    {
      const int order[ 6 ] = { 4, 2, 0, 1, 3, 5 };
      return order[ k ];
    }
    const int order[ 4 ] = { 2, 0, 1, 3 };
    return order[ k ];
  }

  template< int DIM >
  __global__ void __launch_bounds__( 128 ) k_update ( Grid g, const unsigned char *__restrict__ flags, const int *__restrict__ mixedList,
                                                       const int *__restrict__ slot, const StepState *state, int capacity,
                                                       const FluxRecord *__restrict__ records, double *__restrict__ target, int accumulate,
                                                       StepState *stateRW, int *__restrict__ changedIdx, double *__restrict__ changedVal, int changedCapacity,
                                                       unsigned char *__restrict__ marks, int segsPerRow, CommPeers peers )
  {
    pdl_prologue();
    constexpr int LANES = ( DIM == 3 ) ? 8 : 8;
    constexpr int NFACES = 2 * DIM;
    const int count = mixed_count( state, capacity );
    const int groupsPerBlock = blockDim.x / LANES;
    const int lane = threadIdx.x % LANES;
    for( int idx = blockIdx.x * groupsPerBlock + threadIdx.x / LANES; idx < count; idx += gridDim.x * groupsPerBlock )
    {
      const long long m = mixedList[ idx ];
      int mijk[ 3 ];
      cell_ijk( g, m, mijk );
      if( lane > NFACES )
        continue;
      int x[ 3 ] = { mijk[ 0 ], mijk[ 1 ], mijk[ 2 ] };
      if( lane > 0 )
      {
        const int face = lane - 1;
        x[ face >> 1 ] += ( face & 1 ) ? 1 : -1;
        if( !cell_exists( g, x ) )
          continue;
      }
      const int la = x[ DIM - 1 ];
      if( la < g.own0 || la >= g.own1 )
        continue;                                  // not owned: another rank updates it
      const long long X = cell_index( g, x );
      const unsigned char flagX = flags[ X ];
      // The six neighbour records of X are fetched level by level (flags, slots, records: independent loads in flight)
      // and only then summed in the reference's order; the dependent flag -> slot -> record chains used to run one after
      // the other (long-scoreboard stalls: 16 of 22 cycles per issue).  k runs over z-,y-,x-,x+,y+,z+.
      long long Yk[ NFACES ];
      bool mixedK[ NFACES ];
      unsigned char flagK[ NFACES ];
#pragma unroll
      for( int k = 0; k < NFACES; ++k )
      {
        const int f = gather_face< DIM >( k );
        int y[ 3 ] = { x[ 0 ], x[ 1 ], x[ 2 ] };
        y[ f >> 1 ] += ( k < DIM ) ? -1 : 1;
        const bool exists = cell_exists( g, y );
        Yk[ k ] = exists ? cell_index( g, y ) : -1;
      }
#pragma unroll
      for( int k = 0; k < NFACES; ++k )
        flagK[ k ] = ( Yk[ k ] >= 0 ) ? flags[ Yk[ k ] ] : (unsigned char) 0;
      if( lane > 0 )
      {
        if( is_mixed( flagX ) )
          continue;                                // a mixed cell gathers for itself (lane 0 of its own group)
        // is m the first mixed neighbour of X in gather order?
        bool first = true;
#pragma unroll
        for( int k = 0; k < NFACES; ++k )
        {
          if( Yk[ k ] < 0 )
            continue;
          if( Yk[ k ] == m )
            break;
          const int f = gather_face< DIM >( k );
          const int ly = x[ DIM - 1 ] + ( ( ( f >> 1 ) == DIM - 1 ) ? ( ( k < DIM ) ? -1 : 1 ) : 0 );
          if( is_mixed( flagK[ k ] ) && ly >= g.list0 && ly < g.list1 )
          {
            first = false;
            break;
          }
        }
        if( !first )
          continue;
      }

      int slotK[ NFACES ];
#pragma unroll
      for( int k = 0; k < NFACES; ++k )
      {
        mixedK[ k ] = is_mixed( flagK[ k ] );
        slotK[ k ] = mixedK[ k ] ? slot[ Yk[ k ] ] : 0;
      }
      const int slotX = is_mixed( flagX ) ? slot[ X ] : 0;
      double nbK[ NFACES ];
      bool hasK[ NFACES ];
#pragma unroll
      for( int k = 0; k < NFACES; ++k )
      {
        const int f = gather_face< DIM >( k );
        const int toward = ( k < DIM ) ? f + 1 : f - 1;   // Y's face towards X: the opposite face of the same axis
        hasK[ k ] = false;
        nbK[ k ] = 0.0;
        if( mixedK[ k ] )
        {
          const FluxRecord &r = records[ slotK[ k ] ];
          hasK[ k ] = ( r.nbMask & ( 1u << toward ) ) != 0;
          nbK[ k ] = r.nb[ toward ];
        }
      }

      double acc = 0.0;
      // mixed neighbours with smaller index: z-, y-, x-
#pragma unroll
      for( int k = 0; k < DIM; ++k )
        if( hasK[ k ] )
          acc += nbK[ k ];
      if( is_mixed( flagX ) )
      {
        const FluxRecord &r = records[ slotX ];
        const unsigned selfMask = r.selfMask;
        double selfK[ NFACES ];
#pragma unroll
        for( int f = 0; f < NFACES; ++f )
          selfK[ f ] = r.self[ f ];
#pragma unroll
        for( int f = 0; f < NFACES; ++f )
          if( selfMask & ( 1u << f ) )
            acc += selfK[ f ];
      }
#pragma unroll
      for( int k = DIM; k < NFACES; ++k )
        if( hasK[ k ] )
          acc += nbK[ k ];

      if( accumulate )
      {
        const double value = target[ X ] + 1.0 * acc;   // uh.axpy( 1.0, update ), colorfunction.hh:31-37
        target[ X ] = value;
        // distributed step: the neighbour ranks' halo copies of this cell are kept up to date by direct stores into their
        // memory (NVLink peer mapping) -- the colour function only ever changes here, so no halo exchange is needed
        if( peers.uLo && la < peers.pushLo1 )
          peers.uLo[ X + peers.offLo ] = value;
        if( peers.uHi && la >= peers.pushHi0 )
          peers.uHi[ X + peers.offHi ] = value;
        if( changedIdx )
        {
          // every cell is updated by exactly one lane per step: the list of (cell, new value) pairs is what a host
          // colour function needs to be brought up to date (vofx_step_host); one atomic per group of converged lanes
          const unsigned active = __activemask();
          const int leader = __ffs( active ) - 1;
          const int laneInWarp = threadIdx.x & 31;
          int base = 0;
          if( laneInWarp == leader )
            base = atomicAdd( &stateRW->changedCount, __popc( active ) );
          base = __shfl_sync( active, base, leader );
          const int pos = base + __popc( active & ( ( 1u << laneInWarp ) - 1 ) );
          if( pos < changedCapacity )
          {
            changedIdx[ pos ] = int( X );
            changedVal[ pos ] = value;
          }
        }
      }
      else
        target[ X ] = acc;                         // the dense `update` of the operator-level API
    }
  }

  // T1: algorithm.hh:85-93
  __device__ __forceinline__ void finalize_state ( StepState *state, double cfl );

  // ---- signals of the distributed step (one block of 32 threads; thread r talks to rank r) -------------------------------
  // `what`: 1 halo wait, 2 flux wait, 3 finalize wait (reported with the peer rank and the pass number on a time-out)
  __device__ __forceinline__ void comm_wait ( StepState *state, volatile int *flag, int atLeast, int what, int peer )
  {
    const long long t0 = clock64();
    while( *flag < atLeast )
    {
      if( clock64() - t0 > 20000000000LL )       // ~10 s: a peer died or the call sequences of the ranks differ
      {
        state->commError = what | ( peer << 4 ) | ( atLeast << 8 );
        break;
      }
      __nanosleep( 64 );
    }
  }

  // before the first read of the colour halos in a pass: the neighbours' updates of the previous pass have landed
  static __global__ void k_comm_wait_halo ( CommPeers P, StepState *state )
  {
    const int r = threadIdx.x;
    if( r < P.world && ( r == P.rank - 1 || r == P.rank + 1 ) )
      comm_wait( state, P.slots[ P.rank ]->updateDone + r, state->epoch - 1, 1, r );
  }

  // after the fluxes: publish my time-step estimate and "fluxes done" to every rank, then wait until the neighbours have
  // finished theirs (they no longer read the halo copies my update is about to overwrite)
  static __global__ void k_comm_flux_done ( CommPeers P, StepState *state )
  {
    const int r = threadIdx.x;
    const int e = state->epoch;
    if( r < P.world )
    {
      P.slots[ r ]->dtEst[ e & 1 ][ P.rank ] = state->dtEst;
      __threadfence_system();
      *( (volatile int *) ( P.slots[ r ]->fluxDone + P.rank ) ) = e;
    }
    if( r < P.world && ( r == P.rank - 1 || r == P.rank + 1 ) )
      comm_wait( state, P.slots[ P.rank ]->fluxDone + r, e, 2, r );
  }

  // T1 of the distributed step: tell the neighbours that my update (and its pushes into their halos) is done, take the
  // MIN of all ranks' time-step estimates (gridView().comm().min, evolution/evolution.hh:75), then k_finalize's work
  static __global__ void k_finalize_dist ( CommPeers P, StepState *state, double cfl )
  {
    const int r = threadIdx.x;
    const int e = state->epoch;
    if( r < P.world && ( r == P.rank - 1 || r == P.rank + 1 ) )
    {
      __threadfence_system();
      *( (volatile int *) ( P.slots[ r ]->updateDone + P.rank ) ) = e;
    }
    double mine = DBL_MAX;
    if( r < P.world )
    {
      comm_wait( state, P.slots[ P.rank ]->fluxDone + r, e, 3, r );
      mine = *( (volatile double *) &P.slots[ P.rank ]->dtEst[ e & 1 ][ r ] );
    }
#pragma unroll
    for( int o = 16; o > 0; o >>= 1 )
    {
      const double other = __shfl_xor_sync( 0xffffffffu, mine, o );
      mine = ( other < mine ) ? other : mine;
    }
    if( r == 0 )
    {
      state->dtEst = mine;
      finalize_state( state, cfl );
    }
  }

  // bulk copy of my boundary layers into the neighbours' halos (start of a loop, after an upload): layers are contiguous
  static __global__ void __launch_bounds__( 256 ) k_halo_push ( CommPeers P, const double *__restrict__ u, long long layerCells, int own0, int own1 )
  {
    const long long nLo = P.uLo ? (long long) ( P.pushLo1 - own0 ) * layerCells : 0;
    const long long nHi = P.uHi ? (long long) ( own1 - P.pushHi0 ) * layerCells : 0;
    for( long long t = blockIdx.x * (long long) blockDim.x + threadIdx.x; t < nLo + nHi; t += (long long) gridDim.x * blockDim.x )
    {
      if( t < nLo )
      {
        const long long X = own0 * layerCells + t;
        P.uLo[ X + P.offLo ] = u[ X ];
      }
      else
      {
        const long long X = P.pushHi0 * layerCells + ( t - nLo );
        P.uHi[ X + P.offHi ] = u[ X ];
      }
    }
  }

  static __global__ void k_finalize ( StepState *state, double cfl )
  {
    pdl_prologue();
    finalize_state( state, cfl );
  }

  __device__ __forceinline__ void finalize_state ( StepState *state, double cfl )
  {
    state->epoch += 1;
    state->overflowLast = state->overflow;       // sticky (vofx_step_begin clears it): the pass was a no-op, see mixed_count()
    if( !state->overflow )
    {
      if( state->dt > 0.0 )
        state->time += state->dt;
      state->dt = state->dtEst * cfl;
      state->dtEstLast = state->dtEst;
    }
    state->dtEst = DBL_MAX;
    state->mixedCountLast = state->mixedCount;
    state->mixedCount = 0;
    state->changedCountLast = state->changedCount;
    state->changedCount = 0;
    state->taskCount = 0;
    state->serialCount = 0;
    state->serialLocate = 0;
    state->serialLocateMore[ 0 ] = state->serialLocateMore[ 1 ] = state->serialLocateMore[ 2 ] = 0;
    state->rootCount = 0;
    for( int c = 0; c < 32; ++c )
      state->costHist[ c ] = state->costFill[ c ] = 0;
  }

  static __global__ void k_reset_counters ( StepState *state )
  {
    state->overflow = 0;          // operator-level calls are self-contained: each reports its own overflow
    state->mixedCount = 0;
    state->taskCount = 0;
    state->serialCount = 0;
    state->serialLocate = 0;
    state->serialLocateMore[ 0 ] = state->serialLocateMore[ 1 ] = state->serialLocateMore[ 2 ] = 0;
    state->rootCount = 0;
    state->changedCount = 0;
    for( int c = 0; c < 32; ++c )
      state->costHist[ c ] = state->costFill[ c ] = 0;
    state->dtEst = DBL_MAX;
  }

  // Processing order of the band for the plane-constant kernel.  The kernel writes a cost class per cell (costClass: the
  // bracket 0 .. 6 its plane-constant solve ended in, 7 if lane 0 walked the topology for tied node heights); the next step
  // sorts the band list by that class, expensive first (a counting sort: k_cost_hist, then k_cost_order places the entries
  // behind the prefix of the histogram).  Eight cells share a warp and leave the bracket loop one by one: with cells of one
  // class per warp the lanes of finished cells no longer idle (the bracket phases ran at 13 .. 21 of 32 lanes).  Temporal
  // coherence does the prediction; a pure permutation of the band list: no result depends on it.
  // Band parts (BandPart): part p owns the list entries p, p + parts, ...; the order array holds part 0's entries first, each
  // part's entries sorted by class.  bucket = part * 8 + class.
  static __global__ void __launch_bounds__( 256 ) k_cost_hist ( const int *__restrict__ mixedList, StepState *state, int capacity,
                                                                 const unsigned char *__restrict__ costClass, int parts )
  {
    pdl_prologue();
    __shared__ int hist[ 32 ];
    if( threadIdx.x < 32 )
      hist[ threadIdx.x ] = 0;
    __syncthreads();
    const int count = mixed_count( state, capacity );
    for( int idx = blockIdx.x * 256 + threadIdx.x; idx < count; idx += gridDim.x * 256 )
      atomicAdd( &hist[ ( idx % parts ) * 8 + ( costClass[ mixedList[ idx ] ] & 7 ) ], 1 );
    __syncthreads();
    if( threadIdx.x < 32 && hist[ threadIdx.x ] )
      atomicAdd( &state->costHist[ threadIdx.x ], hist[ threadIdx.x ] );
  }

  static __global__ void __launch_bounds__( 256 ) k_cost_order ( const int *__restrict__ mixedList, StepState *state, int capacity,
                                                                  const unsigned char *__restrict__ costClass, int *__restrict__ order, int parts )
  {
    pdl_prologue();
    __shared__ int warpCount[ 32 ][ 8 ], tileBase[ 32 ], start[ 32 ];
    const int count = mixed_count( state, capacity );
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if( threadIdx.x == 0 )
    {
      // part by part; within a part class 7 first, class 0 last
      int at = 0;
      for( int p = 0; p < 4; ++p )
        for( int c = 7; c >= 0; --c )
        {
          start[ p * 8 + c ] = at;
          at += state->costHist[ p * 8 + c ];
        }
    }
    __syncthreads();
    const int nBuckets = parts * 8;
    const int rounded = ( count + 255 ) & ~255;
    for( int first = blockIdx.x * 256; first < rounded; first += gridDim.x * 256 )
    {
      const int idx = first + threadIdx.x;
      const bool valid = idx < count;
      const int bucket = valid ? ( idx % parts ) * 8 + ( costClass[ mixedList[ idx ] ] & 7 ) : 32 + warp;   // padding lanes: a bucket of their own
      warpCount[ threadIdx.x >> 3 ][ threadIdx.x & 7 ] = 0;
      __syncthreads();
      // the lanes of a warp that share a bucket find each other in one instruction
      const unsigned peers = __match_any_sync( 0xffffffffu, bucket );
      const int below = __popc( peers & ( ( 1u << lane ) - 1u ) );
      if( valid && below == 0 )
        warpCount[ bucket ][ warp ] = __popc( peers );
      __syncthreads();
      if( threadIdx.x < nBuckets )
      {
        // one atomic per bucket and 256 entries; the warps' shares become exclusive prefixes
        const int b = threadIdx.x;
        int n = 0;
        for( int w = 0; w < 8; ++w )
        {
          const int a = warpCount[ b ][ w ];
          warpCount[ b ][ w ] = n;
          n += a;
        }
        tileBase[ b ] = n ? atomicAdd( &state->costFill[ b ], n ) : 0;
      }
      __syncthreads();
      if( valid )
      {
        const int pos = start[ bucket ] + tileBase[ bucket ] + warpCount[ bucket ][ warp ] + below;
        if( pos < count )                // the histogram and the classes were read from the same arrays: always true
          order[ pos ] = idx;
      }
      __syncthreads();
    }
  }

  // U1 (operator-level API): ColorFunction::axpy, colorfunction.hh:31-37
  static __global__ void __launch_bounds__( 256 ) k_axpy ( long long n, double *__restrict__ u, double a, const double *__restrict__ x )
  {
    for( long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x )
      u[ i ] += a * x[ i ];
  }

  // vofx_step_host_sparse: the cells the caller changed on the host since the previous call, written into the device mirror
  static __global__ void __launch_bounds__( 256 ) k_scatter_values ( long long n, const int *__restrict__ idx, const double *__restrict__ val, double *__restrict__ u )
  {
    for( long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x )
      u[ idx[ i ] ] = val[ i ];
  }

} // namespace vofx

namespace vofx
{

  // ------------------------------------------------------------------------------------------------------------
  // C1: height-function curvature, curvature/cartesianheightfunctioncurvature.hh:54-187
  // ------------------------------------------------------------------------------------------------------------

  // effectiveTdown, heightfunctionstencil.hh:47-53
  template< int DIM >
  __device__ __forceinline__ int hf_effective_tdown ( const Grid &g, const int *ijk, int i, int j, int tup )
  {
    constexpr int noc = ( DIM == 2 ) ? 3 : 9;
    for( int t = -1; t >= -tup; --t )
      if( hf_cell< DIM >( g, ijk, i, j, ( noc - 1 ) / 2, t ) < 0 )
        return -1 - t;
    return tup;
  }

  // pass 1 (applyLocal, :94-123): per mixed cell the raw curvature and whether the stencil satisfies the constraint
  template< int DIM >
  __global__ void __launch_bounds__( 128 ) k_curvature_local ( Grid g, const double *__restrict__ u, const double *__restrict__ hs,
                                                                const int *__restrict__ mixedList, const StepState *state, int capacity,
                                                                double *__restrict__ curv1, unsigned char *__restrict__ ok1 )
  {
    constexpr int noc = ( DIM == 2 ) ? 3 : 9;
    const int count = mixed_count( state, capacity );
    for( int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x )
    {
      const long long c = mixedList[ idx ];
      int ijk[ 3 ];
      cell_ijk( g, c, ijk );
      const double *rec = hs + c * ( DIM + 1 );
      const double normal[ 3 ] = { rec[ 0 ], rec[ 1 ], DIM == 3 ? rec[ 2 ] : 0.0 };
      int dir, sign;
      hf_orientation< DIM >( normal, dir, sign );
      double heights[ 9 ];
      hf_heights< DIM >( g, u, ijk, dir, sign, 1, heights );

      double kappa = 0.0;
      unsigned char ok = 0;
      const double uMid = heights[ ( noc - 1 ) / 2 ];
      const int effTdown = hf_effective_tdown< DIM >( g, ijk, dir, sign, 1 );
      if( !( uMid < effTdown || uMid > effTdown + 1 ) )
      {
        const double deltaX = g.h[ DIM - dir - 1 ];   // jacobianTransposed[dim-o-1][dim-o-1], :111-113
        bool anyZero = false;
        for( int i = 0; i < noc; ++i )
          if( heights[ i ] == 0.0 )
            anyZero = true;
        if( !anyZero )
        {
          ok = 1;
          if( DIM == 2 )
          {
            const double Hx = ( heights[ 2 ] - heights[ 0 ] ) / 2.0;
            const double Hxx = ( heights[ 2 ] - 2 * heights[ 1 ] + heights[ 0 ] ) / deltaX;
            kappa = -Hxx / pow( 1.0 + Hx * Hx, 3.0 / 2.0 );
          }
          else
          {
            const double Hx = ( heights[ 5 ] - heights[ 3 ] ) / 2.0;
            const double Hy = ( heights[ 7 ] - heights[ 1 ] ) / 2.0;
            const double Hxx = ( heights[ 5 ] - 2.0 * heights[ 4 ] + heights[ 3 ] ) / deltaX;
            const double Hyy = ( heights[ 7 ] - 2.0 * heights[ 4 ] + heights[ 1 ] ) / deltaX;
            const double Hxy = ( heights[ 8 ] - heights[ 2 ] - heights[ 6 ] + heights[ 0 ] ) / ( 4.0 * deltaX );
            kappa = -( Hxx + Hyy + Hxx * Hy * Hy + Hyy * Hx * Hx - 2.0 * Hxy * Hx * Hy ) / ( pow( 1.0 + Hx * Hx + Hy * Hy, 3.0 / 2.0 ) );
          }
        }
      }
      curv1[ idx ] = kappa;
      ok1[ idx ] = ok;
    }
  }

  // pass 2 (averageCurvature with firstTurn, :156-187): keep the own value if valid, else the mean over the valid
  // mixed vertex neighbours (ascending index order)
  template< int DIM >
  __global__ void __launch_bounds__( 128 ) k_curvature_average ( Grid g, const unsigned char *__restrict__ flags, const int *__restrict__ mixedList,
                                                                  const int *__restrict__ slot, const StepState *state, int capacity,
                                                                  const double *__restrict__ curv1, const unsigned char *__restrict__ ok1,
                                                                  double *__restrict__ kappa, unsigned char *__restrict__ valid )
  {
    const int count = mixed_count( state, capacity );
    for( int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x )
    {
      const long long c = mixedList[ idx ];
      if( valid )
        valid[ c ] = ok1[ idx ];
      if( ok1[ idx ] )
      {
        kappa[ c ] = curv1[ idx ];
        continue;
      }
      int ijk[ 3 ];
      cell_ijk( g, c, ijk );
      double sum = 0.0;
      int n = 0;
      for( int dz = ( DIM == 3 ? -1 : 0 ); dz <= ( DIM == 3 ? 1 : 0 ); ++dz )
        for( int dy = -1; dy <= 1; ++dy )
          for( int dx = -1; dx <= 1; ++dx )
          {
            if( dx == 0 && dy == 0 && dz == 0 )
              continue;
            const int nb[ 3 ] = { ijk[ 0 ] + dx, ijk[ 1 ] + dy, ijk[ 2 ] + dz };
            if( !cell_exists( g, nb ) )
              continue;
            const long long cn = cell_index( g, nb );
            const int ln = nb[ DIM - 1 ];
            if( !is_mixed( flags[ cn ] ) || ln < g.list0 || ln >= g.list1 )
              continue;
            const int s = slot[ cn ];
            if( ok1[ s ] )
            {
              sum += curv1[ s ];
              n++;
            }
          }
      if( n > 0 )
        sum /= (double) n;
      kappa[ c ] = sum;
    }
  }

  // ------------------------------------------------------------------------------------------------------------
  // SwartzCurvature (2D), curvature/swartzcurvature.hh:41-173
  // ------------------------------------------------------------------------------------------------------------

  // interface( entity, reconstructions ) of the cell ijk: centroid and length of the segment, 2d/polygon.hh:183-197
  __device__ __forceinline__ void swc_interface ( const Grid &g, const double *__restrict__ hs, const int *ijk, double *centroid, double &length )
  {
    Poly2 p;
    make_cell( cell_lower( g, 0, ijk[ 0 ] ), cell_lower( g, 1, ijk[ 1 ] ), g.h[ 0 ], g.h[ 1 ], p );
    const double *rec = hs + cell_index( g, ijk ) * 3;
    const Hs2 H = { rec[ 0 ], rec[ 1 ], rec[ 2 ] };
    double lx[ 2 ], ly[ 2 ];
    plane_cut( p, H, lx, ly );
    centroid[ 0 ] = ( lx[ 0 ] + lx[ 1 ] ) * 0.5;
    centroid[ 1 ] = ( ly[ 0 ] + ly[ 1 ] ) * 0.5;
    length = norm2( lx[ 0 ] - lx[ 1 ], ly[ 0 ] - ly[ 1 ] );
  }

  // pass 1 (applyLocal, :75-152): one thread per mixed cell; the interfaces of the <= 8 mixed neighbours are cut once
  static __global__ void __launch_bounds__( 128 ) k_swartz_curv_local ( Grid g, const double *__restrict__ hs, const unsigned char *__restrict__ flags,
                                                                         const int *__restrict__ mixedList, const StepState *state, int capacity,
                                                                         double *__restrict__ curv1 )
  {
    const int count = mixed_count( state, capacity );
    for( int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x )
    {
      const long long c = mixedList[ idx ];
      int ijk[ 3 ];
      cell_ijk( g, c, ijk );
      const double nEn[ 2 ] = { hs[ c * 3 ], hs[ c * 3 + 1 ] };
      double cEn[ 2 ], muEn;
      swc_interface( g, hs, ijk, cEn, muEn );
      muEn *= muEn / 12.0;

      double cen[ 8 ][ 2 ], mu[ 8 ];
      bool mixedNb[ 8 ];
      for( int s = 0; s < 8; ++s )
      {
        const int m = ( s < 4 ) ? s : s + 1;
        const int nb[ 3 ] = { ijk[ 0 ] + ( m % 3 ) - 1, ijk[ 1 ] + ( m / 3 ) - 1, 0 };
        mixedNb[ s ] = cell_exists( g, nb ) && is_mixed( flags[ cell_index( g, nb ) ] );
        cen[ s ][ 0 ] = cen[ s ][ 1 ] = mu[ s ] = 0.0;
        if( mixedNb[ s ] )
        {
          swc_interface( g, hs, nb, cen[ s ], mu[ s ] );
          mu[ s ] *= mu[ s ] / 12.0;
        }
      }

      double curv = 0.0, weights = 0.0;
      for( int s1 = 0; s1 < 8; ++s1 )
      {
        if( !mixedNb[ s1 ] )
          continue;
        const double midway1[ 2 ] = { ( cen[ s1 ][ 0 ] + cEn[ 0 ] ) * 0.5, ( cen[ s1 ][ 1 ] + cEn[ 1 ] ) * 0.5 };
        const double triangle1 = norm2( cen[ s1 ][ 0 ] - cEn[ 0 ], cen[ s1 ][ 1 ] - cEn[ 1 ] );
        double tan1[ 2 ] = { cEn[ 0 ] - cen[ s1 ][ 0 ], cEn[ 1 ] - cen[ s1 ][ 1 ] };
        const double t1 = norm2( tan1[ 0 ], tan1[ 1 ] );
        tan1[ 0 ] /= t1;
        tan1[ 1 ] /= t1;
        const double nu12[ 2 ] = { -tan1[ 1 ], tan1[ 0 ] };
        if( dot2( nu12[ 0 ], nu12[ 1 ], nEn[ 0 ], nEn[ 1 ] ) < 0 )
          continue;
        for( int s2 = 0; s2 < 8; ++s2 )
        {
          if( !mixedNb[ s2 ] || s2 == s1 )
            continue;
          const double midway2[ 2 ] = { ( cen[ s2 ][ 0 ] + cEn[ 0 ] ) * 0.5, ( cen[ s2 ][ 1 ] + cEn[ 1 ] ) * 0.5 };
          const double triangle2 = norm2( cen[ s2 ][ 0 ] - cEn[ 0 ], cen[ s2 ][ 1 ] - cEn[ 1 ] );
          const double md[ 2 ] = { midway2[ 0 ] - midway1[ 0 ], midway2[ 1 ] - midway1[ 1 ] };
          const double deltaS = norm2( md[ 0 ], md[ 1 ] );
          double tan2[ 2 ] = { cen[ s2 ][ 0 ] - cEn[ 0 ], cen[ s2 ][ 1 ] - cEn[ 1 ] };
          const double t2 = norm2( tan2[ 0 ], tan2[ 1 ] );
          tan2[ 0 ] /= t2;
          tan2[ 1 ] /= t2;
          const double nu23[ 2 ] = { -tan2[ 1 ], tan2[ 0 ] };
          if( dot2( nu23[ 0 ], nu23[ 1 ], nEn[ 0 ], nEn[ 1 ] ) < 0 )
            continue;
          if( dot2( md[ 0 ], md[ 1 ], tan1[ 0 ] + tan2[ 0 ], tan1[ 1 ] + tan2[ 1 ] ) < 0 )
            continue;
          const double M = ( ( mu[ s2 ] - muEn ) / triangle2 - ( muEn - mu[ s1 ] ) / triangle1 ) / ( 2.0 * deltaS );
          const double weight = 1.0;
          curv += weight * ( nu12[ 0 ] * nu23[ 1 ] - nu12[ 1 ] * nu23[ 0 ] ) / ( ( 1.0 + M ) * deltaS );
          weights += weight;
        }
      }
      if( weights > 0 )
        curv /= weights;
      curv1[ idx ] = curv;
    }
  }

  // pass 2 (:57-66,154-170): cells whose curvature is exactly 0 take the mean over all mixed vertex neighbours
  static __global__ void __launch_bounds__( 128 ) k_swartz_curv_average ( Grid g, const unsigned char *__restrict__ flags, const int *__restrict__ mixedList,
                                                                           const int *__restrict__ slot, const StepState *state, int capacity,
                                                                           const double *__restrict__ curv1, double *__restrict__ kappa )
  {
    const int count = mixed_count( state, capacity );
    for( int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x )
    {
      const long long c = mixedList[ idx ];
      double value = curv1[ idx ];
      if( value == 0.0 )
      {
        int ijk[ 3 ], n = 0;
        cell_ijk( g, c, ijk );
        for( int s = 0; s < 8; ++s )
        {
          const int m = ( s < 4 ) ? s : s + 1;
          const int nb[ 3 ] = { ijk[ 0 ] + ( m % 3 ) - 1, ijk[ 1 ] + ( m / 3 ) - 1, 0 };
          if( !cell_exists( g, nb ) )
            continue;
          const long long cn = cell_index( g, nb );
          if( !is_mixed( flags[ cn ] ) )
            continue;
          value += curv1[ slot[ cn ] ];
          n++;
        }
        if( n > 0 )
          value /= (double) n;
      }
      kappa[ c ] = value;
    }
  }

  // ------------------------------------------------------------------------------------------------------------
  // initial data: RecursiveInterpolationCube depth 3 (interpolation/recursiveinterpolationcube.hh:33-82) of the
  // built-in problems' indicator functions (test/problems, Problem::evaluate).  Transcendental time factors are
  // evaluated on the host and passed in, so the device only does + - * sqrt and comparisons.
  // ------------------------------------------------------------------------------------------------------------


  __device__ __forceinline__ int indicator_eval ( const Indicator &I, const double *x )
  {
    double uu = 0.0;
    if( I.kind == 0 )
    {
      // rotatingcircle.hh:81-90, LeVeque (oracle/problems_extra.hh): ( x - center ).two_norm() < r
      double s = 0.0;
      for( int k = 0; k < I.dim; ++k )
      {
        const double d = x[ k ] - I.c[ k ];
        s += d * d;
      }
      uu = ( sqrt( s ) < I.r ) ? 1.0 : 0.0;
    }
    else if( I.kind == 1 )
    {
      // sflow.hh:26-33,63-71: y = c - x; y.two_norm2() <= r*r
      double s = 0.0;
      for( int k = 0; k < I.dim; ++k )
      {
        const double d = I.c[ k ] - x[ k ];
        s += d * d;
      }
      uu = ( s <= I.r * I.r ) ? 1.0 : 0.0;
    }
    else if( I.kind == 2 )
    {
      // slottedcylinder.hh:22-45
      const double cx = 0.5, cy = 0.5;
      double xpx = cx, xpy = cy;
      const double tx = x[ 0 ] - cx, ty = x[ 1 ] - cy;
      xpx += I.cs * tx;
      xpy += I.cs * ty;
      const double r0 = -ty, r1 = tx;
      xpx += I.sn * r0;
      xpy += I.sn * r1;
      const double slotWidth = 0.075;
      double dx = xpx - cx, dy = xpy - cy;
      if( norm2( dx, dy ) < 0.4 )
        uu = 1.0;
      const double ccx = cx + 0.0, ccy = cy + slotWidth;
      dx = xpx - ccx;
      dy = xpy - ccy;
      if( norm2( dx, dy ) < slotWidth )
        uu = 0.0;
      if( xpy > cy + slotWidth && fabs( xpx - cx ) < slotWidth )
        uu = 0.0;
    }
    else
      uu = ( x[ 0 ] <= I.c[ 0 ] ) ? 1.0 : 0.0;   // linearwall.hh:17-20
    return (int) ( uu + 0.3 );
  }

  // LEVEL is a template parameter so that the recursion depth (3) is static: no device call stack needed
  template< int DIM, int LEVEL >
  __device__ double recursive_evaluation ( const Grid &g, const Indicator &I, const double *cellLo, const double *origin, double h )
  {
    constexpr int corners = 1 << DIM;
    int sum = 0;
    for( int i = 0; i < corners; ++i )
    {
      double x[ 3 ] = { 0.0, 0.0, 0.0 };
      for( int j = 0; j < DIM; ++j )
      {
        double corner = origin[ j ];
        corner += h * (double) ( ( i >> j ) & 1 );
        x[ j ] = cellLo[ j ];
        x[ j ] += corner * g.h[ j ];
      }
      sum += indicator_eval( I, x );
    }
    double vol = 1.0;
    for( int i = 0; i < DIM; ++i )
      vol *= h;
    if( LEVEL == 0 )
      return vol * ( (double) sum / (double) corners );
    if( sum == corners )
      return vol;
    if( sum == 0 )
      return 0.0;
    h /= 2;
    double value = 0.0;
    for( int i = 0; i < corners; ++i )
    {
      double childOrigin[ 3 ] = { 0.0, 0.0, 0.0 };
      for( int j = 0; j < DIM; ++j )
      {
        childOrigin[ j ] = origin[ j ];
        childOrigin[ j ] += h * (double) ( ( i >> j ) & 1 );
      }
      value += recursive_evaluation< DIM, ( LEVEL > 0 ? LEVEL - 1 : 0 ) >( g, I, cellLo, childOrigin, h );
    }
    return value;
  }

  template< int DIM >
  __global__ void __launch_bounds__( 128 ) k_init ( Grid g, Indicator I, double *__restrict__ u )
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
      double lo[ 3 ] = { 0.0, 0.0, 0.0 };
      for( int d = 0; d < DIM; ++d )
        lo[ d ] = cell_lower( g, d, ijk[ d ] );
      const double origin[ 3 ] = { 0.0, 0.0, 0.0 };
      u[ c ] = recursive_evaluation< DIM, 3 >( g, I, lo, origin, 1.0 );
    }
  }

  // ------------------------------------------------------------------------------------------------------------
  // interface extraction.  getInterfaceVertices (utility.hh:19-48) visits the mixed cells in
  // ASCENDING cell index and appends interface( entity, reconstructions ) of each: a CSR ( offsets, vertices ), with
  // MixedCellMapper::update's numbering (mixedcellmapper.hh:82-86).  The step's own mixed list is unordered (atomics),
  // so the order comes from a deterministic exclusive scan of the mixed predicate.
  // ------------------------------------------------------------------------------------------------------------

  constexpr int kScanBlock = 1024;   // items per block of the scan kernels (256 threads x 4)

  // a mixed cell this context owns (halo layers belong to the neighbouring slab)
  template< int DIM >
  __device__ __forceinline__ int owned_mixed ( const Grid &g, const unsigned char *__restrict__ flags, long long i )
  {
    if( !is_mixed( flags[ i ] ) )
      return 0;
    const long long layer = (long long) g.ns[ 0 ] * ( DIM == 3 ? g.ns[ 1 ] : 1 );
    const int la = int( i / layer );
    return ( la >= g.own0 && la < g.own1 ) ? 1 : 0;
  }

  // pass 1: per block, the number of selected items ( mode 0: is_mixed( flags[i] ) inside the listed layers; mode 1: counts[i] )
  template< int DIM >
  __global__ void __launch_bounds__( 256 ) k_scan_block_sums ( Grid g, long long n, const unsigned char *__restrict__ flags, const int *__restrict__ counts,
                                                                int *__restrict__ blockSums )
  {
    __shared__ int warpSums[ 8 ];
    const long long base = (long long) blockIdx.x * kScanBlock;
    int mine = 0;
    for( int k = 0; k < 4; ++k )
    {
      const long long i = base + threadIdx.x * 4 + k;
      if( i < n )
        mine += counts ? counts[ i ] : owned_mixed< DIM >( g, flags, i );
    }
    for( int o = 16; o > 0; o >>= 1 )
      mine += __shfl_xor_sync( 0xffffffffu, mine, o );
    if( ( threadIdx.x & 31 ) == 0 )
      warpSums[ threadIdx.x >> 5 ] = mine;
    __syncthreads();
    if( threadIdx.x == 0 )
    {
      int s = 0;
      for( int w = 0; w < 8; ++w )
        s += warpSums[ w ];
      blockSums[ blockIdx.x ] = s;
    }
  }

  // pass 2: exclusive scan of the block sums by one block; total -> *total
  static __global__ void __launch_bounds__( 1024 ) k_scan_block_offsets ( int nBlocks, int *__restrict__ blockSums, long long *__restrict__ total )
  {
    __shared__ long long partial[ 1024 ];
    const int per = ( nBlocks + 1023 ) / 1024;
    const int lo = threadIdx.x * per, hi = ( lo + per < nBlocks ) ? lo + per : nBlocks;
    long long s = 0;
    for( int i = lo; i < hi; ++i )
      s += blockSums[ i ];
    partial[ threadIdx.x ] = s;
    __syncthreads();
    if( threadIdx.x == 0 )
    {
      long long run = 0;
      for( int t = 0; t < 1024; ++t )
      {
        const long long v = partial[ t ];
        partial[ t ] = run;
        run += v;
      }
      *total = run;
    }
    __syncthreads();
    long long run = partial[ threadIdx.x ];
    for( int i = lo; i < hi; ++i )
    {
      const int v = blockSums[ i ];
      blockSums[ i ] = (int) run;
      run += v;
    }
  }

  // pass 3 (mode 0): ordered list of the mixed cells; (mode 1): exclusive offsets of counts
  template< int DIM >
  __global__ void __launch_bounds__( 256 ) k_scan_scatter ( Grid g, long long n, const unsigned char *__restrict__ flags, const int *__restrict__ counts,
                                                             const int *__restrict__ blockOffsets, int *__restrict__ out )
  {
    __shared__ int warpSums[ 8 ];
    const long long base = (long long) blockIdx.x * kScanBlock;
    int v[ 4 ], mine = 0;
    for( int k = 0; k < 4; ++k )
    {
      const long long i = base + threadIdx.x * 4 + k;
      v[ k ] = ( i < n ) ? ( counts ? counts[ i ] : owned_mixed< DIM >( g, flags, i ) ) : 0;
      mine += v[ k ];
    }
    int incl = mine;
    const int lane = threadIdx.x & 31;
    for( int o = 1; o < 32; o <<= 1 )
    {
      const int t = __shfl_up_sync( 0xffffffffu, incl, o );
      if( lane >= o )
        incl += t;
    }
    if( lane == 31 )
      warpSums[ threadIdx.x >> 5 ] = incl;
    __syncthreads();
    int before = blockOffsets[ blockIdx.x ];
    for( int w = 0; w < ( threadIdx.x >> 5 ); ++w )
      before += warpSums[ w ];
    int pos = before + incl - mine;
    for( int k = 0; k < 4; ++k )
    {
      const long long i = base + threadIdx.x * 4 + k;
      if( i >= n )
        break;
      if( counts )
        out[ i ] = pos;                       // exclusive offset of item i
      else if( v[ k ] )
        out[ pos ] = (int) i;                 // the pos-th mixed cell
      pos += v[ k ];
    }
  }

  // interface polygon of the p-th mixed cell: vertex count and vertices into a fixed-stride scratch
  template< int DIM >
  __global__ void __launch_bounds__( 128 ) k_interface_cut ( Grid g, const double *__restrict__ hs, const int *__restrict__ cells, long long nMixed,
                                                              int *__restrict__ nverts, double *__restrict__ verts )
  {
    constexpr int STRIDE = ( DIM == 2 ) ? 2 : 12;
    for( long long p = blockIdx.x * (long long) blockDim.x + threadIdx.x; p < nMixed; p += (long long) gridDim.x * blockDim.x )
    {
      const long long c = cells[ p ];
      int ijk[ 3 ];
      cell_ijk( g, c, ijk );
      const double *rec = hs + c * ( DIM + 1 );
      double *out = verts + p * STRIDE * DIM;
      if( DIM == 2 )
      {
        Poly2 poly;
        make_cell( cell_lower( g, 0, ijk[ 0 ] ), cell_lower( g, 1, ijk[ 1 ] ), g.h[ 0 ], g.h[ 1 ], poly );
        const Hs2 H = { rec[ 0 ], rec[ 1 ], rec[ 2 ] };
        double lx[ 2 ], ly[ 2 ];
        plane_cut( poly, H, lx, ly );
        nverts[ p ] = 2;
        out[ 0 ] = lx[ 0 ]; out[ 1 ] = ly[ 0 ]; out[ 2 ] = lx[ 1 ]; out[ 3 ] = ly[ 1 ];
      }
      else
      {
        const double lo[ 3 ] = { cell_lower( g, 0, ijk[ 0 ] ), cell_lower( g, 1, ijk[ 1 ] ), cell_lower( g, 2, ijk[ 2 ] ) };
        Hex cell;
        make_cell( lo, g.h, cell );
        const Hs3 H = { { rec[ 0 ], rec[ 1 ], rec[ 2 ] }, rec[ 3 ] };
        Face3 f;
        plane_cut( cell, H, f );
        nverts[ p ] = f.n;
        for( int i = 0; i < f.n; ++i )
          for( int k = 0; k < 3; ++k )
            out[ 3 * i + k ] = f.v[ i ][ k ];
      }
    }
  }

  template< int DIM >
  __global__ void __launch_bounds__( 128 ) k_interface_gather ( long long nMixed, const int *__restrict__ nverts, const int *__restrict__ offsets,
                                                                 const double *__restrict__ verts, double *__restrict__ csr )
  {
    constexpr int STRIDE = ( DIM == 2 ) ? 2 : 12;
    for( long long p = blockIdx.x * (long long) blockDim.x + threadIdx.x; p < nMixed; p += (long long) gridDim.x * blockDim.x )
    {
      const double *in = verts + p * STRIDE * DIM;
      double *out = csr + (long long) offsets[ p ] * DIM;
      for( int i = 0; i < nverts[ p ] * DIM; ++i )
        out[ i ] = in[ i ];
    }
  }

  // ------------------------------------------------------------------------------------------------------------
  // single-cell primitives on small batches (parity tests of G1/G2/G5/G6/G7 on the device)
  // ------------------------------------------------------------------------------------------------------------

  template< int DIM >
  __global__ void k_test_locate ( long long count, const double *lo, const double *h, const double *normal, const double *fraction, double *out )
  {
    for( long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < count; i += (long long) gridDim.x * blockDim.x )
    {
      double d;
      if( DIM == 2 )
      {
        Poly2 p;
        make_cell( lo[ 2 * i ], lo[ 2 * i + 1 ], h[ 2 * i ], h[ 2 * i + 1 ], p );
        d = locate( p, normal[ 2 * i ], normal[ 2 * i + 1 ], fraction[ i ] );
      }
      else
      {
        Hex c;
        make_cell( lo + 3 * i, h + 3 * i, c );
        d = locate( c, normal + 3 * i, fraction[ i ] );
      }
      for( int k = 0; k < DIM; ++k )
        out[ ( DIM + 1 ) * i + k ] = normal[ DIM * i + k ];
      out[ ( DIM + 1 ) * i + DIM ] = d;
    }
  }

  template< int DIM >
  __global__ void k_test_flux ( long long count, const double *lo, const double *h, const int *face, const double *v, const double *hs, double *out )
  {
    for( long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < count; i += (long long) gridDim.x * blockDim.x )
    {
      if( DIM == 2 )
      {
        const Hs2 H = { hs[ 3 * i ], hs[ 3 * i + 1 ], hs[ 3 * i + 2 ] };
        out[ i ] = flux( lo[ 2 * i ], lo[ 2 * i + 1 ], h[ 2 * i ], h[ 2 * i + 1 ], face[ i ], v[ 2 * i ], v[ 2 * i + 1 ], H );
      }
      else
      {
        const Hs3 H = { { hs[ 4 * i ], hs[ 4 * i + 1 ], hs[ 4 * i + 2 ] }, hs[ 4 * i + 3 ] };
        out[ i ] = flux( lo + 3 * i, h + 3 * i, face[ i ], v + 3 * i, H );
      }
    }
  }

  template< int DIM >
  __global__ void k_test_interface ( long long count, const double *lo, const double *h, const double *hs, int *nverts, double *centroid, double *measure )
  {
    for( long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < count; i += (long long) gridDim.x * blockDim.x )
    {
      if( DIM == 2 )
      {
        Poly2 p;
        make_cell( lo[ 2 * i ], lo[ 2 * i + 1 ], h[ 2 * i ], h[ 2 * i + 1 ], p );
        const Hs2 H = { hs[ 3 * i ], hs[ 3 * i + 1 ], hs[ 3 * i + 2 ] };
        double lx[ 2 ], ly[ 2 ];
        plane_cut( p, H, lx, ly );
        nverts[ i ] = 2;
        centroid[ 2 * i ] = ( lx[ 0 ] + lx[ 1 ] ) * 0.5;
        centroid[ 2 * i + 1 ] = ( ly[ 0 ] + ly[ 1 ] ) * 0.5;
        measure[ i ] = norm2( lx[ 0 ] - lx[ 1 ], ly[ 0 ] - ly[ 1 ] );
      }
      else
      {
        Hex c;
        make_cell( lo + 3 * i, h + 3 * i, c );
        const Hs3 H = { { hs[ 4 * i ], hs[ 4 * i + 1 ], hs[ 4 * i + 2 ] }, hs[ 4 * i + 3 ] };
        Face3 f;
        plane_cut( c, H, f );
        nverts[ i ] = f.n;
        if( f.n > 2 )
        {
          face_centroid( f, centroid + 3 * i );
          measure[ i ] = face_measure( f );
        }
        else
        {
          centroid[ 3 * i ] = centroid[ 3 * i + 1 ] = centroid[ 3 * i + 2 ] = 0.0;
          measure[ i ] = 0.0;
        }
      }
    }
  }

  static __global__ void k_test_face_velocity ( Grid g, const Velocity *velocity, double time, long long cell, int face, double *out )
  {
    int ijk[ 3 ];
    cell_ijk( g, cell, ijk );
    face_velocity( g, *velocity, ijk, face, time, out );
  }

} // namespace vofx
