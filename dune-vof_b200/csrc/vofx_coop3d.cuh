// vofx: 3D band geometry for GROUPS OF 8 LANES per work item (B200: the band holds ~1e5 mixed cells at 512^3, far too
// few for one thread per cell to fill 148 SMs, and the serial formulation diverges 4-way inside a warp).
//
// Every function here is a PHASE executed by lane 0..7 of a group on a scratch block in shared memory; the kernel
// separates the phases by __syncwarp( groupMask ).  Phases only communicate through the scratch, so the host test
// harness (tests/host_geom) runs the very same code by looping over the lanes phase by phase, and the CPU test-suite
// checks it bit-for-bit against the oracle without a GPU.
//
// The ARITHMETIC is still the reference's, operation for operation (vofx_math.cuh): independent terms (one face of
// the clipped polyhedron, one edge of the sweep polygon, one stencil neighbour) are evaluated by different lanes,
// and every SUM is then taken by a single lane in the reference's order.
#pragma once

#include "vofx_geom3d.cuh"

namespace vofx
{
  namespace coop
  {

    constexpr int kLanes = 8;

    // ---- G5: clip of a hexahedron, topology looked up by the inner mask ----------------------------------------------

    // One entry per inner mask when no node lies on the plane: which cut nodes exist and, face by face in the
    // reference's summation order, the cyclic node sequence.  ref < 8: original node; ref >= 8: cut node ref - 8.
    struct ClipCase
    {
      unsigned char nCut;            // 0xFF: not representable here (take the serial path)
      unsigned char nFaces;
      unsigned char cut[ 6 ];        // a | b << 3: Edge( a, b ).intersection( plane ), operand order as in the walk
      unsigned char faceLen[ 7 ];
      unsigned char pad0;
      unsigned char ref[ 7 ][ 6 ];   // first node of the k-th edge of the face; its second node is ref[ (k+1) % len ]
      unsigned char pad1[ 6 ];
    };
    static_assert( sizeof( ClipCase ) == 64, "ClipCase must be 64 bytes" );

    // table generator: runs the reference's walk (clip_walk) for every inner mask
    inline void make_clip_table ( ClipCase *table )
    {
      for( unsigned mask = 0; mask < 256; ++mask )
      {
        ClipCase &c = table[ mask ];
        for( unsigned i = 0; i < sizeof( ClipCase ); ++i )
          reinterpret_cast< unsigned char * >( &c )[ i ] = 0;
        ClipTopo t;
        if( !clip_walk( mask, 0u, t ) )
        {
          c.nCut = 0;
          c.nFaces = 0;      // empty result (only mask 0)
          continue;
        }
        bool regular = t.nNew <= 6 && t.nFaces <= 7;
        for( int f = 0; regular && f < t.nFaces; ++f )
        {
          const int len = t.flen[ f ];
          regular = len <= 6;
          for( int k = 0; regular && k < len; ++k )
            regular = t.fe1[ f ][ k ] == t.fe0[ f ][ ( k + 1 ) % len ];
        }
        if( !regular )
        {
          c.nCut = 0xFF;
          continue;
        }
        c.nCut = (unsigned char) t.nNew;
        c.nFaces = (unsigned char) t.nFaces;
        for( int j = 0; j < t.nNew; ++j )
          c.cut[ j ] = (unsigned char) ( t.cutA[ j ] | ( t.cutB[ j ] << 3 ) );
        for( int f = 0; f < t.nFaces; ++f )
        {
          c.faceLen[ f ] = (unsigned char) t.flen[ f ];
          for( int k = 0; k < t.flen[ f ]; ++k )
          {
            const int id = t.fe0[ f ][ k ];
            c.ref[ f ][ k ] = (unsigned char) ( id < t.n ? t.innerNode[ id ] : 8 + ( id - t.n ) );
          }
        }
      }
    }

    struct ClipScratch
    {
      double x[ 3 ][ 16 ];             // coordinate d of ref r: 0..7 nodes of the hexahedron, 8..13 cut nodes
      double term[ 8 ];                // face terms of Polyhedron::volume
      unsigned char inner[ 8 ], bnd[ 8 ];
    };

    // phase A: lane i classifies node i (3d/intersect.hh:99-112) and publishes it
    VOFX_HD inline void clip_phase_classify ( int lane, ClipScratch &S, const double *xn, const Hs3 &hs )
    {
      const double ls = level_set( hs, xn );
      const bool inner = ls > -kEps;
      S.inner[ lane ] = inner ? (unsigned char) ( 1u << lane ) : (unsigned char) 0;
      S.bnd[ lane ] = ( inner && fabs( ls ) < kEps ) ? (unsigned char) ( 1u << lane ) : (unsigned char) 0;
      for( int d = 0; d < 3; ++d )
        S.x[ d ][ lane ] = xn[ d ];
    }

    VOFX_HD inline void clip_masks ( const ClipScratch &S, unsigned &innerMask, unsigned &boundaryMask )
    {
      innerMask = boundaryMask = 0;
      for( int i = 0; i < 8; ++i )
      {
        innerMask |= S.inner[ i ];
        boundaryMask |= S.bnd[ i ];
      }
    }

    // phase B: lane j creates cut node j.  Returns 0 table path, 1 empty result, 2 serial path (the same for all lanes)
    VOFX_HD inline int clip_phase_cut ( int lane, ClipScratch &S, const Hs3 &hs, const ClipCase *table, const ClipCase *&cs )
    {
      unsigned innerMask, boundaryMask;
      clip_masks( S, innerMask, boundaryMask );
      cs = table + innerMask;
      if( innerMask == 0 )
        return 1;
      if( boundaryMask != 0 || cs->nCut == 0xFF )
        return 2;
      if( lane < cs->nCut )
      {
        const int a = cs->cut[ lane ] & 7, b = cs->cut[ lane ] >> 3;
        const double xa[ 3 ] = { S.x[ 0 ][ a ], S.x[ 1 ][ a ], S.x[ 2 ][ a ] };
        const double xb[ 3 ] = { S.x[ 0 ][ b ], S.x[ 1 ][ b ], S.x[ 2 ][ b ] };
        double out[ 3 ];
        edge_plane_point( xa, xb, hs, out );
        for( int d = 0; d < 3; ++d )
          S.x[ d ][ 8 + lane ] = out[ d ];
      }
      return 0;
    }

    // phase C: lane f evaluates face f of the clipped polyhedron (3d/polyhedron.hh:187-222 as in clipped_face_term)
    VOFX_HD inline void clip_phase_faces ( int lane, ClipScratch &S, const ClipCase &cs )
    {
      if( lane >= cs.nFaces )
        return;
      const int len = cs.faceLen[ lane ];
      double x[ 6 ][ 3 ];
#pragma unroll
      for( int k = 0; k < 6; ++k )
      {
        const int r = cs.ref[ lane ][ k < len ? k : 0 ];
#pragma unroll
        for( int d = 0; d < 3; ++d )
          x[ k ][ d ] = S.x[ d ][ r ];
      }
      double center[ 3 ] = { 0.0, 0.0, 0.0 }, fn[ 3 ];
#pragma unroll
      for( int k = 0; k < 6; ++k )
        if( k < len )
#pragma unroll
          for( int d = 0; d < 3; ++d )
            center[ d ] += x[ k ][ d ];
      const double inv = 1.0 / len;
#pragma unroll
      for( int d = 0; d < 3; ++d )
        center[ d ] *= inv;
      outer_normal3( x[ 0 ], x[ 1 ], x[ 2 ], fn );
      double fv = 0.0;
#pragma unroll
      for( int k = 0; k < 6; ++k )
        if( k < len )
          fv += face_edge_term( x[ k ], x[ ( k + 1 < 6 ) ? k + 1 : 0 ], fn );   // x[k+1] holds x[0] where k+1 == len
      S.term[ lane ] = dot3( center, fn ) * fabs( fv / 2.0 );
    }

    // phase D (every lane): Polyhedron::volume, 3d/polyhedron.hh:317-324
    VOFX_HD inline double clip_phase_sum ( const ClipScratch &S, const ClipCase &cs )
    {
      double vol = 0.0;
      for( int f = 0; f < cs.nFaces; ++f )
        vol += S.term[ f ];
      return fabs( vol / 3.0 );
    }

    // serial path for the rare cases (a node on the plane, irregular mask): rebuild the hexahedron from the scratch
    VOFX_HD inline double clip_serial_from_scratch ( const ClipScratch &S, const Hs3 &hs )
    {
      Hex poly;
      for( int i = 0; i < 8; ++i )
        for( int d = 0; d < 3; ++d )
          poly.x[ i ][ d ] = S.x[ d ][ i ];
      return clip_volume( poly, hs );
    }

    // node `lane` of upwindPolygon3d( face, v ) for the cell with lower corner lo (same arithmetic as vofx::flux)
    VOFX_HD inline void upwind_node ( int lane, const double *lo, const double *h, int face, const double *v, double *xn )
    {
      const int axis = face >> 1;
      double o[ 3 ] = { lo[ 0 ], lo[ 1 ], lo[ 2 ] };
      if( face & 1 )
        o[ axis ] += h[ axis ];
      const int a0 = ( axis == 0 ) ? 1 : 0;
      const int a1 = ( axis == 2 ) ? 1 : 2;
      double c[ 4 ][ 3 ];
      for( int i = 0; i < 4; ++i )
      {
        for( int d = 0; d < 3; ++d )
          c[ i ][ d ] = o[ d ];
        if( i & 1 )
          c[ i ][ a0 ] += h[ a0 ];
        if( i & 2 )
          c[ i ][ a1 ] += h[ a1 ];
      }
      double e1[ 3 ], e2[ 3 ], cr[ 3 ];
      for( int d = 0; d < 3; ++d )
      {
        e1[ d ] = c[ 1 ][ d ] - c[ 0 ][ d ];
        e2[ d ] = c[ 2 ][ d ] - c[ 0 ][ d ];
      }
      cross3( e1, e2, cr );
      const bool shiftedFirst = dot3( cr, v ) < 0.0;
      const int i = lane & 3;
      const bool moved = ( lane < 4 ) == shiftedFirst;
      for( int d = 0; d < 3; ++d )
      {
        double cd = o[ d ];
        if( d == a0 && ( i & 1 ) )
          cd += h[ a0 ];
        if( d == a1 && ( i & 2 ) )
          cd += h[ a1 ];
        xn[ d ] = moved ? cd - v[ d ] : cd;
      }
    }

  } // namespace coop
} // namespace vofx
