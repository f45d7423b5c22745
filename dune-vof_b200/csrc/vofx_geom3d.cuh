// vofx 3D geometry device functions for hexahedral cells and swept face boxes:
//   locate()       plane-constant solve     = locateHalfSpace( Polyhedron ), geometry/algorithm.hh:119-236,
//                                             3d/rotation.hh:17-53, 3d/polygonwithdirections.hh:27-316
//   clip_volume()  |polyhedron n half-space| = __impl::intersect( Polyhedron, HalfSpace ), 3d/intersect.hh:72-263
//                                             followed by Polyhedron::volume(), 3d/polyhedron.hh:187-222,317-324
//   flux()         truncVolume( upwindPolygon3d( face, v ), hs ), 3d/upwindpolygon.hh:24-74
//   plane_cut()/face_centroid()  interface polygon and its centroid, 3d/intersect.hh:274-355, 3d/face.hh:14-119
//
// B200 design notes.  The reference builds std::vector/std::map based polyhedra; here every polytope has the
// topology of a cube (8 nodes, 24 directed edges, 6 quadrilateral faces), so the topology lives in bit-packed
// 64-bit immediates (no memory traffic, no constant-cache serialisation under divergent indices) and the
// per-thread working set is a handful of fixed-size arrays.  The ARITHMETIC, however, is the reference's,
// operation for operation and in its summation order (see vofx_math.cuh): node/edge/face numbering below is the
// reference's makePolyhedron table because it fixes that order.
#pragma once

#include "vofx_math.cuh"

namespace vofx
{

  namespace cube
  {
    // directed edges (node0,node1) and faces (4 edge ids) of makePolyhedron( cube ), 3d/polyhedron.hh:387-401
    // edges: (0,4)(1,5)(2,6)(3,7)(0,2)(1,3)(0,1)(2,3)(4,6)(5,7)(4,5)(6,7)(4,0)(5,1)(6,2)(7,3)(2,0)(3,1)(1,0)(3,2)(6,4)(7,5)(5,4)(7,6)
    // faces: {0,8,14,16} {5,3,21,13} {6,1,22,12} {19,2,11,15} {4,7,17,18} {10,9,23,20}
    constexpr int kEdgeN0[ 24 ] = { 0, 1, 2, 3, 0, 1, 0, 2, 4, 5, 4, 6, 4, 5, 6, 7, 2, 3, 1, 3, 6, 7, 5, 7 };
    constexpr int kEdgeN1[ 24 ] = { 4, 5, 6, 7, 2, 3, 1, 3, 6, 7, 5, 7, 0, 1, 2, 3, 0, 1, 0, 2, 4, 5, 4, 6 };
    constexpr int kFaceEdge[ 6 ][ 4 ] = { { 0, 8, 14, 16 }, { 5, 3, 21, 13 }, { 6, 1, 22, 12 }, { 19, 2, 11, 15 }, { 4, 7, 17, 18 }, { 10, 9, 23, 20 } };

    // ---- compile-time packing of the derived look-up tables into 64-bit words --------------------------------
    constexpr uint64_t packEdgeNodes ( int word )          // 8 edges per word, 6 bits each: n0 | n1 << 3
    {
      uint64_t w = 0;
      for( int i = 0; i < 8; ++i )
        w |= uint64_t( kEdgeN0[ 8 * word + i ] | ( kEdgeN1[ 8 * word + i ] << 3 ) ) << ( 6 * i );
      return w;
    }
    constexpr int faceOfEdgeCT ( int e )
    {
      for( int f = 0; f < 6; ++f )
        for( int k = 0; k < 4; ++k )
          if( kFaceEdge[ f ][ k ] == e )
            return f;
      return 7;
    }
    constexpr uint64_t packFaceOfEdge ( int word )         // 12 edges per word, 3 bits each
    {
      uint64_t w = 0;
      for( int i = 0; i < 12; ++i )
        w |= uint64_t( faceOfEdgeCT( 12 * word + i ) ) << ( 3 * i );
      return w;
    }
    constexpr uint64_t packFaceEdges ( int word )          // 3 faces per word, 4 x 5 bits each
    {
      uint64_t w = 0;
      for( int f = 0; f < 3; ++f )
        for( int k = 0; k < 4; ++k )
          w |= uint64_t( kFaceEdge[ 3 * word + f ][ k ] ) << ( 20 * f + 5 * k );
      return w;
    }
    constexpr int outEdgeCT ( int node, int which )        // which-th edge leaving `node`, in edge order
    {
      int c = 0;
      for( int e = 0; e < 24; ++e )
        if( kEdgeN0[ e ] == node )
        {
          if( c == which )
            return e;
          ++c;
        }
      return 31;
    }
    constexpr uint64_t packOutEdges ( int word )           // 4 nodes per word, 3 x 5 bits each
    {
      uint64_t w = 0;
      for( int n = 0; n < 4; ++n )
        for( int k = 0; k < 3; ++k )
          w |= uint64_t( outEdgeCT( 4 * word + n, k ) ) << ( 15 * n + 5 * k );
      return w;
    }
    constexpr int reverseCT ( int e )
    {
      for( int r = 0; r < 24; ++r )
        if( kEdgeN0[ r ] == kEdgeN1[ e ] && kEdgeN1[ r ] == kEdgeN0[ e ] )
          return r;
      return 31;
    }
    constexpr uint64_t packReverse ( int word )            // 12 edges per word, 5 bits each
    {
      uint64_t w = 0;
      for( int i = 0; i < 12; ++i )
        w |= uint64_t( reverseCT( 12 * word + i ) ) << ( 5 * i );
      return w;
    }

    constexpr uint64_t kEN0 = packEdgeNodes( 0 ), kEN1 = packEdgeNodes( 1 ), kEN2 = packEdgeNodes( 2 );
    constexpr uint64_t kFOE0 = packFaceOfEdge( 0 ), kFOE1 = packFaceOfEdge( 1 );
    constexpr uint64_t kFE0 = packFaceEdges( 0 ), kFE1 = packFaceEdges( 1 );
    constexpr uint64_t kOE0 = packOutEdges( 0 ), kOE1 = packOutEdges( 1 );
    constexpr uint64_t kREV0 = packReverse( 0 ), kREV1 = packReverse( 1 );

    VOFX_INLINE int edgeNodes ( int e )
    {
      const uint64_t w = ( e < 8 ) ? kEN0 : ( ( e < 16 ) ? kEN1 : kEN2 );
      return int( ( w >> ( 6 * ( e & 7 ) ) ) & 63 );
    }
    VOFX_INLINE int edgeN0 ( int e ) { return edgeNodes( e ) & 7; }
    VOFX_INLINE int edgeN1 ( int e ) { return edgeNodes( e ) >> 3; }
    VOFX_INLINE int faceOfEdge ( int e )                   // Polyhedron::attachedFaceToEdge, 3d/polyhedron.hh:335-344
    {
      return ( e < 12 ) ? int( ( kFOE0 >> ( 3 * e ) ) & 7 ) : int( ( kFOE1 >> ( 3 * ( e - 12 ) ) ) & 7 );
    }
    VOFX_INLINE int faceEdge ( int f, int k )
    {
      return ( f < 3 ) ? int( ( kFE0 >> ( 20 * f + 5 * k ) ) & 31 ) : int( ( kFE1 >> ( 20 * ( f - 3 ) + 5 * k ) ) & 31 );
    }
    VOFX_INLINE int faceNode ( int f, int k ) { return edgeN0( faceEdge( f, k ) ); }
    VOFX_INLINE int outEdge ( int node, int k )            // Polyhedron::attachedEdgesToNode, 3d/polyhedron.hh:346-356
    {
      return ( node < 4 ) ? int( ( kOE0 >> ( 15 * node + 5 * k ) ) & 31 ) : int( ( kOE1 >> ( 15 * ( node - 4 ) + 5 * k ) ) & 31 );
    }
    VOFX_INLINE int reverseEdge ( int e )
    {
      return ( e < 12 ) ? int( ( kREV0 >> ( 5 * e ) ) & 31 ) : int( ( kREV1 >> ( 5 * ( e - 12 ) ) ) & 31 );
    }
    VOFX_INLINE int edgeId ( int a, int b )                // directed edge a->b or -1
    {
      for( int k = 0; k < 3; ++k )
      {
        const int e = outEdge( a, k );
        if( edgeN1( e ) == b )
          return e;
      }
      return -1;
    }
    // comparator PolygonWithDirections::operator(), 3d/polygonwithdirections.hh:293-306: lhs < rhs iff the face attached
    // to lhs contains the reverse of rhs (every directed edge of a cube lies in exactly one face)
    VOFX_INLINE bool dirLess ( int lhs, int rhs ) { return faceOfEdge( lhs ) == faceOfEdge( reverseEdge( rhs ) ); }
  } // namespace cube

  // ---- topology policies ------------------------------------------------------------------------------------------------
  // The serial geometry below (clip walk, sweep, plane-constant solve) is written against a policy T: T = topo::Cube is the
  // table above (cells of Cartesian grids, hexahedra, swept quadrilateral faces); topo::Tet is makePolyhedron( simplex ),
  // 3d/polyhedron.hh:374-385; topo::Prism the swept triangle of upwindPolygon3d, 3d/upwindpolygon.hh:47-57.  A polytope
  // keeps its nodes in a Hex (at most 8).  The derived look-ups of Tet and Prism are searches: they serve the unstructured
  // path (vofx_mesh.cu), one thread per cell.
  namespace topo
  {
    struct Cube
    {
      static constexpr int kNodes = 8, kEdges = 24, kFaces = 6;
      VOFX_INLINE static int faceLen ( int ) { return 4; }
      VOFX_INLINE static int edgeN0 ( int e ) { return cube::edgeN0( e ); }
      VOFX_INLINE static int edgeN1 ( int e ) { return cube::edgeN1( e ); }
      VOFX_INLINE static int faceEdge ( int f, int k ) { return cube::faceEdge( f, k ); }
      VOFX_INLINE static int faceNode ( int f, int k ) { return cube::faceNode( f, k ); }
      VOFX_INLINE static int faceOfEdge ( int e ) { return cube::faceOfEdge( e ); }
      VOFX_INLINE static int outEdge ( int node, int k ) { return cube::outEdge( node, k ); }
      VOFX_INLINE static int reverseEdge ( int e ) { return cube::reverseEdge( e ); }
      VOFX_INLINE static int edgeId ( int a, int b ) { return cube::edgeId( a, b ); }
      VOFX_INLINE static bool dirLess ( int lhs, int rhs ) { return cube::dirLess( lhs, rhs ); }
    };

    // the look-ups every policy derives from its edge and face tables
    template< class B >
    struct Derived : B
    {
      VOFX_INLINE static int faceNode ( int f, int k ) { return B::edgeN0( B::faceEdge( f, k ) ); }
      VOFX_INLINE static int faceOfEdge ( int e )            // Polyhedron::attachedFaceToEdge: the first face holding the edge
      {
        for( int f = 0; f < B::kFaces; ++f )
          for( int k = 0; k < B::faceLen( f ); ++k )
            if( B::faceEdge( f, k ) == e )
              return f;
        return 7;
      }
      VOFX_INLINE static int outEdge ( int node, int which ) // Polyhedron::attachedEdgesToNode: edges leaving the node, in edge order
      {
        int c = 0;
        for( int e = 0; e < B::kEdges; ++e )
          if( B::edgeN0( e ) == node )
          {
            if( c == which )
              return e;
            ++c;
          }
        return 31;
      }
      VOFX_INLINE static int edgeId ( int a, int b )
      {
        for( int e = 0; e < B::kEdges; ++e )
          if( B::edgeN0( e ) == a && B::edgeN1( e ) == b )
            return e;
        return -1;
      }
      VOFX_INLINE static int reverseEdge ( int e ) { const int r = edgeId( B::edgeN1( e ), B::edgeN0( e ) ); return r < 0 ? 31 : r; }
      VOFX_INLINE static bool dirLess ( int lhs, int rhs ) { return faceOfEdge( lhs ) == faceOfEdge( reverseEdge( rhs ) ); }
    };

    struct TetTables
    {
      static constexpr int kNodes = 4, kEdges = 12, kFaces = 4;
      VOFX_INLINE static int faceLen ( int ) { return 3; }
      VOFX_INLINE static int edgeN0 ( int e ) { const signed char t[ 12 ] = { 0, 1, 0, 2, 1, 2, 0, 3, 1, 3, 3, 2 }; return t[ e ]; }
      VOFX_INLINE static int edgeN1 ( int e ) { const signed char t[ 12 ] = { 1, 0, 2, 0, 2, 1, 3, 0, 3, 1, 2, 3 }; return t[ e ]; }
      VOFX_INLINE static int faceEdge ( int f, int k ) { const signed char t[ 4 ][ 3 ] = { { 2, 5, 1 }, { 0, 8, 7 }, { 6, 10, 3 }, { 4, 11, 9 } }; return t[ f ][ k ]; }
    };
    using Tet = Derived< TetTables >;

    struct PrismTables
    {
      static constexpr int kNodes = 6, kEdges = 18, kFaces = 5;
      VOFX_INLINE static int faceLen ( int f ) { return f < 3 ? 4 : 3; }
      VOFX_INLINE static int edgeN0 ( int e ) { const signed char t[ 18 ] = { 0, 1, 2, 0, 0, 1, 3, 3, 4, 3, 4, 5, 1, 2, 2, 4, 5, 5 }; return t[ e ]; }
      VOFX_INLINE static int edgeN1 ( int e ) { const signed char t[ 18 ] = { 3, 4, 5, 1, 2, 2, 4, 5, 5, 0, 1, 2, 0, 0, 1, 3, 3, 4 }; return t[ e ]; }
      VOFX_INLINE static int faceEdge ( int f, int k )
      {
        const signed char t[ 5 ][ 4 ] = { { 9, 3, 1, 15 }, { 0, 7, 11, 13 }, { 5, 2, 17, 10 }, { 4, 14, 12, 0 }, { 6, 8, 16, 0 } };
        return t[ f ][ k ];
      }
    };
    using Prism = Derived< PrismTables >;
  } // namespace topo

  struct Hs3
  {
    double n[ 3 ];
    double d;
  };

  VOFX_INLINE bool hs_valid ( const Hs3 &h ) { return norm2_3( h.n ) > 0.5; }
  VOFX_INLINE double level_set ( const Hs3 &h, const double *x ) { return dot3( h.n, x ) + h.d; }

  // eight nodes of a polytope with cube topology
  struct Hex
  {
    double x[ 8 ][ 3 ];
  };

  // makePolyhedron( cube geometry ): node i = geometry.corner( i ), bit d of i selects lower + h
  VOFX_HD inline void make_cell ( const double *lo, const double *h, Hex &c )
  {
    for( int i = 0; i < 8; ++i )
      for( int d = 0; d < 3; ++d )
        c.x[ i ][ d ] = ( ( i >> d ) & 1 ) ? lo[ d ] + h[ d ] : lo[ d ];
  }

  // ---- face formulas on an arbitrary node table, 3d/polyhedron.hh:187-222 ----------------------------------------

  // Face::outerNormal from the first three nodes: cross( n0-n1, n2-n1 ) / ( -|.| )
  VOFX_HD inline void outer_normal3 ( const double *n0, const double *n1, const double *n2, double *normal )
  {
    double v1[ 3 ], v2[ 3 ];
    for( int k = 0; k < 3; ++k )
    {
      v1[ k ] = n0[ k ] - n1[ k ];
      v2[ k ] = n2[ k ] - n1[ k ];
    }
    cross3( v1, v2, normal );
    const double s = -1.0 * norm3( normal );
    vdiv3( normal[ 0 ], normal[ 1 ], normal[ 2 ], s );
  }

  // one edge's term of Face::volume: ( edge.center() . edge.outerNormal( faceNormal ) ) * edge.volume()
  VOFX_HD inline double face_edge_term ( const double *x0, const double *x1, const double *fn )
  {
    double center[ 3 ], v1[ 3 ], en[ 3 ];
    for( int k = 0; k < 3; ++k )
    {
      center[ k ] = ( x0[ k ] + x1[ k ] ) * 0.5;
      v1[ k ] = x1[ k ] - x0[ k ];
    }
    cross3( v1, fn, en );
    const double nrm = norm3( en );
    vdiv3( en[ 0 ], en[ 1 ], en[ 2 ], nrm );
    return dot3( center, en ) * norm3( v1 );
  }

  // Polyhedron::volume of an (unclipped) hexahedron, 3d/polyhedron.hh:317-324
  template< class T = topo::Cube >
  VOFX_HD inline double hex_volume ( const Hex &c )
  {
    double vol = 0.0;
    for( int f = 0; f < T::kFaces; ++f )
    {
      double center[ 3 ] = { 0.0, 0.0, 0.0 }, fn[ 3 ];
      const int len = T::faceLen( f );
      for( int k = 0; k < len; ++k )
      {
        const double *x = c.x[ T::faceNode( f, k ) ];
        for( int d = 0; d < 3; ++d )
          center[ d ] += x[ d ];
      }
      const double s = 1.0 / len;
      for( int d = 0; d < 3; ++d )
        center[ d ] *= s;
      outer_normal3( c.x[ T::faceNode( f, 0 ) ], c.x[ T::faceNode( f, 1 ) ], c.x[ T::faceNode( f, 2 ) ], fn );
      double fv = 0.0;
      for( int k = 0; k < len; ++k )
      {
        const int e = T::faceEdge( f, k );
        fv += face_edge_term( c.x[ T::edgeN0( e ) ], c.x[ T::edgeN1( e ) ], fn );
      }
      vol += dot3( center, fn ) * fabs( fv / 2.0 );
    }
    return fabs( vol / 3.0 );
  }

  // ---- G5: clip against a half-space and take the volume ----------------------------------------------------------

  // Edge::intersection( hyperplane ), 3d/polyhedron.hh:93-101
  VOFX_HD inline void edge_plane_point ( const double *x0, const double *x1, const Hs3 &hs, double *out )
  {
    double p[ 3 ];
    for( int k = 0; k < 3; ++k )
      p[ k ] = x0[ k ] - x1[ k ];
    const double s = vdiv( dot3( hs.n, x0 ) + hs.d, dot3( hs.n, p ) * -1.0 );
    for( int k = 0; k < 3; ++k )
    {
      p[ k ] *= s;
      p[ k ] += x0[ k ];
      out[ k ] = p[ k ];
    }
  }

  // Topology of the clipped polyhedron (3d/intersect.hh:123-262).  The walk only depends on which nodes are inner /
  // on the boundary, not on coordinates: the serial path runs it per clip, the cooperative path (vofx_coop3d.cuh)
  // looks it up in a table indexed by the inner mask that is generated from this very function.
  struct ClipTopo
  {
    int n, nNew;                              // inner nodes (new ids 0..n-1), cut nodes (new ids n..n+nNew-1)
    signed char innerNode[ 8 ];               // original id of new node k < n (ascending, :252-257)
    signed char cutA[ 12 ], cutB[ 12 ];       // cut node n+j = Edge( cutA[j], cutB[j] ).intersection( plane ), :171-183
    int nFaces;                               // faces that enter Polyhedron::volume, in summation order
    signed char flen[ 7 ];
    signed char fe0[ 7 ][ 16 ], fe1[ 7 ][ 16 ];   // directed edges (new node ids) of each face
  };

  // returns false for the trivial cases of :115-120 (empty result)
  template< class T = topo::Cube >
  VOFX_HD inline bool clip_walk ( unsigned innerMask, unsigned boundaryMask, ClipTopo &t )
  {
    signed char p[ 8 ];
    int n = 0, onBoundary = 0;
    for( int i = 0; i < T::kNodes; ++i )
    {
      if( innerMask & ( 1u << i ) )
      {
        t.innerNode[ n ] = (signed char) i;
        p[ i ] = (signed char) n;
        n++;
        if( boundaryMask & ( 1u << i ) )
          onBoundary++;
      }
      else
        p[ i ] = -1;
    }
    t.n = n;
    t.nNew = 0;
    t.nFaces = 0;
    if( n == 0 || ( n == 1 && onBoundary == 1 ) || ( n == 2 && onBoundary == 2 ) )
      return false;

    int nNew = 0;
    signed char cutNode[ 24 ];                 // new node id created on directed edge e, or -1 (std::map of :93-95)
    for( int e = 0; e < T::kEdges; ++e )
      cutNode[ e ] = -1;
    signed char is0[ 16 ], is1[ 16 ];          // edges of the intersection face
    int nIF = 0;

    for( int f = 0; f < T::kFaces; ++f )
    {
      signed char *n0 = t.fe0[ t.nFaces ], *n1 = t.fe1[ t.nFaces ];
      int nNF = 0;
      int lastIsId = -1;
#define VOFX_FACE_EDGE( a_, b_ ) ( n0[ nNF ] = (signed char) ( a_ ), n1[ nNF ] = (signed char) ( b_ ), nNF++ )
#define VOFX_CAP_EDGE( a_, b_ ) ( is0[ nIF ] = (signed char) ( a_ ), is1[ nIF ] = (signed char) ( b_ ), nIF++ )
      for( int k = 0; k < T::faceLen( f ); ++k )
      {
        const int edge = T::faceEdge( f, k );
        const int a = T::edgeN0( edge ), b = T::edgeN1( edge );
        const bool ia = innerMask & ( 1u << a ), ib = innerMask & ( 1u << b );
        if( ia && ib )
          VOFX_FACE_EDGE( p[ a ], p[ b ] );
        else if( ia != ib )
        {
          if( boundaryMask & ( 1u << a ) )
          {
            if( lastIsId != -1 && lastIsId != p[ a ] )
            {
              VOFX_FACE_EDGE( p[ a ], lastIsId );
              VOFX_CAP_EDGE( lastIsId, p[ a ] );
            }
            lastIsId = p[ a ];
          }
          else if( boundaryMask & ( 1u << b ) )
          {
            if( lastIsId != -1 && lastIsId != p[ b ] )
            {
              VOFX_FACE_EDGE( lastIsId, p[ b ] );
              VOFX_CAP_EDGE( p[ b ], lastIsId );
            }
            lastIsId = p[ b ];
          }
          else
          {
            int isNodeId = cutNode[ T::reverseEdge( edge ) ];
            if( isNodeId == -1 )
            {
              isNodeId = n + nNew;
              t.cutA[ nNew ] = (signed char) a;
              t.cutB[ nNew ] = (signed char) b;
              nNew++;
              cutNode[ edge ] = (signed char) isNodeId;
            }
            if( ia )
            {
              VOFX_FACE_EDGE( p[ a ], isNodeId );
              if( lastIsId != -1 && lastIsId != isNodeId )
              {
                VOFX_FACE_EDGE( isNodeId, lastIsId );
                VOFX_CAP_EDGE( lastIsId, isNodeId );
              }
            }
            else
            {
              if( lastIsId != -1 && lastIsId != isNodeId )
              {
                VOFX_FACE_EDGE( lastIsId, isNodeId );
                VOFX_CAP_EDGE( isNodeId, lastIsId );
              }
              VOFX_FACE_EDGE( isNodeId, p[ b ] );
            }
            lastIsId = isNodeId;
          }
        }
      }
#undef VOFX_FACE_EDGE
#undef VOFX_CAP_EDGE
      // faces are summed in this order, the intersection face last; faces with fewer than 3 edges are dropped, :185-190
      if( nNF > 2 )
      {
        t.flen[ t.nFaces ] = (signed char) nNF;
        t.nFaces++;
      }
    }
    t.nNew = nNew;

    // erase null edges (the reference does not re-test position i after an erase), 3d/intersect.hh:217-222
    for( int i = 0; i < nIF; ++i )
      if( is0[ i ] == is1[ i ] )
      {
        for( int j = i; j + 1 < nIF; ++j )
        {
          is0[ j ] = is0[ j + 1 ];
          is1[ j ] = is1[ j + 1 ];
        }
        nIF--;
      }
    // chain the edges of the intersection face, :224-239
    for( int i = 0; i + 1 < nIF; ++i )
    {
      if( is1[ i ] == is0[ i + 1 ] )
        continue;
      const int index = is1[ i ];
      int pos = -1;
      for( int j = i; j < nIF; ++j )
        if( is0[ j ] == index )
        {
          pos = j;
          break;
        }
      if( pos != -1 )
      {
        signed char s = is0[ i + 1 ];
        is0[ i + 1 ] = is0[ pos ];
        is0[ pos ] = s;
        s = is1[ i + 1 ];
        is1[ i + 1 ] = is1[ pos ];
        is1[ pos ] = s;
      }
    }
    if( nIF > 2 )
    {
      for( int i = 0; i < nIF; ++i )
      {
        t.fe0[ t.nFaces ][ i ] = is0[ i ];
        t.fe1[ t.nFaces ][ i ] = is1[ i ];
      }
      t.flen[ t.nFaces ] = (signed char) nIF;
      t.nFaces++;
    }
    return true;
  }

  // contribution ( face.center() . face.outerNormal() ) * face.volume() of one face given as directed edges between
  // rows of the node table x, 3d/polyhedron.hh:187-222
  VOFX_HD inline double clipped_face_term ( const double ( *x )[ 3 ], const signed char *e0, const signed char *e1, int n )
  {
    double center[ 3 ] = { 0.0, 0.0, 0.0 }, fn[ 3 ];
    for( int i = 0; i < n; ++i )
      for( int d = 0; d < 3; ++d )
        center[ d ] += x[ e0[ i ] ][ d ];
    const double inv = 1.0 / n;
    for( int d = 0; d < 3; ++d )
      center[ d ] *= inv;
    outer_normal3( x[ e0[ 0 ] ], x[ e0[ 1 ] ], x[ e0[ 2 ] ], fn );
    double fv = 0.0;
    for( int i = 0; i < n; ++i )
      fv += face_edge_term( x[ e0[ i ] ], x[ e1[ i ] ], fn );
    return dot3( center, fn ) * fabs( fv / 2.0 );
  }

  // classification of the nodes, 3d/intersect.hh:99-112
  template< class T = topo::Cube >
  VOFX_HD inline void clip_classify ( const Hex &poly, const Hs3 &hs, unsigned &innerMask, unsigned &boundaryMask )
  {
    innerMask = boundaryMask = 0;
    for( int i = 0; i < T::kNodes; ++i )
    {
      const double ls = level_set( hs, poly.x[ i ] );
      if( ls > -kEps )
      {
        innerMask |= 1u << i;
        if( fabs( ls ) < kEps )
          boundaryMask |= 1u << i;
      }
    }
  }

  // volume of { hex } n { hs > 0 }: 3d/intersect.hh:72-263 + 3d/polyhedron.hh:317-324
  template< class T = topo::Cube >
  VOFX_HD inline double clip_volume ( const Hex &poly, const Hs3 &hs )
  {
    if( !hs_valid( hs ) )
      return fabs( 0.0 / 3.0 );
    unsigned innerMask, boundaryMask;
    clip_classify< T >( poly, hs, innerMask, boundaryMask );
    ClipTopo t;
    if( !clip_walk< T >( innerMask, boundaryMask, t ) )
      return fabs( 0.0 / 3.0 );
    double x[ 20 ][ 3 ];
    for( int k = 0; k < t.n; ++k )
      for( int d = 0; d < 3; ++d )
        x[ k ][ d ] = poly.x[ t.innerNode[ k ] ][ d ];
    for( int j = 0; j < t.nNew; ++j )
      edge_plane_point( poly.x[ t.cutA[ j ] ], poly.x[ t.cutB[ j ] ], hs, x[ t.n + j ] );
    double vol = 0.0;
    for( int f = 0; f < t.nFaces; ++f )
      vol += clipped_face_term( x, t.fe0[ f ], t.fe1[ f ], t.flen[ f ] );
    return fabs( vol / 3.0 );
  }

  // getVolumeFraction, geometry/algorithm.hh:39-45
  VOFX_HD inline double volume_fraction ( const Hex &c, const Hs3 &hs ) { return clip_volume( c, hs ) / hex_volume( c ); }

  // truncVolume( upwindPolygon3d( face geometry, v ), hs ), 3d/upwindpolygon.hh:24-74.  face = 2*axis+side of the cell
  // with lower corner lo; corners of the face in dune order (spanned axes ascending); swept box = face u (face - v).
  VOFX_HD inline double flux ( const double *lo, const double *h, int face, const double *v, const Hs3 &hs )
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
    Hex up;
    for( int i = 0; i < 4; ++i )
      for( int d = 0; d < 3; ++d )
      {
        const double moved = c[ i ][ d ] - v[ d ];
        up.x[ shiftedFirst ? i : i + 4 ][ d ] = moved;
        up.x[ shiftedFirst ? i + 4 : i ][ d ] = c[ i ][ d ];
      }
    return clip_volume( up, hs );
  }

  // ---- G2: plane-constant solve -------------------------------------------------------------------------------------

  // rotateToReferenceFrame, 3d/rotation.hh:17-53: rotate so that the outer normal n becomes e_z
  template< class T = topo::Cube >
  VOFX_HD inline void rotate_to_reference_frame ( const double *n, const Hex &cell, Hex &out )
  {
    if( n[ 0 ] == 0.0 && n[ 1 ] == 0.0 && n[ 2 ] == 1.0 )
    {
      out = cell;
      return;
    }
    double R[ 3 ][ 3 ] = { { 0.0, 0.0, 0.0 }, { 0.0, 0.0, 0.0 }, { 0.0, 0.0, 0.0 } };
    if( n[ 0 ] == 0.0 && n[ 1 ] == 0.0 && n[ 2 ] == -1.0 )
    {
      R[ 0 ][ 0 ] = 1.0;
      R[ 1 ][ 1 ] = -1.0;
      R[ 2 ][ 2 ] = -1.0;
    }
    else
    {
      const double gamma = ( 1 - n[ 2 ] ) / ( n[ 0 ] * n[ 0 ] + n[ 1 ] * n[ 1 ] );
      R[ 0 ][ 0 ] = n[ 1 ] * n[ 1 ] * gamma + n[ 2 ];
      R[ 0 ][ 1 ] = -n[ 0 ] * n[ 1 ] * gamma;
      R[ 1 ][ 0 ] = R[ 0 ][ 1 ];
      R[ 0 ][ 2 ] = -n[ 0 ];
      R[ 1 ][ 1 ] = n[ 0 ] * n[ 0 ] * gamma + n[ 2 ];
      R[ 1 ][ 2 ] = -n[ 1 ];
      R[ 2 ][ 0 ] = n[ 0 ];
      R[ 2 ][ 1 ] = n[ 1 ];
      R[ 2 ][ 2 ] = n[ 2 ];
    }
    for( int i = 0; i < T::kNodes; ++i )
      for( int r = 0; r < 3; ++r )
      {
        double y = 0.0;
        for( int c = 0; c < 3; ++c )
          y += R[ r ][ c ] * cell.x[ i ][ c ];
        out.x[ i ][ r ] = y;
      }
  }

  // a list of at most 3 directed edges leaving a node
  struct Dirs
  {
    signed char n;
    signed char e[ 4 ];   // one spare slot: the reference reads one element past the end after an erase (see sweep_init)
  };

  template< class T = topo::Cube >
  VOFX_INLINE void dirs_of_node ( int node, Dirs &d )
  {
    d.n = 3;
    for( int k = 0; k < 3; ++k )
      d.e[ k ] = (signed char) T::outEdge( node, k );
    d.e[ 3 ] = 0;
  }

  VOFX_INLINE void dirs_erase ( Dirs &d, int i )
  {
    for( int j = i; j + 1 < d.n; ++j )
      d.e[ j ] = d.e[ j + 1 ];
    d.n--;
  }

  // std::sort on fewer than 16 elements is libstdc++'s __insertion_sort; the comparator is cyclic (not a strict weak
  // ordering), so the exact algorithm is part of the specification
  template< class T = topo::Cube >
  VOFX_HD inline void dirs_sort ( Dirs &d )
  {
    for( int i = 1; i < d.n; ++i )
    {
      const signed char val = d.e[ i ];
      if( T::dirLess( val, d.e[ 0 ] ) )
      {
        for( int j = i; j > 0; --j )
          d.e[ j ] = d.e[ j - 1 ];
        d.e[ 0 ] = val;
      }
      else
      {
        int j = i;
        while( T::dirLess( val, d.e[ j - 1 ] ) )
        {
          d.e[ j ] = d.e[ j - 1 ];
          --j;
        }
        d.e[ j ] = val;
      }
    }
  }

  constexpr int kMaxSweep = 8;

  // the cross-section polygon swept upwards through the rotated cell (PolygonWithDirections)
  struct Sweep
  {
    int nn;                       // polygon nodes
    int nEdges;                   // polygon edges (i,(i+1)%nn); 0 while the polygon is a single point
    int nCF;
    bool degenerate;              // set where the reference runs into undefined behaviour
    double px[ kMaxSweep ], py[ kMaxSweep ];
    signed char nodeId[ kMaxSweep ];   // parent node or -1
    signed char corrFace[ kMaxSweep ];
    signed char srcNode[ kMaxSweep ];  // where the node came from: the parent node id (srcEdge == -1), or the index of
    signed char srcEdge[ kMaxSweep ];  // the previous polygon's node that slid here along the directed edge srcEdge
    Dirs dirs[ kMaxSweep ];
  };

  // PolygonWithDirections::directionVector, 3d/polygonwithdirections.hh:64-74
  template< class T = topo::Cube >
  VOFX_HD inline void direction_vector ( const Hex &P, int edge, double *dv )
  {
    const int a = T::edgeN0( edge ), b = T::edgeN1( edge );
    for( int c = 0; c < 3; ++c )
      dv[ c ] = P.x[ b ][ c ] - P.x[ a ][ c ];
    const double norm = norm2( dv[ 0 ], dv[ 1 ] );
    if( norm > 0 )
      vdiv3( dv[ 0 ], dv[ 1 ], dv[ 2 ], norm );
  }

  // PolygonWithDirections::initialize, 3d/polygonwithdirections.hh:76-174.  ids = node ids sorted by height, ds = sorted heights
  // POS = false: topology only (the cooperative path evaluates the positions itself); P may then be null
  template< bool POS, class T = topo::Cube >
  VOFX_HD inline void sweep_init ( Sweep &s, const Hex *Pp, const int *ids, const double *ds )
  {
    s.nn = 0;
    s.nEdges = 0;
    s.nCF = 0;
    s.degenerate = false;

    int N = 1;
    for( ; N < T::kNodes && ds[ N ] == ds[ 0 ]; N++ )
      ;

    if( N == 1 )
    {
      if( POS )
      {
        s.px[ 0 ] = Pp->x[ ids[ 0 ] ][ 0 ];
        s.py[ 0 ] = Pp->x[ ids[ 0 ] ][ 1 ];
      }
      s.nodeId[ 0 ] = (signed char) ids[ 0 ];
      dirs_of_node< T >( ids[ 0 ], s.dirs[ 0 ] );
      dirs_sort< T >( s.dirs[ 0 ] );
      s.nn = 1;
    }
    else if( N == 2 )
    {
      for( int k = 0; k < 2; ++k )
      {
        if( POS )
        {
          s.px[ k ] = Pp->x[ ids[ k ] ][ 0 ];
          s.py[ k ] = Pp->x[ ids[ k ] ][ 1 ];
        }
        s.nodeId[ k ] = (signed char) ids[ k ];
      }
      Dirs d1, d2;
      dirs_of_node< T >( ids[ 0 ], d1 );
      dirs_of_node< T >( ids[ 1 ], d2 );
      // :107-114: both loops keep running after the erase; an index equal to the new size reads the stale copy of
      // the former last element that std::vector::erase leaves behind
      for( int i = 0; i < d1.n; ++i )
        for( int j = 0; j < d2.n; ++j )
          if( d1.e[ i ] == T::reverseEdge( d2.e[ j ] ) )
          {
            const signed char stale1 = d1.e[ d1.n - 1 ], stale2 = d2.e[ d2.n - 1 ];
            dirs_erase( d1, i );
            dirs_erase( d2, j );
            d1.e[ d1.n ] = stale1;
            d2.e[ d2.n ] = stale2;
          }
      dirs_sort< T >( d1 );
      dirs_sort< T >( d2 );
      s.dirs[ 0 ] = d1;
      s.dirs[ 1 ] = d2;
      s.nn = 2;
    }
    else
    {
      unsigned lowMask = 0;
      for( int i = 0; i < N; ++i )
        lowMask |= 1u << ids[ i ];
      int found = -1;
      for( int f = 0; f < T::kFaces && found < 0; ++f )
      {
        bool all = true;
        for( int k = 0; k < T::faceLen( f ); ++k )
          all = all && ( lowMask & ( 1u << T::faceNode( f, k ) ) );
        if( all )
          found = f;
      }
      if( found < 0 || N > T::faceLen( found ) )
      {
        s.degenerate = true;   // the reference indexes an empty / too short nodeIds_
        return;
      }
      for( int k = 0; k < T::faceLen( found ); ++k )
      {
        const int id = T::faceNode( found, k );
        s.nodeId[ k ] = (signed char) id;
        if( POS )
        {
          s.px[ k ] = Pp->x[ id ][ 0 ];
          s.py[ k ] = Pp->x[ id ][ 1 ];
        }
      }
      s.nn = T::faceLen( found );
      for( int k = 0; k < N; ++k )
      {
        Dirs dd;
        dirs_of_node< T >( s.nodeId[ k ], dd );
        for( int j = 0; j < N; ++j )
          for( int i = 0; i < dd.n; ++i )
            if( T::edgeN1( dd.e[ i ] ) == s.nodeId[ j ] )
            {
              dirs_erase( dd, i );
              --i;
            }
        dirs_sort< T >( dd );
        s.dirs[ k ] = dd;
      }
    }

    for( int i = 0; i < s.nn; ++i )
    {
      s.srcNode[ i ] = s.nodeId[ i ];
      s.srcEdge[ i ] = -1;
    }
    if( s.nn != 1 )
    {
      for( int i = 0; i < s.nn; ++i )
      {
        const int j = ( i + 1 == s.nn ) ? 0 : i + 1;
        const int e = T::edgeId( s.nodeId[ j ], s.nodeId[ i ] );
        if( e < 0 )
        {
          s.degenerate = true;   // reference: DUNE_THROW( InvalidStateException, "Edge was not found in any face." )
          return;
        }
        s.corrFace[ s.nCF++ ] = (signed char) T::faceOfEdge( e );
      }
      s.nEdges = s.nn;
    }
  }

  // PolygonWithDirections::evolveToNextPolygon, 3d/polygonwithdirections.hh:177-290
  // z: the eight node heights (P.x[ . ][ 2 ]); P only for POS = true
  template< bool POS, class T = topo::Cube >
  VOFX_HD inline void sweep_evolve ( Sweep &s, const Hex *Pp, const double *z, double dk, double dk1, double h )
  {
    double nx[ kMaxSweep ], ny[ kMaxSweep ];
    signed char nid[ kMaxSweep ], nsrcN[ kMaxSweep ], nsrcE[ kMaxSweep ];
    Dirs nd[ kMaxSweep ];
    int nNew = 0;
    const int oldN = s.nn;
    // corrFace is rebuilt while the old polygon is still being read: keep the old node data, overwrite the faces
    s.nCF = 0;
    s.nEdges = 0;

    for( int n = 0; n < oldN; ++n )
      for( int i = 0; i < s.dirs[ n ].n; ++i )
      {
        const int dirEdge = s.dirs[ n ].e[ i ];
        const int target = T::edgeN1( dirEdge );
        const double z1 = z[ target ];

        if( fabs( z1 - dk1 ) < 1e-10 )
        {
          if( nNew > 0 && target == nid[ nNew - 1 ] )
          {
            s.corrFace[ s.nCF - 1 ] = (signed char) T::faceOfEdge( dirEdge );
            continue;
          }
          if( nNew >= kMaxSweep )
          {
            s.degenerate = true;
            return;
          }
          if( POS )
          {
            nx[ nNew ] = Pp->x[ target ][ 0 ];
            ny[ nNew ] = Pp->x[ target ][ 1 ];
          }
          nid[ nNew ] = (signed char) target;
          nsrcN[ nNew ] = (signed char) target;
          nsrcE[ nNew ] = -1;
          Dirs dd;
          dd.n = 0;
          dd.e[ 3 ] = 0;
          for( int a = 0; a < 3; ++a )
          {
            const int e = T::outEdge( target, a );
            if( z[ T::edgeN1( e ) ] > dk )
              dd.e[ dd.n++ ] = (signed char) e;
          }
          dirs_sort< T >( dd );
          nd[ nNew ] = dd;
          nNew++;
          s.corrFace[ s.nCF++ ] = (signed char) T::faceOfEdge( dirEdge );
        }
        else if( z1 > dk1 )
        {
          if( nNew >= kMaxSweep )
          {
            s.degenerate = true;
            return;
          }
          if( POS )
          {
            double dv[ 3 ];
            direction_vector< T >( *Pp, dirEdge, dv );
            const double f = vdiv( h, dv[ 2 ] );
            const double ax = dv[ 0 ] * f, ay = dv[ 1 ] * f;
            nx[ nNew ] = s.px[ n ] + ax;
            ny[ nNew ] = s.py[ n ] + ay;
          }
          nid[ nNew ] = -1;
          nsrcN[ nNew ] = (signed char) n;
          nsrcE[ nNew ] = (signed char) dirEdge;
          nd[ nNew ].n = 1;
          nd[ nNew ].e[ 0 ] = (signed char) dirEdge;
          nNew++;
          s.corrFace[ s.nCF++ ] = (signed char) T::faceOfEdge( dirEdge );
        }
      }

    if( nNew == 0 )
    {
      s.degenerate = true;   // the reference reads newNodeIds[ 0 ] of an empty vector
      return;
    }
    if( nid[ 0 ] == nid[ nNew - 1 ] && nid[ 0 ] != -1 )
    {
      nNew--;
      s.nCF--;
    }

    // drop directions that point to a node of the new polygon, :255-262 (after an erase the reference re-examines
    // the element before the erased one against the remaining node ids; that element never matches)
    for( int a = 0; a < nNew; ++a )
    {
      Dirs &dd = nd[ a ];
      for( int i = 0; i < dd.n; ++i )
        for( int b = 0; b < nNew; ++b )
          if( i >= 0 && T::edgeN1( dd.e[ i ] ) == nid[ b ] )
          {
            dirs_erase( dd, i );
            --i;
          }
    }

    if( nNew != 1 )
      s.nEdges = nNew;

    // two consecutive parent nodes joined by a parent edge: take the face on the other side of that edge, :271-287
    if( nNew > 2 )
      for( int i = 0; i < s.nEdges; ++i )
      {
        const int id0 = nid[ i ];
        const int id1 = nid[ ( i + 1 == nNew ) ? 0 : i + 1 ];
        if( id0 != -1 && id1 != -1 && T::edgeId( id0, id1 ) >= 0 )
          s.corrFace[ i ] = (signed char) T::faceOfEdge( T::edgeId( id1, id0 ) );
      }

    s.nn = nNew;
    for( int a = 0; a < nNew; ++a )
    {
      if( POS )
      {
        s.px[ a ] = nx[ a ];
        s.py[ a ] = ny[ a ];
      }
      s.nodeId[ a ] = nid[ a ];
      s.srcNode[ a ] = nsrcN[ a ];
      s.srcEdge[ a ] = nsrcE[ a ];
      s.dirs[ a ] = nd[ a ];
    }
  }

  struct Cubic
  {
    double A, B, C, target;
    // V(h) of geometry/algorithm.hh:211 (minus the target of :223)
    VOFX_HD double operator() ( double h ) const { return ( C * h + B * h * h / 2.0 + vdiv( A * h * h * h, 6.0 ) ) - target; }
  };

  // locateHalfSpace( Polyhedron, innerNormal, fraction ), geometry/algorithm.hh:119-236.  Returns the plane constant;
  // NaN where the reference's behaviour is undefined (it would read out of bounds or throw).
  template< class T = topo::Cube >
  VOFX_HD inline double locate ( const Hex &cell, const double *innerNormal, double fraction )
  {
    double outerNormal[ 3 ];
    for( int k = 0; k < 3; ++k )
      outerNormal[ k ] = innerNormal[ k ] * -1.0;

    fraction = ( fraction > 1.0 ) ? 1.0 : fraction;
    fraction = ( fraction < 0.0 ) ? 0.0 : fraction;

    const double givenVolume = fraction * hex_volume< T >( cell );

    Hex P;
    rotate_to_reference_frame< T >( outerNormal, cell, P );

    // node heights sorted by insertion sort (std::sort, N = 8 < 16; stable: ties keep ascending node id), :147-158
    int ids[ 8 ];
    double ds[ 8 ];
    for( int i = 0; i < T::kNodes; ++i )
    {
      const double di = P.x[ i ][ 2 ];
      int j = i;
      while( j > 0 && di < ds[ j - 1 ] )
      {
        ds[ j ] = ds[ j - 1 ];
        ids[ j ] = ids[ j - 1 ];
        --j;
      }
      ds[ j ] = di;
      ids[ j ] = i;
    }
    // std::unique with |x - y| < 1e-10 against the last kept element, :160-162
    double dU[ 8 ];
    int nU = 0;
    dU[ nU++ ] = ds[ 0 ];
    for( int i = 1; i < T::kNodes; ++i )
      if( !( fabs( dU[ nU - 1 ] - ds[ i ] ) < 1e-10 ) )
        dU[ nU++ ] = ds[ i ];

    Sweep s;
    sweep_init< true, T >( s, &P, ids, ds );
    if( s.degenerate )
      return NAN;

    double Vd_k = 0.0;
    for( int k = 0; k + 1 < nU; ++k )
    {
      const int Nn = s.nn;
      double A = 0.0;
      for( int n = 0; n < Nn; ++n )
      {
        const int nxt = ( n + 1 == Nn ) ? 0 : n + 1;
        if( s.dirs[ n ].n == 0 || s.dirs[ nxt ].n == 0 )
          return NAN;   // reference: assert( !directions( n ).empty() ), then an out-of-range read
        const int epsn = s.dirs[ n ].n - 1;
        double v1[ 3 ], v2[ 3 ];
        direction_vector< T >( P, s.dirs[ n ].e[ epsn ], v1 );
        direction_vector< T >( P, s.dirs[ nxt ].e[ 0 ], v2 );
        A -= vdiv( v1[ 0 ] * v2[ 1 ] - v1[ 1 ] * v2[ 0 ], v1[ 2 ] * v2[ 2 ] );
        for( int l = 0; l < epsn; ++l )
        {
          double w1[ 3 ], w2[ 3 ];
          direction_vector< T >( P, s.dirs[ n ].e[ l ], w1 );
          direction_vector< T >( P, s.dirs[ n ].e[ l + 1 ], w2 );
          A -= vdiv( w1[ 0 ] * w2[ 1 ] - w1[ 1 ] * w2[ 0 ], w1[ 2 ] * w2[ 2 ] );
        }
      }

      double B = 0.0;
      for( int i = 0; i < s.nEdges; ++i )
      {
        if( i >= s.nCF )
          return NAN;
        const int f = s.corrFace[ i ];
        double normal[ 3 ];
        outer_normal3( P.x[ T::faceNode( f, 0 ) ], P.x[ T::faceNode( f, 1 ) ], P.x[ T::faceNode( f, 2 ) ], normal );
        const double norm = norm2( normal[ 0 ], normal[ 1 ] );
        if( norm > 0 )
        {
          const int j = ( i + 1 == s.nn ) ? 0 : i + 1;
          const double ex = s.px[ j ] - s.px[ i ], ey = s.py[ j ] - s.py[ i ];
          B -= vdiv( norm2( ex, ey ) * normal[ 2 ], norm );
        }
      }

      // PolygonWithDirections::volume, :56-62, with the 2D Edge::outerNormal of 3d/polyhedron.hh:49-60
      double C = 0.0;
      for( int i = 0; i < s.nEdges; ++i )
      {
        const int j = ( i + 1 == s.nn ) ? 0 : i + 1;
        const double ex = s.px[ j ] - s.px[ i ], ey = s.py[ j ] - s.py[ i ];
        double onx = ey, ony = -ex;
        const double nrm = norm2( onx, ony );
        vdiv2( onx, ony, nrm );
        const double cx = ( s.px[ i ] + s.px[ j ] ) * 0.5, cy = ( s.py[ i ] + s.py[ j ] ) * 0.5;
        C += dot2( cx, cy, onx, ony ) * norm2( ex, ey );
      }
      C = fabs( C / 2.0 );

      const double h = dU[ k + 1 ] - dU[ k ];
      Cubic V = { A, B, C, 0.0 };
      const double Vd_k1 = Vd_k + V( h );

      if( fabs( Vd_k1 - givenVolume ) < 1e-14 )
        return dU[ k + 1 ];
      else if( Vd_k1 > givenVolume )
      {
        V.target = givenVolume - Vd_k;
        return dU[ k ] + brent( V, 0.0, h, 1e-14 );
      }
      else
      {
        double zs[ 8 ];
        for( int i = 0; i < T::kNodes; ++i )
          zs[ i ] = P.x[ i ][ 2 ];
        sweep_evolve< true, T >( s, &P, zs, dU[ k ], dU[ k + 1 ], h );
        if( s.degenerate )
          return NAN;
      }
      Vd_k = Vd_k1;
    }
    return dU[ nU - 1 ];
  }

  // ---- cells and intersections of unstructured 3D grids, given by their corners (vofx_mesh.cu) ---------------------------

  // locateHalfSpace( makePolyhedron( geometry ), innerNormal, fraction ) for a tetrahedron (nc = 4, 3d/polyhedron.hh:374-385)
  // or a hexahedron (nc = 8, :387-401); corners in DUNE corner order, [ nc ][ 3 ].  NaN for any other corner count
  // (the reference throws InvalidStateException, "Invalid GeometryType.").
  VOFX_HD inline double locate_corners ( const double *corners, int nc, const double *innerNormal, double fraction )
  {
    Hex cell;
    for( int i = 0; i < nc && i < 8; ++i )
      for( int d = 0; d < 3; ++d )
        cell.x[ i ][ d ] = corners[ 3 * i + d ];
    if( nc == 4 )
      return locate< topo::Tet >( cell, innerNormal, fraction );
    if( nc == 8 )
      return locate< topo::Cube >( cell, innerNormal, fraction );
    return NAN;
  }

  // truncVolume( upwindPolygon3d( iGeometry, v ), hs ), 3d/upwindpolygon.hh:24-74 + geometry/upwindpolygon.hh:58-62, for an
  // intersection with nc = 3 (prism, :47-57) or nc = 4 (cube topology, :59-70) corners c[ nc ][ 3 ]
  VOFX_HD inline double flux_corners ( const double *c, int nc, const double *v, const Hs3 &hs )
  {
    double e1[ 3 ], e2[ 3 ], cr[ 3 ];
    for( int d = 0; d < 3; ++d )
    {
      e1[ d ] = c[ 3 + d ] - c[ d ];
      e2[ d ] = c[ 6 + d ] - c[ d ];
    }
    cross3( e1, e2, cr );
    const bool shiftedFirst = dot3( cr, v ) < 0.0;
    Hex up;
    for( int i = 0; i < nc && i < 4; ++i )
      for( int d = 0; d < 3; ++d )
      {
        const double moved = c[ 3 * i + d ] - v[ d ];
        up.x[ shiftedFirst ? i : i + nc ][ d ] = moved;
        up.x[ shiftedFirst ? i + nc : i ][ d ] = c[ 3 * i + d ];
      }
    if( nc == 3 )
      return clip_volume< topo::Prism >( up, hs );
    if( nc == 4 )
      return clip_volume< topo::Cube >( up, hs );
    return NAN;
  }

  // ---- G6: interface polygon of a cell and its centroid ---------------------------------------------------------------

  constexpr int kMaxIface = 12;

  struct Face3
  {
    int n;
    double v[ kMaxIface ][ 3 ];
    double unitNormal[ 3 ];
  };

  // Dune::VoF::Face constructor, 3d/face.hh:19-24 (vertex( i ) wraps modulo size)
  VOFX_HD inline void face_init_normal ( Face3 &f )
  {
    const int n = f.n;
    double a[ 3 ], b[ 3 ];
    for( int k = 0; k < 3; ++k )
    {
      a[ k ] = f.v[ 2 % n ][ k ] - f.v[ 1 % n ][ k ];
      b[ k ] = f.v[ 0 ][ k ] - f.v[ 1 % n ][ k ];
    }
    cross3( a, b, f.unitNormal );
    const double nrm = norm3( f.unitNormal );
    vdiv3( f.unitNormal[ 0 ], f.unitNormal[ 1 ], f.unitNormal[ 2 ], nrm );
  }

  // __impl::intersect( Polyhedron, HyperPlane ), 3d/intersect.hh:274-355 (eps = 1e-12)
  VOFX_HD inline void plane_cut ( const Hex &poly, const Hs3 &hs, Face3 &out )
  {
    out.n = 0;
    if( !hs_valid( hs ) )
      return;
    const double eps = 1e-12;
    unsigned innerMask = 0;
    int n = 0;
    for( int i = 0; i < 8; ++i )
      if( level_set( hs, poly.x[ i ] ) <= eps )
      {
        innerMask |= 1u << i;
        n++;
      }
    if( n == 0 || n == 8 )
    {
      for( int i = 0; i < 8; ++i )
        if( level_set( hs, poly.x[ i ] ) >= -eps )
        {
          out.n = 1;
          for( int k = 0; k < 3; ++k )
            out.v[ 0 ][ k ] = poly.x[ i ][ k ];
          face_init_normal( out );
          return;
        }
      return;
    }

    double e0[ kMaxIface ][ 3 ], e1[ kMaxIface ][ 3 ];
    int nE = 0;
    for( int f = 0; f < 6; ++f )
    {
      double last[ 3 ] = { 0.0, 0.0, 0.0 };
      for( int k = 0; k < 4; ++k )
      {
        const int edge = cube::faceEdge( f, k );
        const int a = cube::edgeN0( edge ), b = cube::edgeN1( edge );
        if( bool( innerMask & ( 1u << a ) ) != bool( innerMask & ( 1u << b ) ) )
        {
          double isPoint[ 3 ];
          edge_plane_point( poly.x[ a ], poly.x[ b ], hs, isPoint );
          if( !( last[ 0 ] == 0.0 && last[ 1 ] == 0.0 && last[ 2 ] == 0.0 ) )
          {
            if( nE < kMaxIface )
            {
              for( int d = 0; d < 3; ++d )
              {
                e0[ nE ][ d ] = last[ d ];
                e1[ nE ][ d ] = isPoint[ d ];
              }
              nE++;
            }
          }
          else
            for( int d = 0; d < 3; ++d )
              last[ d ] = isPoint[ d ];
        }
      }
    }

    // chain the edges, :317-340
    for( int i = 0; i + 1 < nE; ++i )
    {
      double diff[ 3 ];
      for( int k = 0; k < 3; ++k )
        diff[ k ] = e1[ i ][ k ] - e0[ i + 1 ][ k ];
      if( norm3( diff ) < eps )
        continue;
      int pos = -1;
      for( int j = i + 1; j < nE && pos < 0; ++j )
      {
        for( int k = 0; k < 3; ++k )
          diff[ k ] = e1[ i ][ k ] - e0[ j ][ k ];
        if( norm3( diff ) < eps )
          pos = j;
      }
      if( pos >= 0 )
      {
        for( int k = 0; k < 3; ++k )
        {
          double t = e0[ i + 1 ][ k ]; e0[ i + 1 ][ k ] = e0[ pos ][ k ]; e0[ pos ][ k ] = t;
          t = e1[ i + 1 ][ k ]; e1[ i + 1 ][ k ] = e1[ pos ][ k ]; e1[ pos ][ k ] = t;
        }
        continue;
      }
      for( int j = i + 1; j < nE && pos < 0; ++j )
      {
        for( int k = 0; k < 3; ++k )
          diff[ k ] = e1[ i ][ k ] - e1[ j ][ k ];
        if( norm3( diff ) < eps )
          pos = j;
      }
      if( pos >= 0 )
      {
        for( int k = 0; k < 3; ++k )
        {
          double t = e0[ i + 1 ][ k ]; e0[ i + 1 ][ k ] = e0[ pos ][ k ]; e0[ pos ][ k ] = t;
          t = e1[ i + 1 ][ k ]; e1[ i + 1 ][ k ] = e1[ pos ][ k ]; e1[ pos ][ k ] = t;
          t = e0[ i + 1 ][ k ]; e0[ i + 1 ][ k ] = e1[ i + 1 ][ k ]; e1[ i + 1 ][ k ] = t;
        }
        continue;
      }
    }

    out.n = nE;
    for( int i = 0; i < nE; ++i )
      for( int k = 0; k < 3; ++k )
        out.v[ i ][ k ] = e0[ i ][ k ];
    if( nE > 0 )
      face_init_normal( out );
  }

  // sum over the edges of ( edgeCenter . cross( edge, unitNormal ) ) for the polygon (v0, v1, ..), 3d/face.hh:52-67,92-117
  VOFX_HD inline double face_area_sum ( const double *v0, const double *v1, const double *un )
  {
    double center[ 3 ], edge[ 3 ], en[ 3 ];
    for( int k = 0; k < 3; ++k )
    {
      center[ k ] = ( v1[ k ] + v0[ k ] ) * 0.5;
      edge[ k ] = v1[ k ] - v0[ k ];
    }
    cross3( edge, un, en );
    return dot3( center, en );
  }

  // Face::volume, 3d/face.hh:92-98
  VOFX_HD inline double face_measure ( const Face3 &f )
  {
    double sum = 0.0;
    const int n = f.n;
    for( int i = 0; i < n; ++i )
      sum += face_area_sum( f.v[ i ], f.v[ ( i + 1 == n ) ? 0 : i + 1 ], f.unitNormal );
    return fabs( sum ) / 2.0;
  }

  // Face::centroid, 3d/face.hh:69-90 (NaN for an empty face, as the reference's 0/0)
  VOFX_HD inline void face_centroid ( const Face3 &f, double *centroid )
  {
    const int n = f.n;
    if( n == 0 )
    {
      centroid[ 0 ] = centroid[ 1 ] = centroid[ 2 ] = NAN;
      return;
    }
    double c[ 3 ] = { 0.0, 0.0, 0.0 };
    for( int i = 0; i < n; ++i )
      for( int k = 0; k < 3; ++k )
        c[ k ] += f.v[ i ][ k ];
    for( int k = 0; k < 3; ++k )
      c[ k ] /= (double) n;

    centroid[ 0 ] = centroid[ 1 ] = centroid[ 2 ] = 0.0;
    for( int i = 0; i < n; ++i )
    {
      const double *a = f.v[ i ], *b = f.v[ ( i + 1 == n ) ? 0 : i + 1 ];
      // triangleVolume( { a, b, c } )
      double sum = 0.0;
      sum += face_area_sum( a, b, f.unitNormal );
      sum += face_area_sum( b, c, f.unitNormal );
      sum += face_area_sum( c, a, f.unitNormal );
      const double volTriang = fabs( sum ) / 2.0;
      for( int k = 0; k < 3; ++k )
      {
        double ct = c[ k ];
        ct += a[ k ];
        ct += b[ k ];
        ct /= 3.0;
        centroid[ k ] += volTriang * ct;
      }
    }
    const double vol = face_measure( f );
    for( int k = 0; k < 3; ++k )
      centroid[ k ] /= vol;
  }

} // namespace vofx
