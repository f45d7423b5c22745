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
    // reference's summation order, the cyclic node sequences (concatenated).  ref < 8: original node; ref >= 8: cut node.
    constexpr int kMaxClipEdges = 36;
    struct ClipCase
    {
      unsigned char nCut;            // 0xFF: not representable here (take the serial path)
      unsigned char nFaces;
      unsigned char nEdges;
      unsigned char pad0;
      unsigned char cut[ 6 ];        // a | b << 3: Edge( a, b ).intersection( plane ), operand order as in the walk
      unsigned char faceStart[ 8 ];  // face f owns the edges faceStart[f] .. faceStart[f+1]-1
      unsigned char refs[ kMaxClipEdges ];   // per edge: first node (low nibble) | face (high nibble); second node = next entry (cyclic in the face)
      unsigned char pad1[ 10 ];
    };
    static_assert( sizeof( ClipCase ) == 64, "ClipCase must be 64 bytes" );

    // Nodes ON the plane (|levelSet| < eps, 3d/intersect.hh:104-108) change the walk (no cut node on their edges, the cap
    // passes through them).  They are not rare: a cell whose fraction reached 0 or 1 is located with the plane through one of
    // its nodes, and the swept boxes of its faces share that node - 10 % of the clips of the 512^3 LeVeque run after 60 steps.
    // A second table, indexed by the ternary code of ( outer / inner / on the plane ) per node, covers them.
    constexpr int kClipBoundaryCases = 6561;   // 3^8

    VOFX_HD inline unsigned clip_ternary ( unsigned innerMask, unsigned boundaryMask )
    {
      unsigned code = 0, power = 1;
      for( int i = 0; i < 8; ++i )
      {
        code += ( ( ( innerMask >> i ) & 1u ) + ( ( boundaryMask >> i ) & 1u ) ) * power;
        power *= 3;
      }
      return code;
    }

    // one table entry from the reference's walk (clip_walk); nCut = 0xFE: empty result, 0xFF: not representable (serial path)
    inline void fill_clip_case ( ClipCase &c, unsigned innerMask, unsigned boundaryMask )
    {
      for( unsigned i = 0; i < sizeof( ClipCase ); ++i )
        reinterpret_cast< unsigned char * >( &c )[ i ] = 0;
      ClipTopo t;
      if( !clip_walk( innerMask, boundaryMask, t ) )
      {
        c.nCut = 0xFE;       // empty result, 3d/intersect.hh:115-120
        return;
      }
      bool regular = t.nNew <= 6 && t.nFaces <= 7;
      int nEdges = 0;
      for( int f = 0; regular && f < t.nFaces; ++f )
      {
        const int len = t.flen[ f ];
        nEdges += len;
        regular = len <= 6 && nEdges <= kMaxClipEdges;
        for( int k = 0; regular && k < len; ++k )
          regular = t.fe1[ f ][ k ] == t.fe0[ f ][ ( k + 1 ) % len ];
      }
      if( !regular )
      {
        c.nCut = 0xFF;
        return;
      }
      c.nCut = (unsigned char) t.nNew;
      c.nFaces = (unsigned char) t.nFaces;
      for( int j = 0; j < t.nNew; ++j )
        c.cut[ j ] = (unsigned char) ( t.cutA[ j ] | ( t.cutB[ j ] << 3 ) );
      int e = 0;
      for( int f = 0; f < t.nFaces; ++f )
      {
        c.faceStart[ f ] = (unsigned char) e;
        for( int k = 0; k < t.flen[ f ]; ++k )
        {
          const int id = t.fe0[ f ][ k ];
          c.refs[ e++ ] = (unsigned char) ( ( id < t.n ? t.innerNode[ id ] : 8 + ( id - t.n ) ) | ( f << 4 ) );
        }
      }
      for( int f = t.nFaces; f < 8; ++f )
        c.faceStart[ f ] = (unsigned char) e;
      c.nEdges = (unsigned char) e;
    }

    // table generators: the reference's walk for every inner mask (no node on the plane) ...
    inline void make_clip_table ( ClipCase *table )
    {
      for( unsigned mask = 0; mask < 256; ++mask )
        fill_clip_case( table[ mask ], mask, 0u );
    }

    // ... and for every ( outer / inner / on the plane ) pattern, indexed by clip_ternary
    inline void make_clip_boundary_table ( ClipCase *table )
    {
      for( unsigned code = 0; code < (unsigned) kClipBoundaryCases; ++code )
      {
        unsigned inner = 0, bnd = 0, rest = code;
        for( int i = 0; i < 8; ++i )
        {
          const unsigned digit = rest % 3;
          rest /= 3;
          if( digit >= 1 )
            inner |= 1u << i;
          if( digit == 2 )
            bnd |= 1u << i;
        }
        fill_clip_case( table[ code ], inner, bnd );
      }
    }

    struct ClipScratch
    {
      double x[ 3 ][ 16 ];             // coordinate d of ref r: 0..7 nodes of the hexahedron, 8..13 cut nodes
      double fn[ 3 ][ 8 ];             // outer normals of the faces of the clipped polyhedron
      double cf[ 8 ];                  // face centre . face normal
      double eterm[ kMaxClipEdges ];   // edge terms of Face::volume
      double term[ 8 ];                // face terms of Polyhedron::volume
      unsigned char inner[ 8 ], bnd[ 8 ];
      double bankPad[ 1 ];
    };
    static_assert( sizeof( ClipScratch ) % 8 == 0 && ( sizeof( ClipScratch ) / 8 ) % 2 == 1, "ClipScratch stride must be an odd number of doubles" );

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
    VOFX_HD inline int clip_phase_cut ( int lane, ClipScratch &S, const Hs3 &hs, const ClipCase *table, const ClipCase *&cs,
                                        const ClipCase *boundaryTable = nullptr )
    {
      unsigned innerMask, boundaryMask;
      clip_masks( S, innerMask, boundaryMask );
      cs = table + innerMask;
      if( innerMask == 0 )
        return 1;
      if( boundaryMask != 0 )
      {
        if( !boundaryTable )
          return 2;
        cs = boundaryTable + clip_ternary( innerMask, boundaryMask );
        if( cs->nCut == 0xFE )
          return 1;
      }
      if( cs->nCut == 0xFF )
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

    // phase C1: lane f sets up face f of the clipped polyhedron: centre (mean of the edges' first nodes) and outer
    // normal from the first three nodes, 3d/polyhedron.hh:187-214
    VOFX_HD inline void clip_phase_faces ( int lane, ClipScratch &S, const ClipCase &cs )
    {
      if( lane >= cs.nFaces )
        return;
      const int e0 = cs.faceStart[ lane ], len = cs.faceStart[ lane + 1 ] - e0;
      double center[ 3 ] = { 0.0, 0.0, 0.0 }, x3[ 3 ][ 3 ], fn[ 3 ];
#pragma unroll 1
      for( int k = 0; k < len; ++k )
      {
        const int r = cs.refs[ e0 + k ] & 15;
        for( int d = 0; d < 3; ++d )
          center[ d ] += S.x[ d ][ r ];
      }
      for( int k = 0; k < 3; ++k )
      {
        const int r = cs.refs[ e0 + k ] & 15;
        for( int d = 0; d < 3; ++d )
          x3[ k ][ d ] = S.x[ d ][ r ];
      }
      // 1.0 / nodeIds_.size(): a face of a clipped hexahedron has 3 .. 7 nodes; the constants are the correctly rounded quotients
      const double inv = ( len == 4 ) ? 1.0 / 4 : ( ( len == 3 ) ? 1.0 / 3 : ( ( len == 5 ) ? 1.0 / 5 : ( ( len == 6 ) ? 1.0 / 6 : vdiv( 1.0, (double) len ) ) ) );
      for( int d = 0; d < 3; ++d )
        center[ d ] *= inv;
      outer_normal3( x3[ 0 ], x3[ 1 ], x3[ 2 ], fn );
      for( int d = 0; d < 3; ++d )
        S.fn[ d ][ lane ] = fn[ d ];
      S.cf[ lane ] = dot3( center, fn );
    }

    // phase C2: the edge terms of Face::volume (3d/polyhedron.hh:216-222), spread evenly over the L lanes.  Two edges
    // per iteration in straight-line code: the two dependent fp64 chains (cross, sqrt, three divisions) interleave.
    VOFX_INLINE void clip_edge_operands ( const ClipScratch &S, const ClipCase &cs, int e, double *x0, double *x1, double *fn )
    {
      const int f = cs.refs[ e ] >> 4;
      const int r0 = cs.refs[ e ] & 15;
      const int r1 = cs.refs[ ( e + 1 < cs.faceStart[ f + 1 ] ) ? e + 1 : cs.faceStart[ f ] ] & 15;
      for( int d = 0; d < 3; ++d )
      {
        x0[ d ] = S.x[ d ][ r0 ];
        x1[ d ] = S.x[ d ][ r1 ];
        fn[ d ] = S.fn[ d ][ f ];
      }
    }

    template< int L >
    VOFX_HD inline void clip_phase_edges ( int lane, ClipScratch &S, const ClipCase &cs )
    {
      const int nEdges = cs.nEdges;
#pragma unroll 1
      for( int ea = lane; ea < nEdges; ea += 2 * L )
      {
        const int eb = ea + L;
        const bool hasB = eb < nEdges;
        double xa0[ 3 ], xa1[ 3 ], fna[ 3 ], xb0[ 3 ], xb1[ 3 ], fnb[ 3 ];
        clip_edge_operands( S, cs, ea, xa0, xa1, fna );
        clip_edge_operands( S, cs, hasB ? eb : ea, xb0, xb1, fnb );
        const double ta = face_edge_term( xa0, xa1, fna );
        const double tb = face_edge_term( xb0, xb1, fnb );
        S.eterm[ ea ] = ta;
        if( hasB )
          S.eterm[ eb ] = tb;
      }
    }

    // phase C3: lane f sums the edge terms of face f in order
    VOFX_HD inline void clip_phase_face_sums ( int lane, ClipScratch &S, const ClipCase &cs )
    {
      if( lane >= cs.nFaces )
        return;
      double fv = 0.0;
#pragma unroll 1
      for( int e = cs.faceStart[ lane ]; e < cs.faceStart[ lane + 1 ]; ++e )
        fv += S.eterm[ e ];
      S.term[ lane ] = S.cf[ lane ] * fabs( fv / 2.0 );
    }

    // phase D (every lane): Polyhedron::volume, 3d/polyhedron.hh:317-324
    VOFX_HD inline double clip_phase_sum ( const ClipScratch &S, const ClipCase &cs )
    {
      double vol = 0.0;
#pragma unroll 1
      for( int f = 0; f < cs.nFaces; ++f )
        vol += S.term[ f ];
      return fabs( vdiv( vol, 3.0 ) );
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

    // upwindPolygon3d( face, v ), 3d/upwindpolygon.hh:24-74: is the shifted copy of the face the first half of the nodes?
    VOFX_HD inline bool upwind_shifted_first ( const double *lo, const double *h, int face, const double *v )
    {
      const int axis = face >> 1;
      double o[ 3 ] = { lo[ 0 ], lo[ 1 ], lo[ 2 ] };
      if( face & 1 )
        o[ axis ] += h[ axis ];
      const int a0 = ( axis == 0 ) ? 1 : 0;
      const int a1 = ( axis == 2 ) ? 1 : 2;
      double c1[ 3 ] = { o[ 0 ], o[ 1 ], o[ 2 ] }, c2[ 3 ] = { o[ 0 ], o[ 1 ], o[ 2 ] };
      c1[ a0 ] += h[ a0 ];
      c2[ a1 ] += h[ a1 ];
      double e1[ 3 ], e2[ 3 ], cr[ 3 ];
      for( int d = 0; d < 3; ++d )
      {
        e1[ d ] = c1[ d ] - o[ d ];
        e2[ d ] = c2[ d ] - o[ d ];
      }
      cross3( e1, e2, cr );
      return dot3( cr, v ) < 0.0;
    }

    // node `lane` of that polytope once the orientation is known
    VOFX_HD inline void upwind_node_oriented ( int lane, const double *lo, const double *h, int face, const double *v, bool shiftedFirst, double *xn )
    {
      const int axis = face >> 1;
      const int a0 = ( axis == 0 ) ? 1 : 0;
      const int a1 = ( axis == 2 ) ? 1 : 2;
      const int i = lane & 3;
      const bool moved = ( lane < 4 ) == shiftedFirst;
      for( int d = 0; d < 3; ++d )
      {
        double cd = lo[ d ];
        if( d == axis && ( face & 1 ) )
          cd += h[ d ];
        if( d == a0 && ( i & 1 ) )
          cd += h[ d ];
        if( d == a1 && ( i & 2 ) )
          cd += h[ d ];
        xn[ d ] = moved ? cd - v[ d ] : cd;
      }
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

// ====================================================================================================================
// G2 for groups of lanes: plane-constant solve of a Cartesian cell with the sweep topology looked up by the order of
// the node heights.
//
// For a hexahedron whose eight rotated node heights are pairwise distinct (by more than the reference's 1e-10 merge
// tolerance) everything PolygonWithDirections does between the brackets (3d/polygonwithdirections.hh:76-290) depends
// only on the ORDER of the heights: which polygon nodes exist, which edge each one slides along, the direction lists
// and the faces the polygon edges lie in.  There are 96 such orders for a box (lowest node x order of the three axis
// increments x whether the third increment exceeds the sum of the other two); the table is generated by running the
// reference's walk (sweep_init / sweep_evolve of vofx_geom3d.cuh) on synthetic heights.  Cells with merged or tied
// heights (axis-aligned normals) are left to the serial walk.
// ====================================================================================================================
namespace vofx
{
  namespace coop
  {

    // unroll factor of the lane-strided loops of the plane-constant solve whose iterations are independent fp64 chains
    // (a lane owns up to 3 edges / 2 faces / 3 terms); 1 keeps the code compact (instruction cache), more overlaps the chains
#ifndef VOFX_PHASE_UNROLL
#define VOFX_PHASE_UNROLL 1
#endif
    constexpr int kPhaseUnroll = VOFX_PHASE_UNROLL;
    constexpr int kSweepCases = 96;
    constexpr int kMaxPoly = 8;       // nodes of the cross-section polygon (6 for a generic box; 8 = the walk's own bound)
    constexpr int kMaxATerms = 12;

    // what the arithmetic of one bracket [ d_k, d_k+1 ] needs to know about the sweep polygon
    struct SweepBracket
    {
      unsigned char nn, nEdges, nA, ok;      // ok = 0: the walk ran into a state the reference does not define
      unsigned char srcNode[ kMaxPoly ];     // parent node id (srcEdge == 0xFF) or index of the old polygon node it slides from
      unsigned char srcEdge[ kMaxPoly ];     // directed edge the node slides along, 0xFF: the node IS the parent node
      unsigned char corrFace[ kMaxPoly ];    // face of the hexahedron that holds polygon edge (i, i+1)
      unsigned char aE1[ kMaxATerms ], aE2[ kMaxATerms ];   // pairs of directed edges of the A terms, in summation order
      unsigned char pad[ 4 ];
    };
    static_assert( sizeof( SweepBracket ) == 56, "SweepBracket layout" );

    struct SweepCase
    {
      unsigned char perm[ 8 ];               // node ids by ascending height
      SweepBracket bracket[ 7 ];
    };
    static_assert( sizeof( SweepCase ) == 400, "SweepCase layout" );

    struct SweepTable
    {
      unsigned char caseOf[ 1024 ];          // sweep_key( ids[0..3] ) -> case or 0xFF
      SweepCase cases[ kSweepCases ];
    };

    // the three lowest nodes and whether the fourth is their common "diagonal" (else it is the third axis neighbour)
    VOFX_INLINE int sweep_key ( int i0, int i1, int i2, int i3 ) { return i0 | ( i1 << 3 ) | ( i2 << 6 ) | ( ( i3 == ( i0 ^ i1 ^ i2 ) ) ? 512 : 0 ); }

    // descriptor of the current polygon of the walk: nodes, faces of its edges, and the A terms in the order of
    // geometry/algorithm.hh:177-194
    VOFX_HD inline void sweep_describe ( const Sweep &s, SweepBracket &B )
    {
      B.ok = 0;
      B.nn = (unsigned char) s.nn;
      B.nEdges = (unsigned char) s.nEdges;
      B.nA = 0;
      if( s.degenerate || s.nn > kMaxPoly || s.nEdges > kMaxPoly || s.nEdges > s.nCF )
        return;
      for( int a = 0; a < s.nn; ++a )
      {
        B.srcNode[ a ] = (unsigned char) s.srcNode[ a ];
        B.srcEdge[ a ] = ( s.srcEdge[ a ] < 0 ) ? (unsigned char) 0xFF : (unsigned char) s.srcEdge[ a ];
      }
      for( int i = 0; i < s.nEdges; ++i )
        B.corrFace[ i ] = (unsigned char) s.corrFace[ i ];
      int nA = 0;
      for( int n = 0; n < s.nn; ++n )
      {
        const int nxt = ( n + 1 == s.nn ) ? 0 : n + 1;
        if( s.dirs[ n ].n == 0 || s.dirs[ nxt ].n == 0 )
          return;                                        // reference: assert( !directions( n ).empty() )
        const int epsn = s.dirs[ n ].n - 1;
        if( nA + 1 + epsn > kMaxATerms )
          return;
        B.aE1[ nA ] = (unsigned char) s.dirs[ n ].e[ epsn ];
        B.aE2[ nA ] = (unsigned char) s.dirs[ nxt ].e[ 0 ];
        nA++;
        for( int l = 0; l < epsn; ++l )
        {
          B.aE1[ nA ] = (unsigned char) s.dirs[ n ].e[ l ];
          B.aE2[ nA ] = (unsigned char) s.dirs[ n ].e[ l + 1 ];
          nA++;
        }
      }
      B.nA = (unsigned char) nA;
      B.ok = 1;
    }

    // returns the number of cases generated (96) or -1 if a case does not fit
    inline int make_sweep_table ( SweepTable &T )
    {
      for( int i = 0; i < 1024; ++i )
        T.caseOf[ i ] = 0xFF;
      int nCases = 0;
      for( int low = 0; low < 8; ++low )
        for( int ax1 = 0; ax1 < 3; ++ax1 )
          for( int ax2 = 0; ax2 < 3; ++ax2 )
          {
            if( ax2 == ax1 )
              continue;
            const int ax3 = 3 - ax1 - ax2;
            for( int third = 0; third < 2; ++third )
            {
              double inc[ 3 ];
              inc[ ax1 ] = 1.0;
              inc[ ax2 ] = 2.0;
              inc[ ax3 ] = third ? 2.5 : 4.0;          // 2.5 < 1 + 2 < 4
              double z[ 8 ];
              for( int i = 0; i < 8; ++i )
              {
                z[ i ] = 0.0;
                for( int d = 0; d < 3; ++d )
                  if( ( ( i ^ low ) >> d ) & 1 )
                    z[ i ] += inc[ d ];
              }
              int ids[ 8 ];
              double ds[ 8 ];
              for( int i = 0; i < 8; ++i )
              {
                const double di = z[ i ];
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
              if( nCases >= kSweepCases )
                return -1;
              SweepCase &C = T.cases[ nCases ];
              for( unsigned b = 0; b < sizeof( SweepCase ); ++b )
                reinterpret_cast< unsigned char * >( &C )[ b ] = 0;
              for( int i = 0; i < 8; ++i )
                C.perm[ i ] = (unsigned char) ids[ i ];

              Sweep s;
              sweep_init< false >( s, nullptr, ids, ds );
              for( int k = 0; k < 7; ++k )
              {
                sweep_describe( s, C.bracket[ k ] );
                if( !C.bracket[ k ].ok || C.bracket[ k ].nn > 6 )
                  return -1;
                if( k < 6 )
                  sweep_evolve< false >( s, nullptr, z, ds[ k ], ds[ k + 1 ], ds[ k + 1 ] - ds[ k ] );
              }
              if( T.caseOf[ sweep_key( ids[ 0 ], ids[ 1 ], ids[ 2 ], ids[ 3 ] ) ] != 0xFF )
                return -1;
              T.caseOf[ sweep_key( ids[ 0 ], ids[ 1 ], ids[ 2 ], ids[ 3 ] ) ] = (unsigned char) nCases;
              nCases++;
            }
          }
      return nCases;
    }

  } // namespace coop
} // namespace vofx

namespace vofx
{
  namespace coop
  {

    // A group of L lanes working on one item.  Device: L consecutive lanes of a warp, sync() = __syncwarp on their mask.
    // Host (tests): the lanes are emulated by a loop, phase by phase.
    struct Group
    {
      int lane;
      unsigned mask;
      VOFX_INLINE void sync () const
      {
#if defined( __CUDA_ARCH__ )
        __syncwarp( mask );
#endif
      }
    };

    // VOFX_LANES( G, L, lane ) { phase }: the phase is run by every lane of the group; a group barrier follows.
    // Code between two phases is "uniform": every lane computes the same values from the scratch.
#if defined( __CUDA_ARCH__ )
#define VOFX_LANES( G, L, lane ) for( int lane = ( G ).lane, vofx_once_ = 1; vofx_once_; vofx_once_ = 0, ( G ).sync() )
#else
#define VOFX_LANES( G, L, lane ) for( int lane = 0; lane < ( L ); ++lane )
#endif

    struct LocateScratch
    {
      union
      {
        struct
        {
          double P[ 3 ][ 8 ];           // rotated nodes
          double dv[ 3 ][ 12 ];         // PolygonWithDirections::directionVector of edge e = ( a, b ); the reverse edge e + 12 is
                                        // derived on read (dv_get): 1192 B per cell let 10 blocks of 16 cells share an SM
          double px[ 2 ][ 8 ], py[ 2 ][ 8 ];   // polygon nodes of the current / next bracket
          double aterm[ kMaxATerms ];   // terms of A of the current bracket
          double bterm[ 8 ], cterm[ 8 ];   // terms of B and C of the current bracket
        };
        // before the plane-constant solve: per stencil slot the weighted difference vector d / |d|^2 and
        // ( colour difference ) / |d|^2 of the least-squares fit (youngs_normal_coop); zeros for a slot without a cell
        double lsq[ 26 ][ 4 ];
      };
      union
      {
        struct { double fnz[ 6 ], fnrm[ 6 ]; };   // outer normal of face f of P: z component, norm of the xy part
        double sums[ 12 ];              // AtA (6 unique entries) and Atb (before the plane-constant solve)
      };
      double ds[ 8 ];                   // sorted node heights
      unsigned char ids[ 8 ];
      unsigned char pad8[ 8 ];
      SweepBracket walked;              // descriptor produced by the walk of lane 0 (cells with tied node heights)
    };

    // direction vector component c of the directed edge e; e >= 12 is the reverse of e - 12: ( P_a - P_b ) / norm with the same
    // norm, i.e. the negated component - except that a zero difference is +0 in either direction ( x - x = +0, +0 / norm = +0 )
    VOFX_INLINE double dv_get ( const LocateScratch &S, int c, int e )
    {
      const bool reverse = e >= 12;
      const double x = S.dv[ c ][ reverse ? e - 12 : e ];
      return ( reverse && x != 0.0 ) ? -x : x;
    }
    // consecutive groups of a warp address the same members: an odd stride in 8-byte words spreads them over the banks
    static_assert( sizeof( LocateScratch ) % 8 == 0 && ( sizeof( LocateScratch ) / 8 ) % 2 == 1, "LocateScratch stride must be an odd number of doubles" );

    // the walk of lane 0 for cells with tied node heights: rare, kept out of line so that the hot path stays compact
    // in the instruction cache (the inlined walk is 11k of the kernel's 20k instructions)
    VOFX_NOINLINE static void walk_begin ( Sweep &sweep, SweepBracket &walked, const unsigned char *ids8, const double *ds )
    {
      int ids[ 8 ];
      for( int k = 0; k < 8; ++k )
        ids[ k ] = ids8[ k ];
      sweep_init< false >( sweep, nullptr, ids, ds );
      sweep_describe( sweep, walked );
    }

    VOFX_NOINLINE static void walk_next ( Sweep &sweep, SweepBracket &walked, const double *z, double dk, double dk1 )
    {
      sweep_evolve< false >( sweep, nullptr, z, dk, dk1, dk1 - dk );
      sweep_describe( sweep, walked );
    }

    // ---- the reference's face formulas on an AXIS-ALIGNED box -----------------------------------------------------------
    // Every vector that occurs in Polyhedron::volume of the cell itself (edge vectors, their cross products, the face
    // normals) has exactly one non-zero component c.  Then | . | = sqrt( c*c ) = |c| exactly (radix 2, round to nearest,
    // no over/underflow: S. Boldo, "Stupid is as stupid does: taking the square root of the square of a floating-point
    // number", 2015) and c / |c| = +-1, (+-0) / s = +-0 with the sign of IEEE division: the same bits as
    // outer_normal3 / face_edge_term without a single sqrt or division.  Checked against the oracle in the CPU tests.
    // Carried through Polyhedron::volume (3d/polyhedron.hh:187-222,317-324) this leaves, for the face f normal to axis k with
    // the other two axes a, b (tables below, derived from the node order of makePolyhedron's cube faces):
    //   outerNormal          = fn e_k, fn = -1 for the lower face and +1 for the upper one
    //   center . outerNormal = fn * ( ( ( ( 0 + c ) + c ) + c ) + c ) * 0.25 ), c = lo_k or hi_k (3 c need not be exact)
    //   edge term q          = sign_q * ( c_q * len ), c_q = the coordinate along a or b that the edge's two nodes share
    //                          ( ( c_q + c_q ) * 0.5 = c_q exactly ), len = | hi - lo | along the edge's own axis
    //   Face::volume         = | ( ( ( ( 0 + t_0 ) + t_1 ) + t_2 ) + t_3 ) / 2 |
    // i.e. 4 products and 8 additions per face instead of 5 cross products, 5 norms and 15 divisions; zero terms can only
    // differ in the sign of a zero, which no partial sum keeps (it starts at +0) and the final | . | removes.
    struct BoxFace
    {
      signed char k, upper, a, b;       // axis of the normal, lower / upper face, axes of the even / odd edge terms' coordinates
      signed char sign[ 4 ], hiBit[ 4 ];
    };
    constexpr BoxFace kBoxFace[ 6 ] = {
      { 0, 0, 1, 2, { -1,  1, 1, -1 }, { 0, 1, 1, 0 } },
      { 0, 1, 2, 1, { -1,  1, 1, -1 }, { 0, 1, 1, 0 } },
      { 1, 0, 2, 0, { -1,  1, 1, -1 }, { 0, 1, 1, 0 } },
      { 1, 1, 2, 0, { -1, -1, 1,  1 }, { 0, 0, 1, 1 } },
      { 2, 0, 0, 1, { -1,  1, 1, -1 }, { 0, 1, 1, 0 } },
      { 2, 1, 1, 0, { -1,  1, 1, -1 }, { 0, 1, 1, 0 } } };

    // center . outerNormal * Face::volume of face F of the box [ lo, hi ]; len = hi - lo
    template< int F >
    VOFX_INLINE double box_face_term ( const double *lo, const double *hi, const double *len )
    {
      constexpr BoxFace T = kBoxFace[ F ];
      const double ck = T.upper ? hi[ T.k ] : lo[ T.k ];
      double center = 0.0;
#pragma unroll
      for( int q = 0; q < 4; ++q )
        center += ck;
      center *= 1.0 / 4;
      double fv = 0.0;
#pragma unroll
      for( int q = 0; q < 4; ++q )
      {
        const int m = ( q & 1 ) ? T.b : T.a, j = ( q & 1 ) ? T.a : T.b;
        const double t = ( T.hiBit[ q ] ? hi[ m ] : lo[ m ] ) * len[ j ];
        fv += ( T.sign[ q ] > 0 ) ? t : -t;
      }
      return ( T.upper ? center : -center ) * fabs( fv / 2.0 );
    }

    // the six terms of Polyhedron::volume of the box [ lo, hi ], in face order
    VOFX_INLINE void box_face_terms ( const double *lo, const double *hi, double *term )
    {
      const double len[ 3 ] = { hi[ 0 ] - lo[ 0 ], hi[ 1 ] - lo[ 1 ], hi[ 2 ] - lo[ 2 ] };
      term[ 0 ] = box_face_term< 0 >( lo, hi, len );
      term[ 1 ] = box_face_term< 1 >( lo, hi, len );
      term[ 2 ] = box_face_term< 2 >( lo, hi, len );
      term[ 3 ] = box_face_term< 3 >( lo, hi, len );
      term[ 4 ] = box_face_term< 4 >( lo, hi, len );
      term[ 5 ] = box_face_term< 5 >( lo, hi, len );
    }

    VOFX_INLINE void cell_node ( const double *lo, const double *hi, int i, double *x )
    {
      x[ 0 ] = ( i & 1 ) ? hi[ 0 ] : lo[ 0 ];
      x[ 1 ] = ( i & 2 ) ? hi[ 1 ] : lo[ 1 ];
      x[ 2 ] = ( i & 4 ) ? hi[ 2 ] : lo[ 2 ];
    }

    // locateHalfSpace( Polyhedron( cube geometry ), innerNormal, fraction ), geometry/algorithm.hh:119-236, for a group of
    // L lanes.  Returns true and the plane constant, or false if the cell must take the serial walk (tied node heights).
    // deferred root search of locate_coop: dist = base + brent( V, 0, h, 1e-14 ), geometry/algorithm.hh:221-227
    struct RootTask
    {
      Cubic V;
      double h, base;
      int pending;
      int cell;
    };

#if defined( __CUDACC__ ) && defined( VOFX_COUNT_WALKS )
    static __device__ unsigned long long g_dbg[ 8 ];
#endif
    template< int L >
    VOFX_HD inline bool locate_coop ( const Group &G, LocateScratch &S, const SweepTable &T, const double *lo, const double *h,
                                      const double *innerNormal, double fraction, double &dist, RootTask *root = nullptr,
                                      int *cost = nullptr )
    {
      double hi[ 3 ], outerNormal[ 3 ];
      for( int k = 0; k < 3; ++k )
      {
        hi[ k ] = lo[ k ] + h[ k ];                      // make_cell: lo + h
        outerNormal[ k ] = innerNormal[ k ] * -1.0;
      }
      fraction = ( fraction > 1.0 ) ? 1.0 : fraction;
      fraction = ( fraction < 0.0 ) ? 0.0 : fraction;

      // rotateToReferenceFrame, 3d/rotation.hh:17-53
      const double *n = outerNormal;
      const bool identity = ( n[ 0 ] == 0.0 && n[ 1 ] == 0.0 && n[ 2 ] == 1.0 );
      double R[ 3 ][ 3 ] = { { 0.0, 0.0, 0.0 }, { 0.0, 0.0, 0.0 }, { 0.0, 0.0, 0.0 } };
      if( n[ 0 ] == 0.0 && n[ 1 ] == 0.0 && n[ 2 ] == -1.0 )
      {
        R[ 0 ][ 0 ] = 1.0;
        R[ 1 ][ 1 ] = -1.0;
        R[ 2 ][ 2 ] = -1.0;
      }
      else if( !identity )
      {
        const double gamma = vdiv( 1 - n[ 2 ], n[ 0 ] * n[ 0 ] + n[ 1 ] * n[ 1 ] );
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

      // phase 1: the rotated nodes
      VOFX_LANES( G, L, lane )
      {
#pragma unroll 1
        for( int i = lane; i < 8; i += L )
        {
          double x[ 3 ];
          cell_node( lo, hi, i, x );
          for( int r = 0; r < 3; ++r )
          {
            double y = 0.0;
            for( int c = 0; c < 3; ++c )
              y += R[ r ][ c ] * x[ c ];
            S.P[ r ][ i ] = identity ? x[ r ] : y;
          }
        }
      }

      // the cell volume (hex_volume) in closed form, by every lane: 24 products and 48 additions
      double faceTerm[ 6 ];
      box_face_terms( lo, hi, faceTerm );
      double vol = 0.0;
#pragma unroll
      for( int f = 0; f < 6; ++f )
        vol += faceTerm[ f ];
      const double givenVolume = fraction * fabs( vdiv( vol, 3.0 ) );

      // phase 2: sort the node heights (std::sort = stable insertion sort on 8 elements: rank by (height, id)),
      // direction vectors of all directed edges, face normals of the rotated cell
      VOFX_LANES( G, L, lane )
      {
#pragma unroll 1
        for( int i = lane; i < 8; i += L )
        {
          const double zi = S.P[ 2 ][ i ];
          int rank = 0;
          for( int j = 0; j < 8; ++j )
          {
            const double zj = S.P[ 2 ][ j ];
            rank += ( zj < zi || ( zj == zi && j < i ) ) ? 1 : 0;
          }
          S.ids[ rank ] = (unsigned char) i;
          S.ds[ rank ] = zi;
        }
#pragma unroll kPhaseUnroll
        for( int e = lane; e < 12; e += L )
        {
          // edge e = ( a, b ); its reverse e + 12 = ( b, a ) is not stored (dv_get)
          const int a = cube::edgeN0( e ), b = cube::edgeN1( e );
          double dv[ 3 ];
          for( int c = 0; c < 3; ++c )
            dv[ c ] = S.P[ c ][ b ] - S.P[ c ][ a ];
          const double norm = norm2( dv[ 0 ], dv[ 1 ] );
          if( norm > 0 )
            vdiv3( dv[ 0 ], dv[ 1 ], dv[ 2 ], norm );
          for( int c = 0; c < 3; ++c )
            S.dv[ c ][ e ] = dv[ c ];
        }
#pragma unroll kPhaseUnroll
        for( int f = lane; f < 6; f += L )
        {
          double x[ 3 ][ 3 ], normal[ 3 ];
          for( int k = 0; k < 3; ++k )
          {
            const int id = cube::faceNode( f, k );
            for( int c = 0; c < 3; ++c )
              x[ k ][ c ] = S.P[ c ][ id ];
          }
          outer_normal3( x[ 0 ], x[ 1 ], x[ 2 ], normal );
          S.fnz[ f ] = normal[ 2 ];
          S.fnrm[ f ] = norm2( normal[ 0 ], normal[ 1 ] );
        }
      }

      // uniform: is this one of the tabulated height orders?
      for( int i = 0; i < 8; ++i )
        if( !( fabs( S.P[ 2 ][ i ] ) < DBL_MAX ) )
          return false;                                  // NaN / inf heights: the ranks below are meaningless
      double ds[ 8 ];
      for( int k = 0; k < 8; ++k )
        ds[ k ] = S.ds[ k ];
      // std::unique with | x - y | < 1e-10 against the last kept element, geometry/algorithm.hh:160-162
      double dU[ 8 ];
      int nU = 0;
      dU[ nU++ ] = ds[ 0 ];
      for( int i = 1; i < 8; ++i )
        if( !( fabs( dU[ nU - 1 ] - ds[ i ] ) < 1e-10 ) )
          dU[ nU++ ] = ds[ i ];

      // tabulated height order?  Otherwise (merged / tied heights) lane 0 runs the reference's walk on the topology
      // and publishes one descriptor per bracket; the arithmetic below is the same either way.
      const SweepCase *C = nullptr;
      if( nU == 8 )
      {
        const int ci = T.caseOf[ sweep_key( S.ids[ 0 ], S.ids[ 1 ], S.ids[ 2 ], S.ids[ 3 ] ) ];
        if( ci != 0xFF )
        {
          C = &T.cases[ ci ];
          for( int k = 0; k < 8; ++k )
            if( C->perm[ k ] != S.ids[ k ] )
              C = nullptr;
        }
      }
      const bool walking = ( C == nullptr );
      // cost class for the scheduling of the next step: 7 the topology is walked on lane 0, else the bracket the solve ends in
      if( cost )
        *cost = walking ? 7 : ( nU >= 2 ? ( nU - 2 < 6 ? nU - 2 : 6 ) : 0 );
#if defined( __CUDA_ARCH__ ) && defined( VOFX_COUNT_WALKS )
      if( G.lane == 0 )
      {
        atomicAdd( &g_dbg[ 0 ], 1ull );
        atomicAdd( &g_dbg[ 1 ], walking ? 1ull : 0ull );
        atomicAdd( &g_dbg[ 2 ], (unsigned long long) nU );
        atomicAdd( &g_dbg[ 3 ], ( fraction >= 1.0 ) ? 1ull : 0ull );
        atomicAdd( &g_dbg[ 4 ], ( fraction <= 0.0 ) ? 1ull : 0ull );
      }
#endif
      Sweep sweep;                                       // state of lane 0's walk (topology only)
      if( walking )
      {
        VOFX_LANES( G, L, lane )
        {
          if( lane == 0 )
          {
            double dsCopy[ 8 ];
            for( int k = 0; k < 8; ++k )
              dsCopy[ k ] = S.ds[ k ];
            walk_begin( sweep, S.walked, S.ids, dsCopy );
          }
        }
      }

#if defined( __CUDA_ARCH__ )
      constexpr int kWords = int( sizeof( SweepBracket ) / sizeof( unsigned ) );
      constexpr int kPre = ( kWords + L - 1 ) / L;
      unsigned pre[ kPre ];
      if( !walking )
      {
        const unsigned *src = reinterpret_cast< const unsigned * >( &C->bracket[ 0 ] );
#pragma unroll
        for( int q = 0; q < kPre; ++q )
          pre[ q ] = ( G.lane + q * L < kWords ) ? src[ G.lane + q * L ] : 0u;
      }
#endif

      double Vd_k = 0.0;
      for( int k = 0; k + 1 < nU; ++k )
      {
        if( k > 0 )
          G.sync();                                      // every lane has finished reading the previous descriptor
        if( walking && k > 0 )
        {
          VOFX_LANES( G, L, lane )
          {
            if( lane == 0 )
              walk_next( sweep, S.walked, S.P[ 2 ], dU[ k - 1 ], dU[ k ] );
          }
        }
        if( !walking )
        {
          // stage the tabulated descriptor (56 B in global memory) in the scratch: the phases below read it many times.
          // Its words were fetched one bracket ahead (the load latency hides behind the previous bracket's arithmetic).
          VOFX_LANES( G, L, lane )
          {
#if defined( __CUDA_ARCH__ )
            unsigned *dst = reinterpret_cast< unsigned * >( &S.walked );
#pragma unroll
            for( int q = 0; q < kPre; ++q )
              if( lane + q * L < kWords )
                dst[ lane + q * L ] = pre[ q ];
#else
            (void) lane;
            S.walked = C->bracket[ k ];
#endif
          }
#if defined( __CUDA_ARCH__ )
          if( k + 1 < 7 )
          {
            const unsigned *src = reinterpret_cast< const unsigned * >( &C->bracket[ k + 1 ] );
#pragma unroll
            for( int q = 0; q < kPre; ++q )
              if( G.lane + q * L < kWords )
                pre[ q ] = src[ G.lane + q * L ];
          }
#endif
        }
        const SweepBracket &B = S.walked;
        if( !B.ok )
          return false;                                  // undefined in the reference: left to the serial kernel (NaN)
        const int cur = k & 1;
        const double hPrev = ( k > 0 ) ? dU[ k ] - dU[ k - 1 ] : 0.0;

        // polygon nodes of this bracket (PolygonWithDirections::evolveToNextPolygon, :177-250)
        VOFX_LANES( G, L, lane )
        {
#pragma unroll 1
          for( int a = lane; a < B.nn; a += L )
          {
            if( B.srcEdge[ a ] == 0xFF )
            {
              S.px[ cur ][ a ] = S.P[ 0 ][ B.srcNode[ a ] ];
              S.py[ cur ][ a ] = S.P[ 1 ][ B.srcNode[ a ] ];
            }
            else
            {
              const int e = B.srcEdge[ a ], from = B.srcNode[ a ];
              const double f = vdiv( hPrev, dv_get( S, 2, e ) );
              const double ax = dv_get( S, 0, e ) * f, ay = dv_get( S, 1, e ) * f;
              S.px[ cur ][ a ] = S.px[ cur ^ 1 ][ from ] + ax;
              S.py[ cur ][ a ] = S.py[ cur ^ 1 ][ from ] + ay;
            }
          }
        }

        // the terms of A, B, C (geometry/algorithm.hh:177-208), one per lane
        VOFX_LANES( G, L, lane )
        {
#pragma unroll kPhaseUnroll
          for( int t = lane; t < B.nA; t += L )
          {
            const int e1 = B.aE1[ t ], e2 = B.aE2[ t ];
            const double v1x = dv_get( S, 0, e1 ), v1y = dv_get( S, 1, e1 ), v1z = dv_get( S, 2, e1 );
            const double v2x = dv_get( S, 0, e2 ), v2y = dv_get( S, 1, e2 ), v2z = dv_get( S, 2, e2 );
            S.aterm[ t ] = vdiv( v1x * v2y - v1y * v2x, v1z * v2z );
          }
#pragma unroll kPhaseUnroll
          for( int i = lane; i < B.nEdges; i += L )
          {
            const int j = ( i + 1 == B.nn ) ? 0 : i + 1;
            const double xi = S.px[ cur ][ i ], yi = S.py[ cur ][ i ], xj = S.px[ cur ][ j ], yj = S.py[ cur ][ j ];
            const double ex = xj - xi, ey = yj - yi;
            const int f = B.corrFace[ i ];
            const double norm = S.fnrm[ f ];
            // | edge | is needed three times: as norm2( ex, ey ) in B and C and as the norm of the edge normal ( ey, -ex ),
            // sqrt( ey*ey + ex*ex ): the same bits, because the two squares are added to zero in either order
            const double len = norm2( ex, ey );
            S.bterm[ i ] = ( norm > 0 ) ? vdiv( len * S.fnz[ f ], norm ) : 0.0;
            double onx = ey, ony = -ex;
            vdiv2( onx, ony, len );
            const double cx = ( xi + xj ) * 0.5, cy = ( yi + yj ) * 0.5;
            S.cterm[ i ] = dot2( cx, cy, onx, ony ) * len;
          }
        }

        double A = 0.0, Bc = 0.0, Cc = 0.0;
#pragma unroll 1
        for( int t = 0; t < B.nA; ++t )
          A -= S.aterm[ t ];
#pragma unroll 1
        for( int i = 0; i < B.nEdges; ++i )
        {
          Bc -= S.bterm[ i ];
          Cc += S.cterm[ i ];
        }
        Cc = fabs( Cc / 2.0 );

        const double hk = dU[ k + 1 ] - dU[ k ];
        Cubic V = { A, Bc, Cc, 0.0 };
        const double Vd_k1 = Vd_k + V( hk );
        if( cost && !walking && ( fabs( Vd_k1 - givenVolume ) < 1e-14 || Vd_k1 > givenVolume ) )
          *cost = k < 6 ? k : 6;
        if( fabs( Vd_k1 - givenVolume ) < 1e-14 )
        {
#if defined( __CUDA_ARCH__ ) && defined( VOFX_COUNT_WALKS )
          if( G.lane == 0 )
          {
            atomicAdd( &g_dbg[ 5 ], 1ull );
            if( k + 1 == nU - 1 )
              atomicAdd( &g_dbg[ 6 ], 1ull );
            if( k == 0 )
              atomicAdd( &g_dbg[ 7 ], 1ull );
          }
#endif
          dist = dU[ k + 1 ];
          return true;
        }
        else if( Vd_k1 > givenVolume )
        {
          V.target = givenVolume - Vd_k;
          if( root )
          {
            // the root of V on [0, hk] is found by a one-thread-per-cell kernel (Brent is serial: here it would run
            // redundantly in every lane of the group)
            root->V = V;
            root->h = hk;
            root->base = dU[ k ];
            root->pending = 1;
            return true;
          }
          dist = dU[ k ] + brent( V, 0.0, hk, 1e-14 );
          return true;
        }
        Vd_k = Vd_k1;
      }
      dist = dU[ nU - 1 ];
      return true;
    }

  } // namespace coop
} // namespace vofx

#include "vofx_grid.cuh"

namespace vofx
{
  namespace coop
  {

    // FieldMatrix< double, 3, 3 >::solve of dune-common 2.6 (closed form, the same expression tree as solve_small< 3 >)
    VOFX_HD inline void solve3 ( const double M[ 3 ][ 3 ], const double *b, double *x )
    {
      const double t4 = M[ 0 ][ 0 ] * M[ 1 ][ 1 ];
      const double t6 = M[ 0 ][ 0 ] * M[ 1 ][ 2 ];
      const double t8 = M[ 0 ][ 1 ] * M[ 1 ][ 0 ];
      const double t10 = M[ 0 ][ 2 ] * M[ 1 ][ 0 ];
      const double t12 = M[ 0 ][ 1 ] * M[ 2 ][ 0 ];
      const double t14 = M[ 0 ][ 2 ] * M[ 2 ][ 0 ];
      const double d = ( t4 * M[ 2 ][ 2 ] - t6 * M[ 2 ][ 1 ] - t8 * M[ 2 ][ 2 ] + t10 * M[ 2 ][ 1 ] + t12 * M[ 1 ][ 2 ] - t14 * M[ 1 ][ 1 ] );
      x[ 0 ] = b[ 0 ] * M[ 1 ][ 1 ] * M[ 2 ][ 2 ] - b[ 0 ] * M[ 2 ][ 1 ] * M[ 1 ][ 2 ]
               - b[ 1 ] * M[ 0 ][ 1 ] * M[ 2 ][ 2 ] + b[ 1 ] * M[ 2 ][ 1 ] * M[ 0 ][ 2 ]
               + b[ 2 ] * M[ 0 ][ 1 ] * M[ 1 ][ 2 ] - b[ 2 ] * M[ 1 ][ 1 ] * M[ 0 ][ 2 ];
      x[ 1 ] = M[ 0 ][ 0 ] * b[ 1 ] * M[ 2 ][ 2 ] - M[ 0 ][ 0 ] * b[ 2 ] * M[ 1 ][ 2 ]
               - M[ 1 ][ 0 ] * b[ 0 ] * M[ 2 ][ 2 ] + M[ 1 ][ 0 ] * b[ 2 ] * M[ 0 ][ 2 ]
               + M[ 2 ][ 0 ] * b[ 0 ] * M[ 1 ][ 2 ] - M[ 2 ][ 0 ] * b[ 1 ] * M[ 0 ][ 2 ];
      x[ 2 ] = M[ 0 ][ 0 ] * M[ 1 ][ 1 ] * b[ 2 ] - M[ 0 ][ 0 ] * M[ 2 ][ 1 ] * b[ 1 ]
               - M[ 1 ][ 0 ] * M[ 0 ][ 1 ] * b[ 2 ] + M[ 1 ][ 0 ] * M[ 2 ][ 1 ] * b[ 0 ]
               + M[ 2 ][ 0 ] * M[ 0 ][ 1 ] * b[ 1 ] - M[ 2 ][ 0 ] * M[ 1 ][ 1 ] * b[ 0 ];
      vdiv3( x[ 0 ], x[ 1 ], x[ 2 ], d );
    }

    // ModifiedYoungsReconstruction::applyLocal up to the normal, reconstruction/modifiedyoungs.hh:90-131, for a group of
    // L lanes: slot s < 26 is the s-th vertex neighbour in ascending cell index, slot 26 + f the ghost of wall face f
    // (:112-124); a lane evaluates the contribution of one slot, the sums run over the slots in order.
    // Returns false if the normal degenerates (:128-129).
    template< int L >
    VOFX_HD inline bool youngs_normal_coop ( const Group &G, LocateScratch &S, const Grid &g, const double *u, const int *ijk,
                                             double colorEn, double *normal )
    {
      constexpr int kOwn = ( 9 + L - 1 ) / L;
      double center[ 3 ], lower[ 3 ];
      for( int d = 0; d < 3; ++d )
      {
        lower[ d ] = cell_lower( g, d, ijk[ d ] );
        center[ d ] = lower[ d ] + 0.5 * g.h[ d ];
      }
      // slots 26..31 are the ghosts of wall faces: a cell away from the domain boundary has none (:112-124)
      bool atWall = false;
      for( int d = 0; d < 3; ++d )
      {
        const int gi = ijk[ d ] + g.g0[ d ];
        atWall = atWall || gi == 0 || gi == g.ng[ d ] - 1;
      }

      // phase A: a lane evaluates the slots lane, lane + L, ...: ls_accumulate up to the scaled vector, i.e.
      // weight = 1 / |d|^2, e = d * weight, wc = weight * colorDiff.  A slot without a cell contributes zeros: every sum below
      // starts at +0 and therefore never holds -0, so adding a zero of either sign leaves it unchanged.
      VOFX_LANES( G, L, lane )
      {
#pragma unroll 1
        for( int s = lane; s < 26; s += L )
        {
          const int m = ( s < 13 ) ? s : s + 1;
          const int nb[ 3 ] = { ijk[ 0 ] + ( m % 3 ) - 1, ijk[ 1 ] + ( ( m / 3 ) % 3 ) - 1, ijk[ 2 ] + ( m / 9 ) - 1 };
          double e[ 3 ] = { 0.0, 0.0, 0.0 }, wc = 0.0;
          if( cell_exists( g, nb ) )
          {
            const double colorNb = clamp01( u[ cell_index( g, nb ) ] );
            double n2 = 0.0;
            for( int k = 0; k < 3; ++k )
            {
              e[ k ] = ( cell_lower( g, k, nb[ k ] ) + 0.5 * g.h[ k ] ) - center[ k ];
              n2 += e[ k ] * e[ k ];
            }
            const double weight = vdiv( 1.0, n2 );
            for( int k = 0; k < 3; ++k )
              e[ k ] *= weight;
            wc = weight * ( colorNb - colorEn );
          }
          S.lsq[ s ][ 0 ] = e[ 0 ];
          S.lsq[ s ][ 1 ] = e[ 1 ];
          S.lsq[ s ][ 2 ] = e[ 2 ];
          S.lsq[ s ][ 3 ] = wc;
        }
      }
      // phase B: a lane owns the sums c = lane, lane + L, ... of ( AtA_00, AtA_01, AtA_02, AtA_11, AtA_12, AtA_22, Atb_0, Atb_1,
      // Atb_2 ) and runs over the slots in order: AtA_ij += e_i e_j (outerProduct adds the product to a zero matrix first: the
      // same value, see above), Atb_k += wc e_k.  The operands of sum c are entries ( ia, ib ) of the slot's record.
      VOFX_LANES( G, L, lane )
      {
        double acc[ kOwn ];
        int ia[ kOwn ], ib[ kOwn ];
#pragma unroll
        for( int q = 0; q < kOwn; ++q )
        {
          const int c = lane + q * L;
          acc[ q ] = 0.0;
          // c:  0 1 2 3 4 5 6 7 8  ->  ia: 0 0 0 1 1 2 3 3 3,  ib: 0 1 2 1 2 2 0 1 2
          ia[ q ] = ( c < 3 ) ? 0 : ( ( c < 5 ) ? 1 : ( ( c < 6 ) ? 2 : 3 ) );
          ib[ q ] = ( c < 3 ) ? c : ( ( c < 5 ) ? c - 2 : ( ( c < 6 ) ? 2 : c - 6 ) );
          if( c >= 9 )
            ia[ q ] = ib[ q ] = 0;
        }
#pragma unroll 2
        for( int s = 0; s < 26; ++s )
        {
#pragma unroll
          for( int q = 0; q < kOwn; ++q )
            acc[ q ] += S.lsq[ s ][ ia[ q ] ] * S.lsq[ s ][ ib[ q ] ];
        }
#pragma unroll
        for( int q = 0; q < kOwn; ++q )
          if( lane + q * L < 9 )
            S.sums[ lane + q * L ] = acc[ q ];
      }
      if( atWall )
      {
        // the ghosts of the wall faces follow in face order (:112-124): d = 2 ( face centre - cell centre ), colour 0.5
        VOFX_LANES( G, L, lane )
        {
#pragma unroll 1
          for( int face = lane; face < 6; face += L )
          {
            int nb[ 3 ] = { ijk[ 0 ], ijk[ 1 ], ijk[ 2 ] };
            nb[ face >> 1 ] += ( face & 1 ) ? 1 : -1;
            double e[ 3 ] = { 0.0, 0.0, 0.0 }, wc = 0.0;
            if( !in_domain( g, nb ) )
            {
              double n2 = 0.0;
              for( int k = 0; k < 3; ++k )
              {
                double fc = lower[ k ];
                if( k == ( face >> 1 ) )
                {
                  if( face & 1 )
                    fc += g.h[ k ];
                }
                else
                  fc += 0.5 * g.h[ k ];
                e[ k ] = fc - center[ k ];
                e[ k ] *= 2.0;
                n2 += e[ k ] * e[ k ];
              }
              const double weight = vdiv( 1.0, n2 );
              for( int k = 0; k < 3; ++k )
                e[ k ] *= weight;
              wc = weight * ( 0.5 - colorEn );
            }
            S.lsq[ face ][ 0 ] = e[ 0 ];
            S.lsq[ face ][ 1 ] = e[ 1 ];
            S.lsq[ face ][ 2 ] = e[ 2 ];
            S.lsq[ face ][ 3 ] = wc;
          }
        }
        VOFX_LANES( G, L, lane )
        {
#pragma unroll
          for( int q = 0; q < kOwn; ++q )
          {
            const int c = lane + q * L;
            if( c < 9 )
            {
              const int a = ( c < 3 ) ? 0 : ( ( c < 5 ) ? 1 : ( ( c < 6 ) ? 2 : 3 ) );
              const int b = ( c < 3 ) ? c : ( ( c < 5 ) ? c - 2 : ( ( c < 6 ) ? 2 : c - 6 ) );
              double acc = S.sums[ c ];
              for( int face = 0; face < 6; ++face )
                acc += S.lsq[ face ][ a ] * S.lsq[ face ][ b ];
              S.sums[ c ] = acc;
            }
          }
        }
      }

      const double AtA[ 3 ][ 3 ] = { { S.sums[ 0 ], S.sums[ 1 ], S.sums[ 2 ] }, { S.sums[ 1 ], S.sums[ 3 ], S.sums[ 4 ] }, { S.sums[ 2 ], S.sums[ 4 ], S.sums[ 5 ] } };
      const double Atb[ 3 ] = { S.sums[ 6 ], S.sums[ 7 ], S.sums[ 8 ] };
      normal[ 0 ] = normal[ 1 ] = normal[ 2 ] = 0.0;
      solve3( AtA, Atb, normal );
      double n2 = 0.0;
      for( int k = 0; k < 3; ++k )
        n2 += normal[ k ] * normal[ k ];
      if( n2 < kEps )
        return false;
      const double nrm = vsqrt( n2 );
      vdiv3( normal[ 0 ], normal[ 1 ], normal[ 2 ], nrm );
      return true;
    }

  } // namespace coop
} // namespace vofx

// ====================================================================================================================
// R3 in 3D for a group of L lanes: ModifiedSwartzReconstruction::applyLocal, reconstruction/modifiedswartz.hh:167-243.
// Per iteration the cell and each of its mixed vertex neighbours is located with the cell's OLD normal
// (locate_coop: brackets by all lanes), then the root searches, the interface polygons and their centroids
// (3d/intersect.hh:274-355, 3d/face.hh:69-90) are taken by different lanes for different neighbours, and the
// n(n-1) cross-product terms of the new normal are evaluated row by row and summed in the reference's order.
// ====================================================================================================================
namespace vofx
{
  namespace coop
  {

    constexpr int kMaxSwartzNb = 26;

    template< int L >
    struct SwartzScratch
    {
      union
      {
        LocateScratch loc;                 // while cells are being located
        double row[ kMaxSwartzNb ][ 3 ];   // afterwards: normalised cross products of one row a
      };
      RootTask root[ L ];                  // one batch of L located cells
      double dist[ L ];
      double cen[ kMaxSwartzNb + 1 ][ 3 ]; // interface centroids: entry 0 the cell itself, 1.. its neighbours
      double nsum[ 3 ];
      double bcast;
      unsigned char state[ 8 ];            // per batch entry: 0 plane constant known, 1 root pending, 2 serial walk needed
      unsigned char nbOff[ kMaxSwartzNb ]; // packed offsets ( dx+1 ) | ( dy+1 ) << 2 | ( dz+1 ) << 4 of the mixed neighbours
      unsigned char nbSlot[ 32 ];          // per stencil slot: 1 if it is a mixed neighbour
      unsigned char rowValid[ kMaxSwartzNb ];
      unsigned char pad[ 4 ];
    };

    // interface( cell, ( normal, dist ) ).centroid(): serial, one lane per located cell
    VOFX_NOINLINE static void interface_centroid_serial ( const double *lo, const double *h, const double *normal, double dist, double *centroid )
    {
      Hex cell;
      make_cell( lo, h, cell );
      const Hs3 H = { { normal[ 0 ], normal[ 1 ], normal[ 2 ] }, dist };
      Face3 f;
      plane_cut( cell, H, f );
      face_centroid( f, centroid );
    }

    VOFX_NOINLINE static double locate_serial ( const double *lo, const double *h, const double *normal, double fraction )
    {
      Hex cell;
      make_cell( lo, h, cell );
      return locate( cell, normal, fraction );
    }

    // out-of-line instances: the Swartz routine calls these from several places, and its instruction-cache footprint
    // decides its speed (inlined everywhere the kernel was 96 k instructions and stalled 28 of 33 cycles on fetch)
    template< int L >
    VOFX_NOINLINE static bool locate_coop_call ( const Group &G, LocateScratch &S, const SweepTable &T, const double *lo, const double *h,
                                                 const double *innerNormal, double fraction, double *dist, RootTask *root )
    {
      return locate_coop< L >( G, S, T, lo, h, innerNormal, fraction, *dist, root );
    }

    VOFX_NOINLINE static double root_of ( const RootTask &r ) { return r.base + brent( r.V, 0.0, r.h, 1e-14 ); }

    // rec: in = the Young's reconstruction of the cell, out = the Swartz one (4 doubles)
    template< int L >
    VOFX_HD inline void swartz_cell_coop ( const Group &G, SwartzScratch< L > &S, const SweepTable &T, const Grid &g, const double *u,
                                           const unsigned char *flags, const int *ijk, int maxIterations, double *rec )
    {
      const double colorEn = u[ cell_index( g, ijk ) ];

      // mixed vertex neighbours in ascending cell index (stencil slot order)
      VOFX_LANES( G, L, lane )
      {
        for( int s = lane; s < 26; s += L )
        {
          const int m = ( s < 13 ) ? s : s + 1;
          const int nb[ 3 ] = { ijk[ 0 ] + ( m % 3 ) - 1, ijk[ 1 ] + ( ( m / 3 ) % 3 ) - 1, ijk[ 2 ] + ( m / 9 ) - 1 };
          S.nbSlot[ s ] = ( cell_exists( g, nb ) && is_mixed( flags[ cell_index( g, nb ) ] ) ) ? 1 : 0;
        }
      }
      int nNb = 0;
      VOFX_LANES( G, L, lane )
      {
        if( lane == 0 )
        {
          int n = 0;
          for( int s = 0; s < 26; ++s )
            if( S.nbSlot[ s ] )
            {
              const int m = ( s < 13 ) ? s : s + 1;
              S.nbOff[ n++ ] = (unsigned char) ( ( m % 3 ) | ( ( ( m / 3 ) % 3 ) << 2 ) | ( ( m / 9 ) << 4 ) );
            }
          S.bcast = (double) n;
        }
      }
      nNb = (int) S.bcast;

      double normal[ 3 ] = { rec[ 0 ], rec[ 1 ], rec[ 2 ] };
      int iterations = 0;
      double residuum = 0.0;
      do
      {
        const double oldNormal[ 3 ] = { normal[ 0 ], normal[ 1 ], normal[ 2 ] };

        // interface centroids of the cell (entry 0) and of its mixed neighbours, all with the old normal, in batches of L
        for( int first = 0; first < nNb + 1; first += L )
        {
          const int batch = ( nNb + 1 - first < L ) ? nNb + 1 - first : L;
          for( int j = 0; j < batch; ++j )
          {
            const int a = first + j;
            int cell[ 3 ] = { ijk[ 0 ], ijk[ 1 ], ijk[ 2 ] };
            if( a > 0 )
            {
              const int off = S.nbOff[ a - 1 ];
              cell[ 0 ] += ( off & 3 ) - 1;
              cell[ 1 ] += ( ( off >> 2 ) & 3 ) - 1;
              cell[ 2 ] += ( ( off >> 4 ) & 3 ) - 1;
            }
            const double lo[ 3 ] = { cell_lower( g, 0, cell[ 0 ] ), cell_lower( g, 1, cell[ 1 ] ), cell_lower( g, 2, cell[ 2 ] ) };
            const double color = ( a == 0 ) ? colorEn : u[ cell_index( g, cell ) ];
            RootTask root;
            root.pending = 0;
            double dist = 0.0;
            const bool ok = locate_coop_call< L >( G, S.loc, T, lo, g.h, oldNormal, color, &dist, &root );
            VOFX_LANES( G, L, lane )
            {
              if( lane == 0 )
              {
                S.state[ j ] = (unsigned char) ( !ok ? 2 : ( root.pending ? 1 : 0 ) );
                S.dist[ j ] = dist;
                S.root[ j ] = root;
              }
            }
          }
          // lane j finishes located cell j: root search (or serial walk), interface polygon, centroid
          VOFX_LANES( G, L, lane )
          {
            if( lane < batch )
            {
              const int a = first + lane;
              int cell[ 3 ] = { ijk[ 0 ], ijk[ 1 ], ijk[ 2 ] };
              if( a > 0 )
              {
                const int off = S.nbOff[ a - 1 ];
                cell[ 0 ] += ( off & 3 ) - 1;
                cell[ 1 ] += ( ( off >> 2 ) & 3 ) - 1;
                cell[ 2 ] += ( ( off >> 4 ) & 3 ) - 1;
              }
              const double lo[ 3 ] = { cell_lower( g, 0, cell[ 0 ] ), cell_lower( g, 1, cell[ 1 ] ), cell_lower( g, 2, cell[ 2 ] ) };
              double dist = S.dist[ lane ];
              if( S.state[ lane ] == 1 )
                dist = root_of( S.root[ lane ] );
              else if( S.state[ lane ] == 2 )
                dist = locate_serial( lo, g.h, oldNormal, ( a == 0 ) ? colorEn : u[ cell_index( g, cell ) ] );
              double c[ 3 ];
              interface_centroid_serial( lo, g.h, oldNormal, dist, c );
              for( int k = 0; k < 3; ++k )
                S.cen[ a ][ k ] = c[ k ];
            }
          }
        }

        // normal = sum over ordered pairs ( a, b ), a != b, of the normalised cross products, modifiedswartz.hh:196-229
        VOFX_LANES( G, L, lane )
        {
          if( lane < 3 )
            S.nsum[ lane ] = 0.0;
        }
        for( int a = 0; a < nNb; ++a )
        {
          VOFX_LANES( G, L, lane )
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
                if( !( norm3( cn ) < 1e-12 ) )
                {
                  valid = true;
                  if( dot3( cn, oldNormal ) < 0 )
                    for( int k = 0; k < 3; ++k )
                      cn[ k ] *= -1.0;
                  const double nrm = norm3( cn );
                  for( int k = 0; k < 3; ++k )
                    cn[ k ] = vdiv( cn[ k ], nrm );
                }
              }
              S.rowValid[ b ] = valid ? 1 : 0;
              for( int k = 0; k < 3; ++k )
                S.row[ b ][ k ] = cn[ k ];
            }
          }
          VOFX_LANES( G, L, lane )
          {
            if( lane < 3 )
            {
              double acc = S.nsum[ lane ];
              for( int b = 0; b < nNb; ++b )
                if( S.rowValid[ b ] )
                  acc += 1.0 * S.row[ b ][ lane ];
              S.nsum[ lane ] = acc;
            }
          }
        }
        normal[ 0 ] = S.nsum[ 0 ];
        normal[ 1 ] = S.nsum[ 1 ];
        normal[ 2 ] = S.nsum[ 2 ];

        double n2 = 0.0;
        for( int k = 0; k < 3; ++k )
          n2 += normal[ k ] * normal[ k ];
        const double nrm = sqrt( n2 );
        if( nrm < 0.5 )
          break;   // "return": keep the last reconstruction, modifiedswartz.hh:231-232
        for( int k = 0; k < 3; ++k )
          normal[ k ] = vdiv( normal[ k ], nrm );

        const double lo[ 3 ] = { cell_lower( g, 0, ijk[ 0 ] ), cell_lower( g, 1, ijk[ 1 ] ), cell_lower( g, 2, ijk[ 2 ] ) };
        double dist = 0.0;
        RootTask finalRoot;
        finalRoot.pending = 0;
        const bool ok = locate_coop_call< L >( G, S.loc, T, lo, g.h, normal, colorEn, &dist, &finalRoot );
        if( ok && finalRoot.pending )
          dist = root_of( finalRoot );
        if( !ok )
        {
          VOFX_LANES( G, L, lane )
          {
            if( lane == 0 )
              S.bcast = locate_serial( lo, g.h, normal, colorEn );
          }
          dist = S.bcast;
        }
        G.sync();     // the locate scratch is reused by the next iteration's first locate
        for( int k = 0; k < 3; ++k )
          rec[ k ] = normal[ k ];
        rec[ 3 ] = dist;

        double dotp = 0.0;
        for( int k = 0; k < 3; ++k )
          dotp += normal[ k ] * oldNormal[ k ];
        residuum = 1.0 - dotp;
        ++iterations;
      }
      while( residuum > 1e-12 && iterations < maxIterations );
    }

  } // namespace coop
} // namespace vofx
