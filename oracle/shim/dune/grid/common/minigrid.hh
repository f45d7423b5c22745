// TEST INFRASTRUCTURE ONLY (oracle/): a serial, axis-aligned Cartesian "GridView" model providing exactly the
// slice of the dune-grid / dune-spgrid interface that the reference's hot-path operators touch
// (list in SURVEY.md section 8(c)).  Nothing here is reference code; dune-grid and dune-spgrid are third-party
// dependencies that the reference does not vendor (dune.module:4-5).
//
// Geometry convention (this is the convention the C ABI in include/vofx.h defines as well, so that the
// oracle and the device generate bit-identical coordinates):
//   cell lower corner  x_lo[d] = origin[d] + i[d] * h[d]       (one multiply, one add)
//   cell upper corner  x_lo[d] + h[d]
//   cell centre        x_lo[d] + 0.5 * h[d]
//   cell index         x-fastest lexicographic,  faces in the order x-,x+,y-,y+,z-,z+.
#pragma once
#include <algorithm>
#include <array>
#include <cstddef>
#include <limits>
#include <vector>

#include <dune/common/fmatrix.hh>
#include <dune/common/fvector.hh>
#include <dune/common/power.hh>

namespace Dune
{

  enum InterfaceType { InteriorBorder_InteriorBorder_Interface, InteriorBorder_All_Interface, Overlap_OverlapFront_Interface, Overlap_All_Interface, All_All_Interface };
  enum CommunicationDirection { ForwardCommunication, BackwardCommunication };
  enum PartitionIteratorType { Interior_Partition, InteriorBorder_Partition, Overlap_Partition, OverlapFront_Partition, All_Partition, Ghost_Partition };

  template< class Impl, class T >
  struct CommDataHandleIF { using DataType = T; };

  namespace Partitions
  {
    struct All {};
    struct InteriorBorder {};
    struct Interior {};
    constexpr All all {};
    constexpr InteriorBorder interiorBorder {};
    constexpr Interior interior {};
  }

  struct GeometryType
  {
    int dim;
    bool simplex = false;   // only the unstructured model (meshgrid.hh) has simplices
    bool isCube () const { return !simplex; }
    bool isSimplex () const { return simplex; }
    bool isTriangle () const { return dim == 2 && simplex; }
    bool isQuadrilateral () const { return dim == 2 && !simplex; }
    bool isLine () const { return dim == 1; }
  };

  template< int dim > struct MiniGrid;
  template< int dim > struct MiniGridView;

  // axis-aligned box spanned by mydim axes
  template< int dim, int mydim >
  struct BoxGeometry
  {
    using ctype = double;
    using GlobalCoordinate = FieldVector< double, dim >;
    using LocalCoordinate = FieldVector< double, mydim >;
    static constexpr int mydimension = mydim, coorddimension = dim;

    GlobalCoordinate o;               // lower corner
    FieldVector< double, dim > h;     // mesh widths
    std::array< int, ( mydim > 0 ? mydim : 1 ) > axes {};   // spanned axes, ascending

    GeometryType type () const { return { mydim }; }
    int corners () const { return 1 << mydim; }
    GlobalCoordinate corner ( int i ) const
    {
      GlobalCoordinate c = o;
      for( int j = 0; j < mydim; ++j )
        if( ( i >> j ) & 1 )
          c[ axes[ j ] ] += h[ axes[ j ] ];
      return c;
    }
    GlobalCoordinate center () const
    {
      GlobalCoordinate c = o;
      for( int j = 0; j < mydim; ++j )
        c[ axes[ j ] ] += 0.5 * h[ axes[ j ] ];
      return c;
    }
    GlobalCoordinate global ( const LocalCoordinate &x ) const
    {
      GlobalCoordinate c = o;
      for( int j = 0; j < mydim; ++j )
        c[ axes[ j ] ] += x[ j ] * h[ axes[ j ] ];
      return c;
    }
    double volume () const
    {
      double v = 1;
      for( int j = 0; j < mydim; ++j )
        v *= h[ axes[ j ] ];
      return v;
    }
    double integrationElement ( const LocalCoordinate & ) const { return volume(); }
    FieldMatrix< double, mydim, dim > jacobianTransposed ( const LocalCoordinate & ) const
    {
      FieldMatrix< double, mydim, dim > J( 0.0 );
      for( int j = 0; j < mydim; ++j )
        J[ j ][ axes[ j ] ] = h[ axes[ j ] ];
      return J;
    }
  };

  template< int dim >
  struct MiniEntity
  {
    static constexpr int codimension = 0, dimension = dim;
    using Geometry = BoxGeometry< dim, dim >;

    const MiniGrid< dim > *g = nullptr;
    std::array< int, dim > ijk {};

    Geometry geometry () const;
    bool operator== ( const MiniEntity &o ) const { return ijk == o.ijk; }
    bool operator!= ( const MiniEntity &o ) const { return !( *this == o ); }
  };

  template< int dim >
  struct MiniIntersection
  {
    using Entity = MiniEntity< dim >;
    using Geometry = BoxGeometry< dim, dim - 1 >;
    using LocalCoordinate = FieldVector< double, dim - 1 >;
    using GlobalCoordinate = FieldVector< double, dim >;

    Entity in;
    int face;   // 2*axis + side

    bool neighbor () const;
    bool boundary () const { return !neighbor(); }
    Entity inside () const { return in; }
    Entity outside () const { Entity e = in; e.ijk[ face / 2 ] += ( face % 2 ? 1 : -1 ); return e; }
    Geometry geometry () const;
    GeometryType type () const { return { dim - 1 }; }
    GlobalCoordinate centerUnitOuterNormal () const
    {
      GlobalCoordinate n( 0.0 );
      n[ face / 2 ] = ( face % 2 ? 1.0 : -1.0 );
      return n;
    }
    GlobalCoordinate integrationOuterNormal ( const LocalCoordinate & ) const
    {
      GlobalCoordinate n = centerUnitOuterNormal();
      n *= geometry().volume();
      return n;
    }
  };

  template< int dim >
  struct MiniIndexSet
  {
    using IndexType = std::size_t;
    const MiniGrid< dim > *g;
    std::size_t size ( int codim ) const;
    std::size_t index ( const MiniEntity< dim > &e ) const;
    std::size_t subIndex ( const MiniEntity< dim > &e, int k, int codim ) const;
  };

  struct MiniComm
  {
    template< class T > T min ( T x ) const { return x; }
    template< class T > T max ( T x ) const { return x; }
    template< class T > T sum ( T x ) const { return x; }
    int rank () const { return 0; }
    int size () const { return 1; }
    void barrier () const {}
    template< class T > void allgather ( const T *in, int count, T *out ) const { std::copy( in, in + count, out ); }
  };

  // --- the SPGrid internals that stencil/heightfunctionstencil.hh:28-30,55-66 reaches into ------------------
  template< int dim >
  struct MiniMultiIndex
  {
    std::array< int, dim > a {};   // zero-initialised; heightfunctionstencil.hh:73,89 relies on that
    int &operator[] ( int i ) { return a[ i ]; }
    const int &operator[] ( int i ) const { return a[ i ]; }
    MiniMultiIndex &operator*= ( int s ) { for( auto &x : a ) x *= s; return *this; }
    MiniMultiIndex operator+ ( const MiniMultiIndex &o ) const
    {
      MiniMultiIndex r = *this;
      for( int i = 0; i < dim; ++i )
        r.a[ i ] += o.a[ i ];
      return r;
    }
  };

  template< int dim >
  struct MiniPartitionList
  {
    const MiniGrid< dim > *g;
    bool contains ( const MiniMultiIndex< dim > &id, unsigned ) const;
  };

  template< int dim >
  struct MiniGridLevel
  {
    const MiniGrid< dim > *g;
    template< PartitionIteratorType > MiniPartitionList< dim > partition () const { return { g }; }
  };

  template< int dim >
  struct MiniEntityInfo
  {
    using MultiIndex = MiniMultiIndex< dim >;
    const MiniGrid< dim > *g;
    MultiIndex id_;
    MultiIndex &id () { return id_; }
    const MultiIndex &id () const { return id_; }
    void update () {}
    MiniGridLevel< dim > gridLevel () const { return { g }; }
    unsigned partitionNumber () const { return 0; }
  };

  template< int dim >
  struct MiniEntityImpl
  {
    MiniEntity< dim > e;
    MiniEntityInfo< dim > entityInfo () const
    {
      MiniEntityInfo< dim > i { e.g, {} };
      for( int d = 0; d < dim; ++d )
        i.id_[ d ] = 2 * e.ijk[ d ] + 1;   // SPGrid ids of codim-0 entities are odd
      return i;
    }
  };

  template< int codim, int dim, class Grid >
  struct SPEntity
  {
    using EntityInfo = MiniEntityInfo< dim >;
    EntityInfo info;
    SPEntity ( const EntityInfo &i ) : info( i ) {}
    operator MiniEntity< dim > () const
    {
      MiniEntity< dim > e;
      e.g = info.g;
      for( int d = 0; d < dim; ++d )
        e.ijk[ d ] = ( info.id_[ d ] - 1 ) / 2;
      return e;
    }
  };

  template< int dim >
  struct MiniGrid
  {
    using ctype = double;
    static constexpr int dimension = dim;
    using LeafGridView = MiniGridView< dim >;

    std::array< int, dim > n;
    FieldVector< double, dim > lo, h;

    static MiniEntityImpl< dim > getRealImplementation ( const MiniEntity< dim > &e ) { return { e }; }
    std::size_t cells () const
    {
      std::size_t c = 1;
      for( int d = 0; d < dim; ++d )
        c *= n[ d ];
      return c;
    }
  };

  template< int dim >
  struct MiniElemIt
  {
    using Entity = MiniEntity< dim >;
    const MiniGrid< dim > *g;
    std::size_t i;
    Entity operator* () const
    {
      Entity e;
      e.g = g;
      std::size_t k = i;
      for( int d = 0; d < dim; ++d )
      {
        e.ijk[ d ] = int( k % g->n[ d ] );
        k /= g->n[ d ];
      }
      return e;
    }
    MiniElemIt &operator++ () { ++i; return *this; }
    bool operator!= ( const MiniElemIt &o ) const { return i != o.i; }
  };

  template< int dim >
  struct MiniElemRange
  {
    const MiniGrid< dim > *g;
    MiniElemIt< dim > begin () const { return { g, 0 }; }
    MiniElemIt< dim > end () const { return { g, g->cells() }; }
  };

  template< int dim >
  struct MiniIsIt
  {
    MiniEntity< dim > e;
    int f;
    MiniIntersection< dim > operator* () const { return { e, f }; }
    MiniIsIt &operator++ () { ++f; return *this; }
    bool operator!= ( const MiniIsIt &o ) const { return f != o.f; }
  };

  template< int dim >
  struct MiniIsRange
  {
    MiniEntity< dim > e;
    MiniIsIt< dim > begin () const { return { e, 0 }; }
    MiniIsIt< dim > end () const { return { e, 2 * dim }; }
  };

  template< int dim >
  struct MiniGridView
  {
    using Grid = MiniGrid< dim >;
    using ctype = double;
    static constexpr int dimension = dim, dimensionworld = dim;
    using IndexSet = MiniIndexSet< dim >;
    using Intersection = MiniIntersection< dim >;
    using CollectiveCommunication = MiniComm;
    template< int cd > struct Codim { using Entity = MiniEntity< dim >; using Geometry = BoxGeometry< dim, dim >; };

    const Grid *g;
    IndexSet is;
    MiniComm c;

    MiniGridView ( const Grid &gr ) : g( &gr ), is { &gr } {}
    const IndexSet &indexSet () const { return is; }
    const MiniComm &comm () const { return c; }
    const Grid &grid () const { return *g; }
    template< int cd > MiniElemIt< dim > begin () const { return { g, 0 }; }
    template< int cd > MiniElemIt< dim > end () const { return { g, g->cells() }; }
    template< class H > void communicate ( H &, InterfaceType, CommunicationDirection ) const {}
    int size ( int cd ) const { return int( is.size( cd ) ); }
  };

  template< int dim > MiniElemRange< dim > elements ( const MiniGridView< dim > &gv ) { return { gv.g }; }
  template< int dim, class P > MiniElemRange< dim > elements ( const MiniGridView< dim > &gv, P ) { return { gv.g }; }
  template< int dim > MiniIsRange< dim > intersections ( const MiniGridView< dim > &, const MiniEntity< dim > &e ) { return { e }; }

  // --- implementations --------------------------------------------------------------------------------------
  template< int dim >
  typename MiniEntity< dim >::Geometry MiniEntity< dim >::geometry () const
  {
    Geometry q;
    q.h = g->h;
    for( int d = 0; d < dim; ++d )
    {
      q.o[ d ] = g->lo[ d ] + ijk[ d ] * g->h[ d ];
      q.axes[ d ] = d;
    }
    return q;
  }

  template< int dim >
  bool MiniIntersection< dim >::neighbor () const
  {
    const int a = face / 2;
    const int k = in.ijk[ a ] + ( face % 2 ? 1 : -1 );
    return k >= 0 && k < in.g->n[ a ];
  }

  template< int dim >
  typename MiniIntersection< dim >::Geometry MiniIntersection< dim >::geometry () const
  {
    Geometry q;
    const auto eg = in.geometry();
    q.o = eg.o;
    q.h = eg.h;
    const int a = face / 2;
    if( face % 2 )
      q.o[ a ] += eg.h[ a ];
    int j = 0;
    for( int d = 0; d < dim; ++d )
      if( d != a )
        q.axes[ j++ ] = d;
    return q;
  }

  template< int dim >
  std::size_t MiniIndexSet< dim >::size ( int codim ) const
  {
    if( codim == 0 )
      return g->cells();
    std::size_t c = 1;
    for( int d = 0; d < dim; ++d )
      c *= g->n[ d ] + 1;
    return c;
  }

  template< int dim >
  std::size_t MiniIndexSet< dim >::index ( const MiniEntity< dim > &e ) const
  {
    std::size_t k = 0;
    for( int d = dim - 1; d >= 0; --d )
      k = k * g->n[ d ] + e.ijk[ d ];
    return k;
  }

  template< int dim >
  std::size_t MiniIndexSet< dim >::subIndex ( const MiniEntity< dim > &e, int c, int ) const
  {
    std::size_t k = 0;
    for( int d = dim - 1; d >= 0; --d )
      k = k * ( g->n[ d ] + 1 ) + e.ijk[ d ] + ( ( c >> d ) & 1 );
    return k;
  }

  template< int dim >
  bool MiniPartitionList< dim >::contains ( const MiniMultiIndex< dim > &id, unsigned ) const
  {
    for( int d = 0; d < dim; ++d )
      if( id[ d ] < 1 || id[ d ] > 2 * g->n[ d ] - 1 )
        return false;
    return true;
  }

  template< class ct, int d >
  struct MiniRefElem
  {
    FieldVector< ct, d > position ( int, int ) const { return FieldVector< ct, d >( 0.5 ); }
  };

  template< class ct, int d >
  struct ReferenceElements
  {
    static MiniRefElem< ct, d > general ( const GeometryType & ) { return {}; }
  };

} // namespace Dune
