// TEST INFRASTRUCTURE ONLY (oracle/): a serial, conforming "GridView" model of unstructured grids -- triangles and
// quadrilaterals in 2D, tetrahedra and hexahedra in 3D -- that is nothing but a view of FLATTENED ARRAYS, the same arrays
// include/vofx.h's vofx_mesh_desc carries to the device.
// Every geometric quantity the reference's operators ask a DUNE grid for (corner, center, volume of cells and
// intersections, unit and integration outer normals) is returned from those arrays verbatim: this model does no
// arithmetic, so the reference (on this model), the C restatement and the device all start from identical numbers,
// exactly as they would behind a real dune-grid (whose arithmetic is third party).
// Nothing here is reference code.
#pragma once
#include <cstdint>
#include <vector>

#include <dune/grid/common/minigrid.hh>

namespace Dune
{

  template< int dim >
  struct MeshGridT
  {
    using ctype = double;
    static constexpr int dimension = dim;

    std::int64_t nCells = 0, nVertices = 0;
    const std::int32_t *cornerOffsets = nullptr;      // [nCells + 1]
    const std::int32_t *cornerVertex = nullptr;       // vertex id of every cell corner (DUNE corner order)
    const double *corners = nullptr;                  // [ .. ][ dim ] coordinates of every cell corner (DUNE corner order)
    const double *centers = nullptr;                  // [nCells][ dim ]
    const double *volumes = nullptr;                  // [nCells]
    const std::int32_t *faceOffsets = nullptr;        // [nCells + 1], intersections in the grid's iteration order
    const std::int32_t *faceNeighbor = nullptr;       // outside cell or -1
    const std::int32_t *faceCornerOffsets = nullptr;  // [nFaces + 1]; null: two corners per intersection (2D)
    const double *faceCorners = nullptr;              // [ .. ][ dim ] corners of every intersection (DUNE corner order)
    const double *faceCenters = nullptr;              // [ .. ][ dim ]
    const double *faceUnitNormals = nullptr;          // [ .. ][ dim ]
    const double *faceIntNormals = nullptr;           // [ .. ][ dim ]
    const double *faceVolumes = nullptr;              // [ .. ]

    std::int64_t faceCornerBegin ( std::int64_t slot ) const { return faceCornerOffsets ? faceCornerOffsets[ slot ] : 2 * slot; }
    int faceCornerCount ( std::int64_t slot ) const { return faceCornerOffsets ? int( faceCornerOffsets[ slot + 1 ] - faceCornerOffsets[ slot ] ) : 2; }
  };

  template< int dim, int mydim >
  struct MeshGeometryT
  {
    using ctype = double;
    using GlobalCoordinate = FieldVector< double, dim >;
    using LocalCoordinate = FieldVector< double, mydim >;
    static constexpr int mydimension = mydim, coorddimension = dim;

    int nc = 0;
    const double *x = nullptr;     // nc corners
    const double *ctr = nullptr;
    double vol = 0.0;

    // simplices have mydim + 1 corners, cubes 2^mydim
    GeometryType type () const { return { mydim, mydim >= 2 && nc == mydim + 1 }; }
    int corners () const { return nc; }
    GlobalCoordinate corner ( int i ) const
    {
      GlobalCoordinate c;
      for( int k = 0; k < dim; ++k )
        c[ k ] = x[ dim * i + k ];
      return c;
    }
    GlobalCoordinate center () const
    {
      GlobalCoordinate c;
      for( int k = 0; k < dim; ++k )
        c[ k ] = ctr[ k ];
      return c;
    }
    double volume () const { return vol; }
  };

  template< int dim >
  struct MeshEntityT
  {
    static constexpr int codimension = 0, dimension = dim;
    using Geometry = MeshGeometryT< dim, dim >;

    const MeshGridT< dim > *g = nullptr;
    std::int64_t idx = -1;

    Geometry geometry () const
    {
      Geometry q;
      q.nc = g->cornerOffsets[ idx + 1 ] - g->cornerOffsets[ idx ];
      q.x = g->corners + dim * std::size_t( g->cornerOffsets[ idx ] );
      q.ctr = g->centers + dim * idx;
      q.vol = g->volumes[ idx ];
      return q;
    }
    bool operator== ( const MeshEntityT &o ) const { return idx == o.idx; }
    bool operator!= ( const MeshEntityT &o ) const { return idx != o.idx; }
  };

  template< int dim >
  struct MeshIntersectionT
  {
    using Entity = MeshEntityT< dim >;
    using Geometry = MeshGeometryT< dim, dim - 1 >;
    using LocalCoordinate = FieldVector< double, dim - 1 >;
    using GlobalCoordinate = FieldVector< double, dim >;

    Entity in;
    std::int64_t slot;     // position in the face arrays

    int indexInInside () const { return int( slot - in.g->faceOffsets[ in.idx ] ); }
    bool neighbor () const { return in.g->faceNeighbor[ slot ] >= 0; }
    bool boundary () const { return !neighbor(); }
    Entity inside () const { return in; }
    Entity outside () const { return Entity { in.g, in.g->faceNeighbor[ slot ] }; }
    Geometry geometry () const
    {
      Geometry q;
      q.nc = in.g->faceCornerCount( slot );
      q.x = in.g->faceCorners + dim * in.g->faceCornerBegin( slot );
      q.ctr = in.g->faceCenters + dim * slot;
      q.vol = in.g->faceVolumes[ slot ];
      return q;
    }
    GeometryType type () const { return geometry().type(); }
    GlobalCoordinate centerUnitOuterNormal () const
    {
      GlobalCoordinate n;
      for( int k = 0; k < dim; ++k )
        n[ k ] = in.g->faceUnitNormals[ dim * slot + k ];
      return n;
    }
    GlobalCoordinate integrationOuterNormal ( const LocalCoordinate & ) const
    {
      GlobalCoordinate n;
      for( int k = 0; k < dim; ++k )
        n[ k ] = in.g->faceIntNormals[ dim * slot + k ];
      return n;
    }
  };

  template< int dim >
  struct MeshIndexSetT
  {
    using IndexType = std::size_t;
    const MeshGridT< dim > *g;
    std::size_t size ( int codim ) const { return codim == 0 ? std::size_t( g->nCells ) : std::size_t( g->nVertices ); }
    std::size_t index ( const MeshEntityT< dim > &e ) const { return std::size_t( e.idx ); }
    std::size_t subIndex ( const MeshEntityT< dim > &e, int k, int ) const { return std::size_t( g->cornerVertex[ g->cornerOffsets[ e.idx ] + k ] ); }
  };

  template< int dim >
  struct MeshElemItT
  {
    using Entity = MeshEntityT< dim >;
    const MeshGridT< dim > *g;
    std::int64_t i;
    Entity operator* () const { return Entity { g, i }; }
    MeshElemItT &operator++ () { ++i; return *this; }
    bool operator!= ( const MeshElemItT &o ) const { return i != o.i; }
  };

  template< int dim >
  struct MeshElemRangeT
  {
    const MeshGridT< dim > *g;
    MeshElemItT< dim > begin () const { return { g, 0 }; }
    MeshElemItT< dim > end () const { return { g, g->nCells }; }
  };

  template< int dim >
  struct MeshIsItT
  {
    MeshEntityT< dim > e;
    std::int64_t slot;
    MeshIntersectionT< dim > operator* () const { return { e, slot }; }
    MeshIsItT &operator++ () { ++slot; return *this; }
    bool operator!= ( const MeshIsItT &o ) const { return slot != o.slot; }
  };

  template< int dim >
  struct MeshIsRangeT
  {
    MeshEntityT< dim > e;
    MeshIsItT< dim > begin () const { return { e, e.g->faceOffsets[ e.idx ] }; }
    MeshIsItT< dim > end () const { return { e, e.g->faceOffsets[ e.idx + 1 ] }; }
  };

  template< int dim >
  struct MeshGridViewT
  {
    using Grid = MeshGridT< dim >;
    using ctype = double;
    static constexpr int dimension = dim, dimensionworld = dim;
    using IndexSet = MeshIndexSetT< dim >;
    using Intersection = MeshIntersectionT< dim >;
    using CollectiveCommunication = MiniComm;
    template< int cd > struct Codim { using Entity = MeshEntityT< dim >; using Geometry = MeshGeometryT< dim, dim >; };

    const Grid *g;
    IndexSet is;
    MiniComm c;

    MeshGridViewT ( const Grid &gr ) : g( &gr ), is { &gr } {}
    const IndexSet &indexSet () const { return is; }
    const MiniComm &comm () const { return c; }
    const Grid &grid () const { return *g; }
    template< int cd > MeshElemItT< dim > begin () const { return { g, 0 }; }
    template< int cd > MeshElemItT< dim > end () const { return { g, g->nCells }; }
    template< class H > void communicate ( H &, InterfaceType, CommunicationDirection ) const {}
    int size ( int cd ) const { return int( is.size( cd ) ); }
  };

  template< int dim > MeshElemRangeT< dim > elements ( const MeshGridViewT< dim > &gv ) { return { gv.g }; }
  template< int dim, class P > MeshElemRangeT< dim > elements ( const MeshGridViewT< dim > &gv, P ) { return { gv.g }; }
  template< int dim > MeshIsRangeT< dim > intersections ( const MeshGridViewT< dim > &, const MeshEntityT< dim > &e ) { return { e }; }

  // the 2D names the harnesses were written against
  using MeshGrid2 = MeshGridT< 2 >;
  template< int mydim > using MeshGeometry = MeshGeometryT< 2, mydim >;
  using MeshEntity = MeshEntityT< 2 >;
  using MeshIntersection = MeshIntersectionT< 2 >;
  using MeshIndexSet = MeshIndexSetT< 2 >;
  using MeshGridView = MeshGridViewT< 2 >;
  using MeshGrid3 = MeshGridT< 3 >;
  using MeshEntity3 = MeshEntityT< 3 >;
  using MeshIntersection3 = MeshIntersectionT< 3 >;
  using MeshGridView3 = MeshGridViewT< 3 >;

} // namespace Dune
