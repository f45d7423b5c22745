// TEST INFRASTRUCTURE ONLY (oracle/): a serial, conforming 2D "GridView" model of triangles and quadrilaterals that is
// nothing but a view of FLATTENED ARRAYS - the same arrays include/vofx.h's vofx_mesh_desc carries to the device.
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

  struct MeshGrid2
  {
    using ctype = double;
    static constexpr int dimension = 2;

    std::int64_t nCells = 0, nVertices = 0;
    const std::int32_t *cornerOffsets = nullptr;   // [nCells + 1]
    const std::int32_t *cornerVertex = nullptr;    // vertex id of every cell corner (DUNE corner order)
    const double *corners = nullptr;               // [ .. ][ 2 ] coordinates of every cell corner (DUNE corner order)
    const double *centers = nullptr;               // [nCells][ 2 ]
    const double *volumes = nullptr;               // [nCells]
    const std::int32_t *faceOffsets = nullptr;     // [nCells + 1], intersections in the grid's iteration order
    const std::int32_t *faceNeighbor = nullptr;    // outside cell or -1
    const double *faceCorners = nullptr;           // [ .. ][ 2 ][ 2 ]
    const double *faceCenters = nullptr;           // [ .. ][ 2 ]
    const double *faceUnitNormals = nullptr;       // [ .. ][ 2 ]
    const double *faceIntNormals = nullptr;        // [ .. ][ 2 ]
    const double *faceVolumes = nullptr;           // [ .. ]
  };

  template< int mydim >
  struct MeshGeometry
  {
    using ctype = double;
    using GlobalCoordinate = FieldVector< double, 2 >;
    using LocalCoordinate = FieldVector< double, mydim >;
    static constexpr int mydimension = mydim, coorddimension = 2;

    int nc = 0;
    const double *x = nullptr;     // nc corners
    const double *ctr = nullptr;
    double vol = 0.0;

    GeometryType type () const { return { mydim, mydim == 2 && nc == 3 }; }
    int corners () const { return nc; }
    GlobalCoordinate corner ( int i ) const { return GlobalCoordinate( { x[ 2 * i ], x[ 2 * i + 1 ] } ); }
    GlobalCoordinate center () const { return GlobalCoordinate( { ctr[ 0 ], ctr[ 1 ] } ); }
    double volume () const { return vol; }
  };

  struct MeshEntity
  {
    static constexpr int codimension = 0, dimension = 2;
    using Geometry = MeshGeometry< 2 >;

    const MeshGrid2 *g = nullptr;
    std::int64_t idx = -1;

    Geometry geometry () const
    {
      Geometry q;
      q.nc = g->cornerOffsets[ idx + 1 ] - g->cornerOffsets[ idx ];
      q.x = g->corners + 2 * std::size_t( g->cornerOffsets[ idx ] );
      q.ctr = g->centers + 2 * idx;
      q.vol = g->volumes[ idx ];
      return q;
    }
    bool operator== ( const MeshEntity &o ) const { return idx == o.idx; }
    bool operator!= ( const MeshEntity &o ) const { return idx != o.idx; }
  };

  struct MeshIntersection
  {
    using Entity = MeshEntity;
    using Geometry = MeshGeometry< 1 >;
    using LocalCoordinate = FieldVector< double, 1 >;
    using GlobalCoordinate = FieldVector< double, 2 >;

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
      q.nc = 2;
      q.x = in.g->faceCorners + 4 * slot;
      q.ctr = in.g->faceCenters + 2 * slot;
      q.vol = in.g->faceVolumes[ slot ];
      return q;
    }
    GeometryType type () const { return { 1 }; }
    GlobalCoordinate centerUnitOuterNormal () const
    {
      return GlobalCoordinate( { in.g->faceUnitNormals[ 2 * slot ], in.g->faceUnitNormals[ 2 * slot + 1 ] } );
    }
    GlobalCoordinate integrationOuterNormal ( const LocalCoordinate & ) const
    {
      return GlobalCoordinate( { in.g->faceIntNormals[ 2 * slot ], in.g->faceIntNormals[ 2 * slot + 1 ] } );
    }
  };

  struct MeshIndexSet
  {
    using IndexType = std::size_t;
    const MeshGrid2 *g;
    std::size_t size ( int codim ) const { return codim == 0 ? std::size_t( g->nCells ) : std::size_t( g->nVertices ); }
    std::size_t index ( const MeshEntity &e ) const { return std::size_t( e.idx ); }
    std::size_t subIndex ( const MeshEntity &e, int k, int ) const { return std::size_t( g->cornerVertex[ g->cornerOffsets[ e.idx ] + k ] ); }
  };

  struct MeshElemIt
  {
    using Entity = MeshEntity;
    const MeshGrid2 *g;
    std::int64_t i;
    Entity operator* () const { return Entity { g, i }; }
    MeshElemIt &operator++ () { ++i; return *this; }
    bool operator!= ( const MeshElemIt &o ) const { return i != o.i; }
  };

  struct MeshElemRange
  {
    const MeshGrid2 *g;
    MeshElemIt begin () const { return { g, 0 }; }
    MeshElemIt end () const { return { g, g->nCells }; }
  };

  struct MeshIsIt
  {
    MeshEntity e;
    std::int64_t slot;
    MeshIntersection operator* () const { return { e, slot }; }
    MeshIsIt &operator++ () { ++slot; return *this; }
    bool operator!= ( const MeshIsIt &o ) const { return slot != o.slot; }
  };

  struct MeshIsRange
  {
    MeshEntity e;
    MeshIsIt begin () const { return { e, e.g->faceOffsets[ e.idx ] }; }
    MeshIsIt end () const { return { e, e.g->faceOffsets[ e.idx + 1 ] }; }
  };

  struct MeshGridView
  {
    using Grid = MeshGrid2;
    using ctype = double;
    static constexpr int dimension = 2, dimensionworld = 2;
    using IndexSet = MeshIndexSet;
    using Intersection = MeshIntersection;
    using CollectiveCommunication = MiniComm;
    template< int cd > struct Codim { using Entity = MeshEntity; using Geometry = MeshGeometry< 2 >; };

    const Grid *g;
    IndexSet is;
    MiniComm c;

    MeshGridView ( const Grid &gr ) : g( &gr ), is { &gr } {}
    const IndexSet &indexSet () const { return is; }
    const MiniComm &comm () const { return c; }
    const Grid &grid () const { return *g; }
    template< int cd > MeshElemIt begin () const { return { g, 0 }; }
    template< int cd > MeshElemIt end () const { return { g, g->nCells }; }
    template< class H > void communicate ( H &, InterfaceType, CommunicationDirection ) const {}
    int size ( int cd ) const { return int( is.size( cd ) ); }
  };

  inline MeshElemRange elements ( const MeshGridView &gv ) { return { gv.g }; }
  template< class P > MeshElemRange elements ( const MeshGridView &gv, P ) { return { gv.g }; }
  inline MeshIsRange intersections ( const MeshGridView &, const MeshEntity &e ) { return { e }; }

} // namespace Dune
