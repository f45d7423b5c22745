// TEST INFRASTRUCTURE ONLY (oracle/).  Not part of the product; the product never loads this.
//
// C-ABI harness around the UNMODIFIED reference headers (/root/reference/dune/vof/**) instantiated on the
// unstructured 2D grid model oracle/shim/dune/grid/common/meshgrid.hh (triangles and quadrilaterals as flattened
// arrays).  Built by oracle/Makefile into oracle/_ref/libvofrefmesh.so (git-ignored).  It pins
// vo_mesh_* of oracle/vof_oracle.c and generates tests/golden/mesh2d_*.npz.
//
// The velocity is a SAMPLED field: one vector per intersection, evaluated by the caller at the intersection centre
// (what Velocity< Problem, GridView >::operator() returns for refElement.position( 0, 0 ), velocity.hh:15-20,
// evolution/evolution.hh:107-108).  The class below models the reference's Velocity concept on those samples.
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#define GRIDDIM 2

#include <dune/common/exceptions.hh>
#include <dune/common/fmatrix.hh>
#include <dune/common/fvector.hh>
#include <dune/grid/common/meshgrid.hh>

#include <dune/vof/colorfunction.hh>
#include <dune/vof/flagset.hh>
#include <dune/vof/flagging.hh>
#include <dune/vof/reconstructionset.hh>
#include <dune/vof/evolution.hh>
#include <dune/vof/geometry/algorithm.hh>
#include <dune/vof/geometry/intersect.hh>
#include <dune/vof/geometry/polytope.hh>
#include <dune/vof/geometry/upwindpolygon.hh>
#include <dune/vof/stencil/vertexneighborsstencil.hh>
#include <dune/vof/reconstruction/modifiedyoungs.hh>
#include <dune/vof/reconstruction/modifiedswartz.hh>

namespace
{

  thread_local std::string lastError;

  struct SampledVelocity
  {
    using Coord = Dune::FieldVector< double, 2 >;
    const double *samples;
    const Dune::MeshIntersection *is = nullptr;

    void bind ( const Dune::MeshIntersection &intersection ) { is = &intersection; }
    void unbind () { is = nullptr; }
    Coord operator() ( const Dune::FieldVector< double, 1 > & ) const
    {
      return Coord( { samples[ 2 * is->slot ], samples[ 2 * is->slot + 1 ] } );
    }
  };

  struct Ctx
  {
    using GridView = Dune::MeshGridView;
    using Coord = Dune::FieldVector< double, 2 >;
    using Stencils = Dune::VoF::VertexNeighborsStencil< GridView >;
    using Color = Dune::VoF::ColorFunction< GridView >;
    using Flags = Dune::VoF::FlagSet< GridView >;
    using Recs = Dune::VoF::ReconstructionSet< GridView >;
    using Youngs = Dune::VoF::ModifiedYoungsReconstruction< GridView, Stencils >;
    using Swartz = Dune::VoF::ModifiedSwartzReconstruction< GridView, Stencils, Youngs >;

    // owned copies of the caller's arrays
    std::vector< std::int32_t > cornerOffsets, cornerVertex, faceOffsets, faceNeighbor;
    std::vector< double > corners, centers, volumes, faceCorners, faceCenters, faceUnitNormals, faceIntNormals, faceVolumes;
    Dune::MeshGrid2 grid;
    std::unique_ptr< GridView > gv;
    std::unique_ptr< Stencils > stencils;

    std::size_t cells () const { return std::size_t( grid.nCells ); }

    void loadFlags ( const std::uint8_t *f, Flags &flags ) const
    {
      auto it = flags.begin();
      for( std::size_t i = 0; i < cells(); ++i, ++it )
        *it = static_cast< Dune::VoF::Flag >( static_cast< std::size_t >( f[ i ] ) );
    }
    void loadRecs ( const double *hs, Recs &recs ) const
    {
      auto it = recs.begin();
      for( std::size_t i = 0; i < cells(); ++i, ++it )
        *it = Dune::VoF::HalfSpace< Coord >( Coord( { hs[ 3 * i ], hs[ 3 * i + 1 ] } ), hs[ 3 * i + 2 ] );
    }
  };

  template< class F >
  int guarded ( F &&f )
  {
    try
    {
      return f();
    }
    catch( const std::exception &e )
    {
      lastError = e.what();
      return 2;
    }
    catch( ... )
    {
      lastError = "unknown exception";
      return 3;
    }
  }

} // namespace


extern "C"
{

  const char *refmesh_last_error () { return lastError.c_str(); }

  void *refmesh_create ( std::int64_t nCells, std::int64_t nVertices, const std::int32_t *cornerOffsets, const std::int32_t *cornerVertex,
                         const double *corners, const double *centers, const double *volumes, const std::int32_t *faceOffsets,
                         const std::int32_t *faceNeighbor, const double *faceCorners, const double *faceCenters,
                         const double *faceUnitNormals, const double *faceIntNormals, const double *faceVolumes )
  {
    auto *c = new Ctx;
    const std::size_t nCorners = std::size_t( cornerOffsets[ nCells ] ), nFaces = std::size_t( faceOffsets[ nCells ] );
    c->cornerOffsets.assign( cornerOffsets, cornerOffsets + nCells + 1 );
    c->cornerVertex.assign( cornerVertex, cornerVertex + nCorners );
    c->corners.assign( corners, corners + 2 * nCorners );
    c->centers.assign( centers, centers + 2 * nCells );
    c->volumes.assign( volumes, volumes + nCells );
    c->faceOffsets.assign( faceOffsets, faceOffsets + nCells + 1 );
    c->faceNeighbor.assign( faceNeighbor, faceNeighbor + nFaces );
    c->faceCorners.assign( faceCorners, faceCorners + 4 * nFaces );
    c->faceCenters.assign( faceCenters, faceCenters + 2 * nFaces );
    c->faceUnitNormals.assign( faceUnitNormals, faceUnitNormals + 2 * nFaces );
    c->faceIntNormals.assign( faceIntNormals, faceIntNormals + 2 * nFaces );
    c->faceVolumes.assign( faceVolumes, faceVolumes + nFaces );
    Dune::MeshGrid2 &g = c->grid;
    g.nCells = nCells;
    g.nVertices = nVertices;
    g.cornerOffsets = c->cornerOffsets.data();
    g.cornerVertex = c->cornerVertex.data();
    g.corners = c->corners.data();
    g.centers = c->centers.data();
    g.volumes = c->volumes.data();
    g.faceOffsets = c->faceOffsets.data();
    g.faceNeighbor = c->faceNeighbor.data();
    g.faceCorners = c->faceCorners.data();
    g.faceCenters = c->faceCenters.data();
    g.faceUnitNormals = c->faceUnitNormals.data();
    g.faceIntNormals = c->faceIntNormals.data();
    g.faceVolumes = c->faceVolumes.data();
    c->gv = std::make_unique< Ctx::GridView >( c->grid );
    c->stencils = std::make_unique< Ctx::Stencils >( *c->gv );   // stencil/vertexneighborsstencil.hh:52-94
    return c;
  }

  void refmesh_destroy ( void *handle ) { delete static_cast< Ctx * >( handle ); }

  // the reference's vertex-neighbour stencils as CSR (entries = cell indices in the order the reference stores them)
  std::int64_t refmesh_stencil ( void *handle, std::int32_t *offsets, std::int32_t *entries, std::int64_t capacity )
  {
    Ctx &ctx = *static_cast< Ctx * >( handle );
    std::int64_t k = 0;
    for( std::int64_t i = 0; i < ctx.grid.nCells; ++i )
    {
      offsets[ i ] = std::int32_t( k );
      for( const auto &nb : ( *ctx.stencils )[ Dune::MeshEntity { &ctx.grid, i } ] )
      {
        if( k < capacity )
          entries[ k ] = std::int32_t( nb.idx );
        ++k;
      }
    }
    offsets[ ctx.grid.nCells ] = std::int32_t( k );
    return k;
  }

  // flagging.hh:35-70
  int refmesh_flag ( void *handle, const double *u, double eps, std::uint8_t *flagsOut )
  {
    return guarded( [ & ] {
      Ctx &ctx = *static_cast< Ctx * >( handle );
      Ctx::Color uh( *ctx.gv );
      Ctx::Flags flags( *ctx.gv );
      std::copy( u, u + ctx.cells(), uh.begin() );
      Dune::VoF::FlagOperator< Ctx::GridView > flagOperator( eps );
      flagOperator( uh, flags );
      auto it = static_cast< const Dune::VoF::DataSet< Ctx::GridView, Dune::VoF::Flag > & >( flags ).begin();
      for( std::size_t i = 0; i < ctx.cells(); ++i, ++it )
        flagsOut[ i ] = static_cast< std::uint8_t >( static_cast< std::size_t >( *it ) );
      return 0;
    } );
  }

  // reconstruction/modifiedyoungs.hh:63-76
  int refmesh_youngs ( void *handle, const double *u, const std::uint8_t *flagsIn, double *hs )
  {
    return guarded( [ & ] {
      Ctx &ctx = *static_cast< Ctx * >( handle );
      Ctx::Color uh( *ctx.gv );
      Ctx::Flags flags( *ctx.gv );
      Ctx::Recs recs( *ctx.gv );
      std::copy( u, u + ctx.cells(), uh.begin() );
      ctx.loadFlags( flagsIn, flags );
      Ctx::Youngs youngs( *ctx.stencils );
      youngs( uh, recs, flags );
      auto it = recs.begin();
      for( std::size_t i = 0; i < ctx.cells(); ++i, ++it )
      {
        hs[ 3 * i ] = it->innerNormal()[ 0 ];
        hs[ 3 * i + 1 ] = it->innerNormal()[ 1 ];
        hs[ 3 * i + 2 ] = it->boundary().distance();
      }
      return 0;
    } );
  }

  // reconstruction/modifiedswartz.hh:70-153 (2D), initialised by ModifiedYoungsReconstruction
  int refmesh_swartz ( void *handle, const double *u, const std::uint8_t *flagsIn, double *hs, int maxIterations )
  {
    return guarded( [ & ] {
      Ctx &ctx = *static_cast< Ctx * >( handle );
      Ctx::Color uh( *ctx.gv );
      Ctx::Flags flags( *ctx.gv );
      Ctx::Recs recs( *ctx.gv );
      std::copy( u, u + ctx.cells(), uh.begin() );
      ctx.loadFlags( flagsIn, flags );
      Ctx::Swartz swartz( *ctx.stencils, std::size_t( maxIterations ) );
      swartz( uh, recs, flags );
      auto it = recs.begin();
      for( std::size_t i = 0; i < ctx.cells(); ++i, ++it )
      {
        hs[ 3 * i ] = it->innerNormal()[ 0 ];
        hs[ 3 * i + 1 ] = it->innerNormal()[ 1 ];
        hs[ 3 * i + 2 ] = it->boundary().distance();
      }
      return 0;
    } );
  }

  // evolution/evolution.hh:58-76 with one sampled velocity per intersection
  int refmesh_evolve ( void *handle, const double *hs, const std::uint8_t *flagsIn, const double *faceVelocity, double dt,
                       double *updateOut, double *dtEst )
  {
    return guarded( [ & ] {
      Ctx &ctx = *static_cast< Ctx * >( handle );
      Ctx::Color update( *ctx.gv );
      Ctx::Flags flags( *ctx.gv );
      Ctx::Recs recs( *ctx.gv );
      ctx.loadFlags( flagsIn, flags );
      ctx.loadRecs( hs, recs );
      SampledVelocity velocity { faceVelocity };
      auto evolutionOperator = Dune::VoF::evolution( *ctx.gv );
      *dtEst = evolutionOperator( recs, flags, velocity, dt, update );
      std::copy( update.begin(), update.end(), updateOut );
      return 0;
    } );
  }

  // locateHalfSpace( Polygon, normal, fraction ), geometry/algorithm.hh:68-111, on one cell of the mesh
  int refmesh_locate ( void *handle, std::int64_t cell, const double *normal, double fraction, double *out )
  {
    return guarded( [ & ] {
      Ctx &ctx = *static_cast< Ctx * >( handle );
      const auto polygon = Dune::VoF::makePolytope( Dune::MeshEntity { &ctx.grid, cell }.geometry() );
      const auto hs = Dune::VoF::locateHalfSpace( polygon, Ctx::Coord( { normal[ 0 ], normal[ 1 ] } ), fraction );
      out[ 0 ] = hs.innerNormal()[ 0 ];
      out[ 1 ] = hs.innerNormal()[ 1 ];
      out[ 2 ] = hs.boundary().distance();
      return 0;
    } );
  }

} // extern "C"
