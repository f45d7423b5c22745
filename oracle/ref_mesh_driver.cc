// TEST INFRASTRUCTURE ONLY (oracle/).  Not part of the product; the product never loads this.
//
// C-ABI harness around the UNMODIFIED reference headers (/root/reference/dune/vof/**) instantiated on the
// unstructured grid model oracle/shim/dune/grid/common/meshgrid.hh (flattened arrays): triangles and quadrilaterals
// (MESHDIM 2, the default, symbols refmesh_*) or tetrahedra and hexahedra (-DMESHDIM=3, symbols refmesh3_*).  Built by
// oracle/Makefile into oracle/_ref/libvofrefmesh.so and oracle/_ref/libvofrefmesh3.so (git-ignored).  It pins
// vo_mesh_* of oracle/vof_oracle.c and generates tests/golden/mesh2d_*.npz / mesh3d_*.npz.
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

#ifndef MESHDIM
#define MESHDIM 2
#endif
#define GRIDDIM MESHDIM
#if MESHDIM == 2
#define RM( name ) refmesh_##name
#else
#define RM( name ) refmesh3_##name
#endif

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
#if MESHDIM == 2
#include <dune/vof/reconstruction/modifiedswartz.hh>
#endif

namespace
{

  thread_local std::string lastError;

  constexpr int D = MESHDIM;
  using Coord = Dune::FieldVector< double, D >;
  using Entity = Dune::MeshEntityT< D >;
  using Intersection = Dune::MeshIntersectionT< D >;

  Coord coord ( const double *x )
  {
    Coord c;
    for( int k = 0; k < D; ++k )
      c[ k ] = x[ k ];
    return c;
  }

  struct SampledVelocity
  {
    const double *samples;
    const Intersection *is = nullptr;

    void bind ( const Intersection &intersection ) { is = &intersection; }
    void unbind () { is = nullptr; }
    Coord operator() ( const Dune::FieldVector< double, D - 1 > & ) const { return coord( samples + D * is->slot ); }
  };

  struct Ctx
  {
    using GridView = Dune::MeshGridViewT< D >;
    using Stencils = Dune::VoF::VertexNeighborsStencil< GridView >;
    using Color = Dune::VoF::ColorFunction< GridView >;
    using Flags = Dune::VoF::FlagSet< GridView >;
    using Recs = Dune::VoF::ReconstructionSet< GridView >;
    using Youngs = Dune::VoF::ModifiedYoungsReconstruction< GridView, Stencils >;
#if MESHDIM == 2
    using Swartz = Dune::VoF::ModifiedSwartzReconstruction< GridView, Stencils, Youngs >;
#endif

    // owned copies of the caller's arrays
    std::vector< std::int32_t > cornerOffsets, cornerVertex, faceOffsets, faceNeighbor, faceCornerOffsets;
    std::vector< double > corners, centers, volumes, faceCorners, faceCenters, faceUnitNormals, faceIntNormals, faceVolumes;
    Dune::MeshGridT< D > grid;
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
        *it = Dune::VoF::HalfSpace< Coord >( coord( hs + ( D + 1 ) * i ), hs[ ( D + 1 ) * i + D ] );
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

  const char *RM( last_error ) () { return lastError.c_str(); }

  // faceCornerOffsets: [faces + 1] positions in faceCorners (in corners, not doubles); null = two corners per intersection
  void *RM( create ) ( std::int64_t nCells, std::int64_t nVertices, const std::int32_t *cornerOffsets, const std::int32_t *cornerVertex,
                       const double *corners, const double *centers, const double *volumes, const std::int32_t *faceOffsets,
                       const std::int32_t *faceNeighbor, const double *faceCorners, const double *faceCenters,
                       const double *faceUnitNormals, const double *faceIntNormals, const double *faceVolumes,
                       const std::int32_t *faceCornerOffsets )
  {
    auto *c = new Ctx;
    const std::size_t nCorners = std::size_t( cornerOffsets[ nCells ] ), nFaces = std::size_t( faceOffsets[ nCells ] );
    c->cornerOffsets.assign( cornerOffsets, cornerOffsets + nCells + 1 );
    c->cornerVertex.assign( cornerVertex, cornerVertex + nCorners );
    c->corners.assign( corners, corners + D * nCorners );
    c->centers.assign( centers, centers + D * nCells );
    c->volumes.assign( volumes, volumes + nCells );
    c->faceOffsets.assign( faceOffsets, faceOffsets + nCells + 1 );
    c->faceNeighbor.assign( faceNeighbor, faceNeighbor + nFaces );
    if( faceCornerOffsets )
      c->faceCornerOffsets.assign( faceCornerOffsets, faceCornerOffsets + nFaces + 1 );
    const std::size_t nFaceCorners = faceCornerOffsets ? std::size_t( faceCornerOffsets[ nFaces ] ) : 2 * nFaces;
    c->faceCorners.assign( faceCorners, faceCorners + D * nFaceCorners );
    c->faceCenters.assign( faceCenters, faceCenters + D * nFaces );
    c->faceUnitNormals.assign( faceUnitNormals, faceUnitNormals + D * nFaces );
    c->faceIntNormals.assign( faceIntNormals, faceIntNormals + D * nFaces );
    c->faceVolumes.assign( faceVolumes, faceVolumes + nFaces );
    Dune::MeshGridT< D > &g = c->grid;
    g.nCells = nCells;
    g.nVertices = nVertices;
    g.cornerOffsets = c->cornerOffsets.data();
    g.cornerVertex = c->cornerVertex.data();
    g.corners = c->corners.data();
    g.centers = c->centers.data();
    g.volumes = c->volumes.data();
    g.faceOffsets = c->faceOffsets.data();
    g.faceNeighbor = c->faceNeighbor.data();
    g.faceCornerOffsets = faceCornerOffsets ? c->faceCornerOffsets.data() : nullptr;
    g.faceCorners = c->faceCorners.data();
    g.faceCenters = c->faceCenters.data();
    g.faceUnitNormals = c->faceUnitNormals.data();
    g.faceIntNormals = c->faceIntNormals.data();
    g.faceVolumes = c->faceVolumes.data();
    c->gv = std::make_unique< Ctx::GridView >( c->grid );
    c->stencils = std::make_unique< Ctx::Stencils >( *c->gv );   // stencil/vertexneighborsstencil.hh:52-94
    return c;
  }

  void RM( destroy ) ( void *handle ) { delete static_cast< Ctx * >( handle ); }

  // the reference's vertex-neighbour stencils as CSR (entries = cell indices in the order the reference stores them)
  std::int64_t RM( stencil ) ( void *handle, std::int32_t *offsets, std::int32_t *entries, std::int64_t capacity )
  {
    Ctx &ctx = *static_cast< Ctx * >( handle );
    std::int64_t k = 0;
    for( std::int64_t i = 0; i < ctx.grid.nCells; ++i )
    {
      offsets[ i ] = std::int32_t( k );
      for( const auto &nb : ( *ctx.stencils )[ Entity { &ctx.grid, i } ] )
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
  int RM( flag ) ( void *handle, const double *u, double eps, std::uint8_t *flagsOut )
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
  int RM( youngs ) ( void *handle, const double *u, const std::uint8_t *flagsIn, double *hs )
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
        for( int k = 0; k < D; ++k )
          hs[ ( D + 1 ) * i + k ] = it->innerNormal()[ k ];
        hs[ ( D + 1 ) * i + D ] = it->boundary().distance();
      }
      return 0;
    } );
  }

#if MESHDIM == 2
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
        for( int k = 0; k < D; ++k )
          hs[ ( D + 1 ) * i + k ] = it->innerNormal()[ k ];
        hs[ ( D + 1 ) * i + D ] = it->boundary().distance();
      }
      return 0;
    } );
  }

#endif

  // evolution/evolution.hh:58-76 with one sampled velocity per intersection
  int RM( evolve ) ( void *handle, const double *hs, const std::uint8_t *flagsIn, const double *faceVelocity, double dt,
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

  // locateHalfSpace( Polygon / Polyhedron, normal, fraction ), geometry/algorithm.hh:68-111 / :119-236, on one cell of the mesh
  int RM( locate ) ( void *handle, std::int64_t cell, const double *normal, double fraction, double *out )
  {
    return guarded( [ & ] {
      Ctx &ctx = *static_cast< Ctx * >( handle );
      const auto polytope = Dune::VoF::makePolytope( Entity { &ctx.grid, cell }.geometry() );
      const auto hs = Dune::VoF::locateHalfSpace( polytope, coord( normal ), fraction );
      for( int k = 0; k < D; ++k )
        out[ k ] = hs.innerNormal()[ k ];
      out[ D ] = hs.boundary().distance();
      return 0;
    } );
  }

} // extern "C"
