// Test of include/dune/vofx/unstructured.hh on a 3D grid of TETRAHEDRA (needs a GPU at run time; built on the CPU box).
//
// A conforming mesh of the unit cube (jittered lattice, every lattice cube split into six tetrahedra around its main
// diagonal) is put behind the unstructured grid-view model of oracle/shim (test infrastructure, 3D instance); the adaptor
// flattens that grid view through the DUNE interface and is driven like the reference drives its operators
// (algorithm.hh:57-93).  Results are compared bit-for-bit with the oracle (vo_mesh_*, libvoforacle.so) on the arrays the
// grid view was built from.  With -DVOFX_TEST_WITH_REFERENCE (only where /root/reference is mounted) the containers are the
// reference's own and the reference's operators (makePolyhedron simplex table, upwindPolygon3d prisms) run side by side.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <vector>

#define GRIDDIM 3
#include <dune/common/fvector.hh>
#include <dune/grid/common/meshgrid.hh>

#ifdef VOFX_TEST_WITH_REFERENCE
#include <dune/vof/colorfunction.hh>
#include <dune/vof/flagset.hh>
#include <dune/vof/flagging.hh>
#include <dune/vof/reconstructionset.hh>
#include <dune/vof/evolution.hh>
#include <dune/vof/stencil/vertexneighborsstencil.hh>
#include <dune/vof/reconstruction/modifiedyoungs.hh>
#endif

#include <dune/vofx/unstructured.hh>

extern "C"
{
#include "vof_oracle.h"
}

namespace
{

  using GV = Dune::MeshGridView3;
  using Coord = Dune::FieldVector< double, 3 >;

#ifndef VOFX_TEST_WITH_REFERENCE
  enum class Flag : std::size_t { empty = 0, mixed = 1, mixedfull = 2, full = 3, nan = 99 };

  template< class T >
  struct DataSet   // interface of dune/vof/dataset.hh:26-87
  {
    using DataType = T;
    explicit DataSet ( const GV &gv ) : data_( gv.indexSet().size( 0 ) ) {}
    const T &operator[] ( std::size_t i ) const { return data_[ i ]; }
    T &operator[] ( std::size_t i ) { return data_[ i ]; }
    auto begin () { return data_.begin(); }
    auto end () { return data_.end(); }
    auto begin () const { return data_.begin(); }
    auto end () const { return data_.end(); }
    std::vector< T > data_;
  };
  struct HyperPlane { Coord n; double d; double distance () const { return d; } };
  struct HalfSpace
  {
    using Coordinate = Coord;
    HalfSpace () : n_( 0.0 ), d_( 0.0 ) {}
    HalfSpace ( const Coord &n, double d ) : n_( n ), d_( d ) {}
    const Coord &innerNormal () const { return n_; }
    HyperPlane boundary () const { return { n_, d_ }; }
    Coord n_;
    double d_;
  };
  using ColorFunction = DataSet< double >;
  using FlagSet = DataSet< Flag >;
  using ReconstructionSet = DataSet< HalfSpace >;
#else
  using ColorFunction = Dune::VoF::ColorFunction< GV >;
  using FlagSet = Dune::VoF::FlagSet< GV >;
  using ReconstructionSet = Dune::VoF::ReconstructionSet< GV >;
#endif

  // rigid rotation about the z axis plus a drift, evaluated at the centre of the bound intersection
  struct RotationVelocity
  {
    const Dune::MeshIntersection3 *is = nullptr;
    void bind ( const Dune::MeshIntersection3 &i ) { is = &i; }
    void unbind () { is = nullptr; }
    Coord operator() ( const Dune::FieldVector< double, 2 > & ) const
    {
      const Coord x = is->geometry().center();
      Coord r;
      r[ 0 ] = x[ 1 ] - 0.5;
      r[ 1 ] = 0.5 - x[ 0 ];
      r[ 2 ] = 0.3;
      return r;
    }
  };

  int failures = 0;
  void expect ( bool ok, const char *what )
  {
    std::printf( "%s %s\n", ok ? "ok  " : "FAIL", what );
    if( !ok )
      ++failures;
  }
  bool sameBits ( const double *a, const double *b, std::size_t n )
  {
    for( std::size_t i = 0; i < n; ++i )
      if( std::memcmp( a + i, b + i, sizeof( double ) ) != 0 && !( a[ i ] == 0.0 && b[ i ] == 0.0 ) && !( std::isnan( a[ i ] ) && std::isnan( b[ i ] ) ) )
        return false;
    return true;
  }

  struct Arrays
  {
    std::vector< std::int32_t > cornerOffsets { 0 }, cornerVertex, faceOffsets { 0 }, faceNeighbor, faceCornerOffsets { 0 }, stencilOffsets { 0 }, stencil;
    std::vector< double > corners, centers, volumes, faceCorners, faceCenters, faceUnit, faceInt, faceVolumes;
    std::int64_t nVertices = 0;
  };

  using V3 = std::array< double, 3 >;
  V3 sub ( const V3 &a, const V3 &b ) { return { a[ 0 ] - b[ 0 ], a[ 1 ] - b[ 1 ], a[ 2 ] - b[ 2 ] }; }
  V3 cross ( const V3 &a, const V3 &b ) { return { a[ 1 ] * b[ 2 ] - a[ 2 ] * b[ 1 ], a[ 2 ] * b[ 0 ] - a[ 0 ] * b[ 2 ], a[ 0 ] * b[ 1 ] - a[ 1 ] * b[ 0 ] }; }
  double dot ( const V3 &a, const V3 &b ) { return a[ 0 ] * b[ 0 ] + a[ 1 ] * b[ 1 ] + a[ 2 ] * b[ 2 ]; }

  // jittered lattice, Kuhn's subdivision; DUNE tetrahedron: faces (0,1,2), (0,1,3), (0,2,3), (1,2,3)
  Arrays makeMesh ( int n )
  {
    Arrays A;
    std::uint64_t seed = 424242;
    auto rnd = [ &seed ] { seed = seed * 6364136223846793005ull + 1442695040888963407ull; return double( seed >> 11 ) / 9007199254740992.0; };
    const double h = 1.0 / n;
    auto vid = [ n ] ( int i, int j, int k ) { return ( k * ( n + 1 ) + j ) * ( n + 1 ) + i; };
    std::vector< V3 > v( std::size_t( n + 1 ) * ( n + 1 ) * ( n + 1 ) );
    for( int k = 0; k <= n; ++k )
      for( int j = 0; j <= n; ++j )
        for( int i = 0; i <= n; ++i )
        {
          V3 x = { i * h, j * h, k * h };
          if( i > 0 && i < n && j > 0 && j < n && k > 0 && k < n )
            for( int d = 0; d < 3; ++d )
              x[ d ] += ( rnd() - 0.5 ) * 0.4 * h;
          v[ vid( i, j, k ) ] = x;
        }
    A.nVertices = std::int64_t( v.size() );
    std::vector< std::array< int, 4 > > cells;
    const int path[ 6 ][ 2 ] = { { 1, 3 }, { 3, 2 }, { 2, 6 }, { 6, 4 }, { 4, 5 }, { 5, 1 } };
    for( int k = 0; k < n; ++k )
      for( int j = 0; j < n; ++j )
        for( int i = 0; i < n; ++i )
        {
          int c[ 8 ];
          for( int q = 0; q < 8; ++q )
            c[ q ] = vid( i + ( q & 1 ), j + ( ( q >> 1 ) & 1 ), k + ( ( q >> 2 ) & 1 ) );
          for( const auto &p : path )
          {
            std::array< int, 4 > t = { c[ 0 ], c[ p[ 0 ] ], c[ p[ 1 ] ], c[ 7 ] };
            if( dot( cross( sub( v[ t[ 1 ] ], v[ t[ 0 ] ] ), sub( v[ t[ 2 ] ], v[ t[ 0 ] ] ) ), sub( v[ t[ 3 ] ], v[ t[ 0 ] ] ) ) < 0 )
              std::swap( t[ 1 ], t[ 2 ] );     // positively oriented
            cells.push_back( t );
          }
        }
    const std::size_t N = cells.size();
    const int faces[ 4 ][ 3 ] = { { 0, 1, 2 }, { 0, 1, 3 }, { 0, 2, 3 }, { 1, 2, 3 } };
    std::map< std::array< int, 3 >, std::vector< int > > owners;
    for( std::size_t c = 0; c < N; ++c )
      for( const auto &f : faces )
      {
        std::array< int, 3 > key = { cells[ c ][ f[ 0 ] ], cells[ c ][ f[ 1 ] ], cells[ c ][ f[ 2 ] ] };
        std::sort( key.begin(), key.end() );
        owners[ key ].push_back( int( c ) );
      }
    std::vector< std::vector< int > > cellsOfVertex( v.size() );
    for( std::size_t c = 0; c < N; ++c )
    {
      V3 ctr = { 0, 0, 0 };
      for( int q = 0; q < 4; ++q )
      {
        const V3 &x = v[ cells[ c ][ q ] ];
        A.cornerVertex.push_back( cells[ c ][ q ] );
        for( int d = 0; d < 3; ++d )
        {
          A.corners.push_back( x[ d ] );
          ctr[ d ] += x[ d ];
        }
        cellsOfVertex[ cells[ c ][ q ] ].push_back( int( c ) );
      }
      for( int d = 0; d < 3; ++d )
      {
        ctr[ d ] /= 4;
        A.centers.push_back( ctr[ d ] );
      }
      const V3 &x0 = v[ cells[ c ][ 0 ] ];
      A.volumes.push_back( std::abs( dot( cross( sub( v[ cells[ c ][ 1 ] ], x0 ), sub( v[ cells[ c ][ 2 ] ], x0 ) ), sub( v[ cells[ c ][ 3 ] ], x0 ) ) ) / 6.0 );
      A.cornerOffsets.push_back( std::int32_t( A.cornerVertex.size() ) );
      for( const auto &f : faces )
      {
        std::array< int, 3 > ids = { cells[ c ][ f[ 0 ] ], cells[ c ][ f[ 1 ] ], cells[ c ][ f[ 2 ] ] };
        std::array< int, 3 > key = ids;
        std::sort( key.begin(), key.end() );
        int nb = -1;
        for( int o : owners[ key ] )
          if( o != int( c ) )
            nb = o;
        A.faceNeighbor.push_back( nb );
        V3 fc = { 0, 0, 0 };
        for( int q = 0; q < 3; ++q )
          for( int d = 0; d < 3; ++d )
          {
            A.faceCorners.push_back( v[ ids[ q ] ][ d ] );
            fc[ d ] += v[ ids[ q ] ][ d ];
          }
        A.faceCornerOffsets.push_back( A.faceCornerOffsets.back() + 3 );
        for( int d = 0; d < 3; ++d )
          fc[ d ] /= 3;
        V3 cr = cross( sub( v[ ids[ 1 ] ], v[ ids[ 0 ] ] ), sub( v[ ids[ 2 ] ], v[ ids[ 0 ] ] ) );
        const double area = 0.5 * std::sqrt( dot( cr, cr ) );
        V3 un = { 0.5 * cr[ 0 ] / area, 0.5 * cr[ 1 ] / area, 0.5 * cr[ 2 ] / area };
        if( dot( un, sub( fc, ctr ) ) < 0 )
          un = { -un[ 0 ], -un[ 1 ], -un[ 2 ] };
        for( int d = 0; d < 3; ++d )
        {
          A.faceCenters.push_back( fc[ d ] );
          A.faceUnit.push_back( un[ d ] );
          A.faceInt.push_back( un[ d ] * area );
        }
        A.faceVolumes.push_back( area );
      }
      A.faceOffsets.push_back( std::int32_t( A.faceNeighbor.size() ) );
    }
    for( std::size_t c = 0; c < N; ++c )
    {
      std::vector< int > nb;
      for( int q = 0; q < 4; ++q )
        for( int o : cellsOfVertex[ cells[ c ][ q ] ] )
          if( o != int( c ) )
            nb.push_back( o );
      std::sort( nb.begin(), nb.end() );
      nb.erase( std::unique( nb.begin(), nb.end() ), nb.end() );
      A.stencil.insert( A.stencil.end(), nb.begin(), nb.end() );
      A.stencilOffsets.push_back( std::int32_t( A.stencil.size() ) );
    }
    return A;
  }

} // namespace

int main ()
{
  const Arrays A = makeMesh( 6 );
  const std::size_t N = A.volumes.size(), F = A.faceVolumes.size();

  Dune::MeshGrid3 grid;
  grid.nCells = std::int64_t( N );
  grid.nVertices = A.nVertices;
  grid.cornerOffsets = A.cornerOffsets.data();
  grid.cornerVertex = A.cornerVertex.data();
  grid.corners = A.corners.data();
  grid.centers = A.centers.data();
  grid.volumes = A.volumes.data();
  grid.faceOffsets = A.faceOffsets.data();
  grid.faceNeighbor = A.faceNeighbor.data();
  grid.faceCornerOffsets = A.faceCornerOffsets.data();
  grid.faceCorners = A.faceCorners.data();
  grid.faceCenters = A.faceCenters.data();
  grid.faceUnitNormals = A.faceUnit.data();
  grid.faceIntNormals = A.faceInt.data();
  grid.faceVolumes = A.faceVolumes.data();
  GV gv( grid );

  vo_mesh om{};
  om.dim = 3;
  om.n_cells = std::int64_t( N );
  om.corner_offsets = A.cornerOffsets.data();
  om.corners = A.corners.data();
  om.centers = A.centers.data();
  om.volumes = A.volumes.data();
  om.face_offsets = A.faceOffsets.data();
  om.face_neighbor = A.faceNeighbor.data();
  om.face_corner_offsets = A.faceCornerOffsets.data();
  om.face_corners = A.faceCorners.data();
  om.face_centers = A.faceCenters.data();
  om.face_unit_normals = A.faceUnit.data();
  om.face_int_normals = A.faceInt.data();
  om.face_volumes = A.faceVolumes.data();
  om.stencil_offsets = A.stencilOffsets.data();
  om.stencil = A.stencil.data();

  // a ball, sampled at barycentric lattice points of every tetrahedron (only an input)
  std::vector< double > u0( N );
  for( std::size_t c = 0; c < N; ++c )
  {
    const double *x = A.corners.data() + 12 * c;
    int inside = 0, total = 0;
    const int M = 8;
    for( int a = 0; a <= M; ++a )
      for( int b = 0; a + b <= M; ++b )
        for( int g = 0; a + b + g <= M; ++g )
        {
          const double w[ 4 ] = { double( a ) / M, double( b ) / M, double( g ) / M, double( M - a - b - g ) / M };
          double p[ 3 ] = { 0, 0, 0 };
          for( int q = 0; q < 4; ++q )
            for( int d = 0; d < 3; ++d )
              p[ d ] += w[ q ] * x[ 3 * q + d ];
          ++total;
          inside += std::sqrt( ( p[ 0 ] - 0.5 ) * ( p[ 0 ] - 0.5 ) + ( p[ 1 ] - 0.55 ) * ( p[ 1 ] - 0.55 ) + ( p[ 2 ] - 0.45 ) * ( p[ 2 ] - 0.45 ) ) < 0.28;
        }
    u0[ c ] = double( inside ) / total;
  }

  // the sampled velocity the oracle gets: the same functor evaluated the same way
  std::vector< double > vel( 3 * F );
  for( std::size_t s = 0; s < F; ++s )
  {
    vel[ 3 * s ] = A.faceCenters[ 3 * s + 1 ] - 0.5;
    vel[ 3 * s + 1 ] = 0.5 - A.faceCenters[ 3 * s ];
    vel[ 3 * s + 2 ] = 0.3;
  }

  try
  {
    Dune::VoFx::Unstructured::VertexNeighborsStencil< GV > stencils( gv );
    const auto &ctx = stencils.context();
    expect( ctx.stencilOffsets == A.stencilOffsets && ctx.stencil == A.stencil, "flattened CSR vertex stencil equals the brute-force one" );
    expect( ctx.faceNeighbor == A.faceNeighbor && ctx.corners == A.corners && ctx.faceIntNormals == A.faceInt && ctx.faceCorners == A.faceCorners
            && ctx.faceCornerOffsets == A.faceCornerOffsets, "flattened geometry equals the grid's arrays" );

    Dune::VoFx::Unstructured::FlagOperator< GV > flagOperator( stencils, 1e-6 );
    Dune::VoFx::Unstructured::ModifiedYoungsReconstruction< GV > reconstruction( stencils );
    auto evolutionOperator = Dune::VoFx::Unstructured::evolution( stencils );

    ColorFunction uh( gv ), update( gv );
    FlagSet flags( gv );
    ReconstructionSet recs( gv );
    std::copy( u0.begin(), u0.end(), uh.begin() );

#ifdef VOFX_TEST_WITH_REFERENCE
    Dune::VoF::VertexNeighborsStencil< GV > refStencils( gv );
    Dune::VoF::FlagOperator< GV > refFlag( 1e-6 );
    Dune::VoF::ModifiedYoungsReconstruction< GV, Dune::VoF::VertexNeighborsStencil< GV > > refRecon( refStencils );
    auto refEvolution = Dune::VoF::evolution( gv );
    ColorFunction ruh( gv ), rupdate( gv );
    FlagSet rflags( gv );
    ReconstructionSet rrecs( gv );
    std::copy( u0.begin(), u0.end(), ruh.begin() );
#endif

    std::vector< double > uo( u0 ), hso( 4 * N ), updo( N );
    std::vector< std::uint8_t > flo( N );
    double dt = 0.0, dto = 0.0, rdt = 0.0;
    (void) rdt;
    bool flagsOk = true, recsOk = true, updOk = true, dtOk = true, refOk = true;
    std::size_t mixedSeen = 0;
    for( int pass = 0; pass < 6; ++pass )
    {
      RotationVelocity velocity;
      flagOperator( uh, flags );
      reconstruction( uh, recs, flags );
      const double dtEst = evolutionOperator( recs, flags, velocity, dt, update );
      for( std::size_t i = 0; i < N; ++i )
        uh[ i ] += 1.0 * update[ i ];
      dt = 0.5 * dtEst;

      double esto = 0.0;
      vo_mesh_flag( &om, uo.data(), 1e-6, flo.data() );
      vo_mesh_youngs( &om, uo.data(), flo.data(), hso.data() );
      vo_mesh_evolve( &om, hso.data(), flo.data(), vel.data(), dto, updo.data(), &esto );
      for( std::size_t i = 0; i < N; ++i )
        uo[ i ] += updo[ i ];
      dto = 0.5 * esto;

      std::size_t i = 0;
      for( auto it = flags.begin(); it != flags.end(); ++it, ++i )
      {
        flagsOk &= std::uint8_t( static_cast< std::size_t >( *it ) ) == flo[ i ];
        mixedSeen += ( flo[ i ] == 1 || flo[ i ] == 2 );
      }
      i = 0;
      for( auto it = recs.begin(); it != recs.end(); ++it, ++i )
      {
        const double hs[ 4 ] = { it->innerNormal()[ 0 ], it->innerNormal()[ 1 ], it->innerNormal()[ 2 ], it->boundary().distance() };
        recsOk &= sameBits( hs, hso.data() + 4 * i, 4 );
      }
      for( i = 0; i < N; ++i )
      {
        const double a = update[ i ];
        updOk &= sameBits( &a, &updo[ i ], 1 );
      }
      dtOk &= dtEst == esto;

#ifdef VOFX_TEST_WITH_REFERENCE
      RotationVelocity rvelocity;
      refFlag( ruh, rflags );
      refRecon( ruh, rrecs, rflags );
      const double rEst = refEvolution( rrecs, rflags, rvelocity, rdt, rupdate );
      ruh.axpy( 1.0, rupdate );
      rdt = 0.5 * rEst;
      refOk &= rEst == dtEst;
      {
        auto a = uh.begin();
        for( auto b = ruh.begin(); b != ruh.end(); ++a, ++b )
        {
          const double x = *a, y = *b;
          refOk &= sameBits( &x, &y, 1 );
        }
      }
#endif
    }
#ifdef VOFX_TEST_WITH_REFERENCE
    expect( refOk, "colour function and dtEst of every pass equal the REFERENCE's own operators on the same 3D grid view" );
#else
    (void) refOk;
#endif
    expect( mixedSeen > 500, "the band is not empty" );
    expect( flagsOk, "flags of 6 loop passes equal the oracle's" );
    expect( recsOk, "half-spaces of 6 loop passes equal the oracle's bit for bit" );
    expect( updOk, "updates of 6 loop passes equal the oracle's bit for bit" );
    expect( dtOk, "time-step estimates equal the oracle's" );
    bool colorOk = true;
    double mass0 = 0.0, mass1 = 0.0;
    for( std::size_t i = 0; i < N; ++i )
    {
      const double a = uh[ i ];
      colorOk &= sameBits( &a, &uo[ i ], 1 );
      mass0 += u0[ i ] * A.volumes[ i ];
      mass1 += a * A.volumes[ i ];
    }
    expect( colorOk, "colour function after 6 passes equals the oracle's bit for bit" );
    expect( std::abs( mass1 - mass0 ) < 1e-14, "mass is conserved to round-off" );
  }
  catch( const std::exception &e )
  {
    std::printf( "FAIL exception: %s\n", e.what() );
    ++failures;
  }
  std::printf( "%s (%d failures)\n", failures ? "ADAPTOR MESH3 TEST FAILED" : "ADAPTOR MESH3 TEST PASSED", failures );
  return failures ? 1 : 0;
}
