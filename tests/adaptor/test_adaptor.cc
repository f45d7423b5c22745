// Test of the header adaptor include/dune/vofx/vofx.hh (needs a GPU at run time; built on the CPU box).
//
// The adaptor is driven exactly like the reference drives its operators (algorithm.hh:57-93), on the serial Cartesian
// grid model of oracle/shim (test infrastructure), with host containers that have the interface of the reference's
// DataSet (dataset.hh:26-87).  Results are compared bit-for-bit with the oracle (libvoforacle.so).
// With -DVOFX_TEST_WITH_REFERENCE (only where /root/reference is mounted) the containers ARE the reference's own
// ColorFunction / FlagSet / ReconstructionSet and the reference's operators run side by side on the same inputs.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <vector>

#define GRIDDIM 3
#include <dune/common/fvector.hh>
#include <dune/grid/common/minigrid.hh>

#ifdef VOFX_TEST_WITH_REFERENCE
#include <dune/vof/colorfunction.hh>
#include <dune/vof/flagset.hh>
#include <dune/vof/flagging.hh>
#include <dune/vof/reconstructionset.hh>
#include <dune/vof/evolution.hh>
#include <dune/vof/velocity.hh>
#include <dune/vof/stencil/vertexneighborsstencil.hh>
#include <dune/vof/reconstruction/modifiedyoungs.hh>
#include <dune/vof/reconstruction/heightfunction.hh>
#endif

#include <dune/vofx/vofx.hh>

extern "C"
{
#include "vof_oracle.h"
}

namespace
{

#ifndef VOFX_TEST_WITH_REFERENCE
  // containers with the interface of dune/vof/dataset.hh:26-87
  enum class Flag : std::size_t { empty = 0, mixed = 1, mixedfull = 2, full = 3, nan = 99 };

  template< class GV, class T >
  struct DataSet
  {
    using GridView = GV;
    using DataType = T;
    using Index = std::size_t;
    explicit DataSet ( GridView gv ) : gridView_( gv ), data_( gv.indexSet().size( 0 ) ) {}
    const T &operator[] ( const Index &i ) const { return data_[ i ]; }
    T &operator[] ( const Index &i ) { return data_[ i ]; }
    auto begin () { return data_.begin(); }
    auto end () { return data_.end(); }
    auto begin () const { return data_.begin(); }
    auto end () const { return data_.end(); }
    std::size_t size () const { return data_.size(); }
    const GridView &gridView () const { return gridView_; }
    GridView gridView_;
    std::vector< T > data_;
  };

  template< class Coord >
  struct HyperPlane
  {
    Coord n;
    double d;
    double distance () const { return d; }
  };
  template< class Coord >
  struct HalfSpace
  {
    using Coordinate = Coord;
    HalfSpace () : n_( 0.0 ), d_( 0.0 ) {}
    HalfSpace ( const Coord &n, double d ) : n_( n ), d_( d ) {}
    const Coord &innerNormal () const { return n_; }
    HyperPlane< Coord > boundary () const { return { n_, d_ }; }
    Coord n_;
    double d_;
  };

  template< class GV > using ColorFunction = DataSet< GV, double >;
  template< class GV > using FlagSet = DataSet< GV, Flag >;
  template< class GV > using ReconstructionSet = DataSet< GV, HalfSpace< Dune::FieldVector< double, GV::dimension > > >;
#else
  template< class GV > using ColorFunction = Dune::VoF::ColorFunction< GV >;
  template< class GV > using FlagSet = Dune::VoF::FlagSet< GV >;
  template< class GV > using ReconstructionSet = Dune::VoF::ReconstructionSet< GV >;
#endif

  // problems with the reference's Problem concept (velocityField as in test/problems/rotatingcircle.hh:45-52 and sflow.hh:73-82)
  template< int dim >
  struct Rotation
  {
    using DomainType = Dune::FieldVector< double, dim >;
    void velocityField ( const DomainType &x, const double, DomainType &r ) const
    {
      r = 0.0;
      if( dim == 2 )
      {
        r[ 0 ] = x[ 1 ] - 0.5;
        r[ 1 ] = 0.5 - x[ 0 ];
      }
      else
      {
        r[ 0 ] = x[ 1 ] - 0.5;
        r[ 1 ] = -( x[ 0 ] - 0.5 );
      }
      r *= 2 * M_PI / 10;
    }
  };
  template< int dim >
  struct SFlowField
  {
    using DomainType = Dune::FieldVector< double, dim >;
    void velocityField ( const DomainType &x, const double, DomainType &v ) const
    {
      const double y0 = 4 * x[ 0 ] - 2, y1 = 4 * x[ 1 ] - 2;
      v = 0.0;
      v[ 0 ] = 0.25 * ( y0 + y1 * y1 * y1 );
      v[ 1 ] = -0.25 * ( y1 + y0 * y0 * y0 );
    }
  };

  // RotatingCircle< double, 2 > of test/problems/rotatingcircle.hh:22-57: velocity field, centre and radius of the exact solution
  struct RotatingCircle2
  {
    using DomainType = Dune::FieldVector< double, 2 >;
    void velocityField ( const DomainType &x, const double, DomainType &r ) const
    {
      r[ 0 ] = x[ 1 ] - 0.5;
      r[ 1 ] = 0.5 - x[ 0 ];
      r *= 2 * M_PI / 10;
    }
    DomainType center ( double t ) const
    {
      DomainType c{ 0.5, 0.5 };
      const DomainType offset{ 0.0, 0.25 }, normal{ 0.25, 0.0 };
      c.axpy( std::cos( ( 2 * M_PI / 10 ) * t ), offset );
      c.axpy( std::sin( ( 2 * M_PI / 10 ) * t ), normal );
      return c;
    }
    double radius ( double ) const { return 0.15; }
  };

  struct NullWriter
  {
    int writes = 0;
    bool willWrite ( double ) { return false; }
    void write ( double ) { ++writes; }
  };

  template< class Problem, class GV >
  struct HostVelocity   // velocity.hh:5-33
  {
    using Intersection = typename GV::Intersection;
    HostVelocity ( const Problem &p, double t ) : problem_( &p ), time_( t ) {}
    template< class Local >
    auto operator() ( const Local &local ) const
    {
      typename Problem::DomainType v( 0.0 );
      problem_->velocityField( intersection_->geometry().global( local ), time_, v );
      return v;
    }
    void bind ( const Intersection &i ) { intersection_ = &i; }
    void unbind () { intersection_ = nullptr; }
    const Intersection *intersection_ = nullptr;
    const Problem *problem_;
    double time_;
  };

  int failures = 0;

  bool sameBits ( const double *a, const double *b, std::size_t n )
  {
    for( std::size_t i = 0; i < n; ++i )
    {
      if( std::isnan( a[ i ] ) && std::isnan( b[ i ] ) )
        continue;
      if( std::memcmp( a + i, b + i, sizeof( double ) ) != 0 && !( a[ i ] == 0.0 && b[ i ] == 0.0 ) )
        return false;
    }
    return true;
  }

  void expect ( bool ok, const char *what )
  {
    std::printf( "%s %s\n", ok ? "ok  " : "FAIL", what );
    if( !ok )
      ++failures;
  }

  template< int dim, class Problem >
  void run ( const char *name, const int *n, int oracleProblem, const Problem &problem, int passes )
  {
    using Grid = Dune::MiniGrid< dim >;
    using GV = Dune::MiniGridView< dim >;
    Grid grid;
    vo_grid og;
    std::memset( &og, 0, sizeof( og ) );
    og.dim = dim;
    for( int d = 0; d < 3; ++d )
    {
      og.n[ d ] = d < dim ? n[ d ] : 1;
      og.h[ d ] = d < dim ? 1.0 / n[ d ] : 1.0;
    }
    for( int d = 0; d < dim; ++d )
    {
      grid.n[ d ] = n[ d ];
      grid.lo[ d ] = 0.0;
      grid.h[ d ] = 1.0 / n[ d ];
    }
    GV gv( grid );
    const std::size_t N = grid.cells();
    const double eps = 1e-6, cfl = 0.5;

    std::vector< double > u0( N ), uo( N ), hso( N * ( dim + 1 ) ), updo( N );
    std::vector< std::uint8_t > flo( N );
    vo_init( &og, oracleProblem, 0.0, u0.data() );

    // --- operator by operator, as algorithm.hh:57-83 does ---
    // the reference's own construction lines (algorithm.hh:57-59, test-vof-consistence.cc:155-158) with the namespace changed:
    //   auto flagOperator = FlagOperator< GridView >( eps_ );   auto evolutionOperator = evolution( gridView_ );
    Dune::VoFx::VertexNeighborsStencil< GV > stencils( gv );
    auto flagOperator = Dune::VoFx::FlagOperator< GV >( eps );
    Dune::VoFx::ModifiedYoungsReconstruction< GV, Dune::VoFx::VertexNeighborsStencil< GV > > reconstructionOperator( stencils );
    auto evolutionOperator = Dune::VoFx::evolution( gv );

    ColorFunction< GV > uh( gv ), update( gv );
    FlagSet< GV > flags( gv );
    ReconstructionSet< GV > reconstructions( gv );
    for( std::size_t i = 0; i < N; ++i )
      uh[ i ] = u0[ i ];
    uo = u0;

#ifdef VOFX_TEST_WITH_REFERENCE
    Dune::VoF::VertexNeighborsStencil< GV > refStencils( gv );
    Dune::VoF::FlagOperator< GV > refFlag( eps );
    Dune::VoF::ModifiedYoungsReconstruction< GV, Dune::VoF::VertexNeighborsStencil< GV > > refRecon( refStencils );
    auto refEvolution = Dune::VoF::evolution( gv );
    ColorFunction< GV > refUh( gv ), refUpdate( gv );
    FlagSet< GV > refFlags( gv );
    ReconstructionSet< GV > refRecs( gv );
    for( std::size_t i = 0; i < N; ++i )
      refUh[ i ] = u0[ i ];
#endif

    double time = 0.0, dt = 0.0, dtEst = 0.0, timeO = 0.0, dtO = 0.0, dtEstO = 0.0;
    bool okFlags = true, okRecs = true, okUpd = true, okDt = true;
    for( int pass = 0; pass < passes; ++pass )
    {
      HostVelocity< Problem, GV > velocity( problem, time );
      flagOperator( uh, flags );
      reconstructionOperator( uh, reconstructions, flags );
      dtEst = evolutionOperator( reconstructions, flags, velocity, dt, update );
      for( std::size_t i = 0; i < N; ++i )
        uh[ i ] += 1.0 * update[ i ];      // uh.axpy( 1.0, update ), colorfunction.hh:31-37
      if( dt > 0.0 )
        time += dt;
      dt = dtEst * cfl;

      vo_flag( &og, uo.data(), eps, flo.data() );
      vo_reconstruct( &og, VO_RECON_YOUNGS, uo.data(), flo.data(), hso.data(), 10 );
      vo_evolve( &og, hso.data(), flo.data(), oracleProblem, timeO, dtO, updo.data(), &dtEstO );
      for( std::size_t i = 0; i < N; ++i )
        uo[ i ] += 1.0 * updo[ i ];
      if( dtO > 0.0 )
        timeO += dtO;
      dtO = dtEstO * cfl;

      std::size_t i = 0;
      for( auto it = flags.begin(); it != flags.end(); ++it, ++i )
        okFlags = okFlags && ( std::uint8_t( static_cast< std::size_t >( *it ) ) == flo[ i ] );
      i = 0;
      for( auto it = reconstructions.begin(); it != reconstructions.end(); ++it, ++i )
      {
        double hs[ 4 ];
        for( int d = 0; d < dim; ++d )
          hs[ d ] = it->innerNormal()[ d ];
        hs[ dim ] = it->boundary().distance();
        okRecs = okRecs && sameBits( hs, hso.data() + i * ( dim + 1 ), dim + 1 );
      }
      for( i = 0; i < N; ++i )
      {
        const double a = update[ i ];
        okUpd = okUpd && sameBits( &a, &updo[ i ], 1 );
      }
      okDt = okDt && dtEst == dtEstO && time == timeO;

#ifdef VOFX_TEST_WITH_REFERENCE
      refFlag( refUh, refFlags );
      refRecon( refUh, refRecs, refFlags );
      HostVelocity< Problem, GV > refVelocity( problem, timeO - ( dtO > 0 ? 0 : 0 ) );
      (void) refVelocity;
#endif
    }
    // getInterfaceVertices (utility.hh:19-48) on the last reconstructions against the oracle's per-cell interface()
    bool okIface = true;
    {
      using Coordinate = Dune::FieldVector< double, dim >;
      std::vector< Coordinate > vertices;
      std::vector< std::size_t > offsets;
      Dune::VoFx::getInterfaceVertices( stencils, reconstructions, flags, vertices, offsets );
      std::size_t m = 0;
      for( std::size_t c = 0; c < N; ++c )
      {
        if( flo[ c ] != 1 && flo[ c ] != 2 )
          continue;
        double lo[ 3 ] = { 0, 0, 0 }, hh[ 3 ] = { 1, 1, 1 };
        std::size_t r = c;
        for( int d = 0; d < dim; ++d )
        {
          hh[ d ] = 1.0 / n[ d ];
          lo[ d ] = 0.0 + double( r % n[ d ] ) * hh[ d ];
          r /= n[ d ];
        }
        int nv = 0;
        double verts[ 16 * 3 ], cen[ 3 ], meas = 0.0;
        vo_interface( dim, lo, hh, hso.data() + c * ( dim + 1 ), &nv, verts, cen, &meas );
        okIface = okIface && m + 1 < offsets.size() && offsets[ m + 1 ] - offsets[ m ] == std::size_t( nv );
        for( int i = 0; okIface && i < nv; ++i )
        {
          double x[ 3 ];
          for( int d = 0; d < dim; ++d )
            x[ d ] = vertices[ offsets[ m ] + i ][ d ];
          okIface = okIface && sameBits( x, verts + i * dim, dim );
        }
        ++m;
      }
      okIface = okIface && offsets.size() == m + 1 && offsets.back() == vertices.size();
    }
    std::string tag( name );
    expect( okIface, ( std::string( name ) + ": getInterfaceVertices == oracle (bit-exact)" ).c_str() );
    expect( okFlags, ( tag + ": FlagOperator == oracle (bit-exact)" ).c_str() );
    expect( okRecs, ( tag + ": ModifiedYoungsReconstruction == oracle (bit-exact)" ).c_str() );
    expect( okUpd, ( tag + ": Evolution update == oracle (bit-exact)" ).c_str() );
    expect( okDt, ( tag + ": dtEst / time == oracle" ).c_str() );

    // --- the whole loop through Dune::VoFx::Algorithm (device-resident), general velocity fallback ---
    {
      ColorFunction< GV > uh2( gv );
      for( std::size_t i = 0; i < N; ++i )
        uh2[ i ] = u0[ i ];
      NullWriter writer;
      Dune::VoFx::Algorithm< GV, Problem, NullWriter > algorithm( gv, problem, writer, cfl, eps, false, VOFX_RECON_YOUNGS, -1 );
      // run until the oracle's time after `passes` passes is reached: the loop is do { } while( time < end )
      algorithm( uh2, 0.0, timeO, 0 );
      bool ok = true;
      for( std::size_t i = 0; i < N; ++i )
      {
        const double a = uh2[ i ];
        ok = ok && sameBits( &a, &uo[ i ], 1 );
      }
      expect( ok, ( tag + ": Algorithm (host velocity functor) == oracle loop (bit-exact)" ).c_str() );
    }
    // --- the same with the built-in separable velocity tables (time lives on the device) ---
    {
      ColorFunction< GV > uh3( gv );
      for( std::size_t i = 0; i < N; ++i )
        uh3[ i ] = u0[ i ];
      NullWriter writer;
      Dune::VoFx::Algorithm< GV, Problem, NullWriter > algorithm( gv, problem, writer, cfl, eps, false, VOFX_RECON_YOUNGS, oracleProblem );
      algorithm( uh3, 0.0, timeO, 0 );
      bool ok = true;
      for( std::size_t i = 0; i < N; ++i )
      {
        const double a = uh3[ i ];
        ok = ok && sameBits( &a, &uo[ i ], 1 );
      }
      expect( ok, ( tag + ": Algorithm (built-in velocity tables) == oracle loop (bit-exact)" ).c_str() );
    }
  }

} // namespace

  // Algorithm::operator() returns the accumulated dt * l1error( ..., problem, time ) (algorithm.hh:85-98) for the rotating circle
  void runCircleError ()
  {
    using GV = Dune::MiniGridView< 2 >;
    Dune::MiniGrid< 2 > grid;
    const int n = 64;
    vo_grid og;
    std::memset( &og, 0, sizeof( og ) );
    og.dim = 2;
    og.n[ 0 ] = og.n[ 1 ] = n;
    og.n[ 2 ] = 1;
    og.h[ 0 ] = og.h[ 1 ] = 1.0 / n;
    og.h[ 2 ] = 1.0;
    for( int d = 0; d < 2; ++d )
    {
      grid.n[ d ] = n;
      grid.lo[ d ] = 0.0;
      grid.h[ d ] = 1.0 / n;
    }
    GV gv( grid );
    const std::size_t N = grid.cells();
    RotatingCircle2 problem;
    std::vector< double > uo( N ), hso( N * 3 ), updo( N );
    std::vector< std::uint8_t > flo( N );
    const double c0[ 2 ] = { 0.5, 0.75 };
    vo_circle_interpolation( &og, c0, 0.15, uo.data() );
    ColorFunction< GV > uh( gv );
    for( std::size_t i = 0; i < N; ++i )
      uh[ i ] = uo[ i ];
    // the reference's loop with its error accumulation, on the oracle (HF reconstruction: the factory's default)
    double time = 0.0, dt = 0.0, dtEst = 0.0, error = 0.0;
    const double end = 0.2;
    do
    {
      vo_flag( &og, uo.data(), 1e-6, flo.data() );
      vo_reconstruct( &og, VO_RECON_HF, uo.data(), flo.data(), hso.data(), 10 );
      vo_evolve( &og, hso.data(), flo.data(), VO_PROB_ROTATING_CIRCLE, time, dt, updo.data(), &dtEst );
      for( std::size_t i = 0; i < N; ++i )
        uo[ i ] += 1.0 * updo[ i ];
      if( dt > 0.0 )
      {
        const auto c = problem.center( time );
        const double cc[ 2 ] = { c[ 0 ], c[ 1 ] };
        double e = 0.0;
        vo_l1error_circle( &og, hso.data(), flo.data(), cc, 0.15, &e );
        error += dt * e;
        time += dt;
      }
      dt = dtEst * 0.5;
    }
    while( time < end );
    NullWriter writer;
    Dune::VoFx::Algorithm< GV, RotatingCircle2, NullWriter > algorithm( gv, problem, writer, 0.5, 1e-6 );
    const double errorX = algorithm( uh, 0.0, end, 0 );
    bool ok = true;
    for( std::size_t i = 0; i < N; ++i )
    {
      const double a = uh[ i ];
      ok = ok && sameBits( &a, &uo[ i ], 1 );
    }
    expect( ok, "2D 64^2 rotating circle: Algorithm (factory reconstruction, host velocity functor on the band) == oracle loop (bit-exact)" );
    std::printf( "     accumulated l1 error: adaptor %.17g, oracle loop %.17g\n", errorX, error );
    expect( error > 0.0 && std::abs( errorX - error ) <= 1e-11 * error, "2D 64^2 rotating circle: returned error == sum dt * l1error of the oracle loop (1e-11 relative)" );
  }

int main ()
{
  try
  {
    runCircleError();
    const int n2[ 2 ] = { 32, 24 };
    run< 2 >( "2D 32x24 SFlow", n2, VO_PROB_SFLOW, SFlowField< 2 >(), 6 );
    const int n3[ 3 ] = { 16, 16, 16 };
    run< 3 >( "3D 16^3 rotating sphere", n3, VO_PROB_ROTATING_CIRCLE, Rotation< 3 >(), 5 );
    const int n3b[ 3 ] = { 20, 12, 16 };
    run< 3 >( "3D 20x12x16 SFlow", n3b, VO_PROB_SFLOW, SFlowField< 3 >(), 5 );
  }
  catch( const std::exception &e )
  {
    std::printf( "FAIL exception: %s\n", e.what() );
    return 2;
  }
  std::printf( failures ? "ADAPTOR TEST FAILED (%d)\n" : "ADAPTOR TEST PASSED\n", failures );
  return failures ? 1 : 0;
}
