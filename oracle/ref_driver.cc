// TEST INFRASTRUCTURE ONLY (oracle/).  Not part of the product; the product never loads this.
//
// C-ABI harness around the UNMODIFIED reference headers (/root/reference/dune/vof/**), instantiated on the
// serial Cartesian grid model in oracle/shim.  Built by oracle/Makefile into oracle/_ref/libvofref.so
// (git-ignored).  It is used (a) to pin the plain-C restatement oracle/vof_oracle.c bit-for-bit, (b) to generate
// the golden vectors under tests/golden, and (c) as the "reference" CPU baseline of bench.py.
//
// The time loop replays dune/vof/algorithm.hh:68-95 without l1error and without output (algorithm.hh itself
// cannot be included: it pulls test/errors.hh and with it dune-geometry quadrature rules).
//
// Build: g++ -std=c++17 -O2 -ffp-contract=off -fno-access-control -DNDEBUG -Ioracle/shim -I/root/reference
//   -fno-access-control: __impl::Edge/Face befriend "class Polyhedron" from inside namespace __impl
//                        (3d/polyhedron.hh:23-24,117), which GCC 13 resolves to a different class.
//   -DNDEBUG:            the reference's asserts abort the process; tests drive this from python.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iostream>
#include <limits>
#include <memory>
#include <tuple>
#include <vector>

#define GRIDDIM 3   // only tested for "== 1" by reconstruction/modifiedswartz.hh:74

#include <dune/common/exceptions.hh>
#include <dune/common/fmatrix.hh>
#include <dune/common/fvector.hh>
#include <dune/grid/common/minigrid.hh>

#include <dune/vof/colorfunction.hh>
#include <dune/vof/curvatureset.hh>
#include <dune/vof/flagset.hh>
#include <dune/vof/flagging.hh>
#include <dune/vof/reconstructionset.hh>
#include <dune/vof/evolution.hh>
#include <dune/vof/geometry/algorithm.hh>
#include <dune/vof/geometry/intersect.hh>
#include <dune/vof/geometry/polytope.hh>
#include <dune/vof/geometry/upwindpolygon.hh>
#include <dune/vof/utility.hh>
#include <dune/vof/velocity.hh>
#include <dune/vof/stencil/vertexneighborsstencil.hh>
#include <dune/vof/reconstruction/modifiedyoungs.hh>
#include <dune/vof/reconstruction/heightfunction.hh>
#include <dune/vof/reconstruction/modifiedswartz.hh>
#include <dune/vof/curvature/cartesianheightfunctioncurvature.hh>
#include <dune/vof/curvature/swartzcurvature.hh>
#include <dune/vof/interpolation/circleinterpolation.hh>
#include <dune/vof/interpolation/recursiveinterpolationcube.hh>
#include <dune/vof/test/problems/rotatingcircle.hh>
#include <dune/vof/test/problems/sflow.hh>
#include <dune/vof/test/problems/slottedcylinder.hh>
#include <dune/vof/test/problems/linearwall.hh>

#include "problems_extra.hh"

namespace
{

  enum { RECON_YOUNGS = 0, RECON_HF = 1, RECON_SWARTZ = 2 };
  enum { PROB_ROTATING_CIRCLE = 0, PROB_SFLOW = 1, PROB_SLOTTED_CYLINDER = 2, PROB_LINEAR_WALL = 3, PROB_LEVEQUE = 4 };

  thread_local std::string lastError;

  struct CtxBase
  {
    virtual ~CtxBase () = default;
    int dimension_;
  };

  template< int dim >
  struct Ctx : CtxBase
  {
    using Grid = Dune::MiniGrid< dim >;
    using GridView = Dune::MiniGridView< dim >;
    using Coord = Dune::FieldVector< double, dim >;
    using Stencils = Dune::VoF::VertexNeighborsStencil< GridView >;
    using Color = Dune::VoF::ColorFunction< GridView >;
    using Flags = Dune::VoF::FlagSet< GridView >;
    using Recs = Dune::VoF::ReconstructionSet< GridView >;
    using Curv = Dune::VoF::CurvatureSet< GridView >;
    using Youngs = Dune::VoF::ModifiedYoungsReconstruction< GridView, Stencils >;
    using HF = Dune::VoF::HeightFunctionReconstruction< GridView, Stencils, Youngs >;
    using Swartz = Dune::VoF::ModifiedSwartzReconstruction< GridView, Stencils, Youngs >;
    using HFCurv = Dune::VoF::CartesianHeightFunctionCurvature< GridView, Stencils >;

    Grid grid;
    GridView gv;
    std::unique_ptr< Stencils > stencils;   // built lazily: >= 26 entities per cell

    Ctx ( const int *n, const double *lo, const double *h ) : grid( makeGrid( n, lo, h ) ), gv( grid ) { this->dimension_ = dim; }

    static Grid makeGrid ( const int *n, const double *lo, const double *h )
    {
      Grid g;
      for( int d = 0; d < dim; ++d )
      {
        g.n[ d ] = n[ d ];
        g.lo[ d ] = lo[ d ];
        g.h[ d ] = h[ d ];
      }
      return g;
    }

    const Stencils &vertexStencils ()
    {
      if( !stencils )
        stencils = std::make_unique< Stencils >( gv );
      return *stencils;
    }

    std::size_t cells () const { return grid.cells(); }

    void loadColor ( const double *u, Color &c ) const { std::copy( u, u + cells(), c.begin() ); }
    void loadFlags ( const std::uint8_t *f, Flags &flags ) const
    {
      auto it = flags.begin();
      for( std::size_t i = 0; i < cells(); ++i, ++it )
        *it = static_cast< Dune::VoF::Flag >( static_cast< std::size_t >( f[ i ] ) );
    }
    void storeFlags ( const Flags &flags, std::uint8_t *f ) const
    {
      auto it = static_cast< const Dune::VoF::DataSet< GridView, Dune::VoF::Flag > & >( flags ).begin();
      for( std::size_t i = 0; i < cells(); ++i, ++it )
        f[ i ] = static_cast< std::uint8_t >( static_cast< std::size_t >( *it ) );
    }
    void loadRecs ( const double *hs, Recs &recs ) const
    {
      auto it = recs.begin();
      for( std::size_t i = 0; i < cells(); ++i, ++it )
      {
        Coord nrm;
        for( int d = 0; d < dim; ++d )
          nrm[ d ] = hs[ i * ( dim + 1 ) + d ];
        *it = Dune::VoF::HalfSpace< Coord >( nrm, hs[ i * ( dim + 1 ) + dim ] );
      }
    }
    void storeRecs ( const Recs &recs, double *hs ) const
    {
      auto it = recs.begin();
      for( std::size_t i = 0; i < cells(); ++i, ++it )
      {
        for( int d = 0; d < dim; ++d )
          hs[ i * ( dim + 1 ) + d ] = it->innerNormal()[ d ];
        hs[ i * ( dim + 1 ) + dim ] = it->boundary().distance();
      }
    }
  };

  // dispatch a generic callable on the problem selected at run time
  template< int dim, class F >
  int withProblem ( int problem, F &&f )
  {
    switch( problem )
    {
    case PROB_ROTATING_CIRCLE: { RotatingCircle< double, dim > p; f( p ); return 0; }
    case PROB_SFLOW: { SFlow< double, dim > p; f( p ); return 0; }
    case PROB_LINEAR_WALL: { LinearWall< double, dim > p; f( p ); return 0; }
    case PROB_LEVEQUE: { LeVeque< double, dim > p; f( p ); return 0; }
    case PROB_SLOTTED_CYLINDER:
      if constexpr ( dim == 2 ) { SlottedCylinder< double, dim > p; f( p ); return 0; }
      [[fallthrough]];
    default:
      lastError = "unknown problem id for this dimension";
      return 1;
    }
  }

  // initial data, mirroring the dispatch of test/average.hh:28-101
  template< int dim, class Problem >
  void initialData ( Ctx< dim > &ctx, const Problem &problem, double t, typename Ctx< dim >::Color &uh )
  {
    using GridView = typename Ctx< dim >::GridView;
    if constexpr ( std::is_same< Problem, RotatingCircle< double, 2 > >::value )
      Dune::VoF::circleInterpolation( problem.center( t ), problem.radius( t ), uh );   // average.hh:82-101
    else
    {
      // average.hh:36-54
      auto indicator = [ &problem, t ] ( const auto &x ) { Dune::FieldVector< double, 1 > u; problem.evaluate( x, t, u ); return u; };
      Dune::VoF::RecursiveInterpolationCube< GridView > interpolation( ctx.gv, 3 );
      interpolation( indicator, uh );
    }
  }

  template< int dim, class Recon, class Problem >
  void runLoop ( Ctx< dim > &ctx, const Recon &reconstructionOperator, const Problem &problem, double eps, double cfl,
                 int passes, double *u, double *timeIO, double *dtIO, double *phaseSeconds, std::int64_t *mixedCellSteps )
  {
    using C = Ctx< dim >;
    typename C::Color uh( ctx.gv ), update( ctx.gv );
    typename C::Flags flags( ctx.gv );
    typename C::Recs recs( ctx.gv );
    ctx.loadColor( u, uh );

    auto flagOperator = Dune::VoF::FlagOperator< typename C::GridView >( eps );
    auto evolutionOperator = Dune::VoF::evolution( ctx.gv );

    double time = *timeIO, dt = *dtIO, dtEst = 0.0;
    using clock = std::chrono::steady_clock;
    auto seconds = [] ( clock::time_point a, clock::time_point b ) { return std::chrono::duration< double >( b - a ).count(); };

    // algorithm.hh:68-95
    for( int pass = 0; pass < passes; ++pass )
    {
      Velocity< Problem, typename C::GridView > velocity( problem, time );

      const auto t0 = clock::now();
      flagOperator( uh, flags );
      const auto t1 = clock::now();
      reconstructionOperator( uh, recs, flags );
      const auto t2 = clock::now();
      dtEst = evolutionOperator( recs, flags, velocity, dt, update );
      const auto t3 = clock::now();
      uh.axpy( 1.0, update );
      const auto t4 = clock::now();

      if( phaseSeconds )
      {
        phaseSeconds[ 0 ] += seconds( t0, t1 );
        phaseSeconds[ 1 ] += seconds( t1, t2 );
        phaseSeconds[ 2 ] += seconds( t2, t3 );
        phaseSeconds[ 3 ] += seconds( t3, t4 );
      }
      if( mixedCellSteps )
        for( const auto &entity : elements( ctx.gv ) )
          if( flags.isMixed( entity ) )
            ++*mixedCellSteps;

      if( dt > 0.0 )
        time += dt;
      dt = dtEst * cfl;
    }

    std::copy( uh.begin(), uh.end(), u );
    *timeIO = time;
    *dtIO = dt;
  }

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

  template< class F >
  int withCtx ( void *handle, F &&f )
  {
    CtxBase *base = static_cast< CtxBase * >( handle );
    if( base->dimension_ == 2 )
      return f( *static_cast< Ctx< 2 > * >( base ) );
    if( base->dimension_ == 3 )
      return f( *static_cast< Ctx< 3 > * >( base ) );
    lastError = "unsupported dimension";
    return 1;
  }

} // namespace


extern "C"
{

  const char *ref_last_error () { return lastError.c_str(); }

  void *ref_create ( int dim, const int *n, const double *lo, const double *h )
  {
    if( dim == 2 )
      return static_cast< CtxBase * >( new Ctx< 2 >( n, lo, h ) );
    if( dim == 3 )
      return static_cast< CtxBase * >( new Ctx< 3 >( n, lo, h ) );
    lastError = "unsupported dimension";
    return nullptr;
  }

  void ref_destroy ( void *handle ) { delete static_cast< CtxBase * >( handle ); }

  // flagging.hh:35-70
  int ref_flag ( void *handle, const double *u, double eps, std::uint8_t *flagsOut )
  {
    return guarded( [ & ] { return withCtx( handle, [ & ] ( auto &ctx ) {
      using C = std::decay_t< decltype( ctx ) >;
      typename C::Color uh( ctx.gv );
      typename C::Flags flags( ctx.gv );
      ctx.loadColor( u, uh );
      Dune::VoF::FlagOperator< typename C::GridView > flagOperator( eps );
      flagOperator( uh, flags );
      ctx.storeFlags( flags, flagsOut );
      return 0;
    } ); } );
  }

  // modifiedyoungs.hh:63-76 / heightfunction.hh:70-87 / modifiedswartz.hh:70-87
  int ref_reconstruct ( void *handle, int kind, const double *u, const std::uint8_t *flagsIn, double *hsOut, int maxIterations )
  {
    return guarded( [ & ] { return withCtx( handle, [ & ] ( auto &ctx ) {
      using C = std::decay_t< decltype( ctx ) >;
      typename C::Color uh( ctx.gv );
      typename C::Flags flags( ctx.gv );
      typename C::Recs recs( ctx.gv );
      ctx.loadColor( u, uh );
      ctx.loadFlags( flagsIn, flags );
      const auto &stencils = ctx.vertexStencils();
      switch( kind )
      {
      case RECON_YOUNGS: { typename C::Youngs op( stencils ); op( uh, recs, flags ); break; }
      case RECON_HF: { typename C::HF op( stencils ); op( uh, recs, flags ); break; }
      case RECON_SWARTZ: { typename C::Swartz op( stencils, std::size_t( maxIterations ) ); op( uh, recs, flags ); break; }
      default: lastError = "unknown reconstruction kind"; return 1;
      }
      ctx.storeRecs( recs, hsOut );
      return 0;
    } ); } );
  }

  // evolution/evolution.hh:58-76
  int ref_evolve ( void *handle, const double *hsIn, const std::uint8_t *flagsIn, int problem, double time, double dt,
                   double *updateOut, double *dtEstOut )
  {
    return guarded( [ & ] { return withCtx( handle, [ & ] ( auto &ctx ) {
      using C = std::decay_t< decltype( ctx ) >;
      constexpr int dim = C::GridView::dimension;
      typename C::Flags flags( ctx.gv );
      typename C::Recs recs( ctx.gv );
      typename C::Color update( ctx.gv );
      ctx.loadFlags( flagsIn, flags );
      ctx.loadRecs( hsIn, recs );
      return withProblem< dim >( problem, [ & ] ( const auto &p ) {
        using Problem = std::decay_t< decltype( p ) >;
        Velocity< Problem, typename C::GridView > velocity( p, time );
        *dtEstOut = Dune::VoF::evolution( ctx.gv )( recs, flags, velocity, dt, update );
        std::copy( update.begin(), update.end(), updateOut );
      } );
    } ); } );
  }

  // curvature/swartzcurvature.hh:41-173 (2D only: the reference static_asserts dimension == 2)
  int ref_curvature_swartz ( void *handle, const double *hsIn, const std::uint8_t *flagsIn, double *kappaOut )
  {
    return guarded( [ & ] { return withCtx( handle, [ & ] ( auto &ctx ) {
      using C = std::decay_t< decltype( ctx ) >;
      if constexpr( C::GridView::dimension == 2 )
      {
        typename C::Color uh( ctx.gv );
        typename C::Flags flags( ctx.gv );
        typename C::Recs recs( ctx.gv );
        typename C::Curv curvature( ctx.gv );
        ctx.loadFlags( flagsIn, flags );
        ctx.loadRecs( hsIn, recs );
        Dune::VoF::SwartzCurvature< typename C::GridView, typename C::Stencils > op( ctx.vertexStencils() );
        op( uh, recs, flags, curvature );
        std::copy( curvature.begin(), curvature.end(), kappaOut );
        return 0;
      }
      else
      {
        lastError = "SwartzCurvature is implemented in 2D only";
        return 1;
      }
    } ); } );
  }

  // curvature/cartesianheightfunctioncurvature.hh:54-87 ; valid = satisfiesConstraint_ after the operator
  int ref_curvature ( void *handle, const double *u, const double *hsIn, const std::uint8_t *flagsIn, double *kappaOut, std::uint8_t *validOut )
  {
    return guarded( [ & ] { return withCtx( handle, [ & ] ( auto &ctx ) {
      using C = std::decay_t< decltype( ctx ) >;
      typename C::Color uh( ctx.gv );
      typename C::Flags flags( ctx.gv );
      typename C::Recs recs( ctx.gv );
      typename C::Curv curvature( ctx.gv );
      ctx.loadColor( u, uh );
      ctx.loadFlags( flagsIn, flags );
      ctx.loadRecs( hsIn, recs );
      typename C::HFCurv op( ctx.vertexStencils() );
      op( uh, recs, flags, curvature );
      std::copy( curvature.begin(), curvature.end(), kappaOut );
      if( validOut )
      {
        auto it = op.satisfiesConstraint_.begin();
        for( std::size_t i = 0; i < ctx.cells(); ++i, ++it )
          validOut[ i ] = std::uint8_t( *it );
      }
      return 0;
    } ); } );
  }

  // test/average.hh:28-101
  int ref_init ( void *handle, int problem, double time, double *uOut )
  {
    return guarded( [ & ] { return withCtx( handle, [ & ] ( auto &ctx ) {
      using C = std::decay_t< decltype( ctx ) >;
      constexpr int dim = C::GridView::dimension;
      typename C::Color uh( ctx.gv );
      const int rc = withProblem< dim >( problem, [ & ] ( const auto &p ) { initialData( ctx, p, time, uh ); } );
      std::copy( uh.begin(), uh.end(), uOut );
      return rc;
    } ); } );
  }

  // velocity.hh:15-20 at the centre of face `face` of cell `cell` (as the cell's own intersection sees it)
  int ref_face_velocity ( void *handle, int problem, double time, std::int64_t cell, int face, double *vOut )
  {
    return guarded( [ & ] { return withCtx( handle, [ & ] ( auto &ctx ) {
      using C = std::decay_t< decltype( ctx ) >;
      constexpr int dim = C::GridView::dimension;
      Dune::MiniElemIt< dim > it { &ctx.grid, std::size_t( cell ) };
      Dune::MiniIntersection< dim > is { *it, face };
      return withProblem< dim >( problem, [ & ] ( const auto &p ) {
        using Problem = std::decay_t< decltype( p ) >;
        Velocity< Problem, typename C::GridView > velocity( p, time );
        velocity.bind( is );
        const auto v = velocity( Dune::FieldVector< double, dim - 1 >( 0.5 ) );
        for( int d = 0; d < dim; ++d )
          vOut[ d ] = v[ d ];
      } );
    } ); } );
  }

  // replay of algorithm.hh:68-95; `passes` loop iterations starting from (time, dt) = (*timeIO, *dtIO).
  // phaseSeconds[4] (flag, reconstruct, evolve, axpy) and mixedCellSteps are accumulated if non-null.
  int ref_run ( void *handle, int kind, int problem, double eps, double cfl, int maxIterations, int passes,
                double *u, double *timeIO, double *dtIO, double *phaseSeconds, std::int64_t *mixedCellSteps )
  {
    return guarded( [ & ] { return withCtx( handle, [ & ] ( auto &ctx ) {
      using C = std::decay_t< decltype( ctx ) >;
      constexpr int dim = C::GridView::dimension;
      const auto &stencils = ctx.vertexStencils();
      return withProblem< dim >( problem, [ & ] ( const auto &p ) {
        switch( kind )
        {
        case RECON_YOUNGS: runLoop( ctx, typename C::Youngs { stencils }, p, eps, cfl, passes, u, timeIO, dtIO, phaseSeconds, mixedCellSteps ); break;
        case RECON_HF: runLoop( ctx, typename C::HF { stencils }, p, eps, cfl, passes, u, timeIO, dtIO, phaseSeconds, mixedCellSteps ); break;
        default: runLoop( ctx, typename C::Swartz { stencils, std::size_t( maxIterations ) }, p, eps, cfl, passes, u, timeIO, dtIO, phaseSeconds, mixedCellSteps ); break;
        }
      } );
    } ); } );
  }

  // ---- geometry primitives on a single axis-aligned cell (for fuzzing the restatement and the kernels) ----

  // geometry/algorithm.hh:68-111 (2D) / :119-236 (3D); out = (normal[dim], distance)
  int ref_locate ( int dim, const double *lo, const double *h, const double *normal, double fraction, double *out )
  {
    return guarded( [ & ] {
      if( dim == 2 )
      {
        Dune::BoxGeometry< 2, 2 > geo;
        for( int d = 0; d < 2; ++d ) { geo.o[ d ] = lo[ d ]; geo.h[ d ] = h[ d ]; geo.axes[ d ] = d; }
        Dune::FieldVector< double, 2 > n { normal[ 0 ], normal[ 1 ] };
        const auto hs = Dune::VoF::locateHalfSpace( Dune::VoF::makePolytope( geo ), n, fraction );
        out[ 0 ] = hs.innerNormal()[ 0 ]; out[ 1 ] = hs.innerNormal()[ 1 ]; out[ 2 ] = hs.boundary().distance();
        return 0;
      }
      Dune::BoxGeometry< 3, 3 > geo;
      for( int d = 0; d < 3; ++d ) { geo.o[ d ] = lo[ d ]; geo.h[ d ] = h[ d ]; geo.axes[ d ] = d; }
      Dune::FieldVector< double, 3 > n { normal[ 0 ], normal[ 1 ], normal[ 2 ] };
      const auto hs = Dune::VoF::locateHalfSpace( Dune::VoF::makePolytope( geo ), n, fraction );
      out[ 0 ] = hs.innerNormal()[ 0 ]; out[ 1 ] = hs.innerNormal()[ 1 ]; out[ 2 ] = hs.innerNormal()[ 2 ]; out[ 3 ] = hs.boundary().distance();
      return 0;
    } );
  }

  // getVolumeFraction, geometry/algorithm.hh:31-45
  int ref_volume_fraction ( int dim, const double *lo, const double *h, const double *hs, double *out )
  {
    return guarded( [ & ] {
      if( dim == 2 )
      {
        Dune::BoxGeometry< 2, 2 > geo;
        for( int d = 0; d < 2; ++d ) { geo.o[ d ] = lo[ d ]; geo.h[ d ] = h[ d ]; geo.axes[ d ] = d; }
        Dune::VoF::HalfSpace< Dune::FieldVector< double, 2 > > H( { hs[ 0 ], hs[ 1 ] }, hs[ 2 ] );
        *out = Dune::VoF::getVolumeFraction( Dune::VoF::makePolytope( geo ), H );
        return 0;
      }
      Dune::BoxGeometry< 3, 3 > geo;
      for( int d = 0; d < 3; ++d ) { geo.o[ d ] = lo[ d ]; geo.h[ d ] = h[ d ]; geo.axes[ d ] = d; }
      Dune::VoF::HalfSpace< Dune::FieldVector< double, 3 > > H( { hs[ 0 ], hs[ 1 ], hs[ 2 ] }, hs[ 3 ] );
      *out = Dune::VoF::getVolumeFraction( Dune::VoF::makePolytope( geo ), H );
      return 0;
    } );
  }

  // truncVolume( upwindPolygon( face geometry, v ), halfspace ): evolution/evolution.hh:168-171
  // the face is face `face` (2*axis+side) of the cell with lower corner lo and widths h.
  int ref_flux ( int dim, const double *lo, const double *h, int face, const double *v, const double *hs, double *out )
  {
    return guarded( [ & ] {
      if( dim == 2 )
      {
        Dune::MiniGrid< 2 > g; g.n = { 1, 1 };   // one-cell grid: lower corner reproduced exactly (lo + 0*h)
        for( int d = 0; d < 2; ++d ) { g.lo[ d ] = lo[ d ]; g.h[ d ] = h[ d ]; }
        Dune::MiniEntity< 2 > e; e.g = &g; e.ijk = { 0, 0 };
        Dune::MiniIntersection< 2 > is { e, face };
        using Coord = Dune::FieldVector< double, 2 >;
        Coord vel { v[ 0 ], v[ 1 ] };
        Dune::VoF::HalfSpace< Coord > H( { hs[ 0 ], hs[ 1 ] }, hs[ 2 ] );
        auto upwind = Dune::VoF::upwindPolygon( is.geometry(), vel );
        *out = Dune::VoF::truncVolume( upwind, H );
        return 0;
      }
      Dune::MiniGrid< 3 > g; g.n = { 1, 1, 1 };
      for( int d = 0; d < 3; ++d ) { g.lo[ d ] = lo[ d ]; g.h[ d ] = h[ d ]; }
      Dune::MiniEntity< 3 > e; e.g = &g; e.ijk = { 0, 0, 0 };
      Dune::MiniIntersection< 3 > is { e, face };
      using Coord = Dune::FieldVector< double, 3 >;
      Coord vel { v[ 0 ], v[ 1 ], v[ 2 ] };
      Dune::VoF::HalfSpace< Coord > H( { hs[ 0 ], hs[ 1 ], hs[ 2 ] }, hs[ 3 ] );
      auto upwind = Dune::VoF::upwindPolygon( is.geometry(), vel );
      *out = Dune::VoF::truncVolume( upwind, H );
      return 0;
    } );
  }

  // interface( cell, halfspace ): geometry/polytope.hh:69-75 -> vertices (<= 8), centroid, measure
  // returns the number of vertices in *nverts; verts has room for 8*dim doubles.
  int ref_interface ( int dim, const double *lo, const double *h, const double *hs, int *nverts, double *verts, double *centroid, double *measure )
  {
    return guarded( [ & ] {
      if( dim == 2 )
      {
        using Coord = Dune::FieldVector< double, 2 >;
        Dune::BoxGeometry< 2, 2 > geo;
        for( int d = 0; d < 2; ++d ) { geo.o[ d ] = lo[ d ]; geo.h[ d ] = h[ d ]; geo.axes[ d ] = d; }
        Dune::VoF::HalfSpace< Coord > H( { hs[ 0 ], hs[ 1 ] }, hs[ 2 ] );
        const auto poly = Dune::VoF::makePolytope( geo );
        auto line = Dune::VoF::intersect( poly, H.boundary(), Dune::VoF::eager );
        *nverts = 2;
        for( int i = 0; i < 2; ++i )
          for( int d = 0; d < 2; ++d )
            verts[ 2 * i + d ] = line.vertex( i )[ d ];
        const auto c = line.centroid();
        centroid[ 0 ] = c[ 0 ]; centroid[ 1 ] = c[ 1 ];
        *measure = line.volume();
        return 0;
      }
      using Coord = Dune::FieldVector< double, 3 >;
      Dune::BoxGeometry< 3, 3 > geo;
      for( int d = 0; d < 3; ++d ) { geo.o[ d ] = lo[ d ]; geo.h[ d ] = h[ d ]; geo.axes[ d ] = d; }
      Dune::VoF::HalfSpace< Coord > H( { hs[ 0 ], hs[ 1 ], hs[ 2 ] }, hs[ 3 ] );
      const auto poly = Dune::VoF::makePolytope( geo );
      auto face = Dune::VoF::intersect( poly, H.boundary(), Dune::VoF::eager );
      *nverts = int( face.size() );
      for( std::size_t i = 0; i < face.size() && i < 8; ++i )
        for( int d = 0; d < 3; ++d )
          verts[ 3 * i + d ] = face.vertex( i )[ d ];
      if( face.size() > 2 )
      {
        const auto c = face.centroid();
        centroid[ 0 ] = c[ 0 ]; centroid[ 1 ] = c[ 1 ]; centroid[ 2 ] = c[ 2 ];
        *measure = face.volume();
      }
      else
      {
        centroid[ 0 ] = centroid[ 1 ] = centroid[ 2 ] = 0.0;
        *measure = 0.0;
      }
      return 0;
    } );
  }

  // brents.hh:131-146 on the cubic the 3D plane solve uses (geometry/algorithm.hh:211-224), for unit checks
  double ref_brent_cubic ( double A, double B, double C, double target, double a, double b, double tol )
  {
    auto f = [ = ] ( const double x ) { return C * x + B * x * x / 2.0 + A * x * x * x / 6.0 - target; };
    return Dune::VoF::brentsMethod( f, a, b, tol );
  }

} // extern "C"
