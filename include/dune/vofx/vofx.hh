// dune/vofx/vofx.hh -- header adaptor: dune-vof's operator API on top of the vofx C ABI (include/vofx.h).
//
// Drop-in for the per-timestep hot path of dune-vof: the classes below have the names, template parameters and call
// signatures of the reference's operators, so example/vof.cc and dune/vof/algorithm.hh swap them in by changing
// includes and the namespace (Dune::VoF -> Dune::VoFx), see INTEGRATION.md:
//
//   reference (dune/vof/...)                                          this header (namespace Dune::VoFx)
//   FlagOperator< GV >( eps )( color, flags )                          flagging.hh:24-74          FlagOperator
//   ModifiedYoungsReconstruction< GV, StS >( stencils )( c, r, f )     reconstruction/modifiedyoungs.hh:30-141
//   HeightFunctionReconstruction< GV, StS, IR >( stencils )( c, r, f ) reconstruction/heightfunction.hh:35-287
//   ModifiedSwartzReconstruction< GV, StS, IR >( stencils[, n] )       reconstruction/modifiedswartz.hh:36-252
//   reconstruction( stencils )                                         reconstruction.hh:29-37 (HF with Young's initialiser)
//   Evolution< GV >( gv )( recs, flags, velocity, dt, update ) -> dt   evolution/evolution.hh:37-181 ; evolution( gv ) evolution.hh:25-30
//   CartesianHeightFunctionCurvature< GV, VNST >( stencils )( c, r, f, curvature )   curvature/cartesianheightfunctioncurvature.hh:30-190
//   VertexNeighborsStencil< GV >( gv )                                 stencil/vertexneighborsstencil.hh:25-98 (here: owns the device context)
//   Algorithm< GV, PR, DW >( gv, problem, writer, cfl, eps )( uh, start, end, level )   algorithm.hh:37-120
//
// The grid stays a DUNE grid on the host (YaspGrid / SPGrid: axis-aligned, structured).  It is flattened once into
// the structured descriptor of include/vofx.h; containers stay the host-side DataSets of the reference
// (dataset.hh:26-87: operator[]( index ), size(), gridView()); every operator call stages them through the
// HOST-buffer entry points of the C ABI.  Algorithm keeps the colour function on the device between steps and only
// copies it back when the data writer is going to write.
// Errors of the C ABI are re-thrown as Dune::VoFx::Exception; there is no CPU fallback.
#ifndef DUNE_VOFX_VOFX_HH
#define DUNE_VOFX_VOFX_HH

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include <vofx.h>

namespace Dune
{
  namespace VoFx
  {

    struct Exception : std::runtime_error
    {
      using std::runtime_error::runtime_error;
    };

    inline void check ( int status )
    {
      if( status != VOFX_OK )
        throw Exception( std::string( "vofx: " ) + vofx_last_error() );
    }

    // Context
    // -------
    // The device context of one grid view: the flattened structured grid (include/vofx.h, "Grid convention") and the
    // permutation between the grid view's index set and the x-fastest lexicographic cell index.
    template< class GV >
    class Context
    {
    public:
      using GridView = GV;
      static constexpr int dim = GridView::dimension;

      explicit Context ( const GridView &gridView, int device = -1 )
      : gridView_( gridView )
      {
        // bounding box and mesh width from the cell geometries (the grid must be axis-aligned and uniform)
        double lo[ 3 ] = { 0, 0, 0 }, hi[ 3 ] = { 0, 0, 0 }, h[ 3 ] = { 1, 1, 1 };
        bool first = true;
        std::size_t count = 0;
        for( const auto &entity : elements( gridView_ ) )
        {
          const auto geo = entity.geometry();
          const auto a = geo.corner( 0 );
          const auto b = geo.corner( ( 1 << dim ) - 1 );
          for( int d = 0; d < dim; ++d )
          {
            if( first )
            {
              lo[ d ] = a[ d ];
              hi[ d ] = b[ d ];
              h[ d ] = b[ d ] - a[ d ];
            }
            lo[ d ] = std::min( lo[ d ], double( a[ d ] ) );
            hi[ d ] = std::max( hi[ d ], double( b[ d ] ) );
          }
          first = false;
          ++count;
        }
        vofx_grid_desc desc{};
        desc.dim = dim;
        std::size_t cells = 1;
        for( int d = 0; d < 3; ++d )
        {
          const bool used = d < dim;
          const int n = used ? int( std::lround( ( hi[ d ] - lo[ d ] ) / h[ d ] ) ) : 1;
          desc.n_global[ d ] = desc.n_local[ d ] = n;
          desc.origin[ d ] = used ? lo[ d ] : 0.0;
          desc.h[ d ] = used ? h[ d ] : 1.0;
          desc.lo[ d ] = 0;
          cells *= std::size_t( n );
        }
        desc.halo = 0;
        if( cells != count )
          throw Exception( "vofx: the grid view is not a full structured box (only axis-aligned structured grids are supported)" );
        desc_ = desc;

        // permutation: DUNE index -> lexicographic index (x fastest)
        toLex_.resize( cells );
        for( const auto &entity : elements( gridView_ ) )
        {
          const auto c = entity.geometry().center();
          std::size_t lex = 0;
          for( int d = dim - 1; d >= 0; --d )
          {
            const int i = int( std::floor( ( c[ d ] - desc.origin[ d ] ) / desc.h[ d ] ) );
            lex = lex * std::size_t( desc.n_global[ d ] ) + std::size_t( std::min( std::max( i, 0 ), desc.n_global[ d ] - 1 ) );
          }
          toLex_[ gridView_.indexSet().index( entity ) ] = lex;
        }
        check( vofx_create( &desc_, device, &ctx_ ) );
      }

      Context ( const Context & ) = delete;
      Context &operator= ( const Context & ) = delete;
      ~Context () { vofx_destroy( ctx_ ); }

      vofx_ctx *ctx () const { return ctx_; }
      const GridView &gridView () const { return gridView_; }
      const vofx_grid_desc &desc () const { return desc_; }
      std::size_t size () const { return toLex_.size(); }
      std::size_t lex ( std::size_t duneIndex ) const { return toLex_[ duneIndex ]; }

      // staging buffers in lexicographic order
      std::vector< double > color, update, kappa, halfspaces, faceVelocity;
      std::vector< std::uint8_t > flags, valid;

      template< class DataSet >
      void gatherScalars ( const DataSet &data, std::vector< double > &out ) const
      {
        out.resize( size() );
        for( std::size_t i = 0; i < size(); ++i )
          out[ toLex_[ i ] ] = data[ i ];
      }

      template< class DataSet >
      void scatterScalars ( const std::vector< double > &in, DataSet &data ) const
      {
        for( std::size_t i = 0; i < size(); ++i )
          data[ i ] = in[ toLex_[ i ] ];
      }

      // FlagSet::operator[]( Index ) const returns the flag as double (flagset.hh:66): read through begin()
      template< class FlagSet >
      void gatherFlags ( const FlagSet &flagSet )
      {
        flags.resize( size() );
        std::size_t i = 0;
        for( auto it = flagSet.begin(); it != flagSet.end(); ++it, ++i )
          flags[ toLex_[ i ] ] = std::uint8_t( static_cast< std::size_t >( *it ) );
      }

      template< class FlagSet >
      void scatterFlags ( FlagSet &flagSet ) const
      {
        using Flag = typename FlagSet::DataType;
        std::size_t i = 0;
        for( auto it = flagSet.begin(); it != flagSet.end(); ++it, ++i )
          *it = static_cast< Flag >( static_cast< std::size_t >( flags[ toLex_[ i ] ] ) );
      }

      template< class ReconstructionSet >
      void gatherReconstructions ( const ReconstructionSet &recs )
      {
        halfspaces.resize( size() * ( dim + 1 ) );
        std::size_t i = 0;
        for( auto it = recs.begin(); it != recs.end(); ++it, ++i )
        {
          double *hs = halfspaces.data() + toLex_[ i ] * ( dim + 1 );
          for( int d = 0; d < dim; ++d )
            hs[ d ] = it->innerNormal()[ d ];
          hs[ dim ] = it->boundary().distance();
        }
      }

      template< class ReconstructionSet >
      void scatterReconstructions ( ReconstructionSet &recs ) const
      {
        using HalfSpace = typename ReconstructionSet::DataType;
        using Coordinate = typename HalfSpace::Coordinate;
        std::size_t i = 0;
        for( auto it = recs.begin(); it != recs.end(); ++it, ++i )
        {
          const double *hs = halfspaces.data() + toLex_[ i ] * ( dim + 1 );
          Coordinate normal;
          for( int d = 0; d < dim; ++d )
            normal[ d ] = hs[ d ];
          *it = HalfSpace( normal, hs[ dim ] );
        }
      }

      // sample a velocity functor with the reference's concept (bind( intersection ), operator()( local ), unbind();
      // velocity.hh:15-24) at every face centre and hand it to the device (general fallback of the C ABI)
      template< class Velocity >
      void setVelocity ( Velocity &velocity )
      {
        faceVelocity.assign( size() * 2 * dim * dim, 0.0 );
        for( const auto &entity : elements( gridView_ ) )
        {
          const std::size_t cell = toLex_[ gridView_.indexSet().index( entity ) ];
          int face = 0;
          for( const auto &intersection : intersections( gridView_, entity ) )
          {
            velocity.bind( intersection );
            typename std::decay_t< decltype( intersection ) >::LocalCoordinate local( 0.5 );
            const auto v = velocity( local );
            velocity.unbind();
            // dune-grid numbers cube faces x-,x+,y-,y+,z-,z+; identify the face by its outer normal to be independent of the iteration order
            const auto n = intersection.centerUnitOuterNormal();
            for( int d = 0; d < dim; ++d )
              if( std::abs( n[ d ] ) > 0.5 )
                face = 2 * d + ( n[ d ] > 0 ? 1 : 0 );
            for( int d = 0; d < dim; ++d )
              faceVelocity[ ( cell * 2 * dim + face ) * dim + d ] = v[ d ];
          }
        }
        check( vofx_velocity_set_faces( ctx_, faceVelocity.data() ) );
      }

    private:
      GridView gridView_;
      vofx_grid_desc desc_;
      vofx_ctx *ctx_ = nullptr;
      std::vector< std::size_t > toLex_;
    };


    // VertexNeighborsStencil
    // ----------------------
    // In the reference this stores >= 26 entities per cell (stencil/vertexneighborsstencil.hh:29,94-97); on a
    // structured grid the device derives the neighbours from the index, so the object only owns the device context
    // that the operators constructed from it share (they take `const StencilSet &`, exactly as in the reference).
    template< class GV >
    struct VertexNeighborsStencil
    {
      using GridView = GV;

      explicit VertexNeighborsStencil ( const GridView &gridView, int device = -1 )
      : context_( std::make_shared< Context< GridView > >( gridView, device ) )
      {}

      const GridView &gridView () const { return context_->gridView(); }
      Context< GridView > &context () const { return *context_; }

    private:
      std::shared_ptr< Context< GridView > > context_;
    };


    // FlagOperator, flagging.hh:24-74
    template< class GV >
    struct FlagOperator
    {
      using GridView = GV;

      FlagOperator ( const VertexNeighborsStencil< GV > &stencils, double eps ) : context_( &stencils.context() ), eps_( eps ) {}
      FlagOperator ( Context< GV > &context, double eps ) : context_( &context ), eps_( eps ) {}

      template< class ColorFunction, class FlagSet >
      void operator() ( const ColorFunction &color, FlagSet &flags, bool /*communicate*/ = true ) const
      {
        auto &c = *context_;
        c.gatherScalars( color, c.color );
        c.flags.resize( c.size() );
        check( vofx_flag_host( c.ctx(), c.color.data(), eps_, c.flags.data() ) );
        c.scatterFlags( flags );
      }

    private:
      Context< GV > *context_;
      const double eps_;
    };


    namespace Impl
    {
      template< class GV, int kind >
      struct Reconstruction
      {
        using GridView = GV;
        using StencilSet = VertexNeighborsStencil< GV >;

        explicit Reconstruction ( const StencilSet &stencils, std::size_t maxIterations = 10 )
        : context_( &stencils.context() ), maxIterations_( int( maxIterations ) )
        {}

        template< class ColorFunction, class ReconstructionSet, class Flags >
        void operator() ( const ColorFunction &color, ReconstructionSet &reconstructions, const Flags &flags, bool /*communicate*/ = true ) const
        {
          auto &c = *context_;
          c.gatherScalars( color, c.color );
          c.gatherFlags( flags );
          c.halfspaces.resize( c.size() * ( GV::dimension + 1 ) );
          check( vofx_reconstruct_host( c.ctx(), kind, c.color.data(), c.flags.data(), c.halfspaces.data(), maxIterations_ ) );
          c.scatterReconstructions( reconstructions );
        }

      protected:
        Context< GV > *context_;
        int maxIterations_;
      };
    } // namespace Impl

    // reconstruction/modifiedyoungs.hh:30-141
    template< class GV, class StS = VertexNeighborsStencil< GV > >
    struct ModifiedYoungsReconstruction : Impl::Reconstruction< GV, VOFX_RECON_YOUNGS >
    {
      using Impl::Reconstruction< GV, VOFX_RECON_YOUNGS >::Reconstruction;
    };

    // reconstruction/heightfunction.hh:35-287 (the initial reconstruction is always modified Young's, as in reconstruction.hh:29-37)
    template< class GV, class StS = VertexNeighborsStencil< GV >, class IR = ModifiedYoungsReconstruction< GV, StS > >
    struct HeightFunctionReconstruction : Impl::Reconstruction< GV, VOFX_RECON_HF >
    {
      using Impl::Reconstruction< GV, VOFX_RECON_HF >::Reconstruction;
    };

    // reconstruction/modifiedswartz.hh:36-252
    template< class GV, class StS = VertexNeighborsStencil< GV >, class IR = ModifiedYoungsReconstruction< GV, StS > >
    struct ModifiedSwartzReconstruction : Impl::Reconstruction< GV, VOFX_RECON_SWARTZ >
    {
      using Impl::Reconstruction< GV, VOFX_RECON_SWARTZ >::Reconstruction;
    };

    // reconstruction.hh:29-37
    template< class Stencils >
    static inline auto reconstruction ( Stencils &stencils )
      -> HeightFunctionReconstruction< typename Stencils::GridView, Stencils >
    {
      return HeightFunctionReconstruction< typename Stencils::GridView, Stencils >( stencils );
    }


    // Evolution, evolution/evolution.hh:37-181
    template< class GV >
    struct Evolution
    {
      using GridView = GV;

      explicit Evolution ( const VertexNeighborsStencil< GV > &stencils ) : context_( &stencils.context() ) {}
      explicit Evolution ( Context< GV > &context ) : context_( &context ) {}

      // velocity: the reference's concept (velocity.hh:5-33), frozen at the start-of-step time by its constructor
      template< class ReconstructionSet, class Flags, class Velocity, class DiscreteFunction >
      double operator() ( const ReconstructionSet &reconstructions, const Flags &flags, Velocity &velocity, double deltaT, DiscreteFunction &update ) const
      {
        auto &c = *context_;
        c.setVelocity( velocity );
        c.gatherReconstructions( reconstructions );
        c.gatherFlags( flags );
        c.update.resize( c.size() );
        double dtEst = 0.0;
        check( vofx_evolve_host( c.ctx(), c.halfspaces.data(), c.flags.data(), 0.0, deltaT, c.update.data(), &dtEst ) );
        c.scatterScalars( c.update, update );
        return dtEst;
      }

    private:
      Context< GV > *context_;
    };

    // evolution.hh:25-30 (takes the stencil set instead of the grid view: it carries the device context)
    template< class GV >
    static inline Evolution< GV > evolution ( const VertexNeighborsStencil< GV > &stencils )
    {
      return Evolution< GV >( stencils );
    }


    // CartesianHeightFunctionCurvature, curvature/cartesianheightfunctioncurvature.hh:30-190
    template< class GV, class VNST = VertexNeighborsStencil< GV > >
    struct CartesianHeightFunctionCurvature
    {
      using GridView = GV;

      explicit CartesianHeightFunctionCurvature ( const VNST &stencils ) : context_( &stencils.context() ) {}

      template< class ColorFunction, class ReconstructionSet, class Flags, class CurvatureSet >
      void operator() ( const ColorFunction &color, const ReconstructionSet &reconstructions, const Flags &flags, CurvatureSet &curvature,
                        bool /*communicate*/ = true ) const
      {
        auto &c = *context_;
        c.gatherScalars( color, c.color );
        c.gatherReconstructions( reconstructions );
        c.gatherFlags( flags );
        c.kappa.resize( c.size() );
        c.valid.resize( c.size() );
        check( vofx_curvature_host( c.ctx(), c.color.data(), c.halfspaces.data(), c.flags.data(), c.kappa.data(), c.valid.data() ) );
        c.scatterScalars( c.kappa, curvature );
      }

      // satisfiesConstraint_ of the last call (heightfunction.hh:286), lexicographic order
      const std::vector< std::uint8_t > &satisfiesConstraint () const { return context_->valid; }

    private:
      Context< GV > *context_;
    };


    // Algorithm, algorithm.hh:37-120 (without l1error): the colour function stays on the device between steps.
    // SwartzCurvature (2D)
    // --------------------
    // curvature/swartzcurvature.hh:27-181
    template< class GV, class StS = VertexNeighborsStencil< GV > >
    struct SwartzCurvature
    {
      using GridView = GV;
      explicit SwartzCurvature ( const StS &stencils ) : stencils_( stencils ) {}

      template< class ColorFunction, class ReconstructionSet, class Flags, class CurvatureSet >
      void operator() ( const ColorFunction &, const ReconstructionSet &reconstructions, const Flags &flags, CurvatureSet &curvatureSet,
                        bool /*communicate*/ = true ) const
      {
        auto &c = stencils_.context();
        c.gatherReconstructions( reconstructions );
        c.gatherFlags( flags );
        c.kappa.resize( c.size() );
        check( vofx_curvature_swartz_host( c.ctx(), c.halfspaces.data(), c.flags.data(), c.kappa.data() ) );
        c.scatterScalars( c.kappa, curvatureSet );
      }

    private:
      const StS &stencils_;
    };


    // getInterfaceVertices
    // --------------------
    // utility.hh:19-48: the vertices of interface( entity, reconstructions ) of every mixed cell in the order the grid
    // view visits them, as a CSR.  (On YaspGrid / SPGrid that order is the x-fastest index order the device uses.)
    template< class Stencils, class ReconstructionSet, class Flags >
    void getInterfaceVertices ( const Stencils &stencils, const ReconstructionSet &reconstructions, const Flags &flags,
                                std::vector< typename ReconstructionSet::DataType::Coordinate > &vertices, std::vector< std::size_t > &offsets )
    {
      using Coordinate = typename ReconstructionSet::DataType::Coordinate;
      auto &c = stencils.context();
      constexpr int dim = std::decay_t< decltype( c ) >::dim;
      c.gatherReconstructions( reconstructions );
      c.gatherFlags( flags );
      std::int64_t nMixed = 0, nVertices = 0;
      check( vofx_interface_host( c.ctx(), c.halfspaces.data(), c.flags.data(), &nMixed, &nVertices ) );
      std::vector< std::int32_t > cells( static_cast< std::size_t >( nMixed ) + 1 );
      std::vector< std::int64_t > off( static_cast< std::size_t >( nMixed ) + 1 );
      std::vector< double > flat( static_cast< std::size_t >( nVertices ) * dim + 1 );
      check( vofx_interface_fetch( c.ctx(), cells.data(), off.data(), flat.data() ) );
      offsets.assign( off.begin(), off.end() );
      vertices.resize( static_cast< std::size_t >( nVertices ) );
      for( std::size_t i = 0; i < vertices.size(); ++i )
      {
        Coordinate x;
        for( int d = 0; d < dim; ++d )
          x[ d ] = flat[ i * dim + d ];
        vertices[ i ] = x;
      }
    }


    namespace Impl
    {
      template< class DW, class = void >
      struct WillWrite
      {
        static bool apply ( DW &, double ) { return true; }
      };
      template< class DW >
      struct WillWrite< DW, std::void_t< decltype( std::declval< DW & >().willWrite( 0.0 ) ) > >
      {
        static bool apply ( DW &writer, double time ) { return writer.willWrite( time ); }
      };
    } // namespace Impl

    template< class GV, class PR, class DW >
    struct Algorithm
    {
      using GridView = GV;
      using Problem = PR;
      using DataWriter = DW;
      using Stencils = VertexNeighborsStencil< GridView >;

      Algorithm ( const GridView &gridView, const Problem &problem, DataWriter &dataWriter, double cfl, double eps, const bool verbose = false,
                  int reconstruction = VOFX_RECON_HF, int builtinVelocity = -1 )
      : gridView_( gridView ), problem_( problem ), dataWriter_( dataWriter ), cfl_( cfl ), eps_( eps ), verbose_( verbose ),
        stencils_( gridView ), reconstruction_( reconstruction ), builtinVelocity_( builtinVelocity )
      {}

      // builtinVelocity >= 0 (enum vofx_problem): the separable device tables are used and time lives on the device;
      // otherwise problem.velocityField is sampled on the host at every face each step (general, slower).
      template< class ColorFunction >
      double operator() ( ColorFunction &uh, double start, double end, int /*level*/ = 0 )
      {
        auto &c = stencils_.context();
        const std::size_t N = c.size();
        vofx_step_params params{};
        params.reconstruction = reconstruction_;
        params.max_iterations = 10;
        params.eps = eps_;
        params.cfl = cfl_;
        params.with_curvature = 0;
        if( builtinVelocity_ >= 0 )
          check( vofx_velocity_builtin( c.ctx(), builtinVelocity_ ) );

        c.gatherScalars( uh, c.color );
        c.flags.resize( N );
        check( vofx_color_upload( c.ctx(), c.color.data() ) );
        double time = start, dt = 0.0, dtEst = 0.0;
        check( vofx_step_begin( c.ctx(), time, dt ) );
        do
        {
          if( builtinVelocity_ < 0 )
          {
            HostVelocity velocity( problem_, time );
            c.setVelocity( velocity );
          }
          // one loop pass, algorithm.hh:75-93, entirely on the device-resident colour function
          check( vofx_step_resident( c.ctx(), &params ) );
          check( vofx_step_state( c.ctx(), &time, &dt, &dtEst, nullptr ) );
          if( Impl::WillWrite< DataWriter >::apply( dataWriter_, time ) )
          {
            check( vofx_color_download( c.ctx(), c.color.data(), nullptr ) );
            c.scatterScalars( c.color, uh );
            dataWriter_.write( time );
          }
        }
        while( time < end );
        check( vofx_color_download( c.ctx(), c.color.data(), c.flags.data() ) );
        c.scatterScalars( c.color, uh );
        return 0.0;   // the reference returns the accumulated l1 error (diagnostic, test/errors.hh); not part of the hot path
      }

    private:
      // Velocity< Problem, GridView > of velocity.hh:5-33
      struct HostVelocity
      {
        using Intersection = typename GridView::Intersection;
        HostVelocity ( const Problem &problem, double time ) : problem_( &problem ), time_( time ) {}
        template< class Local >
        auto operator() ( const Local &local ) const
        {
          typename Problem::DomainType v( 0.0 );
          problem_->velocityField( intersection_->geometry().global( local ), time_, v );
          return v;
        }
        void bind ( const Intersection &intersection ) { intersection_ = &intersection; }
        void unbind () { intersection_ = nullptr; }
        const Intersection *intersection_ = nullptr;
        const Problem *problem_;
        double time_;
      };

      const GridView &gridView_;
      const Problem &problem_;
      DataWriter &dataWriter_;
      const double cfl_, eps_;
      const bool verbose_;
      Stencils stencils_;
      int reconstruction_, builtinVelocity_;
    };

  } // namespace VoFx
} // namespace Dune

#endif // #ifndef DUNE_VOFX_VOFX_HH
