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
#include <map>
#include <memory>
#include <mutex>
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

    namespace Impl
    {
      // the reference's containers communicate their overlap (dataset.hh:69-81); plain host containers do not
      template< class DataSet, class = void >
      struct Communicate
      {
        static void apply ( DataSet & ) {}
      };
      template< class DataSet >
      struct Communicate< DataSet, std::void_t< decltype( std::declval< DataSet & >().communicate() ) > >
      {
        static void apply ( DataSet &data ) { data.communicate(); }
      };

      template< class Comm, class = void >
      struct HasAllgather : std::false_type {};
      template< class Comm >
      struct HasAllgather< Comm, std::void_t< decltype( std::declval< const Comm & >().allgather( (char *) nullptr, 1, (char *) nullptr ) ) > > : std::true_type {};
    } // namespace Impl

    // Context
    // -------
    // The device context of one grid view: the flattened structured grid (include/vofx.h, "Grid convention") and the map
    // from the grid view's index set to the storage index of the context (x-fastest lexicographic).
    //
    // Parallel grid views (gridView.comm().size() > 1, one GPU per rank): the rank's INTERIOR cells must form a slab of the
    // global box along the last axis.  The context then owns that slab plus `halo` layers on either side (include/vofx.h,
    // "multi-GPU"), the ranks' windows are connected through gridView.comm().allgather, and Algorithm steps with
    // vofx_step_distributed: the kernels exchange halos and the time-step MIN among themselves.  Cells of the grid view's
    // overlap that lie inside the halo layers are staged like interior cells; results are scattered to interior cells and
    // the container's own communicate() (dataset.hh:69-81) fills its overlap, as after the reference's operators.
    template< class GV >
    class Context
    {
    public:
      using GridView = GV;
      static constexpr int dim = GridView::dimension;
      static constexpr std::size_t npos = ~std::size_t( 0 );

      explicit Context ( const GridView &gridView, int device = -1, int halo = 4 )
      : gridView_( gridView )
      {
        const auto &comm = gridView_.comm();
        world_ = comm.size();
        rank_ = comm.rank();

        // bounding box of the interior cells and mesh width from the cell geometries (axis-aligned, uniform)
        double lo[ 3 ] = { 0, 0, 0 }, hi[ 3 ] = { 0, 0, 0 }, h[ 3 ] = { 1, 1, 1 };
        bool first = true;
        std::size_t interior = 0, all = 0;
        for( const auto &entity : elements( gridView_, Partitions::interior ) )
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
          ++interior;
        }
        double glo[ 3 ], ghi[ 3 ];
        for( int d = 0; d < 3; ++d )
        {
          glo[ d ] = ( world_ > 1 && d < dim ) ? comm.min( lo[ d ] ) : lo[ d ];
          ghi[ d ] = ( world_ > 1 && d < dim ) ? comm.max( hi[ d ] ) : hi[ d ];
        }
        vofx_grid_desc desc{};
        desc.dim = dim;
        std::size_t cells = 1;
        for( int d = 0; d < 3; ++d )
        {
          const bool used = d < dim;
          desc.n_global[ d ] = used ? int( std::lround( ( ghi[ d ] - glo[ d ] ) / h[ d ] ) ) : 1;
          desc.n_local[ d ] = used ? int( std::lround( ( hi[ d ] - lo[ d ] ) / h[ d ] ) ) : 1;
          desc.lo[ d ] = used ? int( std::lround( ( lo[ d ] - glo[ d ] ) / h[ d ] ) ) : 0;
          desc.origin[ d ] = used ? glo[ d ] : 0.0;
          desc.h[ d ] = used ? h[ d ] : 1.0;
          cells *= std::size_t( desc.n_local[ d ] );
          if( used && d != dim - 1 && desc.n_local[ d ] != desc.n_global[ d ] )
            throw Exception( "vofx: a parallel grid view must be decomposed along the last axis only (slabs)" );
        }
        desc.halo = ( world_ > 1 ) ? halo : 0;
        if( cells != interior )
          throw Exception( "vofx: the interior cells of the grid view are not a full structured box (only axis-aligned structured grids are supported)" );
        if( world_ > 1 && desc.n_local[ dim - 1 ] < desc.halo )
          throw Exception( "vofx: the slab of this rank is thinner than the halo" );
        desc_ = desc;
        check( vofx_create( &desc_, device, &ctx_ ) );
        cells_ = std::size_t( vofx_cells( ctx_ ) );

        // map: DUNE index (all partitions) -> storage index; npos for cells outside the context's halo
        const std::size_t duneCells = std::size_t( gridView_.indexSet().size( 0 ) );
        toLex_.assign( duneCells, npos );
        interior_.assign( duneCells, 0 );
        const int la = dim - 1;
        auto storageIndex = [ & ] ( const auto &entity ) {
          const auto c = entity.geometry().center();
          std::size_t lex = 0;
          for( int d = dim - 1; d >= 0; --d )
          {
            int i = int( std::floor( ( c[ d ] - desc_.origin[ d ] ) / desc_.h[ d ] ) );
            i = std::min( std::max( i, 0 ), desc_.n_global[ d ] - 1 );
            int extent = desc_.n_local[ d ];
            if( d == la )
            {
              i += desc_.halo - desc_.lo[ d ];
              extent += 2 * desc_.halo;
              if( i < 0 || i >= extent )
                return npos;
            }
            lex = lex * std::size_t( extent ) + std::size_t( i );
          }
          return lex;
        };
        for( const auto &entity : elements( gridView_ ) )
        {
          toLex_[ gridView_.indexSet().index( entity ) ] = storageIndex( entity );
          ++all;
        }
        for( const auto &entity : elements( gridView_, Partitions::interior ) )
          interior_[ gridView_.indexSet().index( entity ) ] = 1;
        (void) all;

        if( world_ > 1 )
          connect();
      }

      Context ( const Context & ) = delete;
      Context &operator= ( const Context & ) = delete;
      ~Context ()
      {
        if( world_ > 1 )
        {
          // an exported window must not be freed while another rank still maps it
          vofx_comm_disconnect( ctx_ );
          gridView_.comm().barrier();
        }
        vofx_destroy( ctx_ );
      }

      vofx_ctx *ctx () const { return ctx_; }
      const GridView &gridView () const { return gridView_; }
      const vofx_grid_desc &desc () const { return desc_; }
      std::size_t size () const { return toLex_.size(); }       // cells of the grid view (what the DataSets index)
      std::size_t cells () const { return cells_; }             // cells of the context's storage (incl. halo layers)
      std::size_t lex ( std::size_t duneIndex ) const { return toLex_[ duneIndex ]; }
      bool parallel () const { return world_ > 1; }
      int rank () const { return rank_; }

      // staging buffers in storage order
      std::vector< double > color, update, kappa, halfspaces, faceVelocity;
      std::vector< std::uint8_t > flags, valid;
      std::vector< std::int32_t > bandCells;

      template< class DataSet >
      void gatherScalars ( const DataSet &data, std::vector< double > &out ) const
      {
        out.assign( cells_, 0.0 );
        for( std::size_t i = 0; i < size(); ++i )
          if( toLex_[ i ] != npos )
            out[ toLex_[ i ] ] = data[ i ];
      }

      // results go to the interior cells; the container's communicate() fills its overlap (as after the reference's operators)
      template< class DataSet >
      void scatterScalars ( const std::vector< double > &in, DataSet &data, bool communicate = true ) const
      {
        for( std::size_t i = 0; i < size(); ++i )
          if( interior_[ i ] )
            data[ i ] = in[ toLex_[ i ] ];
        if( communicate && world_ > 1 )
          Impl::Communicate< DataSet >::apply( data );
      }

      // FlagSet::operator[]( Index ) const returns the flag as double (flagset.hh:66): read through begin()
      template< class FlagSet >
      void gatherFlags ( const FlagSet &flagSet )
      {
        flags.assign( cells_, 0 );
        std::size_t i = 0;
        for( auto it = flagSet.begin(); it != flagSet.end(); ++it, ++i )
          if( toLex_[ i ] != npos )
            flags[ toLex_[ i ] ] = std::uint8_t( static_cast< std::size_t >( *it ) );
      }

      template< class FlagSet >
      void scatterFlags ( FlagSet &flagSet, bool communicate = true ) const
      {
        using Flag = typename FlagSet::DataType;
        std::size_t i = 0;
        for( auto it = flagSet.begin(); it != flagSet.end(); ++it, ++i )
          if( interior_[ i ] )
            *it = static_cast< Flag >( static_cast< std::size_t >( flags[ toLex_[ i ] ] ) );
        if( communicate && world_ > 1 )
          Impl::Communicate< FlagSet >::apply( flagSet );
      }

      template< class ReconstructionSet >
      void gatherReconstructions ( const ReconstructionSet &recs )
      {
        halfspaces.assign( cells_ * ( dim + 1 ), 0.0 );
        std::size_t i = 0;
        for( auto it = recs.begin(); it != recs.end(); ++it, ++i )
        {
          if( toLex_[ i ] == npos )
            continue;
          double *hs = halfspaces.data() + toLex_[ i ] * ( dim + 1 );
          for( int d = 0; d < dim; ++d )
            hs[ d ] = it->innerNormal()[ d ];
          hs[ dim ] = it->boundary().distance();
        }
      }

      template< class ReconstructionSet >
      void scatterReconstructions ( ReconstructionSet &recs, bool communicate = true ) const
      {
        using HalfSpace = typename ReconstructionSet::DataType;
        using Coordinate = typename HalfSpace::Coordinate;
        std::size_t i = 0;
        for( auto it = recs.begin(); it != recs.end(); ++it, ++i )
        {
          if( !interior_[ i ] )
            continue;
          const double *hs = halfspaces.data() + toLex_[ i ] * ( dim + 1 );
          Coordinate normal;
          for( int d = 0; d < dim; ++d )
            normal[ d ] = hs[ d ];
          *it = HalfSpace( normal, hs[ dim ] );
        }
        if( communicate && world_ > 1 )
          Impl::Communicate< ReconstructionSet >::apply( recs );
      }

      // centre of `face` (x-,x+,y-,y+,z-,z+) of the storage cell `cell`, by the coordinate formulas of include/vofx.h
      template< class Coordinate >
      Coordinate faceCenter ( std::size_t cell, int face ) const
      {
        Coordinate x;
        std::size_t r = cell;
        for( int d = 0; d < dim; ++d )
        {
          const int extent = desc_.n_local[ d ] + ( d == dim - 1 ? 2 * desc_.halo : 0 );
          const int i = int( r % std::size_t( extent ) ) + desc_.lo[ d ] - ( d == dim - 1 ? desc_.halo : 0 );
          r /= std::size_t( extent );
          const double lower = desc_.origin[ d ] + i * desc_.h[ d ];
          x[ d ] = ( d == ( face >> 1 ) ) ? ( ( face & 1 ) ? lower + desc_.h[ d ] : lower ) : lower + 0.5 * desc_.h[ d ];
        }
        return x;
      }

      // sample a velocity functor with the reference's concept (bind( intersection ), operator()( local ), unbind();
      // velocity.hh:15-24) at every face centre of every cell of the grid view and hand it to the device: what the
      // operator-level Evolution needs, whose caller supplies flags it computed itself.  (Algorithm samples the faces of
      // the band only, see there.)
      template< class Velocity >
      void setVelocity ( Velocity &velocity )
      {
        faceVelocity.assign( cells_ * 2 * dim * dim, 0.0 );
        for( const auto &entity : elements( gridView_ ) )
        {
          const std::size_t cell = toLex_[ gridView_.indexSet().index( entity ) ];
          if( cell == npos )
            continue;
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
      void connect ()
      {
        const auto &comm = gridView_.comm();
        using Comm = std::decay_t< decltype( comm ) >;
        if constexpr ( Impl::HasAllgather< Comm >::value )
        {
          std::vector< char > mine( VOFX_COMM_HANDLE_BYTES ), all( std::size_t( VOFX_COMM_HANDLE_BYTES ) * world_ );
          check( vofx_comm_export( ctx_, mine.data() ) );
          comm.allgather( mine.data(), VOFX_COMM_HANDLE_BYTES, all.data() );
          std::vector< std::int32_t > layers( world_ );
          std::int32_t myLayers = desc_.n_local[ dim - 1 ];
          comm.allgather( &myLayers, 1, layers.data() );
          check( vofx_comm_connect( ctx_, rank_, world_, layers.data(), all.data() ) );
          comm.barrier();
        }
        else
          throw Exception( "vofx: the grid view's communication object has no allgather: cannot connect the ranks' device windows" );
      }

      GridView gridView_;
      vofx_grid_desc desc_;
      vofx_ctx *ctx_ = nullptr;
      std::size_t cells_ = 0;
      int world_ = 1, rank_ = 0;
      std::vector< std::size_t > toLex_;
      std::vector< std::uint8_t > interior_;
    };


    // one context per grid view (keyed by its index set), shared by every operator constructed on it: the reference's
    // operators are constructed from an eps, a grid view or a stencil set (flagging.hh:30, evolution.hh:26,
    // modifiedyoungs.hh:48) -- none of which can carry a device context, so it is looked up
    template< class GV >
    struct ContextRegistry
    {
      static std::shared_ptr< Context< GV > > get ( const GV &gridView, int device = -1 )
      {
        static std::mutex mutex;
        static std::map< const void *, std::weak_ptr< Context< GV > > > contexts;
        std::unique_lock< std::mutex > lock( mutex );
        const void *key = static_cast< const void * >( &gridView.indexSet() );
        auto sp = contexts[ key ].lock();
        if( sp )
          return sp;
        lock.unlock();        // the constructor of a parallel context communicates: never under the lock
        sp = std::make_shared< Context< GV > >( gridView, device );
        lock.lock();
        contexts[ key ] = sp;
        return sp;
      }
    };


    // VertexNeighborsStencil
    // ----------------------
    // In the reference this stores >= 26 entities per cell (stencil/vertexneighborsstencil.hh:29,94-97); on a
    // structured grid the device derives the neighbours from the index, so the object only holds the device context of
    // its grid view.
    template< class GV >
    struct VertexNeighborsStencil
    {
      using GridView = GV;

      explicit VertexNeighborsStencil ( const GridView &gridView, int device = -1 )
      : context_( ContextRegistry< GridView >::get( gridView, device ) )
      {}

      const GridView &gridView () const { return context_->gridView(); }
      Context< GridView > &context () const { return *context_; }

    private:
      std::shared_ptr< Context< GridView > > context_;
    };


    // FlagOperator, flagging.hh:24-74.  Constructed from eps alone, as in the reference (flagging.hh:30): the context is
    // the one of the colour function's grid view.
    template< class GV >
    struct FlagOperator
    {
      using GridView = GV;

      explicit FlagOperator ( double eps ) : eps_( eps ) {}
      FlagOperator ( const VertexNeighborsStencil< GV > &stencils, double eps ) : context_( ContextRegistry< GV >::get( stencils.gridView() ) ), eps_( eps ) {}

      template< class ColorFunction, class FlagSet >
      void operator() ( const ColorFunction &color, FlagSet &flags, bool communicate = true ) const
      {
        if( !context_ )
          context_ = ContextRegistry< GV >::get( color.gridView() );
        auto &c = *context_;
        c.gatherScalars( color, c.color );
        c.flags.assign( c.cells(), 0 );
        check( vofx_flag_host( c.ctx(), c.color.data(), eps_, c.flags.data() ) );
        c.scatterFlags( flags, communicate );
      }

    private:
      mutable std::shared_ptr< Context< GV > > context_;
      const double eps_;
    };


    namespace Impl
    {
      template< class GV, int kind >
      struct Reconstruction
      {
        using GridView = GV;

        // any stencil set with gridView(): this header's VertexNeighborsStencil or the reference's own
        // (stencil/vertexneighborsstencil.hh:25-98), whose entity lists are simply not read
        template< class StencilSet >
        explicit Reconstruction ( const StencilSet &stencils, std::size_t maxIterations = 10 )
        : context_( ContextRegistry< GV >::get( stencils.gridView() ) ), maxIterations_( int( maxIterations ) )
        {}

        // ( stencils, initializer[, maxIterations] ) of heightfunction.hh:57-60 / modifiedswartz.hh:54-56: the initial
        // reconstruction is always modified Young's here (as in the factory, reconstruction.hh:29-37)
        template< class StencilSet, class Initializer, class = std::enable_if_t< !std::is_integral< Initializer >::value > >
        Reconstruction ( const StencilSet &stencils, const Initializer &, std::size_t maxIterations = 10 )
        : context_( ContextRegistry< GV >::get( stencils.gridView() ) ), maxIterations_( int( maxIterations ) )
        {}

        template< class ColorFunction, class ReconstructionSet, class Flags >
        void operator() ( const ColorFunction &color, ReconstructionSet &reconstructions, const Flags &flags, bool communicate = true ) const
        {
          auto &c = *context_;
          c.gatherScalars( color, c.color );
          c.gatherFlags( flags );
          c.halfspaces.assign( c.cells() * ( GV::dimension + 1 ), 0.0 );
          check( vofx_reconstruct_host( c.ctx(), kind, c.color.data(), c.flags.data(), c.halfspaces.data(), maxIterations_ ) );
          c.scatterReconstructions( reconstructions, communicate );
        }

      protected:
        std::shared_ptr< Context< GV > > context_;
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


    // Evolution, evolution/evolution.hh:37-181.  Constructed from the grid view, as in the reference (evolution.hh:48).
    template< class GV >
    struct Evolution
    {
      using GridView = GV;

      explicit Evolution ( const GridView &gridView ) : context_( ContextRegistry< GV >::get( gridView ) ) {}
      explicit Evolution ( const VertexNeighborsStencil< GV > &stencils ) : context_( ContextRegistry< GV >::get( stencils.gridView() ) ) {}

      // velocity: the reference's concept (velocity.hh:5-33), frozen at the start-of-step time by its constructor
      template< class ReconstructionSet, class Flags, class Velocity, class DiscreteFunction >
      double operator() ( const ReconstructionSet &reconstructions, const Flags &flags, Velocity &velocity, double deltaT, DiscreteFunction &update ) const
      {
        auto &c = *context_;
        c.setVelocity( velocity );
        c.gatherReconstructions( reconstructions );
        c.gatherFlags( flags );
        c.update.assign( c.cells(), 0.0 );
        double dtEst = 0.0;
        check( vofx_evolve_host( c.ctx(), c.halfspaces.data(), c.flags.data(), 0.0, deltaT, c.update.data(), &dtEst ) );
        // update.communicate( All_All_Interface, Add ) of evolution.hh:73 is not needed: every rank gathers the complete update of
        // its interior cells (owner-computes); the overlap copies are refreshed instead
        c.scatterScalars( c.update, update );
        return c.parallel() ? c.gridView().comm().min( dtEst ) : dtEst;     // evolution.hh:75
      }

    private:
      std::shared_ptr< Context< GV > > context_;
    };

    // evolution.hh:25-30
    template< class GV >
    static inline Evolution< GV > evolution ( const GV &gridView )
    {
      return Evolution< GV >( gridView );
    }
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

      explicit CartesianHeightFunctionCurvature ( const VNST &stencils ) : context_( ContextRegistry< GV >::get( stencils.gridView() ) ) {}

      template< class ColorFunction, class ReconstructionSet, class Flags, class CurvatureSet >
      void operator() ( const ColorFunction &color, const ReconstructionSet &reconstructions, const Flags &flags, CurvatureSet &curvature,
                        bool communicate = true ) const
      {
        auto &c = *context_;
        c.gatherScalars( color, c.color );
        c.gatherReconstructions( reconstructions );
        c.gatherFlags( flags );
        c.kappa.assign( c.cells(), 0.0 );
        c.valid.assign( c.cells(), 0 );
        check( vofx_curvature_host( c.ctx(), c.color.data(), c.halfspaces.data(), c.flags.data(), c.kappa.data(), c.valid.data() ) );
        c.scatterScalars( c.kappa, curvature, communicate );
      }

      // satisfiesConstraint_ of the last call (heightfunction.hh:286), storage order
      const std::vector< std::uint8_t > &satisfiesConstraint () const { return context_->valid; }

    private:
      std::shared_ptr< Context< GV > > context_;
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
        c.kappa.assign( c.cells(), 0.0 );
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

      // l1error( gridView, reconstructions, flags, problem, time ) of test/errors.hh: evaluated on the device for problems
      // whose exact interface is a circle ( center( t ), radius( t ): RotatingCircle< double, 2 >, errors.hh:62-95 ) from the
      // flags and reconstructions of the pass that just ran; 0 for every other problem (the quadrature variant,
      // errors.hh:27-58, needs dune-geometry's rules and stays on the host)
      template< class Problem, int dim, class = void >
      struct L1Error
      {
        template< class Ctx >
        static double apply ( Ctx &, const Problem &, double ) { return 0.0; }
      };
      template< class Problem >
      struct L1Error< Problem, 2, std::void_t< decltype( std::declval< const Problem & >().center( 0.0 ) ), decltype( std::declval< const Problem & >().radius( 0.0 ) ) > >
      {
        template< class Ctx >
        static double apply ( Ctx &c, const Problem &problem, double time )
        {
          const auto center = problem.center( time );
          const double xy[ 2 ] = { double( center[ 0 ] ), double( center[ 1 ] ) };
          double error = 0.0;
          check( vofx_l1error_circle( c.ctx(), xy, double( problem.radius( time ) ), &error ) );
          return error;
        }
      };
    } // namespace Impl

    template< class GV, class PR, class DW >
    struct Algorithm
    {
      using GridView = GV;
      using Problem = PR;
      using DataWriter = DW;
      using Stencils = VertexNeighborsStencil< GridView >;
      static constexpr int dim = GridView::dimension;

      // the reference's signature (algorithm.hh:46) plus two optional arguments: the reconstruction (the reference
      // hard-wires the factory's height functions, reconstruction.hh:29-37) and a built-in velocity (enum vofx_problem)
      Algorithm ( const GridView &gridView, const Problem &problem, DataWriter &dataWriter, double cfl, double eps, const bool verbose = false,
                  int reconstruction = VOFX_RECON_HF, int builtinVelocity = -1 )
      : gridView_( gridView ), problem_( problem ), dataWriter_( dataWriter ), cfl_( cfl ), eps_( eps ), verbose_( verbose ),
        stencils_( gridView ), reconstruction_( reconstruction ), builtinVelocity_( builtinVelocity )
      {}

      // builtinVelocity >= 0: the separable device tables are used.  Otherwise problem.velocityField (what the reference's
      // Velocity< Problem, GridView > evaluates, velocity.hh:15-20) is sampled on the host at the face centres of the
      // MIXED cells of each pass only: the pass is issued in two halves (flags | everything else) and the band list is
      // read back in between -- 2*dim*dim doubles per mixed cell travel to the device instead of per cell of the grid.
      // Returns the accumulated dt * l1error of the reference (algorithm.hh:85-98) for problems with an exact interface
      // the device knows (RotatingCircle< double, 2 >: circle.center( t ), circle.radius( t )), else 0.
      template< class ColorFunction >
      double operator() ( ColorFunction &uh, double start, double end, int /*level*/ = 0 )
      {
        auto &c = stencils_.context();
        vofx_step_params params{};
        params.reconstruction = reconstruction_;
        params.max_iterations = 10;
        params.eps = eps_;
        params.cfl = cfl_;
        params.with_curvature = 0;
        if( builtinVelocity_ >= 0 )
          check( vofx_velocity_builtin( c.ctx(), builtinVelocity_ ) );

        c.gatherScalars( uh, c.color );
        check( vofx_color_upload( c.ctx(), c.color.data() ) );
        double time = start, dt = 0.0, dtEst = 0.0, error = 0.0;
        check( vofx_step_begin( c.ctx(), time, dt ) );
        if( c.parallel() )
        {
          // the halo layers of the device copies are filled from the neighbours' interior values: two barriers per run
          gridView_.comm().barrier();
          check( vofx_halo_push( c.ctx() ) );
          gridView_.comm().barrier();
        }
        do
        {
          check( vofx_step_distributed_phase( c.ctx(), &params, 0 ) );          // flagOperator( uh, flags_ ), algorithm.hh:75
          if( builtinVelocity_ < 0 )
            sampleBandVelocity( c, time );                                       // VelocityField velocity( problem_, time ), :73
          check( vofx_step_distributed_phase( c.ctx(), &params, 1 ) );          // reconstruction, evolution, axpy, :78-83
          const double passTime = time, passDt = dt;
          check( vofx_step_state( c.ctx(), &time, &dt, &dtEst, nullptr ) );
          if( passDt > 0.0 )
            error += passDt * Impl::L1Error< Problem, dim >::apply( c, problem_, passTime );      // :85-89
          if( Impl::WillWrite< DataWriter >::apply( dataWriter_, time ) )
          {
            check( vofx_color_download( c.ctx(), c.color.data(), nullptr ) );
            c.scatterScalars( c.color, uh );
            dataWriter_.write( time );
          }
        }
        while( time < end );
        if( time == start )
          error += Impl::L1Error< Problem, dim >::apply( c, problem_, time );                     // :97-98
        c.flags.assign( c.cells(), 0 );
        check( vofx_color_download( c.ctx(), c.color.data(), c.flags.data() ) );
        c.scatterScalars( c.color, uh );
        return c.parallel() ? gridView_.comm().sum( error ) : error;
      }

    private:
      void sampleBandVelocity ( Context< GridView > &c, double time )
      {
        using Domain = typename Problem::DomainType;
        std::int64_t n = 0;
        check( vofx_band_list_fetch_in_flight( c.ctx(), nullptr, 0, &n ) );
        c.bandCells.resize( std::size_t( n ) + 1 );
        if( n > 0 )
          check( vofx_band_list_fetch_in_flight( c.ctx(), c.bandCells.data(), n, &n ) );
        c.faceVelocity.assign( std::size_t( n ) * 2 * dim * dim + 1, 0.0 );
        for( std::int64_t i = 0; i < n; ++i )
          for( int face = 0; face < 2 * dim; ++face )
          {
            const Domain x = c.template faceCenter< Domain >( std::size_t( c.bandCells[ std::size_t( i ) ] ), face );
            Domain v( 0.0 );
            problem_.velocityField( x, time, v );
            for( int d = 0; d < dim; ++d )
              c.faceVelocity[ ( std::size_t( i ) * 2 * dim + face ) * dim + d ] = v[ d ];
          }
        check( vofx_velocity_set_band_faces( c.ctx(), n, c.faceVelocity.data() ) );
      }

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
