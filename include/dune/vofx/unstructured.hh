// vofx header adaptor for UNSTRUCTURED 2D grid views (triangles and quadrilaterals, e.g. ALUGrid / UGGrid leaf views).
//
// Same roles as include/dune/vofx/vofx.hh, in namespace Dune::VoFx::Unstructured: the stencil object flattens the grid
// view ONCE into the arrays of vofx_mesh_desc (include/vofx.h) - cells in leaf-index order, intersections in the grid
// view's iteration order, all geometry copied verbatim from the grid (corner, center, volume, centerUnitOuterNormal,
// integrationOuterNormal) - and builds the CSR vertex-neighbour stencil the way
// stencil/vertexneighborsstencil.hh:52-94 does (cells sharing a vertex, sorted, unique, self excluded).  The operators
// then call the C ABI (vofx_mesh_*_host):
//
//   reference                                                    here
//   VertexNeighborsStencil< GV >( gv )                           Unstructured::VertexNeighborsStencil< GV >( gv )
//   FlagOperator< GV >( eps )( uh, flags )                       Unstructured::FlagOperator< GV >( stencils, eps )( uh, flags )
//   ModifiedYoungsReconstruction< GV, StS >( st )( uh, r, f )    Unstructured::ModifiedYoungsReconstruction< GV >( stencils )( uh, r, f )
//   ModifiedSwartzReconstruction< GV, StS, IR >( st, n )(..)     Unstructured::ModifiedSwartzReconstruction< GV >( stencils, n )( uh, r, f )
//   evolution( gv )( recs, flags, velocity, dt, update )         Unstructured::evolution( stencils )( recs, flags, velocity, dt, update )
//
// Containers: anything with the interface of dune/vof/dataset.hh:26-87 (begin()/end() in leaf-index order, operator[]( index )).
// Height functions and their curvature are Cartesian-only in the reference as well (heightfunctionstencil.hh reaches
// into SPGrid); they are not offered here.
#ifndef DUNE_VOFX_UNSTRUCTURED_HH
#define DUNE_VOFX_UNSTRUCTURED_HH

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

extern "C"
{
#include <vofx.h>
}

namespace Dune
{
  namespace VoFx
  {
    namespace Unstructured
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

      // the flattened grid view and the device object built from it
      template< class GV >
      class Context
      {
      public:
        using GridView = GV;
        static constexpr int D = GridView::dimension;
        static_assert( D == 2 || D == 3, "the unstructured path is built for 2D (triangles, quadrilaterals) and 3D (tetrahedra, hexahedra) grids" );

        explicit Context ( const GridView &gridView, int device = -1 )
        : gridView_( gridView )
        {
          const auto &indexSet = gridView_.indexSet();
          const std::size_t n = indexSet.size( 0 );
          std::vector< int > nCorners( n, 0 ), nFaces( n, 0 );
          for( const auto &entity : elements( gridView_ ) )
          {
            const std::size_t i = indexSet.index( entity );
            nCorners[ i ] = entity.geometry().corners();
            for( const auto &intersection : intersections( gridView_, entity ) )
            {
              (void) intersection;
              ++nFaces[ i ];
            }
          }
          cornerOffsets.assign( n + 1, 0 );
          faceOffsets.assign( n + 1, 0 );
          for( std::size_t i = 0; i < n; ++i )
          {
            cornerOffsets[ i + 1 ] = cornerOffsets[ i ] + nCorners[ i ];
            faceOffsets[ i + 1 ] = faceOffsets[ i ] + nFaces[ i ];
          }
          const std::size_t nc = std::size_t( cornerOffsets[ n ] ), nf = std::size_t( faceOffsets[ n ] );
          corners.resize( D * nc );
          centers.resize( D * n );
          volumes.resize( n );
          faceNeighbor.assign( nf, -1 );
          faceCornerOffsets.assign( nf + 1, 0 );
          faceCorners.clear();
          faceCenters.resize( D * nf );
          faceUnitNormals.resize( D * nf );
          faceIntNormals.resize( D * nf );
          faceVolumes.resize( nf );

          std::vector< std::vector< std::int32_t > > cellsOfVertex( indexSet.size( D ) );
          for( const auto &entity : elements( gridView_ ) )
          {
            const std::size_t i = indexSet.index( entity );
            const auto geometry = entity.geometry();
            for( int k = 0; k < nCorners[ i ]; ++k )
            {
              const auto x = geometry.corner( k );
              for( int d = 0; d < D; ++d )
                corners[ D * ( cornerOffsets[ i ] + k ) + d ] = x[ d ];
              cellsOfVertex[ indexSet.subIndex( entity, k, D ) ].push_back( std::int32_t( i ) );
            }
            const auto center = geometry.center();
            for( int d = 0; d < D; ++d )
              centers[ D * i + d ] = center[ d ];
            volumes[ i ] = geometry.volume();
            std::size_t s = std::size_t( faceOffsets[ i ] );
            for( const auto &intersection : intersections( gridView_, entity ) )
            {
              if( intersection.neighbor() )
                faceNeighbor[ s ] = std::int32_t( indexSet.index( intersection.outside() ) );
              const auto ig = intersection.geometry();
              const auto c = ig.center();
              const auto un = intersection.centerUnitOuterNormal();
              // refElement.position( 0, 0 ): the centre of the intersection's reference element (evolution.hh:107-108)
              typename std::decay_t< decltype( intersection ) >::LocalCoordinate local( 0.5 );
              if( D == 3 && ig.corners() == 3 )
                local = 1.0 / 3.0;
              const auto in = intersection.integrationOuterNormal( local );
              const int nic = ig.corners();          // 2 in 2D; 3 (triangle) or 4 (quadrilateral) in 3D
              faceCornerOffsets[ s + 1 ] = faceCornerOffsets[ s ] + nic;
              for( int k = 0; k < nic; ++k )
              {
                const auto x = ig.corner( k );
                for( int d = 0; d < D; ++d )
                  faceCorners.push_back( x[ d ] );
              }
              for( int d = 0; d < D; ++d )
              {
                faceCenters[ D * s + d ] = c[ d ];
                faceUnitNormals[ D * s + d ] = un[ d ];
                faceIntNormals[ D * s + d ] = in[ d ];
              }
              faceVolumes[ s ] = ig.volume();
              ++s;
            }
          }

          // CSR vertex-neighbour stencil
          stencilOffsets.assign( n + 1, 0 );
          std::vector< std::int32_t > scratch;
          for( const auto &entity : elements( gridView_ ) )
          {
            const std::size_t i = indexSet.index( entity );
            scratch.clear();
            for( int k = 0; k < nCorners[ i ]; ++k )
              for( std::int32_t other : cellsOfVertex[ indexSet.subIndex( entity, k, D ) ] )
                if( std::size_t( other ) != i )
                  scratch.push_back( other );
            std::sort( scratch.begin(), scratch.end() );
            scratch.erase( std::unique( scratch.begin(), scratch.end() ), scratch.end() );
            perCell_.push_back( { i, scratch } );
          }
          std::sort( perCell_.begin(), perCell_.end(), [] ( const auto &l, const auto &r ) { return l.first < r.first; } );
          for( const auto &entry : perCell_ )
          {
            stencilOffsets[ entry.first + 1 ] = stencilOffsets[ entry.first ] + std::int32_t( entry.second.size() );
            stencil.insert( stencil.end(), entry.second.begin(), entry.second.end() );
          }
          perCell_.clear();
          perCell_.shrink_to_fit();

          vofx_mesh_desc desc{};
          desc.dim = D;
          desc.n_cells = std::int64_t( n );
          desc.corner_offsets = cornerOffsets.data();
          desc.corners = corners.data();
          desc.centers = centers.data();
          desc.volumes = volumes.data();
          desc.face_offsets = faceOffsets.data();
          desc.face_neighbor = faceNeighbor.data();
          desc.face_corners = faceCorners.data();
          desc.face_centers = faceCenters.data();
          desc.face_unit_normals = faceUnitNormals.data();
          desc.face_int_normals = faceIntNormals.data();
          desc.face_volumes = faceVolumes.data();
          desc.stencil_offsets = stencilOffsets.data();
          desc.stencil = stencil.data();
          desc.face_corner_offsets = faceCornerOffsets.data();
          check( vofx_mesh_create( &desc, device, &mesh_ ) );
        }

        Context ( const Context & ) = delete;
        Context &operator= ( const Context & ) = delete;
        ~Context () { vofx_mesh_destroy( mesh_ ); }

        vofx_mesh *mesh () const { return mesh_; }
        const GridView &gridView () const { return gridView_; }
        std::size_t size () const { return volumes.size(); }
        std::size_t faces () const { return faceVolumes.size(); }

        // the flattened grid (leaf-index order)
        std::vector< std::int32_t > cornerOffsets, faceOffsets, faceNeighbor, faceCornerOffsets, stencilOffsets, stencil;
        std::vector< double > corners, centers, volumes, faceCorners, faceCenters, faceUnitNormals, faceIntNormals, faceVolumes;
        // staging
        std::vector< double > color, update, halfspaces, faceVelocity;
        std::vector< std::uint8_t > flags;

        template< class DataSet >
        void gatherScalars ( const DataSet &data, std::vector< double > &out ) const
        {
          out.resize( size() );
          for( std::size_t i = 0; i < size(); ++i )
            out[ i ] = data[ i ];
        }

        template< class FlagSet >
        void gatherFlags ( const FlagSet &flagSet )
        {
          flags.resize( size() );
          std::size_t i = 0;
          for( auto it = flagSet.begin(); it != flagSet.end(); ++it, ++i )
            flags[ i ] = std::uint8_t( static_cast< std::size_t >( *it ) );
        }

        template< class ReconstructionSet >
        void gatherReconstructions ( const ReconstructionSet &recs )
        {
          halfspaces.resize( ( D + 1 ) * size() );
          std::size_t i = 0;
          for( auto it = recs.begin(); it != recs.end(); ++it, ++i )
          {
            for( int d = 0; d < D; ++d )
              halfspaces[ ( D + 1 ) * i + d ] = it->innerNormal()[ d ];
            halfspaces[ ( D + 1 ) * i + D ] = it->boundary().distance();
          }
        }

        // the reference's Velocity concept (velocity.hh:5-33) sampled where Evolution::applyLocal evaluates it
        // (evolution.hh:106-108): at refElement.position( 0, 0 ) of every intersection
        template< class Velocity >
        void sampleVelocity ( Velocity &velocity )
        {
          faceVelocity.assign( D * faces(), 0.0 );
          const auto &indexSet = gridView_.indexSet();
          for( const auto &entity : elements( gridView_ ) )
          {
            std::size_t s = std::size_t( faceOffsets[ indexSet.index( entity ) ] );
            for( const auto &intersection : intersections( gridView_, entity ) )
            {
              velocity.bind( intersection );
              typename std::decay_t< decltype( intersection ) >::LocalCoordinate local( 0.5 );
              if( D == 3 && intersection.geometry().corners() == 3 )
                local = 1.0 / 3.0;             // refElement.position( 0, 0 ) of a triangle
              const auto v = velocity( local );
              velocity.unbind();
              for( int d = 0; d < D; ++d )
                faceVelocity[ D * s + d ] = v[ d ];
              ++s;
            }
          }
        }

      private:
        GridView gridView_;
        vofx_mesh *mesh_ = nullptr;
        std::vector< std::pair< std::size_t, std::vector< std::int32_t > > > perCell_;
      };


      // stencil/vertexneighborsstencil.hh:25-98: owns the flattened grid and the device object the operators share
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


      // flagging.hh:24-74
      template< class GV >
      struct FlagOperator
      {
        FlagOperator ( const VertexNeighborsStencil< GV > &stencils, double eps ) : context_( &stencils.context() ), eps_( eps ) {}

        template< class ColorFunction, class FlagSet >
        void operator() ( const ColorFunction &color, FlagSet &flags, bool /*communicate*/ = true ) const
        {
          auto &c = *context_;
          c.gatherScalars( color, c.color );
          c.flags.resize( c.size() );
          check( vofx_mesh_flag_host( c.mesh(), c.color.data(), eps_, c.flags.data() ) );
          using Flag = typename FlagSet::DataType;
          std::size_t i = 0;
          for( auto it = flags.begin(); it != flags.end(); ++it, ++i )
            *it = static_cast< Flag >( static_cast< std::size_t >( c.flags[ i ] ) );
        }

      private:
        Context< GV > *context_;
        const double eps_;
      };


      // reconstruction/modifiedyoungs.hh:29-139
      template< class GV, class StS = VertexNeighborsStencil< GV > >
      struct ModifiedYoungsReconstruction
      {
        explicit ModifiedYoungsReconstruction ( const StS &stencils ) : context_( &stencils.context() ) {}

        template< class ColorFunction, class ReconstructionSet, class Flags >
        void operator() ( const ColorFunction &color, ReconstructionSet &reconstructions, const Flags &flags, bool /*communicate*/ = true ) const
        {
          auto &c = *context_;
          c.gatherScalars( color, c.color );
          c.gatherFlags( flags );
          c.halfspaces.resize( ( Context< GV >::D + 1 ) * c.size() );
          check( vofx_mesh_reconstruct_host( c.mesh(), c.color.data(), c.flags.data(), c.halfspaces.data() ) );
          using HalfSpace = typename ReconstructionSet::DataType;
          using Coordinate = typename HalfSpace::Coordinate;
          std::size_t i = 0;
          for( auto it = reconstructions.begin(); it != reconstructions.end(); ++it, ++i )
          {
            Coordinate normal;
            constexpr int D = Context< GV >::D;
            for( int d = 0; d < D; ++d )
              normal[ d ] = c.halfspaces[ ( D + 1 ) * i + d ];
            *it = HalfSpace( normal, c.halfspaces[ ( D + 1 ) * i + D ] );
          }
        }

      private:
        Context< GV > *context_;
      };


      // reconstruction/modifiedswartz.hh:34-153 (2D): Young's initialiser, then the fixed-point iteration on the normal
      template< class GV, class StS = VertexNeighborsStencil< GV >, class IR = ModifiedYoungsReconstruction< GV, StS > >
      struct ModifiedSwartzReconstruction
      {
        explicit ModifiedSwartzReconstruction ( const StS &stencils, std::size_t maxIterations = 10 )
        : context_( &stencils.context() ), maxIterations_( maxIterations )
        {}

        template< class ColorFunction, class ReconstructionSet, class Flags >
        void operator() ( const ColorFunction &color, ReconstructionSet &reconstructions, const Flags &flags ) const
        {
          auto &c = *context_;
          c.gatherScalars( color, c.color );
          c.gatherFlags( flags );
          c.halfspaces.resize( ( Context< GV >::D + 1 ) * c.size() );
          check( vofx_mesh_reconstruct_swartz_host( c.mesh(), c.color.data(), c.flags.data(), c.halfspaces.data(), int( maxIterations_ ) ) );
          using HalfSpace = typename ReconstructionSet::DataType;
          using Coordinate = typename HalfSpace::Coordinate;
          std::size_t i = 0;
          for( auto it = reconstructions.begin(); it != reconstructions.end(); ++it, ++i )
          {
            Coordinate normal;
            constexpr int D = Context< GV >::D;
            for( int d = 0; d < D; ++d )
              normal[ d ] = c.halfspaces[ ( D + 1 ) * i + d ];
            *it = HalfSpace( normal, c.halfspaces[ ( D + 1 ) * i + D ] );
          }
        }

      private:
        Context< GV > *context_;
        std::size_t maxIterations_;
      };


      // evolution/evolution.hh:37-181
      template< class GV >
      struct Evolution
      {
        explicit Evolution ( const VertexNeighborsStencil< GV > &stencils ) : context_( &stencils.context() ) {}

        template< class ReconstructionSet, class Flags, class Velocity, class DiscreteFunction >
        double operator() ( const ReconstructionSet &reconstructions, const Flags &flags, Velocity &velocity, double deltaT, DiscreteFunction &update ) const
        {
          auto &c = *context_;
          c.sampleVelocity( velocity );
          c.gatherReconstructions( reconstructions );
          c.gatherFlags( flags );
          c.update.resize( c.size() );
          double dtEst = 0.0;
          check( vofx_mesh_evolve_host( c.mesh(), c.halfspaces.data(), c.flags.data(), c.faceVelocity.data(), deltaT, c.update.data(), &dtEst ) );
          for( std::size_t i = 0; i < c.size(); ++i )
            update[ i ] = c.update[ i ];
          return dtEst;
        }

      private:
        Context< GV > *context_;
      };

      // evolution.hh:25-30 (takes the stencil set instead of the grid view: it carries the device object)
      template< class GV >
      static inline Evolution< GV > evolution ( const VertexNeighborsStencil< GV > &stencils )
      {
        return Evolution< GV >( stencils );
      }

    } // namespace Unstructured
  } // namespace VoFx
} // namespace Dune

#endif // #ifndef DUNE_VOFX_UNSTRUCTURED_HH
