// Test of the header adaptor on a PARALLEL grid view (gridView.comm().size() > 1): the ranks are threads of this process,
// each with a grid view of its own slab (interior layers + overlap, the cells a DUNE overlap decomposition gives a rank) and
// a collective-communication object with min / max / sum / allgather / barrier.  Every rank runs Dune::VoFx::Algorithm on its
// slab exactly as example/vof.cc runs the reference's; the interior values of all ranks must carry the bits of the oracle's
// run on the whole grid.  The device contexts connect through vofx_comm_export / vofx_comm_connect (include/vofx.h): with
// threads of one process the windows are plain pointers, with MPI ranks CUDA IPC handles -- the adaptor code is the same.
// (Needs a GPU at run time; all contexts sit on device 0.)
#include <algorithm>
#include <array>
#include <condition_variable>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <dune/common/fvector.hh>
#include <dune/grid/common/minigrid.hh>
#include <dune/vofx/vofx.hh>

extern "C"
{
#include "vof_oracle.h"
}

namespace
{

  // ---- collective communication among threads ---------------------------------------------------------------------
  struct ThreadWorld
  {
    explicit ThreadWorld ( int n ) : size( n ), slots( std::size_t( n ) ) {}
    int size;
    std::mutex mutex;
    std::condition_variable cv;
    int arrived = 0;
    long generation = 0;
    std::vector< std::vector< char > > slots;

    void barrier ()
    {
      std::unique_lock< std::mutex > lock( mutex );
      const long gen = generation;
      if( ++arrived == size )
      {
        arrived = 0;
        ++generation;
        cv.notify_all();
      }
      else
        cv.wait( lock, [ & ] { return generation != gen; } );
    }
  };

  struct ThreadComm
  {
    ThreadWorld *world = nullptr;
    int myRank = 0;
    int rank () const { return myRank; }
    int size () const { return world->size; }
    void barrier () const { world->barrier(); }
    template< class T >
    void allgather ( const T *in, int count, T *out ) const
    {
      world->slots[ std::size_t( myRank ) ].assign( reinterpret_cast< const char * >( in ), reinterpret_cast< const char * >( in + count ) );
      world->barrier();
      for( int r = 0; r < world->size; ++r )
        std::memcpy( out + std::size_t( r ) * count, world->slots[ std::size_t( r ) ].data(), sizeof( T ) * std::size_t( count ) );
      world->barrier();
    }
    template< class T, class Op >
    T reduce ( T x, Op op ) const
    {
      std::vector< T > all( std::size_t( world->size ) );
      allgather( &x, 1, all.data() );
      T r = all[ 0 ];
      for( int k = 1; k < world->size; ++k )
        r = op( r, all[ std::size_t( k ) ] );
      return r;
    }
    template< class T > T min ( T x ) const { return reduce( x, [] ( T a, T b ) { return std::min( a, b ); } ); }
    template< class T > T max ( T x ) const { return reduce( x, [] ( T a, T b ) { return std::max( a, b ); } ); }
    template< class T > T sum ( T x ) const { return reduce( x, [] ( T a, T b ) { return a + b; } ); }
  };

  // ---- the grid view of one rank: layers [ a0, a1 ) of the global box along the last axis, of which [ k0, k1 ) are interior --
  template< int dim >
  struct SlabGrid
  {
    std::array< int, dim > n;                 // global cells per axis
    Dune::FieldVector< double, dim > lo, h;
    int a0, a1, k0, k1;
    std::size_t layerCells () const
    {
      std::size_t c = 1;
      for( int d = 0; d + 1 < dim; ++d )
        c *= std::size_t( n[ d ] );
      return c;
    }
    std::size_t cells () const { return layerCells() * std::size_t( a1 - a0 ); }
  };

  template< int dim >
  struct SlabEntity
  {
    using Geometry = Dune::BoxGeometry< dim, dim >;
    static constexpr int codimension = 0;
    const SlabGrid< dim > *g = nullptr;
    std::array< int, dim > ijk {};            // GLOBAL multi-index
    Geometry geometry () const
    {
      Geometry q;
      q.h = g->h;
      for( int d = 0; d < dim; ++d )
      {
        q.o[ d ] = g->lo[ d ] + ijk[ d ] * g->h[ d ];
        q.axes[ d ] = d;
      }
      return q;
    }
  };

  template< int dim >
  struct SlabIt
  {
    const SlabGrid< dim > *g;
    std::size_t i;
    SlabEntity< dim > operator* () const
    {
      SlabEntity< dim > e;
      e.g = g;
      std::size_t k = i;
      for( int d = 0; d < dim; ++d )
      {
        const std::size_t extent = ( d + 1 < dim ) ? std::size_t( g->n[ d ] ) : std::size_t( g->a1 - g->a0 );
        e.ijk[ d ] = int( k % extent ) + ( d + 1 < dim ? 0 : g->a0 );
        k /= extent;
      }
      return e;
    }
    SlabIt &operator++ () { ++i; return *this; }
    bool operator!= ( const SlabIt &o ) const { return i != o.i; }
  };

  template< int dim >
  struct SlabRange
  {
    const SlabGrid< dim > *g;
    std::size_t first, last;
    SlabIt< dim > begin () const { return { g, first }; }
    SlabIt< dim > end () const { return { g, last }; }
  };

  template< int dim >
  struct SlabIndexSet
  {
    const SlabGrid< dim > *g;
    std::size_t size ( int ) const { return g->cells(); }
    std::size_t index ( const SlabEntity< dim > &e ) const
    {
      std::size_t k = std::size_t( e.ijk[ dim - 1 ] - g->a0 );
      for( int d = dim - 2; d >= 0; --d )
        k = k * std::size_t( g->n[ d ] ) + std::size_t( e.ijk[ d ] );
      return k;
    }
  };

  template< int dim >
  struct SlabGridView
  {
    static constexpr int dimension = dim, dimensionworld = dim;
    using ctype = double;
    using IndexSet = SlabIndexSet< dim >;
    const SlabGrid< dim > *g;
    IndexSet is;
    ThreadComm c;
    SlabGridView ( const SlabGrid< dim > &grid, ThreadComm comm ) : g( &grid ), is { &grid }, c( comm ) {}
    const IndexSet &indexSet () const { return is; }
    const ThreadComm &comm () const { return c; }
  };

  template< int dim > SlabRange< dim > elements ( const SlabGridView< dim > &gv ) { return { gv.g, 0, gv.g->cells() }; }
  template< int dim > SlabRange< dim > elements ( const SlabGridView< dim > &gv, Dune::Partitions::Interior )
  {
    return { gv.g, gv.g->layerCells() * std::size_t( gv.g->k0 - gv.g->a0 ), gv.g->layerCells() * std::size_t( gv.g->k1 - gv.g->a0 ) };
  }

  // host container with the interface of dataset.hh:26-87 (no communicate(): only interior values are compared)
  template< class GV >
  struct ColorFunction
  {
    using GridView = GV;
    explicit ColorFunction ( const GV &gv ) : gridView_( gv ), data_( gv.indexSet().size( 0 ), 0.0 ) {}
    const double &operator[] ( std::size_t i ) const { return data_[ i ]; }
    double &operator[] ( std::size_t i ) { return data_[ i ]; }
    std::size_t size () const { return data_.size(); }
    const GV &gridView () const { return gridView_; }
    GV gridView_;
    std::vector< double > data_;
  };

  // problems with the reference's Problem concept
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

  struct NullWriter
  {
    bool willWrite ( double ) { return false; }
    void write ( double ) {}
  };

  std::mutex printMutex;
  int failures = 0;
  void expect ( bool ok, const std::string &what )
  {
    std::lock_guard< std::mutex > lock( printMutex );
    std::printf( "%s %s\n", ok ? "ok  " : "FAIL", what.c_str() );
    if( !ok )
      ++failures;
  }

  template< int dim >
  void runCase ( const char *name, const std::array< int, dim > &n, const std::vector< int > &layers, int overlap, int oracleProblem, int passes, bool hostVelocity )
  {
    const int world = int( layers.size() );
    // the oracle on the whole grid
    vo_grid og;
    std::memset( &og, 0, sizeof( og ) );
    og.dim = dim;
    std::size_t N = 1;
    for( int d = 0; d < 3; ++d )
    {
      og.n[ d ] = d < dim ? n[ d ] : 1;
      og.h[ d ] = d < dim ? 1.0 / n[ d ] : 1.0;
      N *= std::size_t( og.n[ d ] );
    }
    std::vector< double > u0( N ), uo( N );
    vo_init( &og, oracleProblem, 0.0, u0.data() );
    uo = u0;
    double timeO = 0.0, dtO = 0.0, phases[ 4 ];
    int64_t mixed = 0;
    vo_run( &og, VO_RECON_YOUNGS, oracleProblem, 1e-6, 0.5, 10, passes, uo.data(), &timeO, &dtO, phases, &mixed );

    ThreadWorld worldState( world );
    std::vector< int > okRank( static_cast< std::size_t >( world ), 0 );
    std::vector< std::string > errors{ static_cast< std::size_t >( world ), std::string() };
    std::vector< std::thread > threads;
    for( int rank = 0; rank < world; ++rank )
      threads.emplace_back( [ &, rank ] {
        try
        {
          SlabGrid< dim > grid;
          grid.n = n;
          for( int d = 0; d < dim; ++d )
          {
            grid.lo[ d ] = 0.0;
            grid.h[ d ] = 1.0 / n[ d ];
          }
          grid.k0 = 0;
          for( int r = 0; r < rank; ++r )
            grid.k0 += layers[ std::size_t( r ) ];
          grid.k1 = grid.k0 + layers[ std::size_t( rank ) ];
          grid.a0 = std::max( 0, grid.k0 - overlap );
          grid.a1 = std::min( n[ dim - 1 ], grid.k1 + overlap );
          using GV = SlabGridView< dim >;
          GV gv( grid, ThreadComm{ &worldState, rank } );
          ColorFunction< GV > uh( gv );
          const std::size_t layer = grid.layerCells();
          for( std::size_t i = 0; i < uh.size(); ++i )
            uh[ i ] = u0[ std::size_t( grid.a0 ) * layer + i ];
          SFlowField< dim > problem;
          NullWriter writer;
          {
            Dune::VoFx::Algorithm< GV, SFlowField< dim >, NullWriter > algorithm( gv, problem, writer, 0.5, 1e-6, false, VOFX_RECON_YOUNGS,
                                                                                  hostVelocity ? -1 : oracleProblem );
            algorithm( uh, 0.0, timeO, 0 );
          }
          bool ok = true;
          for( std::size_t i = std::size_t( grid.k0 - grid.a0 ) * layer; i < std::size_t( grid.k1 - grid.a0 ) * layer; ++i )
          {
            const double a = uh[ i ], b = uo[ std::size_t( grid.a0 ) * layer + i ];
            ok = ok && ( std::memcmp( &a, &b, sizeof( double ) ) == 0 || ( a == 0.0 && b == 0.0 ) );
          }
          okRank[ std::size_t( rank ) ] = ok ? 1 : 0;
        }
        catch( const std::exception &e )
        {
          errors[ std::size_t( rank ) ] = e.what();
        }
      } );
    for( auto &t : threads )
      t.join();
    bool ok = true;
    for( int rank = 0; rank < world; ++rank )
    {
      ok = ok && okRank[ std::size_t( rank ) ] == 1;
      if( !errors[ std::size_t( rank ) ].empty() )
        std::printf( "     rank %d: %s\n", rank, errors[ std::size_t( rank ) ].c_str() );
    }
    expect( ok, std::string( name ) + ( hostVelocity ? " [host velocity functor on the band]" : " [built-in velocity]" )
                  + ": Algorithm on " + std::to_string( world ) + " ranks == oracle on the whole grid (bit-exact)" );
  }

} // namespace

int main ()
{
  // ranks as THREADS of one process share one CUDA context: a kernel that is loaded lazily at its first launch can make that
  // launch wait for the device while another thread's kernel spins on a signal only this thread can send.  MPI ranks are
  // processes and never meet this; here every kernel is loaded up front.
  setenv( "CUDA_MODULE_LOADING", "EAGER", 1 );
  // ... and every rank's stream needs a hardware queue of its own (streams that share one are served in order: a waiting
  // kernel of one rank would block the kernels of another that it waits for)
  setenv( "CUDA_DEVICE_MAX_CONNECTIONS", "32", 1 );
  runCase< 2 >( "2D 48x64 SFlow", { 48, 64 }, { 30, 34 }, 2, VO_PROB_SFLOW, 8, false );
  runCase< 2 >( "2D 48x64 SFlow", { 48, 64 }, { 30, 34 }, 2, VO_PROB_SFLOW, 8, true );
  runCase< 3 >( "3D 20x12x32 SFlow", { 20, 12, 32 }, { 10, 12, 10 }, 1, VO_PROB_SFLOW, 5, false );
  runCase< 3 >( "3D 20x12x32 SFlow", { 20, 12, 32 }, { 10, 12, 10 }, 1, VO_PROB_SFLOW, 5, true );
  std::printf( failures ? "ADAPTOR PARALLEL TEST FAILED (%d)\n" : "ADAPTOR PARALLEL TEST PASSED\n", failures );
  return failures ? 1 : 0;
}
