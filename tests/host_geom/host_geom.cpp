// TEST-ONLY host build of the product's DEVICE geometry headers (dune-vof_b200/csrc/vofx_geom{2d,3d}.cuh).
// The same functions run on the GPU in the product; compiling them with g++ lets the CPU test-suite check their
// logic bit-for-bit against the oracle where no GPU exists.  The product never loads this library.
#include <cstdint>
#include <cstring>

#include "vofx_geom2d.cuh"
#include "vofx_geom3d.cuh"
#include "vofx_coop3d.cuh"

using namespace vofx;

extern "C"
{

  // unstructured 3D grids: cells / intersections given by their corners (topology policies of vofx_geom3d.cuh)
  int hg_locate_corners ( int nc, const double *corners, const double *normal, double fraction, double *out )
  {
    for( int k = 0; k < 3; ++k )
      out[ k ] = normal[ k ];
    out[ 3 ] = locate_corners( corners, nc, normal, fraction );
    return 0;
  }

  int hg_flux_corners ( int nc, const double *corners, const double *v, const double *hs, double *out )
  {
    const Hs3 H = { { hs[ 0 ], hs[ 1 ], hs[ 2 ] }, hs[ 3 ] };
    *out = flux_corners( corners, nc, v, H );
    return 0;
  }

  int hg_locate ( int dim, const double *lo, const double *h, const double *normal, double fraction, double *out )
  {
    if( dim == 2 )
    {
      Poly2 p;
      make_cell( lo[ 0 ], lo[ 1 ], h[ 0 ], h[ 1 ], p );
      out[ 0 ] = normal[ 0 ];
      out[ 1 ] = normal[ 1 ];
      out[ 2 ] = locate( p, normal[ 0 ], normal[ 1 ], fraction );
      return 0;
    }
    Hex c;
    make_cell( lo, h, c );
    for( int k = 0; k < 3; ++k )
      out[ k ] = normal[ k ];
    out[ 3 ] = locate( c, normal, fraction );
    return 0;
  }

  int hg_volume_fraction ( int dim, const double *lo, const double *h, const double *hs, double *out )
  {
    if( dim == 2 )
    {
      Poly2 p;
      make_cell( lo[ 0 ], lo[ 1 ], h[ 0 ], h[ 1 ], p );
      const Hs2 H = { hs[ 0 ], hs[ 1 ], hs[ 2 ] };
      *out = volume_fraction( p, H );
      return 0;
    }
    Hex c;
    make_cell( lo, h, c );
    const Hs3 H = { { hs[ 0 ], hs[ 1 ], hs[ 2 ] }, hs[ 3 ] };
    *out = volume_fraction( c, H );
    return 0;
  }

  int hg_flux ( int dim, const double *lo, const double *h, int face, const double *v, const double *hs, double *out )
  {
    if( dim == 2 )
    {
      const Hs2 H = { hs[ 0 ], hs[ 1 ], hs[ 2 ] };
      *out = flux( lo[ 0 ], lo[ 1 ], h[ 0 ], h[ 1 ], face, v[ 0 ], v[ 1 ], H );
      return 0;
    }
    const Hs3 H = { { hs[ 0 ], hs[ 1 ], hs[ 2 ] }, hs[ 3 ] };
    *out = flux( lo, h, face, v, H );
    return 0;
  }

  int hg_interface ( int dim, const double *lo, const double *h, const double *hs, int *nverts, double *verts, double *centroid, double *measure )
  {
    if( dim == 2 )
    {
      Poly2 p;
      make_cell( lo[ 0 ], lo[ 1 ], h[ 0 ], h[ 1 ], p );
      const Hs2 H = { hs[ 0 ], hs[ 1 ], hs[ 2 ] };
      double lx[ 2 ], ly[ 2 ];
      plane_cut( p, H, lx, ly );
      *nverts = 2;
      for( int i = 0; i < 2; ++i )
      {
        verts[ 2 * i ] = lx[ i ];
        verts[ 2 * i + 1 ] = ly[ i ];
      }
      centroid[ 0 ] = ( lx[ 0 ] + lx[ 1 ] ) * 0.5;
      centroid[ 1 ] = ( ly[ 0 ] + ly[ 1 ] ) * 0.5;
      *measure = norm2( lx[ 0 ] - lx[ 1 ], ly[ 0 ] - ly[ 1 ] );
      return 0;
    }
    Hex c;
    make_cell( lo, h, c );
    const Hs3 H = { { hs[ 0 ], hs[ 1 ], hs[ 2 ] }, hs[ 3 ] };
    Face3 f;
    plane_cut( c, H, f );
    *nverts = f.n;
    for( int i = 0; i < f.n && i < 8; ++i )
      for( int d = 0; d < 3; ++d )
        verts[ 3 * i + d ] = f.v[ i ][ d ];
    if( f.n > 2 )
    {
      face_centroid( f, centroid );
      *measure = face_measure( f );
    }
    else
    {
      centroid[ 0 ] = centroid[ 1 ] = centroid[ 2 ] = 0.0;
      *measure = 0.0;
    }
    return 0;
  }

  double hg_det_cospi ( double s ) { return det_cospi( s ); }

  // the 8-lane cooperative flux (vofx_coop3d.cuh), lanes emulated phase by phase.  *path: 0 table, 1 empty, 2 serial
  int hg_flux_coop ( const double *lo, const double *h, int face, const double *v, const double *hs, double *out, int *path )
  {
    static coop::ClipCase table[ 256 + coop::kClipBoundaryCases ];
    static bool have = false;
    if( !have )
    {
      coop::make_clip_table( table );
      coop::make_clip_boundary_table( table + 256 );
      have = true;
    }
    const Hs3 H = { { hs[ 0 ], hs[ 1 ], hs[ 2 ] }, hs[ 3 ] };
    if( !hs_valid( H ) )
    {
      *out = fabs( 0.0 / 3.0 );
      *path = 1;
      return 0;
    }
    coop::ClipScratch S;
    for( int lane = 0; lane < 8; ++lane )
    {
      double xn[ 3 ], xo[ 3 ];
      coop::upwind_node( lane, lo, h, face, v, xn );
      coop::upwind_node_oriented( lane, lo, h, face, v, coop::upwind_shifted_first( lo, h, face, v ), xo );
      if( std::memcmp( xn, xo, sizeof( xn ) ) != 0 )
        return 3;
      coop::clip_phase_classify( lane, S, xo, H );
    }
    const coop::ClipCase *cs = nullptr;
    int status = 0;
    for( int lane = 0; lane < 8; ++lane )
      status = coop::clip_phase_cut( lane, S, H, table, cs, table + 256 );
    *path = status;
    if( status == 1 )
      *out = fabs( 0.0 / 3.0 );
    else if( status == 2 )
      *out = coop::clip_serial_from_scratch( S, H );
    else
    {
      for( int lane = 0; lane < 8; ++lane )
        coop::clip_phase_faces( lane, S, *cs );
      for( int lane = 0; lane < 8; ++lane )
        coop::clip_phase_edges< 8 >( lane, S, *cs );
      for( int lane = 0; lane < 8; ++lane )
        coop::clip_phase_face_sums( lane, S, *cs );
      *out = coop::clip_phase_sum( S, *cs );
    }
    return 0;
  }

  // entries of the boundary-pattern table: *irregular left to the serial path, *empty empty results
  int hg_clip_boundary_table ( int *irregular, int *empty )
  {
    static coop::ClipCase table[ coop::kClipBoundaryCases ];
    coop::make_clip_boundary_table( table );
    *irregular = *empty = 0;
    for( int m = 0; m < coop::kClipBoundaryCases; ++m )
    {
      *irregular += table[ m ].nCut == 0xFF;
      *empty += table[ m ].nCut == 0xFE;
    }
    return coop::kClipBoundaryCases;
  }

  int hg_clip_table_irregular ( void )
  {
    coop::ClipCase table[ 256 ];
    coop::make_clip_table( table );
    int n = 0;
    for( int m = 0; m < 256; ++m )
      n += table[ m ].nCut == 0xFF;
    return n;
  }

} // extern "C"

extern "C" int hg_sweep_table_cases ( void )
{
  static coop::SweepTable T;
  return coop::make_sweep_table( T );
}

// Young's normal + plane constant of one cell by the cooperative path (lanes emulated).  status: 0 located by the table
// path, 1 degenerate normal (zero half-space), 2 left to the serial walk (then located serially here)
template< int L >
static int youngs_coop_cell ( const int *n, const double *u, const int *ijk, double *out )
{
  static coop::SweepTable T;
  static bool have = false;
  if( !have )
  {
    if( coop::make_sweep_table( T ) != coop::kSweepCases )
      return -1;
    have = true;
  }
  Grid g;
  g.dim = 3;
  g.cells = 1;
  for( int d = 0; d < 3; ++d )
  {
    g.ns[ d ] = g.ng[ d ] = n[ d ];
    g.g0[ d ] = 0;
    g.origin[ d ] = 0.0;
    g.h[ d ] = 1.0 / n[ d ];
    g.cells *= n[ d ];
  }
  g.own0 = g.list0 = 0;
  g.own1 = g.list1 = n[ 2 ];
  coop::LocateScratch S;
  coop::Group G = { 0, 0u };
  const double colorEn = u[ cell_index( g, ijk ) ];
  double normal[ 3 ];
  if( !coop::youngs_normal_coop< L >( G, S, g, u, ijk, colorEn, normal ) )
  {
    out[ 0 ] = out[ 1 ] = out[ 2 ] = out[ 3 ] = 0.0;
    return 1;
  }
  const double lo[ 3 ] = { cell_lower( g, 0, ijk[ 0 ] ), cell_lower( g, 1, ijk[ 1 ] ), cell_lower( g, 2, ijk[ 2 ] ) };
  for( int k = 0; k < 3; ++k )
    out[ k ] = normal[ k ];
  double dist;
  if( coop::locate_coop< L >( G, S, T, lo, g.h, normal, colorEn, dist ) )
  {
    out[ 3 ] = dist;
    return 0;
  }
  Hex c;
  make_cell( lo, g.h, c );
  out[ 3 ] = locate( c, normal, colorEn );
  return 2;
}

extern "C" int hg_youngs_coop ( int lanes, const int *n, const double *u, const int *ijk, double *out )
{
  return lanes == 4 ? youngs_coop_cell< 4 >( n, u, ijk, out ) : youngs_coop_cell< 8 >( n, u, ijk, out );
}

// plane constant only, for fuzzing against the oracle: status 0 table path, 2 serial walk needed
extern "C" int hg_locate_coop ( int lanes, const double *lo, const double *h, const double *normal, double fraction, double *dist )
{
  static coop::SweepTable T;
  static bool have = false;
  if( !have )
  {
    coop::make_sweep_table( T );
    have = true;
  }
  coop::LocateScratch S;
  coop::Group G = { 0, 0u };
  const bool ok = lanes == 4 ? coop::locate_coop< 4 >( G, S, T, lo, h, normal, fraction, *dist )
                             : coop::locate_coop< 8 >( G, S, T, lo, h, normal, fraction, *dist );
  return ok ? 0 : 2;
}

// ModifiedSwartz on one cell by the cooperative path (lanes emulated); rec: in Young's, out Swartz
extern "C" int hg_swartz_coop ( int lanes, const int *n, const double *u, const unsigned char *flags, const int *ijk, int maxIterations, double *rec )
{
  static coop::SweepTable T;
  static bool have = false;
  if( !have )
  {
    coop::make_sweep_table( T );
    have = true;
  }
  Grid g;
  g.dim = 3;
  g.cells = 1;
  for( int d = 0; d < 3; ++d )
  {
    g.ns[ d ] = g.ng[ d ] = n[ d ];
    g.g0[ d ] = 0;
    g.origin[ d ] = 0.0;
    g.h[ d ] = 1.0 / n[ d ];
    g.cells *= n[ d ];
  }
  g.own0 = g.list0 = 0;
  g.own1 = g.list1 = n[ 2 ];
  static coop::SwartzScratch< 4 > S4;
  static coop::SwartzScratch< 8 > S8;
  coop::Group G = { 0, 0u };
  if( lanes == 4 )
    coop::swartz_cell_coop< 4 >( G, S4, T, g, u, flags, ijk, maxIterations, rec );
  else
    coop::swartz_cell_coop< 8 >( G, S8, T, g, u, flags, ijk, maxIterations, rec );
  return 0;
}
