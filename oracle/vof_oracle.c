/* TEST INFRASTRUCTURE ONLY (oracle/) -- see vof_oracle.h.
 *
 * Plain-C restatement of the reference's hot path.  Arithmetic follows the reference operation for operation
 * (same association, same summation order, same tolerances), because in 3D the reference's own results are
 * dominated by amplified round-off and parity therefore means "same bits".
 * Compile WITHOUT -march/-ffast-math and with -ffp-contract=off (oracle/Makefile).
 *
 * All "file:line" citations are relative to /root/reference/dune/vof/.
 */
#define _DEFAULT_SOURCE 1
#define _POSIX_C_SOURCE 200809L
#include "vof_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define VO_EPS DBL_EPSILON

/* ------------------------------------------------------------------------------------------------------------
 * small vector helpers; accumulation order = dune-common DenseVector (left to right, starting from 0)
 * ---------------------------------------------------------------------------------------------------------- */

static double dotn ( const double *a, const double *b, int n )
{
  double s = 0;
  for( int i = 0; i < n; ++i )
    s += a[ i ] * b[ i ];
  return s;
}

static double norm2n ( const double *a, int n )
{
  double s = 0;
  for( int i = 0; i < n; ++i )
    s += a[ i ] * a[ i ];
  return s;
}

static double normn ( const double *a, int n ) { return sqrt( norm2n( a, n ) ); }

static void cross3 ( const double *v, const double *w, double *r )   /* geometry/utility.hh:119-138 */
{
  r[ 0 ] = v[ 1 ] * w[ 2 ] - v[ 2 ] * w[ 1 ];
  r[ 1 ] = v[ 2 ] * w[ 0 ] - v[ 0 ] * w[ 2 ];
  r[ 2 ] = v[ 0 ] * w[ 1 ] - v[ 1 ] * w[ 0 ];
}

static double clampd ( double v, double lo, double hi )   /* utility.hh:51-55: max( min( v, hi ), lo ) */
{
  const double m = ( hi < v ) ? hi : v;     /* std::min( v, hi ) */
  return ( m < lo ) ? lo : m;               /* std::max( m, lo ) */
}

/* half-space validity, geometry/halfspace.hh:104-107 */
static int hs_valid ( const double *hs, int dim ) { return norm2n( hs, dim ) > 0.5; }

/* geometry/halfspace.hh:84 */
static double level_set ( const double *hs, const double *x, int dim ) { return dotn( hs, x, dim ) + hs[ dim ]; }

/* ------------------------------------------------------------------------------------------------------------
 * G3  Brent's method, brents.hh:34-117 (pair form) and :131-146 (scalar form)
 * ---------------------------------------------------------------------------------------------------------- */

typedef double ( *vo_fn ) ( double, void * );

static double brent_pair ( vo_fn f, void *ctx, double a1, double a2, double b1, double b2, double tolerance )
{
  const double half = 1.0 / 2.0;
  double c1 = a1, c2 = a2;
  double d = b1 - a1;
  double e = d;

  for( ;; )
  {
    if( b2 * c2 > 0.0 )
    {
      c1 = a1; c2 = a2;
      d = ( b1 - a1 );
      e = d;
    }
    if( fabs( c2 ) < fabs( b2 ) )
    {
      a1 = b1; a2 = b2;
      b1 = c1; b2 = c2;
      c1 = a1; c2 = a2;
    }

    const double htol = VO_EPS * fabs( b1 ) + tolerance;
    const double tol = 2.0 * htol;
    const double tm = c1 - b1;
    const double m = half * tm;
    if( ( fabs( tm ) <= 4.0 * htol ) || ( b2 == 0.0 ) )
      return b1;

    if( ( fabs( e ) >= tol ) && ( fabs( a2 ) > fabs( b2 ) ) )
    {
      double p, q, r;
      double s = b2 / a2;
      if( a1 == c1 )
      {
        p = tm * s;
        q = 1.0 - s;
      }
      else
      {
        q = a2 / c2;
        r = b2 / c2;
        p = s * ( tm * q * ( q - r ) - ( b1 - a1 ) * ( r - 1.0 ) );
        q = ( q - 1 ) * ( r - 1 ) * ( s - 1 );
      }
      if( p > 0 )
        q = -q;
      else
        p = -p;

      s = e;
      e = d;
      if( ( 2.0 * p < 3.0 * m * q - fabs( tol * q ) ) && ( p < fabs( s * q * half ) ) )
        d = p / q;
      else
        d = e = m;
    }
    else
      d = e = m;

    a1 = b1; a2 = b2;
    if( fabs( d ) > tol )
      b1 += d;
    else
      b1 += ( m > 0 ? tol : -tol );
    b2 = f( b1, ctx );
  }
}

static double brent ( vo_fn f, void *ctx, double a, double b, double tolerance )
{
  const double fa = f( a, ctx );
  if( fabs( fa ) < tolerance )
    return a;
  const double fb = f( b, ctx );
  if( fabs( fb ) < tolerance )
    return b;
  return brent_pair( f, ctx, a, fa, b, fb, tolerance );
}

typedef struct { double A, B, C, target; } cubic_ctx;

/* V(h) of geometry/algorithm.hh:211 minus the target of :223 */
static double cubic_eval ( double h, void *p )
{
  const cubic_ctx *c = (const cubic_ctx *) p;
  return ( c->C * h + c->B * h * h / 2.0 + c->A * h * h * h / 6.0 ) - c->target;
}

double vo_brent_cubic ( double A, double B, double C, double target, double a, double b, double tol )
{
  cubic_ctx c = { A, B, C, target };
  return brent( cubic_eval, &c, a, b, tol );
}

/* ------------------------------------------------------------------------------------------------------------
 * G4  2D polygons: geometry/2d/polygon.hh, geometry/2d/intersect.hh
 * ---------------------------------------------------------------------------------------------------------- */

#define VO_MAXPV 16

typedef struct
{
  int n;
  double v[ VO_MAXPV ][ 2 ];
} vo_polygon;

/* 2d/polygon.hh:99-106 (signed) */
static double polygon_volume ( const vo_polygon *p )
{
  double sum = 0;
  const int n = p->n;
  for( int i = 0; i < n; i++ )
    sum += ( p->v[ i ][ 1 ] + p->v[ ( i + 1 ) % n ][ 1 ] ) * ( p->v[ i ][ 0 ] - p->v[ ( i + 1 ) % n ][ 0 ] );
  return sum / 2.0;
}

/* cut point of 2d/intersect.hh:86-89 and 3d/intersect.hh:49-52: point = 0; point.axpy(-l1/(l0-l1), v0); point.axpy(l0/(l0-l1), v1) */
static void cut_point ( const double *v0, const double *v1, double l0, double l1, double *out, int dim )
{
  const double a = -l1 / ( l0 - l1 );
  const double b = l0 / ( l0 - l1 );
  for( int k = 0; k < dim; ++k )
  {
    double x = 0.0;
    x += a * v0[ k ];
    x += b * v1[ k ];
    out[ k ] = x;
  }
}

/* 2d/intersect.hh:67-99 */
static void polygon_clip ( const vo_polygon *p, const double *hs, vo_polygon *out )
{
  out->n = 0;
  if( !hs_valid( hs, 2 ) )
    return;
  const int n = p->n;
  for( int i = 0; i < n; ++i )
  {
    const double *v0 = p->v[ i ];
    const double *v1 = p->v[ ( i + 1 ) % n ];
    const double l0 = level_set( hs, v0, 2 );
    const double l1 = level_set( hs, v1, 2 );
    if( l0 > 0.0 )
    {
      out->v[ out->n ][ 0 ] = v0[ 0 ];
      out->v[ out->n ][ 1 ] = v0[ 1 ];
      out->n++;
    }
    if( ( l0 > 0.0 ) ^ ( l1 > 0.0 ) )
    {
      cut_point( v0, v1, l0, l1, out->v[ out->n ], 2 );
      out->n++;
    }
  }
}

/* geometry/algorithm.hh:31-37 */
static double volume_fraction_2d ( const vo_polygon *p, const double *hs )
{
  vo_polygon c;
  polygon_clip( p, hs, &c );
  return polygon_volume( &c ) / polygon_volume( p );
}

/* makePolygon for a quadrilateral, 2d/polygon.hh:230-241: corners 0,1,3,2 */
static void make_polygon_cell ( const double *lo, const double *h, vo_polygon *p )
{
  const double x0 = lo[ 0 ], x1 = lo[ 0 ] + h[ 0 ];
  const double y0 = lo[ 1 ], y1 = lo[ 1 ] + h[ 1 ];
  p->n = 4;
  p->v[ 0 ][ 0 ] = x0; p->v[ 0 ][ 1 ] = y0;
  p->v[ 1 ][ 0 ] = x1; p->v[ 1 ][ 1 ] = y0;
  p->v[ 2 ][ 0 ] = x1; p->v[ 2 ][ 1 ] = y1;
  p->v[ 3 ][ 0 ] = x0; p->v[ 3 ][ 1 ] = y1;
}

typedef struct { const vo_polygon *p; double n[ 2 ]; double fraction; } locate2d_ctx;

static double locate2d_eval ( double dist, void *q )
{
  const locate2d_ctx *c = (const locate2d_ctx *) q;
  const double hs[ 3 ] = { c->n[ 0 ], c->n[ 1 ], dist };
  return volume_fraction_2d( c->p, hs ) - c->fraction;
}

/* G1  locateHalfSpace( Polygon ), geometry/algorithm.hh:68-111 */
static void locate_2d ( const vo_polygon *polygon, const double *normal, double fraction, double *out )
{
  double pMin = 0, pMax = 0;
  double volume, volMin = -DBL_MAX, volMax = DBL_MAX;

  for( int i = 0; i < polygon->n; ++i )
  {
    const double dist = -( dotn( normal, polygon->v[ i ], 2 ) );
    const double hs[ 3 ] = { normal[ 0 ], normal[ 1 ], dist };
    volume = volume_fraction_2d( polygon, hs );
    if( ( volume <= volMax ) && ( volume >= fraction ) )
    {
      pMax = dist;
      volMax = volume;
    }
    if( ( volume >= volMin ) && ( volume <= fraction ) )
    {
      pMin = dist;
      volMin = volume;
    }
  }

  out[ 0 ] = normal[ 0 ];
  out[ 1 ] = normal[ 1 ];
  if( volMax == DBL_MAX )
    out[ 2 ] = pMin;
  else if( volMin == -DBL_MAX )
    out[ 2 ] = pMax;
  else if( volMin == volMax )
    out[ 2 ] = pMin;
  else
  {
    locate2d_ctx c = { polygon, { normal[ 0 ], normal[ 1 ] }, fraction };
    out[ 2 ] = brent( locate2d_eval, &c, pMin, pMax, 1e-12 );
  }
}

/* G6 (2D)  intersect( Polygon, HyperPlane ) -> Line, 2d/intersect.hh:110-148; returns 0 if the line is empty */
static int polygon_plane_cut ( const vo_polygon *p, const double *hs, double line[ 2 ][ 2 ] )
{
  line[ 0 ][ 0 ] = line[ 0 ][ 1 ] = line[ 1 ][ 0 ] = line[ 1 ][ 1 ] = 0.0;
  if( !hs_valid( hs, 2 ) )
    return 0;
  int j = 0;
  const int n = p->n;
  for( int i = 0; i < n; ++i )
  {
    if( j == 2 )
      break;
    const double *v0 = p->v[ i ];
    const double *v1 = p->v[ ( i + 1 ) % n ];
    const double l0 = level_set( hs, v0, 2 );
    const double l1 = level_set( hs, v1, 2 );
    if( ( l0 > 0.0 && l1 < 0.0 ) || ( l0 < 0.0 && l1 > 0.0 ) )
    {
      cut_point( v0, v1, l0, l1, line[ j ], 2 );
      j++;
    }
    else if( fabs( l0 ) < VO_EPS )
    {
      line[ j ][ 0 ] = v0[ 0 ];
      line[ j ][ 1 ] = v0[ 1 ];
      j++;
    }
  }
  if( j == 0 )
    return 0;
  else if( j == 1 )
  {
    line[ 1 ][ 0 ] = line[ 0 ][ 0 ];
    line[ 1 ][ 1 ] = line[ 0 ][ 1 ];
  }
  return 1;
}

/* ------------------------------------------------------------------------------------------------------------
 * 3D polyhedra: geometry/3d/polyhedron.hh
 * ---------------------------------------------------------------------------------------------------------- */

#define VO_MAXN 24
#define VO_MAXE 96
#define VO_MAXF 12
#define VO_MAXFE 16

typedef struct
{
  int nn, ne, nf;
  double x[ VO_MAXN ][ 3 ];
  int e[ VO_MAXE ][ 2 ];            /* directed edges (node ids) */
  int fne[ VO_MAXF ];               /* edges per face */
  int fe[ VO_MAXF ][ VO_MAXFE ];    /* edge ids per face, in order; face node i = e[ fe[f][i] ][0] */
} vo_polyhedron;

/* cube topology of makePolyhedron, 3d/polyhedron.hh:387-401 (also used by upwindPolygon3d, 3d/upwindpolygon.hh:59-70) */
static const int cube_edges[ 24 ][ 2 ] = {
  { 0, 4 }, { 1, 5 }, { 2, 6 }, { 3, 7 }, { 0, 2 }, { 1, 3 }, { 0, 1 }, { 2, 3 }, { 4, 6 }, { 5, 7 },
  { 4, 5 }, { 6, 7 }, { 4, 0 }, { 5, 1 }, { 6, 2 }, { 7, 3 }, { 2, 0 }, { 3, 1 }, { 1, 0 }, { 3, 2 },
  { 6, 4 }, { 7, 5 }, { 5, 4 }, { 7, 6 }
};
static const int cube_faces[ 6 ][ 4 ] = {
  { 0, 8, 14, 16 }, { 5, 3, 21, 13 }, { 6, 1, 22, 12 }, { 19, 2, 11, 15 }, { 4, 7, 17, 18 }, { 10, 9, 23, 20 }
};

static void make_cube_topology ( vo_polyhedron *p )
{
  p->nn = 8;
  p->ne = 24;
  p->nf = 6;
  for( int i = 0; i < 24; ++i )
  {
    p->e[ i ][ 0 ] = cube_edges[ i ][ 0 ];
    p->e[ i ][ 1 ] = cube_edges[ i ][ 1 ];
  }
  for( int f = 0; f < 6; ++f )
  {
    p->fne[ f ] = 4;
    for( int i = 0; i < 4; ++i )
      p->fe[ f ][ i ] = cube_faces[ f ][ i ];
  }
}

/* makePolyhedron( cube geometry ): nodes = geometry.corner(0..7) in DUNE order (bit d of i selects the upper side) */
static void make_polyhedron_cell ( const double *lo, const double *h, vo_polyhedron *p )
{
  make_cube_topology( p );
  for( int i = 0; i < 8; ++i )
    for( int d = 0; d < 3; ++d )
      p->x[ i ][ d ] = ( ( i >> d ) & 1 ) ? lo[ d ] + h[ d ] : lo[ d ];
}

/* __impl::Face::outerNormal, 3d/polyhedron.hh:187-204 */
static void face_outer_normal ( const vo_polyhedron *p, int f, double *normal )
{
  const double *n0 = p->x[ p->e[ p->fe[ f ][ 0 ] ][ 0 ] ];
  const double *n1 = p->x[ p->e[ p->fe[ f ][ 1 ] ][ 0 ] ];
  const double *n2 = p->x[ p->e[ p->fe[ f ][ 2 ] ][ 0 ] ];
  double v1[ 3 ], v2[ 3 ];
  for( int k = 0; k < 3; ++k )
  {
    v1[ k ] = n0[ k ] - n1[ k ];
    v2[ k ] = n2[ k ] - n1[ k ];
  }
  cross3( v1, v2, normal );
  const double s = -1.0 * normn( normal, 3 );
  for( int k = 0; k < 3; ++k )
    normal[ k ] /= s;
}

/* __impl::Face::center, 3d/polyhedron.hh:206-214 */
static void face_center ( const vo_polyhedron *p, int f, double *c )
{
  c[ 0 ] = c[ 1 ] = c[ 2 ] = 0.0;
  const int n = p->fne[ f ];
  for( int i = 0; i < n; ++i )
  {
    const double *x = p->x[ p->e[ p->fe[ f ][ i ] ][ 0 ] ];
    for( int k = 0; k < 3; ++k )
      c[ k ] += x[ k ];
  }
  const double s = 1.0 / n;
  for( int k = 0; k < 3; ++k )
    c[ k ] *= s;
}

/* __impl::Face::volume, 3d/polyhedron.hh:216-222, with Edge::center :80-86, Edge::outerNormal( planeNormal ) :62-78, Edge::volume :88-91 */
static double face_volume ( const vo_polyhedron *p, int f )
{
  double vol = 0;
  const int n = p->fne[ f ];
  for( int i = 0; i < n; ++i )
  {
    const int *e = p->e[ p->fe[ f ][ i ] ];
    const double *x0 = p->x[ e[ 0 ] ];
    const double *x1 = p->x[ e[ 1 ] ];
    double fn[ 3 ];
    face_outer_normal( p, f, fn );
    double center[ 3 ], v1[ 3 ], en[ 3 ];
    for( int k = 0; k < 3; ++k )
    {
      center[ k ] = ( x0[ k ] + x1[ k ] ) * 0.5;
      v1[ k ] = x1[ k ] - x0[ k ];
    }
    cross3( v1, fn, en );
    const double nrm = normn( en, 3 );
    for( int k = 0; k < 3; ++k )
      en[ k ] /= nrm;
    vol += dotn( center, en, 3 ) * normn( v1, 3 );
  }
  return fabs( vol / 2.0 );
}

/* Polyhedron::volume, 3d/polyhedron.hh:317-324 */
static double polyhedron_volume ( const vo_polyhedron *p )
{
  double vol = 0.0;
  for( int f = 0; f < p->nf; ++f )
  {
    double c[ 3 ], n[ 3 ];
    face_center( p, f, c );
    face_outer_normal( p, f, n );
    vol += dotn( c, n, 3 ) * face_volume( p, f );
  }
  return fabs( vol / 3.0 );
}

/* __impl::Edge::intersection( hyperplane ), 3d/polyhedron.hh:93-101 */
static void edge_plane_point ( const double *x0, const double *x1, const double *hs, double *out )
{
  double p[ 3 ];
  for( int k = 0; k < 3; ++k )
    p[ k ] = x0[ k ] - x1[ k ];
  const double s = ( dotn( hs, x0, 3 ) + hs[ 3 ] ) / ( dotn( hs, p, 3 ) * -1.0 );
  for( int k = 0; k < 3; ++k )
  {
    p[ k ] *= s;
    p[ k ] += x0[ k ];
    out[ k ] = p[ k ];
  }
}

/* G5  __impl::intersect( Polyhedron, HalfSpace ), 3d/intersect.hh:72-263.  Returns 0 and an empty result if the
 * intersection is empty or degenerate (no faces -> volume 0). */
static void polyhedron_clip ( const vo_polyhedron *poly, const double *hs, vo_polyhedron *out )
{
  out->nn = out->ne = out->nf = 0;
  if( !hs_valid( hs, 3 ) )
    return;

  const double eps = VO_EPS;
  int isInner[ VO_MAXN ], isBoundary[ VO_MAXN ], p[ VO_MAXN ];
  double newNodes[ VO_MAXN ][ 3 ];
  int nNew = 0;
  int n = 0, onBoundary = 0;

  /* map (edge.node0, edge.node1) -> new node id, 3d/intersect.hh:93-95,171-183 (std::map replaced by a small table) */
  int mapKey[ VO_MAXE ][ 2 ], mapVal[ VO_MAXE ], nMap = 0;

  int intersectionFace[ VO_MAXE ], nIF = 0;

  for( int i = 0; i < poly->nn; ++i )
  {
    isInner[ i ] = isBoundary[ i ] = 0;
    const double ls = level_set( hs, poly->x[ i ], 3 );
    if( ls > -eps )
    {
      isInner[ i ] = 1;
      p[ i ] = n;
      n++;
      if( fabs( ls ) < eps )
      {
        isBoundary[ i ] = 1;
        onBoundary++;
      }
    }
    else
      p[ i ] = -1;
  }

  /* trivial cases, :115-120 (a point or an edge: no faces, volume 0) */
  if( n == 0 )
    return;
  if( n == 1 && onBoundary == 1 )
    return;
  if( n == 2 && onBoundary == 2 )
    return;

  for( int f = 0; f < poly->nf; ++f )
  {
    int newFace[ VO_MAXFE ], nNF = 0;
    int lastIsId = -1;

    for( int k = 0; k < poly->fne[ f ]; ++k )
    {
      const int *edge = poly->e[ poly->fe[ f ][ k ] ];
      const int a = edge[ 0 ], b = edge[ 1 ];

      if( isInner[ a ] && isInner[ b ] )
      {
        out->e[ out->ne ][ 0 ] = p[ a ]; out->e[ out->ne ][ 1 ] = p[ b ]; out->ne++;
        newFace[ nNF++ ] = out->ne - 1;
      }
      else if( isInner[ a ] ^ isInner[ b ] )
      {
        if( isBoundary[ a ] )
        {
          if( lastIsId != -1 && lastIsId != p[ a ] )
          {
            out->e[ out->ne ][ 0 ] = p[ a ]; out->e[ out->ne ][ 1 ] = lastIsId; out->ne++;
            newFace[ nNF++ ] = out->ne - 1;
            out->e[ out->ne ][ 0 ] = lastIsId; out->e[ out->ne ][ 1 ] = p[ a ]; out->ne++;
            intersectionFace[ nIF++ ] = out->ne - 1;
          }
          lastIsId = p[ a ];
        }
        else if( isBoundary[ b ] )
        {
          if( lastIsId != -1 && lastIsId != p[ b ] )
          {
            out->e[ out->ne ][ 0 ] = lastIsId; out->e[ out->ne ][ 1 ] = p[ b ]; out->ne++;
            newFace[ nNF++ ] = out->ne - 1;
            out->e[ out->ne ][ 0 ] = p[ b ]; out->e[ out->ne ][ 1 ] = lastIsId; out->ne++;
            intersectionFace[ nIF++ ] = out->ne - 1;
          }
          lastIsId = p[ b ];
        }
        else
        {
          int isNodeId = -1;
          /* look up the reverse edge (b,a) */
          for( int m = 0; m < nMap; ++m )
            if( mapKey[ m ][ 0 ] == b && mapKey[ m ][ 1 ] == a )
              isNodeId = mapVal[ m ];
          if( isNodeId == -1 )
          {
            edge_plane_point( poly->x[ a ], poly->x[ b ], hs, newNodes[ nNew ] );
            nNew++;
            isNodeId = n - 1 + nNew;
            mapKey[ nMap ][ 0 ] = a; mapKey[ nMap ][ 1 ] = b; mapVal[ nMap ] = isNodeId; nMap++;
          }

          if( isInner[ a ] )
          {
            out->e[ out->ne ][ 0 ] = p[ a ]; out->e[ out->ne ][ 1 ] = isNodeId; out->ne++;
            newFace[ nNF++ ] = out->ne - 1;
            if( lastIsId != -1 && lastIsId != isNodeId )
            {
              out->e[ out->ne ][ 0 ] = isNodeId; out->e[ out->ne ][ 1 ] = lastIsId; out->ne++;
              newFace[ nNF++ ] = out->ne - 1;
              out->e[ out->ne ][ 0 ] = lastIsId; out->e[ out->ne ][ 1 ] = isNodeId; out->ne++;
              intersectionFace[ nIF++ ] = out->ne - 1;
            }
          }
          else
          {
            if( lastIsId != -1 && lastIsId != isNodeId )
            {
              out->e[ out->ne ][ 0 ] = lastIsId; out->e[ out->ne ][ 1 ] = isNodeId; out->ne++;
              newFace[ nNF++ ] = out->ne - 1;
              out->e[ out->ne ][ 0 ] = isNodeId; out->e[ out->ne ][ 1 ] = lastIsId; out->ne++;
              intersectionFace[ nIF++ ] = out->ne - 1;
            }
            out->e[ out->ne ][ 0 ] = isNodeId; out->e[ out->ne ][ 1 ] = p[ b ]; out->ne++;
            newFace[ nNF++ ] = out->ne - 1;
          }
          lastIsId = isNodeId;
        }
      }
    }
    if( nNF > 2 )
    {
      out->fne[ out->nf ] = nNF;
      for( int i = 0; i < nNF; ++i )
        out->fe[ out->nf ][ i ] = newFace[ i ];
      out->nf++;
    }
  }

  /* erase null edges, :217-222 (note: the reference does not re-test index i after an erase) */
  for( int i = 0; i < nIF; ++i )
  {
    const int *edge = out->e[ intersectionFace[ i ] ];
    if( edge[ 0 ] == edge[ 1 ] )
    {
      for( int j = i; j + 1 < nIF; ++j )
        intersectionFace[ j ] = intersectionFace[ j + 1 ];
      nIF--;
    }
  }

  /* sort edges of the intersection face, :224-239 */
  for( int i = 0; i + 1 < nIF; ++i )
  {
    if( out->e[ intersectionFace[ i ] ][ 1 ] == out->e[ intersectionFace[ i + 1 ] ][ 0 ] )
      continue;
    const int index = out->e[ intersectionFace[ i ] ][ 1 ];
    int pos = -1;
    for( int j = i; j < nIF; ++j )
      if( out->e[ intersectionFace[ j ] ][ 0 ] == index )
      {
        pos = j;
        break;
      }
    if( pos != -1 )
    {
      const int t = intersectionFace[ i + 1 ];
      intersectionFace[ i + 1 ] = intersectionFace[ pos ];
      intersectionFace[ pos ] = t;
    }
  }

  if( nIF > 2 )
  {
    out->fne[ out->nf ] = nIF;
    for( int i = 0; i < nIF; ++i )
      out->fe[ out->nf ][ i ] = intersectionFace[ i ];
    out->nf++;
  }

  /* nodes: inner nodes in order, then the new nodes, :252-257 */
  for( int i = 0; i < poly->nn; ++i )
    if( isInner[ i ] )
    {
      memcpy( out->x[ out->nn ], poly->x[ i ], 3 * sizeof( double ) );
      out->nn++;
    }
  for( int i = 0; i < nNew; ++i )
  {
    memcpy( out->x[ out->nn ], newNodes[ i ], 3 * sizeof( double ) );
    out->nn++;
  }
}

/* geometry/algorithm.hh:39-45 */
static double volume_fraction_3d ( const vo_polyhedron *p, const double *hs )
{
  vo_polyhedron c;
  polyhedron_clip( p, hs, &c );
  return polyhedron_volume( &c ) / polyhedron_volume( p );
}

/* ------------------------------------------------------------------------------------------------------------
 * G2  locateHalfSpace( Polyhedron ): geometry/algorithm.hh:119-236, 3d/rotation.hh:17-53,
 *     3d/polygonwithdirections.hh:27-316
 * ---------------------------------------------------------------------------------------------------------- */

/* rotateToReferenceFrame, 3d/rotation.hh:17-53 (topology is kept, nodes are rotated) */
static void rotate_to_reference_frame ( const double *n, const vo_polyhedron *cell, vo_polyhedron *out )
{
  *out = *cell;
  if( n[ 0 ] == 0.0 && n[ 1 ] == 0.0 && n[ 2 ] == 1.0 )
    return;

  double R[ 3 ][ 3 ] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
  if( n[ 0 ] == 0.0 && n[ 1 ] == 0.0 && n[ 2 ] == -1.0 )
  {
    R[ 0 ][ 0 ] = 1.0;
    R[ 1 ][ 1 ] = -1.0;
    R[ 2 ][ 2 ] = -1.0;
  }
  else
  {
    const double gamma = ( 1 - n[ 2 ] ) / ( n[ 0 ] * n[ 0 ] + n[ 1 ] * n[ 1 ] );
    R[ 0 ][ 0 ] = n[ 1 ] * n[ 1 ] * gamma + n[ 2 ];
    R[ 0 ][ 1 ] = -n[ 0 ] * n[ 1 ] * gamma;
    R[ 1 ][ 0 ] = R[ 0 ][ 1 ];
    R[ 0 ][ 2 ] = -n[ 0 ];
    R[ 1 ][ 1 ] = n[ 0 ] * n[ 0 ] * gamma + n[ 2 ];
    R[ 1 ][ 2 ] = -n[ 1 ];
    R[ 2 ][ 0 ] = n[ 0 ];
    R[ 2 ][ 1 ] = n[ 1 ];
    R[ 2 ][ 2 ] = n[ 2 ];
  }
  for( int i = 0; i < cell->nn; ++i )
    for( int r = 0; r < 3; ++r )
    {
      double y = 0;
      for( int c = 0; c < 3; ++c )
        y += R[ r ][ c ] * cell->x[ i ][ c ];
      out->x[ i ][ r ] = y;
    }
}

/* Polyhedron::attachedFaceToEdge, 3d/polyhedron.hh:335-344: first face containing the directed edge (a,b); -1 if none
 * (the reference throws InvalidStateException) */
static int attached_face_to_edge ( const vo_polyhedron *p, int a, int b )
{
  for( int f = 0; f < p->nf; ++f )
    for( int k = 0; k < p->fne[ f ]; ++k )
    {
      const int *e = p->e[ p->fe[ f ][ k ] ];
      if( e[ 0 ] == a && e[ 1 ] == b )
        return f;
    }
  return -1;
}

#define VO_MAXDIR 8
typedef struct { int n; int e[ VO_MAXDIR ]; } vo_dirs;   /* edge ids into the parent's edge list */

/* Polyhedron::attachedEdgesToNode, 3d/polyhedron.hh:346-356 */
static void attached_edges_to_node ( const vo_polyhedron *p, int node, vo_dirs *d )
{
  d->n = 0;
  for( int i = 0; i < p->ne; ++i )
    if( p->e[ i ][ 0 ] == node )
      d->e[ d->n++ ] = i;
}

static void dirs_erase ( vo_dirs *d, int i )
{
  for( int j = i; j + 1 < d->n; ++j )
    d->e[ j ] = d->e[ j + 1 ];
  d->n--;
}

/* comparator PolygonWithDirections::operator(), 3d/polygonwithdirections.hh:293-306 */
static int dir_less ( const vo_polyhedron *p, int lhs, int rhs )
{
  const int *L = p->e[ lhs ], *Rr = p->e[ rhs ];
  int f = attached_face_to_edge( p, L[ 0 ], L[ 1 ] );
  if( f >= 0 )
    for( int k = 0; k < p->fne[ f ]; ++k )
    {
      const int *e = p->e[ p->fe[ f ][ k ] ];
      if( e[ 0 ] == Rr[ 1 ] && e[ 1 ] == Rr[ 0 ] )
        return 1;
    }
  f = attached_face_to_edge( p, Rr[ 0 ], Rr[ 1 ] );
  if( f >= 0 )
    for( int k = 0; k < p->fne[ f ]; ++k )
    {
      const int *e = p->e[ p->fe[ f ][ k ] ];
      if( e[ 0 ] == L[ 1 ] && e[ 1 ] == L[ 0 ] )
        return 0;
    }
  return 0;
}

/* std::sort for fewer than 16 elements = libstdc++ __insertion_sort (bits/stl_algo.h); the comparator above is not
 * a strict weak ordering (it is cyclic around a node), so the exact algorithm matters. */
static void dirs_sort ( const vo_polyhedron *p, vo_dirs *d )
{
  for( int i = 1; i < d->n; ++i )
  {
    const int val = d->e[ i ];
    if( dir_less( p, val, d->e[ 0 ] ) )
    {
      for( int j = i; j > 0; --j )
        d->e[ j ] = d->e[ j - 1 ];
      d->e[ 0 ] = val;
    }
    else
    {
      int j = i;
      while( dir_less( p, val, d->e[ j - 1 ] ) )
      {
        d->e[ j ] = d->e[ j - 1 ];
        --j;
      }
      d->e[ j ] = val;
    }
  }
}

#define VO_MAXPN 16
typedef struct
{
  const vo_polyhedron *parent;     /* the rotated cell */
  int nn;                          /* polygon nodes */
  double node[ VO_MAXPN ][ 2 ];
  int nodeId[ VO_MAXPN ];          /* parent node id or -1 */
  vo_dirs dirs[ VO_MAXPN ];
  int nEdges;                      /* polygon edges (i,(i+1)%nn); 0 if nn == 1 */
  int nCF;
  int corrFace[ VO_MAXPN ];
  int degenerate;                  /* set where the reference would run into undefined behaviour */
} vo_sweep;

/* PolygonWithDirections::directionVector, 3d/polygonwithdirections.hh:64-74 */
static void direction_vector ( const vo_sweep *s, int n, int k, double *dv )
{
  const int *e = s->parent->e[ s->dirs[ n ].e[ k ] ];
  for( int c = 0; c < 3; ++c )
    dv[ c ] = s->parent->x[ e[ 1 ] ][ c ] - s->parent->x[ e[ 0 ] ][ c ];
  const double norm = normn( dv, 2 );
  if( norm > 0 )
    for( int c = 0; c < 3; ++c )
      dv[ c ] /= norm;
}

/* PolygonWithDirections::volume, 3d/polygonwithdirections.hh:56-62 with the 2D Edge::outerNormal of 3d/polyhedron.hh:49-60 */
static double sweep_polygon_volume ( const vo_sweep *s )
{
  double vol = 0;
  for( int i = 0; i < s->nEdges; ++i )
  {
    const double *x0 = s->node[ i ];
    const double *x1 = s->node[ ( i + 1 ) % s->nn ];
    double diff[ 2 ] = { x1[ 0 ] - x0[ 0 ], x1[ 1 ] - x0[ 1 ] };
    double normal[ 2 ] = { diff[ 1 ], -diff[ 0 ] };
    const double nrm = normn( normal, 2 );
    normal[ 0 ] /= nrm;
    normal[ 1 ] /= nrm;
    const double center[ 2 ] = { ( x0[ 0 ] + x1[ 0 ] ) * 0.5, ( x0[ 1 ] + x1[ 1 ] ) * 0.5 };
    vol += dotn( center, normal, 2 ) * normn( diff, 2 );
  }
  return fabs( vol / 2.0 );
}

static int in_first ( const int *ids, int N, int id )
{
  for( int i = 0; i < N; ++i )
    if( ids[ i ] == id )
      return 1;
  return 0;
}

/* PolygonWithDirections::initialize, 3d/polygonwithdirections.hh:76-174 */
static void sweep_initialize ( vo_sweep *s, const int *ids, const double *d, int nd )
{
  const vo_polyhedron *P = s->parent;
  s->nn = 0;
  s->nEdges = 0;
  s->nCF = 0;

  int N = 1;
  for( ; N < nd && d[ N ] == d[ 0 ]; N++ )
    ;

  switch( N )
  {
  case 1:
  {
    s->node[ 0 ][ 0 ] = P->x[ ids[ 0 ] ][ 0 ];
    s->node[ 0 ][ 1 ] = P->x[ ids[ 0 ] ][ 1 ];
    s->nodeId[ 0 ] = ids[ 0 ];
    attached_edges_to_node( P, ids[ 0 ], &s->dirs[ 0 ] );
    dirs_sort( P, &s->dirs[ 0 ] );
    s->nn = 1;
    break;
  }
  case 2:
  {
    for( int k = 0; k < 2; ++k )
    {
      s->node[ k ][ 0 ] = P->x[ ids[ k ] ][ 0 ];
      s->node[ k ][ 1 ] = P->x[ ids[ k ] ][ 1 ];
      s->nodeId[ k ] = ids[ k ];
    }
    vo_dirs d1, d2;
    attached_edges_to_node( P, ids[ 0 ], &d1 );
    attached_edges_to_node( P, ids[ 1 ], &d2 );
    /* :107-114; indices keep running after an erase exactly as in the reference (reads past size hit the stale
     * copy of the former last element, which is what a std::vector of trivially copyable edges leaves behind) */
    for( int i = 0; i < d1.n; ++i )
      for( int j = 0; j < d2.n; ++j )
        if( P->e[ d1.e[ i ] ][ 0 ] == P->e[ d2.e[ j ] ][ 1 ] && P->e[ d1.e[ i ] ][ 1 ] == P->e[ d2.e[ j ] ][ 0 ] )
        {
          const int stale1 = d1.e[ d1.n - 1 ], stale2 = d2.e[ d2.n - 1 ];
          dirs_erase( &d1, i );
          dirs_erase( &d2, j );
          d1.e[ d1.n ] = stale1;
          d2.e[ d2.n ] = stale2;
        }
    dirs_sort( P, &d1 );
    dirs_sort( P, &d2 );
    s->dirs[ 0 ] = d1;
    s->dirs[ 1 ] = d2;
    s->nn = 2;
    break;
  }
  default:
  {
    int found = 0;
    for( int f = 0; f < P->nf && !found; ++f )
    {
      int allFound = 1;
      for( int k = 0; k < P->fne[ f ]; ++k )
        allFound = allFound && in_first( ids, N, P->e[ P->fe[ f ][ k ] ][ 0 ] );
      if( allFound )
      {
        for( int k = 0; k < P->fne[ f ]; ++k )
        {
          const int id = P->e[ P->fe[ f ][ k ] ][ 0 ];
          s->nodeId[ s->nn ] = id;
          s->node[ s->nn ][ 0 ] = P->x[ id ][ 0 ];
          s->node[ s->nn ][ 1 ] = P->x[ id ][ 1 ];
          s->nn++;
        }
        found = 1;
      }
    }
    if( !found || s->nn < N )
    {
      s->degenerate = 1;   /* the reference indexes nodeIds_[k] for k < N: undefined behaviour */
      return;
    }
    for( int k = 0; k < N; ++k )
    {
      vo_dirs dd;
      attached_edges_to_node( P, s->nodeId[ k ], &dd );
      for( int j = 0; j < N; ++j )
        for( int i = 0; i < dd.n; ++i )
          if( P->e[ dd.e[ i ] ][ 1 ] == s->nodeId[ j ] )
          {
            dirs_erase( &dd, i );
            --i;
          }
      dirs_sort( P, &dd );
      s->dirs[ k ] = dd;
    }
    /* note: directions_ gets N entries while nodes_ has face-size entries; equal for a cube face (4) */
    break;
  }
  }

  if( s->nn != 1 )
  {
    for( int i = 0; i < s->nn; ++i )
    {
      const int f = attached_face_to_edge( P, s->nodeId[ ( i + 1 ) % s->nn ], s->nodeId[ i ] );
      if( f < 0 )
        s->degenerate = 1;   /* reference: DUNE_THROW */
      s->corrFace[ s->nCF++ ] = f;
    }
    s->nEdges = s->nn;
  }
}

/* PolygonWithDirections::evolveToNextPolygon, 3d/polygonwithdirections.hh:177-290 */
static void sweep_evolve ( vo_sweep *s, int k, double h, const double *dUnique )
{
  const vo_polyhedron *P = s->parent;
  double newNodes[ VO_MAXPN ][ 2 ];
  int newNodeIds[ VO_MAXPN ];
  vo_dirs newDirs[ VO_MAXPN ];
  int nNew = 0;
  s->nCF = 0;
  s->nEdges = 0;

  for( int n = 0; n < s->nn; ++n )
    for( int i = 0; i < s->dirs[ n ].n; ++i )
    {
      const int dirEdge = s->dirs[ n ].e[ i ];
      const int *de = P->e[ dirEdge ];
      const double z1 = P->x[ de[ 1 ] ][ 2 ];

      if( fabs( z1 - dUnique[ k + 1 ] ) < 1e-10 )
      {
        const int nodeId = de[ 1 ];
        if( nNew > 0 )
          if( nodeId == newNodeIds[ nNew - 1 ] )
          {
            s->nCF--;
            s->corrFace[ s->nCF++ ] = attached_face_to_edge( P, de[ 0 ], de[ 1 ] );
            continue;
          }
        if( nNew >= VO_MAXPN - 1 ) { s->degenerate = 1; return; }
        newNodes[ nNew ][ 0 ] = P->x[ nodeId ][ 0 ];
        newNodes[ nNew ][ 1 ] = P->x[ nodeId ][ 1 ];
        newNodeIds[ nNew ] = nodeId;
        vo_dirs all, dd;
        attached_edges_to_node( P, nodeId, &all );
        dd.n = 0;
        for( int a = 0; a < all.n; ++a )
          if( P->x[ P->e[ all.e[ a ] ][ 1 ] ][ 2 ] > dUnique[ k ] )
            dd.e[ dd.n++ ] = all.e[ a ];
        dirs_sort( P, &dd );
        newDirs[ nNew ] = dd;
        nNew++;
        s->corrFace[ s->nCF++ ] = attached_face_to_edge( P, de[ 0 ], de[ 1 ] );
      }
      else if( z1 > dUnique[ k + 1 ] )
      {
        if( nNew >= VO_MAXPN - 1 ) { s->degenerate = 1; return; }
        double dv[ 3 ];
        direction_vector( s, n, i, dv );
        double dv2[ 2 ] = { dv[ 0 ], dv[ 1 ] };
        const double f = h / dv[ 2 ];
        dv2[ 0 ] *= f;
        dv2[ 1 ] *= f;
        newNodes[ nNew ][ 0 ] = s->node[ n ][ 0 ] + dv2[ 0 ];
        newNodes[ nNew ][ 1 ] = s->node[ n ][ 1 ] + dv2[ 1 ];
        newNodeIds[ nNew ] = -1;
        newDirs[ nNew ].n = 1;
        newDirs[ nNew ].e[ 0 ] = dirEdge;
        nNew++;
        s->corrFace[ s->nCF++ ] = attached_face_to_edge( P, de[ 0 ], de[ 1 ] );
      }
    }

  if( nNew == 0 )
  {
    s->degenerate = 1;   /* the reference reads newNodeIds[0] of an empty vector */
    return;
  }

  if( newNodeIds[ 0 ] == newNodeIds[ nNew - 1 ] && newNodeIds[ 0 ] != -1 )
  {
    nNew--;
    s->nCF--;
  }

  for( int a = 0; a < nNew; ++a )
  {
    vo_dirs *dd = &newDirs[ a ];
    for( int i = 0; i < dd->n; ++i )
      for( int b = 0; b < nNew; ++b )
        if( i >= 0 && P->e[ dd->e[ i ] ][ 1 ] == newNodeIds[ b ] )
        {
          dirs_erase( dd, i );
          --i;
        }
  }

  if( nNew != 1 )
    s->nEdges = nNew;

  if( nNew > 2 )
    for( int i = 0; i < s->nEdges; ++i )
    {
      const int id0 = newNodeIds[ i ];
      const int id1 = newNodeIds[ ( i + 1 ) % nNew ];
      if( id0 != -1 && id1 != -1 )
      {
        int isEdge = 0;
        for( int e = 0; e < P->ne; ++e )
          if( P->e[ e ][ 0 ] == id0 && P->e[ e ][ 1 ] == id1 )
            isEdge = 1;
        if( isEdge )
          s->corrFace[ i ] = attached_face_to_edge( P, id1, id0 );
      }
    }

  s->nn = nNew;
  for( int a = 0; a < nNew; ++a )
  {
    s->node[ a ][ 0 ] = newNodes[ a ][ 0 ];
    s->node[ a ][ 1 ] = newNodes[ a ][ 1 ];
    s->nodeId[ a ] = newNodeIds[ a ];
    s->dirs[ a ] = newDirs[ a ];
  }
}

/* locateHalfSpace( Polyhedron ), geometry/algorithm.hh:119-236; out = (innerNormal[3], distance) */
static int locate_3d ( const vo_polyhedron *cell, const double *innerNormal, double fraction, double *out )
{
  double outerNormal[ 3 ];
  for( int k = 0; k < 3; ++k )
  {
    outerNormal[ k ] = innerNormal[ k ] * -1.0;
    out[ k ] = innerNormal[ k ];
  }

  fraction = ( fraction > 1.0 ) ? 1.0 : fraction;
  fraction = ( fraction < 0.0 ) ? 0.0 : fraction;

  const double givenVolume = fraction * polyhedron_volume( cell );

  vo_polyhedron polyhedron;
  rotate_to_reference_frame( outerNormal, cell, &polyhedron );

  const int N = polyhedron.nn;
  double d[ VO_MAXN ];
  int sortedNodesIds[ VO_MAXN ];
  for( int i = 0; i < N; ++i )
  {
    d[ i ] = polyhedron.x[ i ][ 2 ];
    sortedNodesIds[ i ] = i;
  }
  /* std::sort with N < 16 = insertion sort (stable): ties keep ascending node id, :157 */
  for( int i = 1; i < N; ++i )
  {
    const int val = sortedNodesIds[ i ];
    int j = i;
    while( j > 0 && d[ val ] < d[ sortedNodesIds[ j - 1 ] ] )
    {
      sortedNodesIds[ j ] = sortedNodesIds[ j - 1 ];
      --j;
    }
    sortedNodesIds[ j ] = val;
  }
  double ds[ VO_MAXN ] = { 0 };
  for( int i = 0; i < N; ++i )
    ds[ i ] = d[ sortedNodesIds[ i ] ];     /* == std::sort( d ), :158 */

  /* std::unique with |x-y| < 1e-10 against the last kept element, :160-162 */
  double dUnique[ VO_MAXN ];
  int nU = 0;
  dUnique[ nU++ ] = ds[ 0 ];
  for( int i = 1; i < N; ++i )
    if( !( fabs( dUnique[ nU - 1 ] - ds[ i ] ) < 1e-10 ) )
      dUnique[ nU++ ] = ds[ i ];

  vo_sweep polygon;
  memset( &polygon, 0, sizeof( polygon ) );
  polygon.parent = &polyhedron;
  sweep_initialize( &polygon, sortedNodesIds, ds, N );
  if( polygon.degenerate )
  {
    out[ 3 ] = NAN;
    return 1;
  }

  double Vd_k = 0, Vd_k1;
  for( int k = 0; k + 1 < nU; ++k )
  {
    const int Nn = polygon.nn;
    double A = 0.0;
    for( int n = 0; n < Nn; ++n )
    {
      if( polygon.dirs[ n ].n == 0 || polygon.dirs[ ( n + 1 ) % Nn ].n == 0 )
      {
        out[ 3 ] = NAN;   /* reference: assert( !directions( n ).empty() ) / out-of-range read */
        return 1;
      }
      const int epsn = polygon.dirs[ n ].n - 1;
      double v1[ 3 ], v2[ 3 ];
      direction_vector( &polygon, n, epsn, v1 );
      direction_vector( &polygon, ( n + 1 ) % Nn, 0, v2 );
      A -= ( v1[ 0 ] * v2[ 1 ] - v1[ 1 ] * v2[ 0 ] ) / ( v1[ 2 ] * v2[ 2 ] );
      for( int l = 0; l < epsn; ++l )
      {
        double w1[ 3 ], w2[ 3 ];
        direction_vector( &polygon, n, l, w1 );
        direction_vector( &polygon, n, l + 1, w2 );
        A -= ( w1[ 0 ] * w2[ 1 ] - w1[ 1 ] * w2[ 0 ] ) / ( w1[ 2 ] * w2[ 2 ] );
      }
    }

    double B = 0.0;
    for( int i = 0; i < polygon.nEdges; ++i )
    {
      if( i >= polygon.nCF || polygon.corrFace[ i ] < 0 )
      {
        out[ 3 ] = NAN;
        return 1;
      }
      double normal[ 3 ];
      face_outer_normal( &polyhedron, polygon.corrFace[ i ], normal );
      const double norm = normn( normal, 2 );
      if( norm > 0 )
      {
        const double *x0 = polygon.node[ i ];
        const double *x1 = polygon.node[ ( i + 1 ) % polygon.nn ];
        const double diff[ 2 ] = { x1[ 0 ] - x0[ 0 ], x1[ 1 ] - x0[ 1 ] };
        B -= normn( diff, 2 ) * normal[ 2 ] / norm;
      }
    }

    const double C = polygon.nEdges ? sweep_polygon_volume( &polygon ) : fabs( 0.0 / 2.0 );

    const double h = dUnique[ k + 1 ] - dUnique[ k ];
    cubic_ctx cc = { A, B, C, 0.0 };
    Vd_k1 = Vd_k + cubic_eval( h, &cc );   /* target 0: x - 0.0 == x */

    if( fabs( Vd_k1 - givenVolume ) < 1e-14 )
    {
      out[ 3 ] = dUnique[ k + 1 ];
      return 0;
    }
    else if( Vd_k1 > givenVolume )
    {
      cc.target = givenVolume - Vd_k;
      out[ 3 ] = dUnique[ k ] + brent( cubic_eval, &cc, 0.0, h, 1e-14 );
      return 0;
    }
    else
    {
      sweep_evolve( &polygon, k, h, dUnique );
      if( polygon.degenerate )
      {
        out[ 3 ] = NAN;
        return 1;
      }
    }
    Vd_k = Vd_k1;
  }

  out[ 3 ] = dUnique[ nU - 1 ];
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * G6 (3D)  intersect( Polyhedron, HyperPlane ) -> Dune::VoF::Face, 3d/intersect.hh:274-355; 3d/face.hh
 * ---------------------------------------------------------------------------------------------------------- */

#define VO_MAXIV 16
typedef struct
{
  int n;
  double v[ VO_MAXIV ][ 3 ];
  double unitNormal[ 3 ];
} vo_face;

/* Face constructor, 3d/face.hh:19-24 */
static void face_init_normal ( vo_face *f )
{
  const int n = f->n;
  double a[ 3 ], b[ 3 ];
  for( int k = 0; k < 3; ++k )
  {
    a[ k ] = f->v[ 2 % n ][ k ] - f->v[ 1 % n ][ k ];
    b[ k ] = f->v[ 0 ][ k ] - f->v[ 1 % n ][ k ];
  }
  cross3( a, b, f->unitNormal );
  const double nrm = normn( f->unitNormal, 3 );
  for( int k = 0; k < 3; ++k )
    f->unitNormal[ k ] /= nrm;
}

static int vec3_is_zero ( const double *x ) { return x[ 0 ] == 0.0 && x[ 1 ] == 0.0 && x[ 2 ] == 0.0; }

static void polyhedron_plane_cut ( const vo_polyhedron *poly, const double *hs, vo_face *out )
{
  out->n = 0;
  if( !hs_valid( hs, 3 ) )
    return;

  const double eps = 1e-12;
  int isInner[ VO_MAXN ];
  double edges[ VO_MAXIV ][ 2 ][ 3 ];
  int nE = 0;

  int n = 0;
  for( int i = 0; i < poly->nn; ++i )
  {
    isInner[ i ] = 0;
    if( level_set( hs, poly->x[ i ], 3 ) <= eps )
    {
      isInner[ i ] = 1;
      n++;
    }
  }

  if( n == 0 || n == poly->nn )
  {
    for( int i = 0; i < poly->nn; ++i )
      if( level_set( hs, poly->x[ i ], 3 ) >= -eps )
      {
        out->n = 1;
        memcpy( out->v[ 0 ], poly->x[ i ], 3 * sizeof( double ) );
        face_init_normal( out );
        return;
      }
    return;
  }

  for( int f = 0; f < poly->nf; ++f )
  {
    double last[ 3 ] = { 0.0, 0.0, 0.0 };
    for( int k = 0; k < poly->fne[ f ]; ++k )
    {
      const int *e = poly->e[ poly->fe[ f ][ k ] ];
      if( isInner[ e[ 0 ] ] ^ isInner[ e[ 1 ] ] )
      {
        double isPoint[ 3 ];
        edge_plane_point( poly->x[ e[ 0 ] ], poly->x[ e[ 1 ] ], hs, isPoint );
        if( !vec3_is_zero( last ) )
        {
          memcpy( edges[ nE ][ 0 ], last, sizeof( last ) );
          memcpy( edges[ nE ][ 1 ], isPoint, sizeof( isPoint ) );
          nE++;
        }
        else
          memcpy( last, isPoint, sizeof( last ) );
      }
    }
  }

  /* sort edges, :317-340 */
  for( int i = 0; i + 1 < nE; ++i )
  {
    double diff[ 3 ];
    for( int k = 0; k < 3; ++k )
      diff[ k ] = edges[ i ][ 1 ][ k ] - edges[ i + 1 ][ 0 ][ k ];
    if( normn( diff, 3 ) < eps )
      continue;

    const double *point = edges[ i ][ 1 ];
    int pos = -1;
    for( int j = i + 1; j < nE && pos < 0; ++j )
    {
      for( int k = 0; k < 3; ++k )
        diff[ k ] = point[ k ] - edges[ j ][ 0 ][ k ];
      if( normn( diff, 3 ) < eps )
        pos = j;
    }
    if( pos >= 0 )
    {
      double t[ 2 ][ 3 ];
      memcpy( t, edges[ i + 1 ], sizeof( t ) );
      memcpy( edges[ i + 1 ], edges[ pos ], sizeof( t ) );
      memcpy( edges[ pos ], t, sizeof( t ) );
      continue;
    }
    for( int j = i + 1; j < nE && pos < 0; ++j )
    {
      for( int k = 0; k < 3; ++k )
        diff[ k ] = point[ k ] - edges[ j ][ 1 ][ k ];
      if( normn( diff, 3 ) < eps )
        pos = j;
    }
    if( pos >= 0 )
    {
      double t[ 2 ][ 3 ];
      memcpy( t, edges[ i + 1 ], sizeof( t ) );
      memcpy( edges[ i + 1 ], edges[ pos ], sizeof( t ) );
      memcpy( edges[ pos ], t, sizeof( t ) );
      double s[ 3 ];
      memcpy( s, edges[ i + 1 ][ 0 ], sizeof( s ) );
      memcpy( edges[ i + 1 ][ 0 ], edges[ i + 1 ][ 1 ], sizeof( s ) );
      memcpy( edges[ i + 1 ][ 1 ], s, sizeof( s ) );
      continue;
    }
  }

  out->n = nE;
  for( int i = 0; i < nE; ++i )
    memcpy( out->v[ i ], edges[ i ][ 0 ], 3 * sizeof( double ) );
  if( nE > 0 )
    face_init_normal( out );
}

/* Face::volume, 3d/face.hh:92-98 with edgeCenter :111-117 and edgeOuterNormal :104-109 */
static double iface_volume ( const vo_face *f )
{
  double sum = 0;
  const int n = f->n;
  for( int i = 0; i < n; ++i )
  {
    const double *v0 = f->v[ i % n ], *v1 = f->v[ ( i + 1 ) % n ];
    double center[ 3 ], edge[ 3 ], en[ 3 ];
    for( int k = 0; k < 3; ++k )
    {
      center[ k ] = ( v1[ k ] + v0[ k ] ) * 0.5;
      edge[ k ] = v1[ k ] - v0[ k ];
    }
    cross3( edge, f->unitNormal, en );
    sum += dotn( center, en, 3 );
  }
  return fabs( sum ) / 2.0;
}

/* Face::triangleVolume, 3d/face.hh:52-67 */
static double iface_triangle_volume ( const vo_face *f, const double n[ 3 ][ 3 ] )
{
  double sum = 0;
  for( int i = 0; i < 3; ++i )
  {
    double center[ 3 ], edge[ 3 ], en[ 3 ];
    for( int k = 0; k < 3; ++k )
    {
      center[ k ] = ( n[ ( i + 1 ) % 3 ][ k ] + n[ i ][ k ] ) * 0.5;
      edge[ k ] = n[ ( i + 1 ) % 3 ][ k ] - n[ i ][ k ];
    }
    cross3( edge, f->unitNormal, en );
    sum += dotn( center, en, 3 );
  }
  return fabs( sum ) / 2.0;
}

/* Face::centroid, 3d/face.hh:69-90 */
static void iface_centroid ( const vo_face *f, double *centroid )
{
  const int n = f->n;
  double c[ 3 ] = { 0.0, 0.0, 0.0 };
  for( int i = 0; i < n; ++i )
    for( int k = 0; k < 3; ++k )
      c[ k ] += f->v[ i ][ k ];
  for( int k = 0; k < 3; ++k )
    c[ k ] /= (double) n;

  centroid[ 0 ] = centroid[ 1 ] = centroid[ 2 ] = 0.0;
  for( int i = 0; i < n; ++i )
  {
    double tri[ 3 ][ 3 ];
    memcpy( tri[ 0 ], f->v[ i % n ], sizeof( tri[ 0 ] ) );
    memcpy( tri[ 1 ], f->v[ ( i + 1 ) % n ], sizeof( tri[ 1 ] ) );
    memcpy( tri[ 2 ], c, sizeof( tri[ 2 ] ) );
    const double volTriang = iface_triangle_volume( f, tri );
    double cTriang[ 3 ];
    for( int k = 0; k < 3; ++k )
    {
      cTriang[ k ] = c[ k ];
      cTriang[ k ] += f->v[ i % n ][ k ];
      cTriang[ k ] += f->v[ ( i + 1 ) % n ][ k ];
      cTriang[ k ] /= 3.0;
      centroid[ k ] += volTriang * cTriang[ k ];
    }
  }
  const double vol = iface_volume( f );
  for( int k = 0; k < 3; ++k )
    centroid[ k ] /= vol;
}

/* ------------------------------------------------------------------------------------------------------------
 * G7  upwind polytopes and fluxes: geometry/upwindpolygon.hh:33-62, 2d/upwindpolygon.hh:24-30, 3d/upwindpolygon.hh:24-74
 * ---------------------------------------------------------------------------------------------------------- */

/* corners of face `face` (2*axis+side) of the cell [lo, lo+h]: intersection.geometry().corner(i), spanned axes ascending */
static int face_corners ( int dim, const double *lo, const double *h, int face, double c[ 4 ][ 3 ] )
{
  const int a = face / 2;
  double o[ 3 ] = { 0, 0, 0 };
  for( int d = 0; d < dim; ++d )
    o[ d ] = lo[ d ];
  if( face % 2 )
    o[ a ] += h[ a ];
  int axes[ 2 ], j = 0;
  for( int d = 0; d < dim; ++d )
    if( d != a )
      axes[ j++ ] = d;
  const int nc = 1 << ( dim - 1 );
  for( int i = 0; i < nc; ++i )
  {
    for( int d = 0; d < dim; ++d )
      c[ i ][ d ] = o[ d ];
    for( int jj = 0; jj < dim - 1; ++jj )
      if( ( i >> jj ) & 1 )
        c[ i ][ axes[ jj ] ] += h[ axes[ jj ] ];
  }
  return nc;
}

/* truncVolume( upwindPolygon2d( face, v ), hs ): 2d/upwindpolygon.hh:24-30 + geometry/upwindpolygon.hh:58-62 */
static double flux_2d ( const double *lo, const double *h, int face, const double *v, const double *hs )
{
  double c[ 4 ][ 3 ];
  face_corners( 2, lo, h, face, c );
  const double d[ 2 ] = { c[ 1 ][ 0 ] - c[ 0 ][ 0 ], c[ 1 ][ 1 ] - c[ 0 ][ 1 ] };
  const double cr[ 2 ] = { -d[ 1 ], d[ 0 ] };   /* generalizedCrossProduct 2D, geometry/utility.hh:82-85 */
  vo_polygon up, cl;
  up.n = 4;
  if( dotn( cr, v, 2 ) < 0.0 )
  {
    up.v[ 0 ][ 0 ] = c[ 0 ][ 0 ]; up.v[ 0 ][ 1 ] = c[ 0 ][ 1 ];
    up.v[ 1 ][ 0 ] = c[ 1 ][ 0 ]; up.v[ 1 ][ 1 ] = c[ 1 ][ 1 ];
    up.v[ 2 ][ 0 ] = c[ 1 ][ 0 ] - v[ 0 ]; up.v[ 2 ][ 1 ] = c[ 1 ][ 1 ] - v[ 1 ];
    up.v[ 3 ][ 0 ] = c[ 0 ][ 0 ] - v[ 0 ]; up.v[ 3 ][ 1 ] = c[ 0 ][ 1 ] - v[ 1 ];
  }
  else
  {
    up.v[ 0 ][ 0 ] = c[ 1 ][ 0 ]; up.v[ 0 ][ 1 ] = c[ 1 ][ 1 ];
    up.v[ 1 ][ 0 ] = c[ 0 ][ 0 ]; up.v[ 1 ][ 1 ] = c[ 0 ][ 1 ];
    up.v[ 2 ][ 0 ] = c[ 0 ][ 0 ] - v[ 0 ]; up.v[ 2 ][ 1 ] = c[ 0 ][ 1 ] - v[ 1 ];
    up.v[ 3 ][ 0 ] = c[ 1 ][ 0 ] - v[ 0 ]; up.v[ 3 ][ 1 ] = c[ 1 ][ 1 ] - v[ 1 ];
  }
  polygon_clip( &up, hs, &cl );
  return polygon_volume( &cl );
}

/* truncVolume( upwindPolygon3d( face, v ), hs ): 3d/upwindpolygon.hh:24-74 + geometry/upwindpolygon.hh:58-62 */
static double flux_3d ( const double *lo, const double *h, int face, const double *v, const double *hs )
{
  double c[ 4 ][ 3 ];
  face_corners( 3, lo, h, face, c );
  double a[ 3 ], b[ 3 ], cr[ 3 ];
  for( int k = 0; k < 3; ++k )
  {
    a[ k ] = c[ 1 ][ k ] - c[ 0 ][ k ];
    b[ k ] = c[ 2 ][ k ] - c[ 0 ][ k ];
  }
  cross3( a, b, cr );
  vo_polyhedron up, cl;
  make_cube_topology( &up );
  const int shiftedFirst = dotn( cr, v, 3 ) < 0.0;
  for( int i = 0; i < 4; ++i )
    for( int k = 0; k < 3; ++k )
    {
      const double moved = c[ i ][ k ] - v[ k ];
      up.x[ shiftedFirst ? i : i + 4 ][ k ] = moved;
      up.x[ shiftedFirst ? i + 4 : i ][ k ] = c[ i ][ k ];
    }
  polyhedron_clip( &up, hs, &cl );
  return polyhedron_volume( &cl );
}

/* ------------------------------------------------------------------------------------------------------------
 * grid helpers (conventions of vof_oracle.h; identical to oracle/shim/dune/grid/common/minigrid.hh)
 * ---------------------------------------------------------------------------------------------------------- */

static int64_t grid_cells ( const vo_grid *g )
{
  int64_t c = 1;
  for( int d = 0; d < g->dim; ++d )
    c *= g->n[ d ];
  return c;
}

static void cell_ijk ( const vo_grid *g, int64_t cell, int *ijk )
{
  ijk[ 0 ] = ijk[ 1 ] = ijk[ 2 ] = 0;
  for( int d = 0; d < g->dim; ++d )
  {
    ijk[ d ] = (int) ( cell % g->n[ d ] );
    cell /= g->n[ d ];
  }
}

static int64_t cell_index ( const vo_grid *g, const int *ijk )
{
  int64_t k = 0;
  for( int d = g->dim - 1; d >= 0; --d )
    k = k * g->n[ d ] + ijk[ d ];
  return k;
}

static int in_domain ( const vo_grid *g, const int *ijk )
{
  for( int d = 0; d < g->dim; ++d )
    if( ijk[ d ] < 0 || ijk[ d ] >= g->n[ d ] )
      return 0;
  return 1;
}

static void cell_lower ( const vo_grid *g, const int *ijk, double *lo )
{
  for( int d = 0; d < g->dim; ++d )
    lo[ d ] = g->origin[ d ] + ijk[ d ] * g->h[ d ];
}

static void cell_center ( const vo_grid *g, const int *ijk, double *c )
{
  for( int d = 0; d < g->dim; ++d )
  {
    c[ d ] = g->origin[ d ] + ijk[ d ] * g->h[ d ];
    c[ d ] += 0.5 * g->h[ d ];
  }
}

static double cell_volume ( const vo_grid *g )
{
  double v = 1;
  for( int d = 0; d < g->dim; ++d )
    v *= g->h[ d ];
  return v;
}

static double face_volume_of ( const vo_grid *g, int face )
{
  double v = 1;
  for( int d = 0; d < g->dim; ++d )
    if( d != face / 2 )
      v *= g->h[ d ];
  return v;
}

static int is_mixed ( uint8_t f ) { return f == 1 || f == 2; }   /* flagset.hh:71-72,78 */

static void locate_cell ( const vo_grid *g, const int *ijk, const double *normal, double fraction, double *hs )
{
  double lo[ 3 ];
  cell_lower( g, ijk, lo );
  if( g->dim == 2 )
  {
    vo_polygon p;
    make_polygon_cell( lo, g->h, &p );
    locate_2d( &p, normal, fraction, hs );
  }
  else
  {
    vo_polyhedron p;
    make_polyhedron_cell( lo, g->h, &p );
    locate_3d( &p, normal, fraction, hs );
  }
}

/* ------------------------------------------------------------------------------------------------------------
 * F1  FlagOperator, flagging.hh:35-70
 * ---------------------------------------------------------------------------------------------------------- */

int vo_flag ( const vo_grid *g, const double *u, double eps, uint8_t *flags )
{
  const int64_t N = grid_cells( g );
  for( int64_t c = 0; c < N; ++c )
  {
    const double colorEn = u[ c ];
    uint8_t flag;
    if( colorEn < eps )
      flag = 0;
    else if( colorEn <= ( 1 - eps ) )
      flag = 1;
    else
    {
      if( isnan( colorEn ) )
        flag = 99;
      else
      {
        flag = 3;
        int ijk[ 3 ];
        cell_ijk( g, c, ijk );
        for( int face = 0; face < 2 * g->dim; ++face )
        {
          int nb[ 3 ] = { ijk[ 0 ], ijk[ 1 ], ijk[ 2 ] };
          nb[ face / 2 ] += ( face % 2 ) ? 1 : -1;
          if( in_domain( g, nb ) && u[ cell_index( g, nb ) ] < eps )
          {
            flag = 2;
            break;
          }
        }
      }
    }
    flags[ c ] = flag;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * R1  ModifiedYoungsReconstruction, reconstruction/modifiedyoungs.hh:63-134
 *     (vertex stencil: stencil/vertexneighborsstencil.hh:52-94 = 3^d neighbourhood, ascending index, self excluded)
 * ---------------------------------------------------------------------------------------------------------- */

/* FieldMatrix::solve: dune-common 2.6 densematrix.hh closed forms (third-party; parity unpinned, see header) */
static void solve_small ( int dim, double M[ 3 ][ 3 ], const double *b, double *x )
{
  if( dim == 2 )
  {
    double detinv = M[ 0 ][ 0 ] * M[ 1 ][ 1 ] - M[ 0 ][ 1 ] * M[ 1 ][ 0 ];
    detinv = 1.0 / detinv;
    x[ 0 ] = detinv * ( M[ 1 ][ 1 ] * b[ 0 ] - M[ 0 ][ 1 ] * b[ 1 ] );
    x[ 1 ] = detinv * ( M[ 0 ][ 0 ] * b[ 1 ] - M[ 1 ][ 0 ] * b[ 0 ] );
    return;
  }
  const double t4 = M[ 0 ][ 0 ] * M[ 1 ][ 1 ];
  const double t6 = M[ 0 ][ 0 ] * M[ 1 ][ 2 ];
  const double t8 = M[ 0 ][ 1 ] * M[ 1 ][ 0 ];
  const double t10 = M[ 0 ][ 2 ] * M[ 1 ][ 0 ];
  const double t12 = M[ 0 ][ 1 ] * M[ 2 ][ 0 ];
  const double t14 = M[ 0 ][ 2 ] * M[ 2 ][ 0 ];
  const double d = ( t4 * M[ 2 ][ 2 ] - t6 * M[ 2 ][ 1 ] - t8 * M[ 2 ][ 2 ] + t10 * M[ 2 ][ 1 ] + t12 * M[ 1 ][ 2 ] - t14 * M[ 1 ][ 1 ] );
  x[ 0 ] = ( b[ 0 ] * M[ 1 ][ 1 ] * M[ 2 ][ 2 ] - b[ 0 ] * M[ 2 ][ 1 ] * M[ 1 ][ 2 ]
             - b[ 1 ] * M[ 0 ][ 1 ] * M[ 2 ][ 2 ] + b[ 1 ] * M[ 2 ][ 1 ] * M[ 0 ][ 2 ]
             + b[ 2 ] * M[ 0 ][ 1 ] * M[ 1 ][ 2 ] - b[ 2 ] * M[ 1 ][ 1 ] * M[ 0 ][ 2 ] ) / d;
  x[ 1 ] = ( M[ 0 ][ 0 ] * b[ 1 ] * M[ 2 ][ 2 ] - M[ 0 ][ 0 ] * b[ 2 ] * M[ 1 ][ 2 ]
             - M[ 1 ][ 0 ] * b[ 0 ] * M[ 2 ][ 2 ] + M[ 1 ][ 0 ] * b[ 2 ] * M[ 0 ][ 2 ]
             + M[ 2 ][ 0 ] * b[ 0 ] * M[ 1 ][ 2 ] - M[ 2 ][ 0 ] * b[ 1 ] * M[ 0 ][ 2 ] ) / d;
  x[ 2 ] = ( M[ 0 ][ 0 ] * M[ 1 ][ 1 ] * b[ 2 ] - M[ 0 ][ 0 ] * M[ 2 ][ 1 ] * b[ 1 ]
             - M[ 1 ][ 0 ] * M[ 0 ][ 1 ] * b[ 2 ] + M[ 1 ][ 0 ] * M[ 2 ][ 1 ] * b[ 0 ]
             + M[ 2 ][ 0 ] * M[ 0 ][ 1 ] * b[ 1 ] - M[ 2 ][ 0 ] * M[ 1 ][ 1 ] * b[ 0 ] ) / d;
}

/* one least-squares contribution, modifiedyoungs.hh:104-108 / :117-123 */
static void ls_accumulate ( int dim, double *d, double colorDiffNbEn, double AtA[ 3 ][ 3 ], double *Atb )
{
  const double weight = 1.0 / norm2n( d, dim );
  for( int k = 0; k < dim; ++k )
    d[ k ] *= weight;
  for( int i = 0; i < dim; ++i )
    for( int j = 0; j < dim; ++j )
    {
      double m = 0.0;             /* outerProduct, geometry/utility.hh:161-169: m[i].axpy( a[i], b ) on a zero matrix */
      m += d[ i ] * d[ j ];
      AtA[ i ][ j ] += m;
    }
  const double w = weight * colorDiffNbEn;
  for( int k = 0; k < dim; ++k )
    Atb[ k ] += w * d[ k ];
}

/* applyLocal, modifiedyoungs.hh:90-134; hs is left untouched if the normal degenerates (:128-129) */
static void youngs_local ( const vo_grid *g, const double *u, int64_t cell, double *hs )
{
  const int dim = g->dim;
  int ijk[ 3 ];
  cell_ijk( g, cell, ijk );
  double center[ 3 ];
  cell_center( g, ijk, center );
  const double colorEn = u[ cell ];

  double AtA[ 3 ][ 3 ] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
  double Atb[ 3 ] = { 0, 0, 0 };

  const int zlo = ( dim == 3 ) ? -1 : 0, zhi = ( dim == 3 ) ? 1 : 0;
  for( int dz = zlo; dz <= zhi; ++dz )
    for( int dy = -1; dy <= 1; ++dy )
      for( int dx = -1; dx <= 1; ++dx )
      {
        if( dx == 0 && dy == 0 && dz == 0 )
          continue;
        int nb[ 3 ] = { ijk[ 0 ] + dx, ijk[ 1 ] + dy, ijk[ 2 ] + dz };
        if( !in_domain( g, nb ) )
          continue;
        const double colorNb = clampd( u[ cell_index( g, nb ) ], 0.0, 1.0 );
        double cnb[ 3 ], d[ 3 ];
        cell_center( g, nb, cnb );
        for( int k = 0; k < dim; ++k )
          d[ k ] = cnb[ k ] - center[ k ];
        ls_accumulate( dim, d, colorNb - colorEn, AtA, Atb );
      }

  /* boundary faces get a ghost value of 0.5, :112-124 */
  double lo[ 3 ];
  cell_lower( g, ijk, lo );
  for( int face = 0; face < 2 * dim; ++face )
  {
    int nb[ 3 ] = { ijk[ 0 ], ijk[ 1 ], ijk[ 2 ] };
    nb[ face / 2 ] += ( face % 2 ) ? 1 : -1;
    if( in_domain( g, nb ) )
      continue;
    /* intersection.geometry().center(): lower corner of the face + 0.5*h along the spanned axes */
    double fc[ 3 ], d[ 3 ];
    for( int k = 0; k < dim; ++k )
      fc[ k ] = lo[ k ];
    if( face % 2 )
      fc[ face / 2 ] += g->h[ face / 2 ];
    for( int k = 0; k < dim; ++k )
      if( k != face / 2 )
        fc[ k ] += 0.5 * g->h[ k ];
    for( int k = 0; k < dim; ++k )
    {
      d[ k ] = fc[ k ] - center[ k ];
      d[ k ] *= 2.0;
    }
    ls_accumulate( dim, d, 0.5 - colorEn, AtA, Atb );
  }

  double normal[ 3 ] = { 0, 0, 0 };
  solve_small( dim, AtA, Atb, normal );

  if( norm2n( normal, dim ) < VO_EPS )
    return;

  const double nrm = normn( normal, dim );
  for( int k = 0; k < dim; ++k )
    normal[ k ] /= nrm;

  locate_cell( g, ijk, normal, colorEn, hs );
}

static void youngs_all ( const vo_grid *g, const double *u, const uint8_t *flags, double *hs )
{
  const int64_t N = grid_cells( g );
  memset( hs, 0, sizeof( double ) * (size_t) N * ( g->dim + 1 ) );   /* reconstructions.clear(), :65 */
  for( int64_t c = 0; c < N; ++c )
    if( is_mixed( flags[ c ] ) )
      youngs_local( g, u, c, hs + c * ( g->dim + 1 ) );
}

/* ------------------------------------------------------------------------------------------------------------
 * S2 + R2  HeightFunctionStencil (stencil/heightfunctionstencil.hh:22-107) and
 *          HeightFunctionReconstruction (reconstruction/heightfunction.hh:70-259)
 * ---------------------------------------------------------------------------------------------------------- */

/* getOrientation, heightfunction.hh:126-144 */
static void hf_orientation ( int dim, const double *normal, int *dir, int *sign )
{
  int dd = 0;
  double max = DBL_MIN;
  for( int i = 0; i < dim; ++i )
    if( fabs( normal[ i ] ) > max )
    {
      dd = i;
      max = fabs( normal[ i ] );
    }
  if( fabs( max - 1.0 / sqrt( 2.0 ) ) < 1e-3 )
    dd = dim - 1;
  *dir = dd;
  *sign = ( normal[ dd ] > 0 ) ? -1 : 1;
}

/* getMultiIndex, heightfunctionstencil.hh:69-100, in cell-index units (the reference works on SPGrid ids 2*i+1) */
static void hf_offset ( int dim, int i, int j, int c, int t, int *m )
{
  m[ 0 ] = m[ 1 ] = m[ 2 ] = 0;
  if( dim == 2 )
  {
    m[ 1 - i ] += ( c - 1 ) * ( i == 0 ? -1 : 1 );
    m[ i ] += t;
    m[ 0 ] *= j;
    m[ 1 ] *= j;
  }
  else
  {
    m[ i ] += j * t;
    m[ ( i + 1 ) % 3 ] += ( ( c % 3 ) - 1 ) * -j;
    m[ ( i + 2 ) % 3 ] += ( ( c / 3 ) - 1 );
  }
}

static int hf_cell ( const vo_grid *g, const int *ijk, int i, int j, int c, int t, int64_t *cell )
{
  int m[ 3 ], nb[ 3 ];
  hf_offset( g->dim, i, j, c, t, m );
  for( int d = 0; d < 3; ++d )
    nb[ d ] = ijk[ d ] + m[ d ];
  if( !in_domain( g, nb ) )
    return 0;          /* valid(), heightfunctionstencil.hh:63-66 */
  *cell = cell_index( g, nb );
  return 1;
}

/* effectiveTdown, heightfunctionstencil.hh:47-53 */
static int hf_effective_tdown ( const vo_grid *g, const int *ijk, int i, int j, int tup )
{
  const int noc = ( g->dim == 2 ) ? 3 : 9;
  int64_t dummy;
  for( int t = -1; t >= -tup; --t )
    if( !hf_cell( g, ijk, i, j, ( noc - 1 ) / 2, t, &dummy ) )
      return -1 - t;
  return tup;
}

/* getHeightValues, heightfunction.hh:147-196 */
static void hf_heights ( const vo_grid *g, const double *u, const int *ijk, int i, int j, int tup, double *heights )
{
  const int noc = ( g->dim == 2 ) ? 3 : 9;
  const double TOL = 1e-10;
  for( int c = 0; c < noc; ++c )
  {
    heights[ c ] = 0.0;
    int64_t cell;
    if( !hf_cell( g, ijk, i, j, c, 0, &cell ) )
      continue;
    const double u0 = clampd( u[ cell ], 0.0, 1.0 );
    heights[ c ] += u0;

    double lastU = u0;
    for( int t = 1; t <= tup; ++t )
    {
      if( !hf_cell( g, ijk, i, j, c, t, &cell ) )
        break;
      const double uu = clampd( u[ cell ], 0.0, 1.0 );
      if( uu > lastU - TOL && !( uu > 1.0 - TOL ) )
        break;
      heights[ c ] += uu;
      lastU = uu;
    }

    lastU = u0;
    for( int t = -1; t >= -tup; --t )
    {
      if( !hf_cell( g, ijk, i, j, c, t, &cell ) )
        break;
      double uu = clampd( u[ cell ], 0.0, 1.0 );
      if( uu < lastU + TOL && !( uu < TOL ) )
        uu = 1.0;
      heights[ c ] += uu;
      lastU = uu;
    }
  }
}

/* computeNormal, heightfunction.hh:200-226 (2D) and :229-259 (3D) */
static void hf_normal ( int dim, const double *heights, int i, int j, double *n )
{
  if( dim == 2 )
  {
    double Hx;
    if( fabs( heights[ 0 ] ) < VO_EPS )
      Hx = ( heights[ 2 ] - heights[ 1 ] );
    else if( fabs( heights[ 2 ] ) < VO_EPS )
      Hx = ( heights[ 1 ] - heights[ 0 ] );
    else
      Hx = ( heights[ 2 ] - heights[ 0 ] ) / 2.0;

    n[ 0 ] = Hx;
    n[ 1 ] = -1.0;
    const double nrm = normn( n, 2 );
    n[ 0 ] /= nrm;
    n[ 1 ] /= nrm;
    if( ( i == 0 && j == 1 ) || ( i == 1 && j == -1 ) )
    {
      n[ 0 ] *= -1.0;
      n[ 1 ] *= -1.0;
    }
    if( ( i == 0 && j == 1 ) || ( i == 0 && j == -1 ) )
    {
      const double a = -n[ 1 ], b = n[ 0 ];
      n[ 0 ] = a;
      n[ 1 ] = b;
    }
    return;
  }

  double Hx, Hy;
  if( heights[ 5 ] == 0.0 )
    Hx = ( heights[ 4 ] - heights[ 3 ] );
  else if( heights[ 3 ] == 0.0 )
    Hx = ( heights[ 5 ] - heights[ 4 ] );
  else
    Hx = ( heights[ 5 ] - heights[ 3 ] ) / 2.0;

  if( heights[ 7 ] == 0.0 )
    Hy = ( heights[ 4 ] - heights[ 1 ] );
  else if( heights[ 1 ] == 0.0 )
    Hy = ( heights[ 7 ] - heights[ 4 ] );
  else
    Hy = ( heights[ 7 ] - heights[ 1 ] ) / 2.0;

  n[ 0 ] = n[ 1 ] = n[ 2 ] = 0.0;
  n[ i ] = -j;
  n[ ( i + 1 ) % 3 ] = Hx * -j;
  n[ ( i + 2 ) % 3 ] = Hy;
  const double nrm = normn( n, 3 );
  for( int k = 0; k < 3; ++k )
    n[ k ] /= nrm;
}

static void hf_all ( const vo_grid *g, const double *u, const uint8_t *flags, double *hs )
{
  const int dim = g->dim;
  const int64_t N = grid_cells( g );
  youngs_all( g, u, flags, hs );                /* initializer_, heightfunction.hh:73 */
  for( int64_t c = 0; c < N; ++c )
  {
    if( !is_mixed( flags[ c ] ) )
      continue;
    double *rec = hs + c * ( dim + 1 );
    int ijk[ 3 ], dir, sign;
    cell_ijk( g, c, ijk );
    hf_orientation( dim, rec, &dir, &sign );
    double heights[ 9 ], newNormal[ 3 ];
    hf_heights( g, u, ijk, dir, sign, 3, heights );
    hf_normal( dim, heights, dir, sign, newNormal );
    locate_cell( g, ijk, newNormal, u[ c ], rec );
  }
}

/* ------------------------------------------------------------------------------------------------------------
 * R3  ModifiedSwartzReconstruction, reconstruction/modifiedswartz.hh:70-243.
 *     The reference recomputes the neighbours' planes inside the pair loop (:209-211); they depend only on
 *     (neighbour, oldNormal), so computing each once per iteration gives identical values.
 * ---------------------------------------------------------------------------------------------------------- */

/* centroid of interface( cell, locateHalfSpace( cell, normal, u ) ); returns 0 if the interface is empty */
static int interface_centroid ( const vo_grid *g, const int *ijk, const double *normal, double color, double *centroid, double *hsOut )
{
  double lo[ 3 ], hs[ 4 ];
  cell_lower( g, ijk, lo );
  locate_cell( g, ijk, normal, color, hs );
  if( hsOut )
    memcpy( hsOut, hs, sizeof( double ) * ( g->dim + 1 ) );
  if( g->dim == 2 )
  {
    vo_polygon p;
    double line[ 2 ][ 2 ];
    make_polygon_cell( lo, g->h, &p );
    const int ok = polygon_plane_cut( &p, hs, line );
    centroid[ 0 ] = ( line[ 0 ][ 0 ] + line[ 1 ][ 0 ] ) * 0.5;   /* Line::centroid, 2d/polygon.hh:183-188 */
    centroid[ 1 ] = ( line[ 0 ][ 1 ] + line[ 1 ][ 1 ] ) * 0.5;
    return ok;
  }
  vo_polyhedron p;
  vo_face f;
  make_polyhedron_cell( lo, g->h, &p );
  polyhedron_plane_cut( &p, hs, &f );
  if( f.n == 0 )
  {
    centroid[ 0 ] = centroid[ 1 ] = centroid[ 2 ] = NAN;   /* reference: centroid of an empty Face = 0/0 */
    return 0;
  }
  iface_centroid( &f, centroid );
  return 1;
}

static void swartz_local ( const vo_grid *g, const double *u, const uint8_t *flags, int64_t cell, double *hsAll, int maxIterations )
{
  const int dim = g->dim;
  double *reconstruction = hsAll + cell * ( dim + 1 );
  int ijk[ 3 ];
  cell_ijk( g, cell, ijk );

  /* mixed vertex neighbours in ascending index order */
  int nbijk[ 26 ][ 3 ];
  int64_t nbcell[ 26 ];
  int nNb = 0;
  const int zlo = ( dim == 3 ) ? -1 : 0, zhi = ( dim == 3 ) ? 1 : 0;
  for( int dz = zlo; dz <= zhi; ++dz )
    for( int dy = -1; dy <= 1; ++dy )
      for( int dx = -1; dx <= 1; ++dx )
      {
        if( dx == 0 && dy == 0 && dz == 0 )
          continue;
        int nb[ 3 ] = { ijk[ 0 ] + dx, ijk[ 1 ] + dy, ijk[ 2 ] + dz };
        if( !in_domain( g, nb ) )
          continue;
        const int64_t c = cell_index( g, nb );
        if( !is_mixed( flags[ c ] ) )
          continue;
        memcpy( nbijk[ nNb ], nb, sizeof( nb ) );
        nbcell[ nNb ] = c;
        nNb++;
      }

  double normal[ 3 ] = { reconstruction[ 0 ], reconstruction[ 1 ], dim == 3 ? reconstruction[ 2 ] : 0.0 };
  int iterations = 0;
  double residuum = 0.0;

  do
  {
    double oldNormal[ 3 ] = { normal[ 0 ], normal[ 1 ], normal[ 2 ] };
    normal[ 0 ] = normal[ 1 ] = normal[ 2 ] = 0.0;

    double cEn[ 3 ] = { 0, 0, 0 };
    interface_centroid( g, ijk, oldNormal, u[ cell ], cEn, NULL );

    double cNb[ 26 ][ 3 ];
    for( int a = 0; a < nNb; ++a )
    {
      cNb[ a ][ 0 ] = cNb[ a ][ 1 ] = cNb[ a ][ 2 ] = 0.0;
      interface_centroid( g, nbijk[ a ], oldNormal, u[ nbcell[ a ] ], cNb[ a ], NULL );
    }

    if( dim == 2 )
    {
      for( int a = 0; a < nNb; ++a )
      {
        const double dd[ 2 ] = { cNb[ a ][ 0 ] - cEn[ 0 ], cNb[ a ][ 1 ] - cEn[ 1 ] };
        double centerNormal[ 2 ] = { -dd[ 1 ], dd[ 0 ] };
        if( dotn( centerNormal, oldNormal, 2 ) < 0 )
        {
          centerNormal[ 0 ] *= -1.0;
          centerNormal[ 1 ] *= -1.0;
        }
        const double nrm = normn( centerNormal, 2 );
        centerNormal[ 0 ] /= nrm;
        centerNormal[ 1 ] /= nrm;
        const double weight = 1.0;
        normal[ 0 ] += weight * centerNormal[ 0 ];
        normal[ 1 ] += weight * centerNormal[ 1 ];
      }
    }
    else
    {
      for( int a = 0; a < nNb; ++a )
        for( int b = 0; b < nNb; ++b )
        {
          if( a == b )
            continue;
          double d1[ 3 ], d2[ 3 ], centerNormal[ 3 ];
          for( int k = 0; k < 3; ++k )
          {
            d1[ k ] = cNb[ a ][ k ] - cEn[ k ];
            d2[ k ] = cNb[ b ][ k ] - cEn[ k ];
          }
          cross3( d1, d2, centerNormal );
          if( normn( centerNormal, 3 ) < 1e-12 )
            continue;
          if( dotn( centerNormal, oldNormal, 3 ) < 0 )
            for( int k = 0; k < 3; ++k )
              centerNormal[ k ] *= -1.0;
          const double nrm = normn( centerNormal, 3 );
          const double weight = 1.0;
          for( int k = 0; k < 3; ++k )
          {
            centerNormal[ k ] /= nrm;
            normal[ k ] += weight * centerNormal[ k ];
          }
        }
    }

    if( normn( normal, dim ) < 0.5 )
      return;

    const double nrm = normn( normal, dim );
    for( int k = 0; k < dim; ++k )
      normal[ k ] /= nrm;

    locate_cell( g, ijk, normal, u[ cell ], reconstruction );

    residuum = 1.0 - dotn( normal, oldNormal, dim );
    ++iterations;
  }
  while( residuum > 1e-12 && iterations < maxIterations );
}

static void swartz_all ( const vo_grid *g, const double *u, const uint8_t *flags, double *hs, int maxIterations )
{
  const int64_t N = grid_cells( g );
  youngs_all( g, u, flags, hs );     /* initializer()(..), modifiedswartz.hh:72 */
  for( int64_t c = 0; c < N; ++c )
    if( is_mixed( flags[ c ] ) )
      swartz_local( g, u, flags, c, hs, maxIterations );
}

int vo_reconstruct ( const vo_grid *g, int kind, const double *u, const uint8_t *flags, double *hs, int max_iterations )
{
  switch( kind )
  {
  case VO_RECON_YOUNGS: youngs_all( g, u, flags, hs ); return 0;
  case VO_RECON_HF: hf_all( g, u, flags, hs ); return 0;
  case VO_RECON_SWARTZ: swartz_all( g, u, flags, hs, max_iterations ); return 0;
  }
  return 1;
}

/* ------------------------------------------------------------------------------------------------------------
 * E2  velocity fields (the headers in test/problems restated; LeVeque from oracle/problems_extra.hh)
 * ---------------------------------------------------------------------------------------------------------- */

/* cos(pi*s) from + - * only; this DEFINES the time factor of the LeVeque problems (oracle/problems_extra.hh) */
double vo_det_cospi ( double s )
{
  if( s < 0.0 )
    s = -s;
  double r = s - 2.0 * floor( s * 0.5 );
  if( r > 1.0 )
    r = 2.0 - r;
  double sign = 1.0;
  if( r > 0.5 )
  {
    r = 1.0 - r;
    sign = -1.0;
  }
  const double pi = 3.14159265358979323846;
  if( r <= 0.25 )
  {
    const double x = pi * r;
    const double z = x * x;
    double p = 1.0 / 20922789888000.0;
    p = p * z - 1.0 / 87178291200.0;
    p = p * z + 1.0 / 479001600.0;
    p = p * z - 1.0 / 3628800.0;
    p = p * z + 1.0 / 40320.0;
    p = p * z - 1.0 / 720.0;
    p = p * z + 1.0 / 24.0;
    p = p * z - 0.5;
    p = p * z + 1.0;
    return sign * p;
  }
  else
  {
    const double x = pi * ( 0.5 - r );
    const double z = x * x;
    double p = 1.0 / 355687428096000.0;
    p = p * z - 1.0 / 1307674368000.0;
    p = p * z + 1.0 / 6227020800.0;
    p = p * z - 1.0 / 39916800.0;
    p = p * z + 1.0 / 362880.0;
    p = p * z - 1.0 / 5040.0;
    p = p * z + 1.0 / 120.0;
    p = p * z - 1.0 / 6.0;
    p = p * z + 1.0;
    return sign * ( p * x );
  }
}

static double sin_sq ( double x ) { const double s = sin( M_PI * x ); return s * s; }
static double sin_two ( double x ) { return sin( 2.0 * M_PI * x ); }

static int velocity_field ( int dim, int problem, const double *x, double t, double *v )
{
  v[ 0 ] = v[ 1 ] = v[ 2 ] = 0.0;
  switch( problem )
  {
  case VO_PROB_ROTATING_CIRCLE:
    if( dim == 2 )
    {
      /* rotatingcircle.hh:45-52 */
      v[ 0 ] = x[ 1 ] - 0.5;
      v[ 1 ] = 0.5 - x[ 0 ];
      v[ 0 ] *= 2 * M_PI / 10;
      v[ 1 ] *= 2 * M_PI / 10;
    }
    else
    {
      /* rotatingcircle.hh:92-101 */
      const double c0 = x[ 0 ] - 0.5, c1 = x[ 1 ] - 0.5;
      v[ 0 ] = c1;
      v[ 1 ] = -c0;
      v[ 2 ] = 0;
      for( int k = 0; k < 3; ++k )
        v[ k ] *= 2 * M_PI / 10;
    }
    return 0;
  case VO_PROB_SFLOW:
  {
    /* sflow.hh:35-42 / :73-82 */
    const double y0 = 4 * x[ 0 ] - 2, y1 = 4 * x[ 1 ] - 2;
    v[ 0 ] = 0.25 * ( y0 + y1 * y1 * y1 );
    v[ 1 ] = -0.25 * ( y1 + y0 * y0 * y0 );
    return 0;
  }
  case VO_PROB_SLOTTED_CYLINDER:
  {
    /* slottedcylinder.hh:47-55 */
    if( dim != 2 )
      return 1;
    double r0 = x[ 0 ] - 0.5, r1 = x[ 1 ] - 0.5;
    const double t0 = r0;
    r0 = r1;
    r1 = t0;
    r1 = -r1;
    v[ 0 ] = r0 * ( 2 * M_PI / 10 );
    v[ 1 ] = r1 * ( 2 * M_PI / 10 );
    return 0;
  }
  case VO_PROB_LINEAR_WALL:
    v[ 0 ] = 0.02;   /* linearwall.hh:27-31 */
    return 0;
  case VO_PROB_LEVEQUE:
    if( dim == 2 )
    {
      const double c = vo_det_cospi( t / 8.0 );
      v[ 0 ] = ( sin_sq( x[ 0 ] ) * sin_two( x[ 1 ] ) ) * c;
      v[ 1 ] = ( ( -sin_two( x[ 0 ] ) ) * sin_sq( x[ 1 ] ) ) * c;
    }
    else
    {
      const double c = vo_det_cospi( t / 3.0 );
      v[ 0 ] = ( ( ( 2.0 * sin_sq( x[ 0 ] ) ) * sin_two( x[ 1 ] ) ) * sin_two( x[ 2 ] ) ) * c;
      v[ 1 ] = ( ( ( -sin_two( x[ 0 ] ) ) * sin_sq( x[ 1 ] ) ) * sin_two( x[ 2 ] ) ) * c;
      v[ 2 ] = ( ( ( -sin_two( x[ 0 ] ) ) * sin_two( x[ 1 ] ) ) * sin_sq( x[ 2 ] ) ) * c;
    }
    return 0;
  }
  return 1;
}

/* Velocity::operator() at the face centre, velocity.hh:15-20: geometry().global( (0.5,..) ) = o + 0.5*h along spanned axes */
static int face_center_velocity ( const vo_grid *g, int problem, double t, const int *ijk, int face, double *v )
{
  double x[ 3 ] = { 0, 0, 0 };
  cell_lower( g, ijk, x );
  if( face % 2 )
    x[ face / 2 ] += g->h[ face / 2 ];
  for( int k = 0; k < g->dim; ++k )
    if( k != face / 2 )
      x[ k ] += 0.5 * g->h[ k ];
  return velocity_field( g->dim, problem, x, t, v );
}

int vo_face_velocity ( const vo_grid *g, int problem, double time, int64_t cell, int face, double *v )
{
  int ijk[ 3 ];
  double vv[ 3 ];
  cell_ijk( g, cell, ijk );
  const int rc = face_center_velocity( g, problem, time, ijk, face, vv );
  for( int k = 0; k < g->dim; ++k )
    v[ k ] = vv[ k ];
  return rc;
}

/* ------------------------------------------------------------------------------------------------------------
 * E1  Evolution, evolution/evolution.hh:58-174 (scatter form, as in the reference)
 * ---------------------------------------------------------------------------------------------------------- */

int vo_evolve ( const vo_grid *g, const double *hs, const uint8_t *flags, int problem, double time, double dt, double *update, double *dt_est )
{
  const int dim = g->dim;
  const int64_t N = grid_cells( g );
  double dtEst = DBL_MAX;
  memset( update, 0, sizeof( double ) * (size_t) N );   /* update.clear(), :62 */

  const double volume = cell_volume( g );
  for( int64_t c = 0; c < N; ++c )
  {
    if( !is_mixed( flags[ c ] ) )
      continue;

    /* applyLocal, :90-145 */
    int ijk[ 3 ];
    double lo[ 3 ];
    cell_ijk( g, c, ijk );
    cell_lower( g, ijk, lo );
    double sumFluxes = 0.0;

    for( int face = 0; face < 2 * dim; ++face )
    {
      int nbijk[ 3 ] = { ijk[ 0 ], ijk[ 1 ], ijk[ 2 ] };
      nbijk[ face / 2 ] += ( face % 2 ) ? 1 : -1;
      if( !in_domain( g, nbijk ) )
        continue;
      const int64_t nb = cell_index( g, nbijk );

      double outerNormal[ 3 ] = { 0, 0, 0 };
      outerNormal[ face / 2 ] = ( face % 2 ) ? 1.0 : -1.0;

      double v[ 3 ];
      if( face_center_velocity( g, problem, time, ijk, face, v ) )
        return 1;

      const double faceVol = face_volume_of( g, face );
      sumFluxes += faceVol * fabs( dotn( outerNormal, v, dim ) );

      for( int k = 0; k < dim; ++k )
        v[ k ] *= dt;

      double flux = 0.0;
      /* integrationOuterNormal = centerUnitOuterNormal * |face| */
      double intNormal[ 3 ];
      for( int k = 0; k < 3; ++k )
        intNormal[ k ] = outerNormal[ k ] * faceVol;

      if( dotn( v, outerNormal, dim ) > 0 )
      {
        /* geometricFlux, :160-174: the cell is mixed here */
        const double *rec = hs + c * ( dim + 1 );
        flux = ( dim == 2 ) ? flux_2d( lo, g->h, face, v, rec ) : flux_3d( lo, g->h, face, v, rec );

        update[ c ] -= flux / volume;

        if( flags[ nb ] == 3 )
          update[ nb ] -= ( ( dotn( v, intNormal, dim ) ) - flux ) / volume;
        if( flags[ nb ] == 0 )
          update[ nb ] += flux / volume;
        if( is_mixed( flags[ nb ] ) )
          update[ nb ] += flux / volume;
      }

      if( flags[ nb ] == 3 )
        if( dotn( v, outerNormal, dim ) < 0 )
          update[ c ] -= ( dotn( v, intNormal, dim ) ) / volume;
    }

    const double est = volume / sumFluxes;
    dtEst = ( est < dtEst ) ? est : dtEst;   /* std::min( dtEst, est ) */
  }
  *dt_est = dtEst;
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * C1  CartesianHeightFunctionCurvature, curvature/cartesianheightfunctioncurvature.hh:54-187
 * ---------------------------------------------------------------------------------------------------------- */

int vo_curvature ( const vo_grid *g, const double *u, const double *hs, const uint8_t *flags, double *kappa, uint8_t *valid )
{
  const int dim = g->dim;
  const int64_t N = grid_cells( g );
  const int noc = ( dim == 2 ) ? 3 : 9;
  double *curv = (double *) malloc( sizeof( double ) * (size_t) N );
  uint8_t *ok = (uint8_t *) malloc( (size_t) N );
  if( !curv || !ok )
    return 1;

  for( int64_t c = 0; c < N; ++c )
  {
    curv[ c ] = 0.0;
    ok[ c ] = 0;
    if( !is_mixed( flags[ c ] ) )
      continue;

    /* applyLocal, :94-123 */
    int ijk[ 3 ], dir, sign;
    cell_ijk( g, c, ijk );
    hf_orientation( dim, hs + c * ( dim + 1 ), &dir, &sign );
    double heights[ 9 ];
    hf_heights( g, u, ijk, dir, sign, 1, heights );

    const double uMid = heights[ ( noc - 1 ) / 2 ];
    const int effTdown = hf_effective_tdown( g, ijk, dir, sign, 1 );
    if( uMid < effTdown || uMid > effTdown + 1 )
      continue;

    const double deltaX = g->h[ dim - dir - 1 ];   /* jacobianTransposed[dim-o-1][dim-o-1], :111-113 */

    int anyZero = 0;
    for( int i = 0; i < noc; ++i )
      if( heights[ i ] == 0.0 )
        anyZero = 1;
    if( anyZero )
      continue;

    ok[ c ] = 1;
    if( dim == 2 )
    {
      /* kappa<2>, :133-139 */
      const double Hx = ( heights[ 2 ] - heights[ 0 ] ) / 2.0;
      const double Hxx = ( heights[ 2 ] - 2 * heights[ 1 ] + heights[ 0 ] ) / deltaX;
      curv[ c ] = -Hxx / pow( 1.0 + Hx * Hx, 3.0 / 2.0 );
    }
    else
    {
      /* kappa<3>, :143-152 */
      const double Hx = ( heights[ 5 ] - heights[ 3 ] ) / 2.0;
      const double Hy = ( heights[ 7 ] - heights[ 1 ] ) / 2.0;
      const double Hxx = ( heights[ 5 ] - 2.0 * heights[ 4 ] + heights[ 3 ] ) / deltaX;
      const double Hyy = ( heights[ 7 ] - 2.0 * heights[ 4 ] + heights[ 1 ] ) / deltaX;
      const double Hxy = ( heights[ 8 ] - heights[ 2 ] - heights[ 6 ] + heights[ 0 ] ) / ( 4.0 * deltaX );
      curv[ c ] = -( Hxx + Hyy + Hxx * Hy * Hy + Hyy * Hx * Hx - 2.0 * Hxy * Hx * Hy ) / ( pow( 1.0 + Hx * Hx + Hy * Hy, 3.0 / 2.0 ) );
    }
  }

  /* averageCurvature with firstTurn == true, :71-83,156-187; newCurvature is value-initialised (0) elsewhere */
  for( int64_t c = 0; c < N; ++c )
  {
    kappa[ c ] = 0.0;
    if( !is_mixed( flags[ c ] ) )
      continue;
    if( ok[ c ] )
    {
      kappa[ c ] = curv[ c ];
      continue;
    }
    int n = 0;
    int ijk[ 3 ];
    cell_ijk( g, c, ijk );
    const int zlo = ( dim == 3 ) ? -1 : 0, zhi = ( dim == 3 ) ? 1 : 0;
    for( int dz = zlo; dz <= zhi; ++dz )
      for( int dy = -1; dy <= 1; ++dy )
        for( int dx = -1; dx <= 1; ++dx )
        {
          if( dx == 0 && dy == 0 && dz == 0 )
            continue;
          int nb[ 3 ] = { ijk[ 0 ] + dx, ijk[ 1 ] + dy, ijk[ 2 ] + dz };
          if( !in_domain( g, nb ) )
            continue;
          const int64_t cn = cell_index( g, nb );
          if( !is_mixed( flags[ cn ] ) )
            continue;
          if( ok[ cn ] )
          {
            kappa[ c ] += curv[ cn ];
            n++;
          }
        }
    if( n > 0 )
      kappa[ c ] /= (double) n;
  }
  if( valid )
    memcpy( valid, ok, (size_t) N );
  free( curv );
  free( ok );
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * initial data: interpolation/recursiveinterpolationcube.hh:33-82, interpolation/circleinterpolation.hh:23-153,
 * dispatch of test/average.hh:28-101
 * ---------------------------------------------------------------------------------------------------------- */

/* Problem::evaluate restated; returns (int)( u + 0.3 ), recursiveinterpolationcube.hh:36 */
static int indicator ( int dim, int problem, const double *x, double t )
{
  double u = 0.0;
  switch( problem )
  {
  case VO_PROB_ROTATING_CIRCLE:
  {
    /* rotatingcircle.hh:81-90 (3D; the 2D problem is initialised by circleInterpolation) */
    double center[ 3 ] = { 0.5, 0.5, 0.5 };
    const double a[ 3 ] = { 0.0, 0.25, 0.0 }, b[ 3 ] = { 0.25, 0.0, 0.0 }, z[ 3 ] = { 0.0, 0.0, 0.0 };
    const double cs = cos( ( 2 * M_PI / 10 ) * t ), sn = sin( ( 2 * M_PI / 10 ) * t ), msn = -sin( ( 2 * M_PI / 10 ) * t );
    double diff[ 3 ];
    for( int k = 0; k < 3; ++k )
    {
      center[ k ] += cs * a[ k ];
      center[ k ] += sn * b[ k ];
      center[ k ] += msn * z[ k ];
      diff[ k ] = x[ k ] - center[ k ];
    }
    u = ( normn( diff, dim ) < 0.15 ) ? 1.0 : 0.0;
    break;
  }
  case VO_PROB_SFLOW:
  {
    /* sflow.hh:26-33 / :63-71 */
    double y[ 3 ] = { 0.5, 0.5, 0.5 };
    for( int k = 0; k < dim; ++k )
      y[ k ] -= x[ k ];
    u = ( norm2n( y, dim ) <= 0.25 * 0.25 ) ? 1.0 : 0.0;
    break;
  }
  case VO_PROB_SLOTTED_CYLINDER:
  {
    /* slottedcylinder.hh:22-45 */
    const double center[ 2 ] = { 0.5, 0.5 };
    double xPrime[ 2 ] = { center[ 0 ], center[ 1 ] };
    double tmp[ 2 ] = { x[ 0 ] - center[ 0 ], x[ 1 ] - center[ 1 ] };
    const double cs = cos( ( 2.0 * M_PI / 10.0 ) * t ), sn = sin( ( 2.0 * M_PI / 10.0 ) * t );
    xPrime[ 0 ] += cs * tmp[ 0 ];
    xPrime[ 1 ] += cs * tmp[ 1 ];
    const double r0 = -tmp[ 1 ], r1 = tmp[ 0 ];
    xPrime[ 0 ] += sn * r0;
    xPrime[ 1 ] += sn * r1;
    u = 0.0;
    const double slotWidth = 0.075;
    double d[ 2 ] = { xPrime[ 0 ] - center[ 0 ], xPrime[ 1 ] - center[ 1 ] };
    if( normn( d, 2 ) < 0.4 )
      u = 1.0;
    const double cc[ 2 ] = { center[ 0 ] + 0.0, center[ 1 ] + slotWidth };
    d[ 0 ] = xPrime[ 0 ] - cc[ 0 ];
    d[ 1 ] = xPrime[ 1 ] - cc[ 1 ];
    if( normn( d, 2 ) < slotWidth )
      u = 0.0;
    if( xPrime[ 1 ] > center[ 1 ] + slotWidth && fabs( xPrime[ 0 ] - center[ 0 ] ) < slotWidth )
      u = 0.0;
    break;
  }
  case VO_PROB_LINEAR_WALL:
    u = ( x[ 0 ] <= 0.3 + t * 0.02 ) ? 1.0 : 0.0;   /* linearwall.hh:17-20 */
    break;
  case VO_PROB_LEVEQUE:
  {
    double y[ 3 ] = { 0.35, 0.35, 0.35 };
    if( dim == 2 )
    {
      y[ 0 ] = 0.5;
      y[ 1 ] = 0.75;
    }
    for( int k = 0; k < dim; ++k )
      y[ k ] -= x[ k ];
    u = ( normn( y, dim ) < 0.15 ) ? 1.0 : 0.0;
    break;
  }
  }
  return (int) ( u + 0.3 );
}

static double cube_volume ( int dim, double h )
{
  double volume = 1.0;
  for( int i = 0; i < dim; ++i )
    volume *= h;
  return volume;
}

/* recursiveEvaluation, recursiveinterpolationcube.hh:42-71; local coordinates mapped by geometry.global( x ) = o + x*h */
static double recursive_evaluation ( const vo_grid *g, int problem, double t, const double *cellLo, const double *origin, double h, int level )
{
  const int dim = g->dim;
  const int corners = 1 << dim;
  int sum = 0;
  for( int i = 0; i < corners; ++i )
  {
    double corner[ 3 ], x[ 3 ] = { 0, 0, 0 };
    for( int j = 0; j < dim; ++j )
    {
      corner[ j ] = origin[ j ];
      corner[ j ] += h * (double) ( ( i >> j ) & 1 );
      x[ j ] = cellLo[ j ];
      x[ j ] += corner[ j ] * g->h[ j ];
    }
    sum += indicator( dim, problem, x, t );
  }

  if( level == 0 )
    return cube_volume( dim, h ) * ( (double) sum / (double) corners );
  if( sum == corners )
    return cube_volume( dim, h );
  if( sum == 0 )
    return 0.0;

  h /= 2;
  double value = 0.0;
  for( int i = 0; i < corners; ++i )
  {
    double childOrigin[ 3 ];
    for( int j = 0; j < dim; ++j )
    {
      childOrigin[ j ] = origin[ j ];
      childOrigin[ j ] += h * (double) ( ( i >> j ) & 1 );
    }
    value += recursive_evaluation( g, problem, t, cellLo, childOrigin, h, level - 1 );
  }
  return value;
}

/* circleIntersection, circleinterpolation.hh:23-72 */
static int circle_intersection ( const double *l0, const double *l1, const double *c, double r, double pts[ 2 ][ 2 ] )
{
  int np = 0;
  const double dd[ 2 ] = { l1[ 0 ] - l0[ 0 ], l1[ 1 ] - l0[ 1 ] };
  const double n[ 2 ] = { -dd[ 1 ], dd[ 0 ] };
  const double p = dotn( n, l0, 2 );
  const double d = p - n[ 0 ] * c[ 0 ] - n[ 1 ] * c[ 1 ];
  const double diskr = r * r * norm2n( n, 2 ) - d * d;
  if( diskr < 0 )
    return 0;
  else if( fabs( diskr ) == 0.0 )
  {
    double v[ 2 ] = { n[ 0 ] * d, n[ 1 ] * d };
    const double nn = norm2n( n, 2 );
    v[ 0 ] /= nn; v[ 1 ] /= nn;
    v[ 0 ] += c[ 0 ]; v[ 1 ] += c[ 1 ];
    const double a[ 2 ] = { l0[ 0 ] - v[ 0 ], l0[ 1 ] - v[ 1 ] }, b[ 2 ] = { l1[ 0 ] - v[ 0 ], l1[ 1 ] - v[ 1 ] };
    if( dotn( a, b, 2 ) <= 0.0 )
    {
      pts[ 0 ][ 0 ] = v[ 0 ]; pts[ 0 ][ 1 ] = v[ 1 ];
      return 1;
    }
    return 0;
  }
  else
  {
    const double sq = sqrt( diskr );
    const double nn = norm2n( n, 2 );
    double v1[ 2 ] = { n[ 0 ] * d + n[ 1 ] * sq, n[ 1 ] * d - n[ 0 ] * sq };
    v1[ 0 ] /= nn; v1[ 1 ] /= nn;
    v1[ 0 ] += c[ 0 ]; v1[ 1 ] += c[ 1 ];
    {
      const double a[ 2 ] = { l0[ 0 ] - v1[ 0 ], l0[ 1 ] - v1[ 1 ] }, b[ 2 ] = { l1[ 0 ] - v1[ 0 ], l1[ 1 ] - v1[ 1 ] };
      if( dotn( a, b, 2 ) <= 0.0 )
      {
        pts[ np ][ 0 ] = v1[ 0 ]; pts[ np ][ 1 ] = v1[ 1 ];
        np++;
      }
    }
    double v2[ 2 ] = { n[ 0 ] * d - n[ 1 ] * sq, n[ 1 ] * d + n[ 0 ] * sq };
    v2[ 0 ] /= nn; v2[ 1 ] /= nn;
    v2[ 0 ] += c[ 0 ]; v2[ 1 ] += c[ 1 ];
    {
      const double a[ 2 ] = { l0[ 0 ] - v2[ 0 ], l0[ 1 ] - v2[ 1 ] }, b[ 2 ] = { l1[ 0 ] - v2[ 0 ], l1[ 1 ] - v2[ 1 ] };
      if( dotn( a, b, 2 ) <= 0.0 )
      {
        pts[ np ][ 0 ] = v2[ 0 ]; pts[ np ][ 1 ] = v2[ 1 ];
        np++;
      }
    }
    if( np == 2 )
    {
      const double e[ 2 ] = { pts[ 1 ][ 0 ] - pts[ 0 ][ 0 ], pts[ 1 ][ 1 ] - pts[ 0 ][ 1 ] };
      const double n1[ 2 ] = { -e[ 1 ], e[ 0 ] };
      if( dotn( n1, n, 2 ) < 0 )
      {
        double t0 = pts[ 0 ][ 0 ], t1 = pts[ 0 ][ 1 ];
        pts[ 0 ][ 0 ] = pts[ 1 ][ 0 ]; pts[ 0 ][ 1 ] = pts[ 1 ][ 1 ];
        pts[ 1 ][ 0 ] = t0; pts[ 1 ][ 1 ] = t1;
      }
    }
    return np;   /* the reference returns 2 here but callers use points.size() */
  }
}

/* intersectionVolume, circleinterpolation.hh:75-138 */
static double circle_polygon_volume ( const vo_polygon *polygon, const double *center, double radius )
{
  double sec[ 16 ][ 2 ];
  vo_polygon poly;
  int nSec = 0;
  poly.n = 0;

  int i = 0, i0 = 0;
  for( ; i < polygon->n; ++i )
  {
    const double d[ 2 ] = { polygon->v[ i ][ 0 ] - center[ 0 ], polygon->v[ i ][ 1 ] - center[ 1 ] };
    if( normn( d, 2 ) > radius )
    {
      i0 = i;
      break;
    }
  }
  if( i == polygon->n )
    return polygon_volume( polygon );

  for( int k = 0; k < polygon->n; ++k )
  {
    const int e = ( k + i0 ) % polygon->n;
    const double *v0 = polygon->v[ e ], *v1 = polygon->v[ ( e + 1 ) % polygon->n ];
    double pts[ 2 ][ 2 ];
    const int np = circle_intersection( v0, v1, center, radius, pts );
    for( int q = 0; q < np; ++q )
    {
      sec[ nSec ][ 0 ] = pts[ q ][ 0 ]; sec[ nSec ][ 1 ] = pts[ q ][ 1 ]; nSec++;
      poly.v[ poly.n ][ 0 ] = pts[ q ][ 0 ]; poly.v[ poly.n ][ 1 ] = pts[ q ][ 1 ]; poly.n++;
    }
    const double d[ 2 ] = { v1[ 0 ] - center[ 0 ], v1[ 1 ] - center[ 1 ] };
    if( normn( d, 2 ) < radius )
    {
      poly.v[ poly.n ][ 0 ] = v1[ 0 ]; poly.v[ poly.n ][ 1 ] = v1[ 1 ]; poly.n++;
    }
  }

  double volume = 0.0;
  if( nSec > 1 )
  {
    if( poly.n > 2 )
      volume += polygon_volume( &poly );
    for( int q = 1; q < nSec; q = q + 2 )
    {
      const double *a = sec[ q ], *b = sec[ ( q + 1 ) % nSec ];
      const double d[ 2 ] = { a[ 0 ] - b[ 0 ], a[ 1 ] - b[ 1 ] };
      const double alpha = 2.0 * asin( normn( d, 2 ) / ( 2.0 * radius ) );
      volume += radius * radius / 2.0 * ( alpha - sin( alpha ) );
    }
  }
  return volume;
}

/* ---- SwartzCurvature, curvature/swartzcurvature.hh:41-173 (2D only, as in the reference) ------------------------------ */

/* interface( entity, reconstructions ) of one cell: centroid and length of the segment (2d/polygon.hh:183-197) */
static void swartz_curv_interface ( const vo_grid *g, const double *hs, const int *ijk, double *centroid, double *length )
{
  double lo[ 3 ];
  vo_polygon p;
  double line[ 2 ][ 2 ];
  cell_lower( g, ijk, lo );
  make_polygon_cell( lo, g->h, &p );
  polygon_plane_cut( &p, hs + cell_index( g, ijk ) * 3, line );
  centroid[ 0 ] = ( line[ 0 ][ 0 ] + line[ 1 ][ 0 ] ) * 0.5;
  centroid[ 1 ] = ( line[ 0 ][ 1 ] + line[ 1 ][ 1 ] ) * 0.5;
  const double d[ 2 ] = { line[ 0 ][ 0 ] - line[ 1 ][ 0 ], line[ 0 ][ 1 ] - line[ 1 ][ 1 ] };
  *length = normn( d, 2 );
}

/* applyLocal, :75-152 */
static double swartz_curv_local ( const vo_grid *g, const double *hs, const uint8_t *flags, int64_t cell )
{
  int ijk[ 3 ];
  cell_ijk( g, cell, ijk );
  const double *normalEn = hs + cell * 3;
  double curv = 0.0, weights = 0.0;
  double centroidEn[ 2 ], muEn;
  swartz_curv_interface( g, hs, ijk, centroidEn, &muEn );
  muEn *= muEn / 12.0;

  for( int s1 = 0; s1 < 9; ++s1 )
  {
    if( s1 == 4 )
      continue;
    const int nb1[ 3 ] = { ijk[ 0 ] + ( s1 % 3 ) - 1, ijk[ 1 ] + ( s1 / 3 ) - 1, 0 };
    if( !in_domain( g, nb1 ) || !is_mixed( flags[ cell_index( g, nb1 ) ] ) )
      continue;
    double centroidNb1[ 2 ], muNb1;
    swartz_curv_interface( g, hs, nb1, centroidNb1, &muNb1 );
    double midwayNb1[ 2 ] = { centroidNb1[ 0 ] + centroidEn[ 0 ], centroidNb1[ 1 ] + centroidEn[ 1 ] };
    midwayNb1[ 0 ] *= 0.5;
    midwayNb1[ 1 ] *= 0.5;
    const double diff1[ 2 ] = { centroidNb1[ 0 ] - centroidEn[ 0 ], centroidNb1[ 1 ] - centroidEn[ 1 ] };
    const double triangleNb1 = normn( diff1, 2 );
    muNb1 *= muNb1 / 12.0;
    double tanNb1[ 2 ] = { centroidEn[ 0 ] - centroidNb1[ 0 ], centroidEn[ 1 ] - centroidNb1[ 1 ] };
    const double t1 = normn( tanNb1, 2 );
    tanNb1[ 0 ] /= t1;
    tanNb1[ 1 ] /= t1;
    const double nu12[ 2 ] = { -tanNb1[ 1 ], tanNb1[ 0 ] };
    if( dotn( nu12, normalEn, 2 ) < 0 )
      continue;

    for( int s2 = 0; s2 < 9; ++s2 )
    {
      if( s2 == 4 || s2 == s1 )
        continue;
      const int nb2[ 3 ] = { ijk[ 0 ] + ( s2 % 3 ) - 1, ijk[ 1 ] + ( s2 / 3 ) - 1, 0 };
      if( !in_domain( g, nb2 ) || !is_mixed( flags[ cell_index( g, nb2 ) ] ) )
        continue;
      double centroidNb2[ 2 ], muNb2;
      swartz_curv_interface( g, hs, nb2, centroidNb2, &muNb2 );
      double midwayNb2[ 2 ] = { centroidNb2[ 0 ] + centroidEn[ 0 ], centroidNb2[ 1 ] + centroidEn[ 1 ] };
      midwayNb2[ 0 ] *= 0.5;
      midwayNb2[ 1 ] *= 0.5;
      const double diff2[ 2 ] = { centroidNb2[ 0 ] - centroidEn[ 0 ], centroidNb2[ 1 ] - centroidEn[ 1 ] };
      const double triangleNb2 = normn( diff2, 2 );
      muNb2 *= muNb2 / 12.0;
      const double midwayDiff[ 2 ] = { midwayNb2[ 0 ] - midwayNb1[ 0 ], midwayNb2[ 1 ] - midwayNb1[ 1 ] };
      const double deltaS = normn( midwayDiff, 2 );
      double tanNb2[ 2 ] = { centroidNb2[ 0 ] - centroidEn[ 0 ], centroidNb2[ 1 ] - centroidEn[ 1 ] };
      const double t2 = normn( tanNb2, 2 );
      tanNb2[ 0 ] /= t2;
      tanNb2[ 1 ] /= t2;
      const double nu23[ 2 ] = { -tanNb2[ 1 ], tanNb2[ 0 ] };
      if( dotn( nu23, normalEn, 2 ) < 0 )
        continue;
      const double tanSum[ 2 ] = { tanNb1[ 0 ] + tanNb2[ 0 ], tanNb1[ 1 ] + tanNb2[ 1 ] };
      if( dotn( midwayDiff, tanSum, 2 ) < 0 )
        continue;
      const double M = ( ( muNb2 - muEn ) / triangleNb2 - ( muEn - muNb1 ) / triangleNb1 ) / ( 2.0 * deltaS );
      const double weight = 1.0;
      curv += weight * ( nu12[ 0 ] * nu23[ 1 ] - nu12[ 1 ] * nu23[ 0 ] ) / ( ( 1.0 + M ) * deltaS );
      weights += weight;
    }
  }
  if( weights > 0 )
    curv /= weights;
  return curv;
}

int vo_curvature_swartz ( const vo_grid *g, const double *hs, const uint8_t *flags, double *kappa )
{
  if( g->dim != 2 )
    return 1;
  const int64_t N = grid_cells( g );
  double *curv = (double *) malloc( sizeof( double ) * (size_t) N );
  for( int64_t c = 0; c < N; ++c )
    curv[ c ] = is_mixed( flags[ c ] ) ? swartz_curv_local( g, hs, flags, c ) : 0.0;
  /* cells whose curvature is exactly 0 take the mean over ALL mixed vertex neighbours, :57-66,154-170 */
  for( int64_t c = 0; c < N; ++c )
  {
    kappa[ c ] = curv[ c ];
    if( !is_mixed( flags[ c ] ) || curv[ c ] != 0.0 )
      continue;
    int ijk[ 3 ], n = 0;
    cell_ijk( g, c, ijk );
    for( int s = 0; s < 9; ++s )
    {
      if( s == 4 )
        continue;
      const int nb[ 3 ] = { ijk[ 0 ] + ( s % 3 ) - 1, ijk[ 1 ] + ( s / 3 ) - 1, 0 };
      if( !in_domain( g, nb ) || !is_mixed( flags[ cell_index( g, nb ) ] ) )
        continue;
      kappa[ c ] += curv[ cell_index( g, nb ) ];
      n++;
    }
    if( n > 0 )
      kappa[ c ] /= (double) n;
  }
  free( curv );
  return 0;
}

/* circleInterpolation( center, radius, uh ), interpolation/circleinterpolation.hh:140-153 (2D) */
int vo_circle_interpolation ( const vo_grid *g, const double *center, double radius, double *u )
{
  if( g->dim != 2 )
    return 1;
  const int64_t N = grid_cells( g );
  for( int64_t c = 0; c < N; ++c )
  {
    int ijk[ 3 ];
    double lo[ 3 ];
    vo_polygon p;
    cell_ijk( g, c, ijk );
    cell_lower( g, ijk, lo );
    make_polygon_cell( lo, g->h, &p );
    u[ c ] = circle_polygon_volume( &p, center, radius ) / cell_volume( g );
  }
  return 0;
}

/* l1error( gridView, reconstructionSet, flags, RotatingCircle< double, 2 > circle, time ), test/errors.hh:62-95, with
 * center = circle.center( time ), radius = circle.radius( time ).  hs: zero half-space = no reconstruction (the reference's
 * reconstruction set is cleared every pass, so only mixed cells carry one: entries of other cells are ignored here).
 * errors.hh itself cannot be compiled here (it includes dune-geometry's quadrature rules): the loop is restated, the
 * primitives it calls (intersect, intersectionVolume, Polygon::volume) are the pinned ones above. */
int vo_l1error_circle ( const vo_grid *g, const double *hs, const uint8_t *flags, const double *center, double radius, double *error )
{
  if( g->dim != 2 )
    return 1;
  const int64_t N = grid_cells( g );
  double l1 = 0.0;
  for( int64_t c = 0; c < N; ++c )
  {
    int ijk[ 3 ];
    double lo[ 3 ];
    vo_polygon p;
    cell_ijk( g, c, ijk );
    cell_lower( g, ijk, lo );
    make_polygon_cell( lo, g->h, &p );
    const double volume = cell_volume( g );
    const double exact = circle_polygon_volume( &p, center, radius ) / cell_volume( g );
    const double T1 = exact * volume;
    const double T0 = volume - T1;
    const int mixed = ( flags[ c ] == 1 || flags[ c ] == 2 );
    const double *h = hs + 3 * c;
    if( !mixed || !( h[ 0 ] * h[ 0 ] + h[ 1 ] * h[ 1 ] > 0.5 ) )
    {
      const double cc = ( flags[ c ] == 3 ) ? 1.0 : 0.0;
      l1 += T0 * fabs( cc ) + T1 * fabs( 1.0 - cc );
    }
    else
    {
      vo_polygon one;
      polygon_clip( &p, h, &one );
      const double shared = circle_polygon_volume( &one, center, radius );
      l1 += ( T1 - shared ) + ( polygon_volume( &one ) - shared );
    }
  }
  *error = l1;
  return 0;
}

int vo_init ( const vo_grid *g, int problem, double time, double *u )
{
  const int dim = g->dim;
  const int64_t N = grid_cells( g );
  if( problem == VO_PROB_ROTATING_CIRCLE && dim == 2 )
  {
    /* average.hh:82-101: circleInterpolation( problem.center( t ), problem.radius( t ), uh ); rotatingcircle.hh:28-43 */
    double center[ 2 ] = { 0.5, 0.5 };
    const double offset[ 2 ] = { 0, 0.25 }, normal[ 2 ] = { 0.25, 0 };
    const double cs = cos( ( 2 * M_PI / 10 ) * time ), sn = sin( ( 2 * M_PI / 10 ) * time );
    for( int k = 0; k < 2; ++k )
    {
      center[ k ] += cs * offset[ k ];
      center[ k ] += sn * normal[ k ];
    }
    for( int64_t c = 0; c < N; ++c )
    {
      int ijk[ 3 ];
      double lo[ 3 ];
      vo_polygon p;
      cell_ijk( g, c, ijk );
      cell_lower( g, ijk, lo );
      make_polygon_cell( lo, g->h, &p );
      u[ c ] = circle_polygon_volume( &p, center, 0.15 ) / cell_volume( g );
    }
    return 0;
  }
  if( problem == VO_PROB_SLOTTED_CYLINDER && dim != 2 )
    return 1;
  if( problem < 0 || problem > VO_PROB_LEVEQUE )
    return 1;
  for( int64_t c = 0; c < N; ++c )
  {
    int ijk[ 3 ];
    double lo[ 3 ];
    const double origin[ 3 ] = { 0, 0, 0 };
    cell_ijk( g, c, ijk );
    cell_lower( g, ijk, lo );
    u[ c ] = recursive_evaluation( g, problem, time, lo, origin, 1.0, 3 );
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * T1  time loop, algorithm.hh:68-95 (without l1error and output)
 * ---------------------------------------------------------------------------------------------------------- */

static double now_seconds ( void )
{
  struct timespec ts;
  clock_gettime( CLOCK_MONOTONIC, &ts );
  return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

int vo_run ( const vo_grid *g, int kind, int problem, double eps, double cfl, int max_iterations, int passes,
             double *u, double *time_io, double *dt_io, double *phase_seconds, int64_t *mixed_cell_steps )
{
  const int64_t N = grid_cells( g );
  uint8_t *flags = (uint8_t *) malloc( (size_t) N );
  double *hs = (double *) malloc( sizeof( double ) * (size_t) N * ( g->dim + 1 ) );
  double *update = (double *) malloc( sizeof( double ) * (size_t) N );
  if( !flags || !hs || !update )
    return 1;

  double time = *time_io, dt = *dt_io, dtEst = 0.0;
  int rc = 0;
  for( int pass = 0; pass < passes && !rc; ++pass )
  {
    const double t0 = now_seconds();
    vo_flag( g, u, eps, flags );
    const double t1 = now_seconds();
    rc = vo_reconstruct( g, kind, u, flags, hs, max_iterations );
    const double t2 = now_seconds();
    rc = rc || vo_evolve( g, hs, flags, problem, time, dt, update, &dtEst );
    const double t3 = now_seconds();
    for( int64_t c = 0; c < N; ++c )
      u[ c ] += 1.0 * update[ c ];   /* colorfunction.hh:31-37 */
    const double t4 = now_seconds();

    if( phase_seconds )
    {
      phase_seconds[ 0 ] += t1 - t0;
      phase_seconds[ 1 ] += t2 - t1;
      phase_seconds[ 2 ] += t3 - t2;
      phase_seconds[ 3 ] += t4 - t3;
    }
    if( mixed_cell_steps )
      for( int64_t c = 0; c < N; ++c )
        if( is_mixed( flags[ c ] ) )
          ++*mixed_cell_steps;

    if( dt > 0.0 )
      time += dt;
    dt = dtEst * cfl;
  }
  free( flags );
  free( hs );
  free( update );
  *time_io = time;
  *dt_io = dt;
  return rc;
}

/* ------------------------------------------------------------------------------------------------------------
 * single-cell primitives for fuzzing
 * ---------------------------------------------------------------------------------------------------------- */

int vo_locate ( int dim, const double *lo, const double *h, const double *normal, double fraction, double *out )
{
  if( dim == 2 )
  {
    vo_polygon p;
    make_polygon_cell( lo, h, &p );
    locate_2d( &p, normal, fraction, out );
    return 0;
  }
  vo_polyhedron p;
  make_polyhedron_cell( lo, h, &p );
  return locate_3d( &p, normal, fraction, out );
}

int vo_volume_fraction ( int dim, const double *lo, const double *h, const double *hs, double *out )
{
  if( dim == 2 )
  {
    vo_polygon p;
    make_polygon_cell( lo, h, &p );
    *out = volume_fraction_2d( &p, hs );
    return 0;
  }
  vo_polyhedron p;
  make_polyhedron_cell( lo, h, &p );
  *out = volume_fraction_3d( &p, hs );
  return 0;
}

int vo_flux ( int dim, const double *lo, const double *h, int face, const double *v, const double *hs, double *out )
{
  *out = ( dim == 2 ) ? flux_2d( lo, h, face, v, hs ) : flux_3d( lo, h, face, v, hs );
  return 0;
}

int vo_interface ( int dim, const double *lo, const double *h, const double *hs, int *nverts, double *verts, double *centroid, double *measure )
{
  if( dim == 2 )
  {
    vo_polygon p;
    double line[ 2 ][ 2 ];
    make_polygon_cell( lo, h, &p );
    polygon_plane_cut( &p, hs, line );
    *nverts = 2;
    for( int i = 0; i < 2; ++i )
      for( int d = 0; d < 2; ++d )
        verts[ 2 * i + d ] = line[ i ][ d ];
    centroid[ 0 ] = ( line[ 0 ][ 0 ] + line[ 1 ][ 0 ] ) * 0.5;
    centroid[ 1 ] = ( line[ 0 ][ 1 ] + line[ 1 ][ 1 ] ) * 0.5;
    const double d[ 2 ] = { line[ 0 ][ 0 ] - line[ 1 ][ 0 ], line[ 0 ][ 1 ] - line[ 1 ][ 1 ] };
    *measure = normn( d, 2 );
    return 0;
  }
  vo_polyhedron p;
  vo_face f;
  make_polyhedron_cell( lo, h, &p );
  polyhedron_plane_cut( &p, hs, &f );
  *nverts = f.n;
  for( int i = 0; i < f.n && i < 8; ++i )
    for( int d = 0; d < 3; ++d )
      verts[ 3 * i + d ] = f.v[ i ][ d ];
  if( f.n > 2 )
  {
    iface_centroid( &f, centroid );
    *measure = iface_volume( &f );
  }
  else
  {
    centroid[ 0 ] = centroid[ 1 ] = centroid[ 2 ] = 0.0;
    *measure = 0.0;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * Unstructured 2D meshes: the same operators on flattened arrays (vo_mesh).  F1 flagging.hh:35-70,
 * R1 reconstruction/modifiedyoungs.hh:63-134 with the CSR vertex stencil of stencil/vertexneighborsstencil.hh:52-94,
 * E1 evolution/evolution.hh:58-145, G7 2d/upwindpolygon.hh:24-30.  Pinned against oracle/_ref/libvofrefmesh.so.
 * ---------------------------------------------------------------------------------------------------------- */

/* makePolygon( geometry ), 2d/polygon.hh:230-241: triangle corners 0,1,2; quadrilateral corners 0,1,3,2 */
static int mesh_cell_polygon ( const vo_mesh *m, int64_t cell, vo_polygon *p )
{
  const int b = m->corner_offsets[ cell ], n = m->corner_offsets[ cell + 1 ] - b;
  static const int tri[ 3 ] = { 0, 1, 2 }, quad[ 4 ] = { 0, 1, 3, 2 };
  if( n != 3 && n != 4 )
    return 1;                                 /* DUNE_THROW( InvalidStateException, "Invalid GeometryType." ) */
  const int *order = ( n == 3 ) ? tri : quad;
  p->n = n;
  for( int k = 0; k < n; ++k )
  {
    p->v[ k ][ 0 ] = m->corners[ 2 * ( b + order[ k ] ) ];
    p->v[ k ][ 1 ] = m->corners[ 2 * ( b + order[ k ] ) + 1 ];
  }
  return 0;
}

int vo_mesh_flag ( const vo_mesh *m, const double *u, double eps, uint8_t *flags )
{
  for( int64_t c = 0; c < m->n_cells; ++c )
  {
    const double colorEn = u[ c ];
    uint8_t flag;
    if( colorEn < eps )
      flag = 0;
    else if( colorEn <= ( 1 - eps ) )
      flag = 1;
    else if( isnan( colorEn ) )
      flag = 99;
    else
    {
      flag = 3;
      for( int s = m->face_offsets[ c ]; s < m->face_offsets[ c + 1 ]; ++s )
        if( m->face_neighbor[ s ] >= 0 && u[ m->face_neighbor[ s ] ] < eps )
        {
          flag = 2;
          break;
        }
    }
    flags[ c ] = flag;
  }
  return 0;
}

static int mesh_cell_polyhedron ( const vo_mesh *m, int64_t cell, vo_polyhedron *p );
static int mesh_flux3 ( const double *c, int nc, const double *v, const double *hs, double *flux );
static int mesh_dim ( const vo_mesh *m );

int vo_mesh_locate ( const vo_mesh *m, int64_t cell, const double *normal, double fraction, double *out )
{
  if( mesh_dim( m ) == 3 )
  {
    vo_polyhedron ph;
    if( mesh_cell_polyhedron( m, cell, &ph ) )
      return 1;
    return locate_3d( &ph, normal, fraction, out );
  }
  vo_polygon p;
  if( mesh_cell_polygon( m, cell, &p ) )
    return 1;
  locate_2d( &p, normal, fraction, out );
  return 0;
}

/* applyLocal, modifiedyoungs.hh:90-134 */
static int mesh_youngs_local ( const vo_mesh *m, const double *u, int64_t cell, double *hs )
{
  const int D = mesh_dim( m );
  const double *center = m->centers + D * cell;
  const double colorEn = u[ cell ];
  double AtA[ 3 ][ 3 ] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
  double Atb[ 3 ] = { 0, 0, 0 };

  for( int s = m->stencil_offsets[ cell ]; s < m->stencil_offsets[ cell + 1 ]; ++s )
  {
    const int64_t nb = m->stencil[ s ];
    const double colorNb = clampd( u[ nb ], 0.0, 1.0 );
    double d[ 3 ] = { 0, 0, 0 };
    for( int k = 0; k < D; ++k )
      d[ k ] = m->centers[ D * nb + k ] - center[ k ];
    ls_accumulate( D, d, colorNb - colorEn, AtA, Atb );
  }
  for( int s = m->face_offsets[ cell ]; s < m->face_offsets[ cell + 1 ]; ++s )
  {
    if( m->face_neighbor[ s ] >= 0 )
      continue;
    double d[ 3 ] = { 0, 0, 0 };
    for( int k = 0; k < D; ++k )
    {
      d[ k ] = m->face_centers[ D * s + k ] - center[ k ];
      d[ k ] *= 2.0;
    }
    ls_accumulate( D, d, 0.5 - colorEn, AtA, Atb );
  }

  double normal[ 3 ] = { 0, 0, 0 };
  solve_small( D, AtA, Atb, normal );
  if( norm2n( normal, D ) < VO_EPS )
    return 0;
  const double nrm = normn( normal, D );
  for( int k = 0; k < D; ++k )
    normal[ k ] /= nrm;
  return vo_mesh_locate( m, cell, normal, colorEn, hs );
}

int vo_mesh_youngs ( const vo_mesh *m, const double *u, const uint8_t *flags, double *hs )
{
  const int D = mesh_dim( m );
  memset( hs, 0, sizeof( double ) * ( D + 1 ) * (size_t) m->n_cells );   /* reconstructions.clear(), :65 */
  for( int64_t c = 0; c < m->n_cells; ++c )
    if( is_mixed( flags[ c ] ) && mesh_youngs_local( m, u, c, hs + ( D + 1 ) * c ) )
      return 1;
  return 0;
}

/* truncVolume( upwindPolygon2d( iGeometry, v ), hs ), 2d/upwindpolygon.hh:24-30 + geometry/upwindpolygon.hh:58-62 */
static double mesh_flux ( const double *c0, const double *c1, const double *v, const double *hs )
{
  const double d[ 2 ] = { c1[ 0 ] - c0[ 0 ], c1[ 1 ] - c0[ 1 ] };
  const double cr[ 2 ] = { -d[ 1 ], d[ 0 ] };   /* generalizedCrossProduct 2D, geometry/utility.hh:82-85 */
  vo_polygon up, cl;
  up.n = 4;
  const double *a = c0, *b = c1;
  if( !( dotn( cr, v, 2 ) < 0.0 ) )
  {
    a = c1;
    b = c0;
  }
  up.v[ 0 ][ 0 ] = a[ 0 ]; up.v[ 0 ][ 1 ] = a[ 1 ];
  up.v[ 1 ][ 0 ] = b[ 0 ]; up.v[ 1 ][ 1 ] = b[ 1 ];
  up.v[ 2 ][ 0 ] = b[ 0 ] - v[ 0 ]; up.v[ 2 ][ 1 ] = b[ 1 ] - v[ 1 ];
  up.v[ 3 ][ 0 ] = a[ 0 ] - v[ 0 ]; up.v[ 3 ][ 1 ] = a[ 1 ] - v[ 1 ];
  polygon_clip( &up, hs, &cl );
  return polygon_volume( &cl );
}

int vo_mesh_evolve ( const vo_mesh *m, const double *hs, const uint8_t *flags, const double *face_velocity, double dt, double *update, double *dt_est )
{
  const int D = mesh_dim( m );
  double dtEst = DBL_MAX;
  memset( update, 0, sizeof( double ) * (size_t) m->n_cells );   /* update.clear(), :62 */
  for( int64_t c = 0; c < m->n_cells; ++c )
  {
    if( !is_mixed( flags[ c ] ) )
      continue;
    double sumFluxes = 0.0;
    const double volume = m->volumes[ c ];
    for( int s = m->face_offsets[ c ]; s < m->face_offsets[ c + 1 ]; ++s )
    {
      const int64_t nb = m->face_neighbor[ s ];
      if( nb < 0 )
        continue;
      const double *outerNormal = m->face_unit_normals + D * s;
      const double *intNormal = m->face_int_normals + D * s;
      double v[ 3 ] = { 0, 0, 0 };
      for( int k = 0; k < D; ++k )
        v[ k ] = face_velocity[ D * s + k ];
      sumFluxes += m->face_volumes[ s ] * fabs( dotn( outerNormal, v, D ) );
      for( int k = 0; k < D; ++k )
        v[ k ] *= dt;
      double flux = 0.0;
      if( dotn( v, outerNormal, D ) > 0 )
      {
        if( D == 2 )
          flux = mesh_flux( m->face_corners + 4 * s, m->face_corners + 4 * s + 2, v, hs + 3 * c );
        else if( mesh_flux3( m->face_corners + 3 * (size_t) m->face_corner_offsets[ s ], m->face_corner_offsets[ s + 1 ] - m->face_corner_offsets[ s ],
                             v, hs + 4 * c, &flux ) )
          return 1;
        update[ c ] -= flux / volume;
        if( flags[ nb ] == 3 )
          update[ nb ] -= ( ( dotn( v, intNormal, D ) ) - flux ) / m->volumes[ nb ];
        if( flags[ nb ] == 0 )
          update[ nb ] += flux / m->volumes[ nb ];
        if( is_mixed( flags[ nb ] ) )
          update[ nb ] += flux / m->volumes[ nb ];
      }
      if( flags[ nb ] == 3 )
        if( dotn( v, outerNormal, D ) < 0 )
          update[ c ] -= ( dotn( v, intNormal, D ) ) / volume;
    }
    const double est = volume / sumFluxes;
    dtEst = ( est < dtEst ) ? est : dtEst;
  }
  *dt_est = dtEst;
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * Unstructured 3D meshes (tetrahedra and hexahedra): the same operators with makePolyhedron( geometry ),
 * 3d/polyhedron.hh:366-404 (simplex and cube tables), locateHalfSpace( Polyhedron ), geometry/algorithm.hh:119-236, and
 * upwindPolygon3d for triangular and quadrilateral intersections, 3d/upwindpolygon.hh:24-74.
 * Pinned against oracle/_ref/libvofrefmesh3.so.
 * ---------------------------------------------------------------------------------------------------------- */

/* makePolyhedron( simplex ), 3d/polyhedron.hh:374-385 */
static const int tet_edges[ 12 ][ 2 ] = {
  { 0, 1 }, { 1, 0 }, { 0, 2 }, { 2, 0 }, { 1, 2 }, { 2, 1 }, { 0, 3 }, { 3, 0 }, { 1, 3 }, { 3, 1 }, { 3, 2 }, { 2, 3 }
};
static const int tet_faces[ 4 ][ 3 ] = { { 2, 5, 1 }, { 0, 8, 7 }, { 6, 10, 3 }, { 4, 11, 9 } };

/* upwindPolygon3d( triangle ), 3d/upwindpolygon.hh:47-57: a prism of 6 nodes, 18 directed edges, 3 + 2 faces */
static const int prism_edges[ 18 ][ 2 ] = {
  { 0, 3 }, { 1, 4 }, { 2, 5 }, { 0, 1 }, { 0, 2 }, { 1, 2 }, { 3, 4 }, { 3, 5 }, { 4, 5 },
  { 3, 0 }, { 4, 1 }, { 5, 2 }, { 1, 0 }, { 2, 0 }, { 2, 1 }, { 4, 3 }, { 5, 3 }, { 5, 4 }
};
static const int prism_face_len[ 5 ] = { 4, 4, 4, 3, 3 };
static const int prism_faces[ 5 ][ 4 ] = { { 9, 3, 1, 15 }, { 0, 7, 11, 13 }, { 5, 2, 17, 10 }, { 4, 14, 12, 0 }, { 6, 8, 16, 0 } };

static void make_tet_topology ( vo_polyhedron *p )
{
  p->nn = 4;
  p->ne = 12;
  p->nf = 4;
  for( int i = 0; i < 12; ++i )
  {
    p->e[ i ][ 0 ] = tet_edges[ i ][ 0 ];
    p->e[ i ][ 1 ] = tet_edges[ i ][ 1 ];
  }
  for( int f = 0; f < 4; ++f )
  {
    p->fne[ f ] = 3;
    for( int i = 0; i < 3; ++i )
      p->fe[ f ][ i ] = tet_faces[ f ][ i ];
  }
}

static void make_prism_topology ( vo_polyhedron *p )
{
  p->nn = 6;
  p->ne = 18;
  p->nf = 5;
  for( int i = 0; i < 18; ++i )
  {
    p->e[ i ][ 0 ] = prism_edges[ i ][ 0 ];
    p->e[ i ][ 1 ] = prism_edges[ i ][ 1 ];
  }
  for( int f = 0; f < 5; ++f )
  {
    p->fne[ f ] = prism_face_len[ f ];
    for( int i = 0; i < prism_face_len[ f ]; ++i )
      p->fe[ f ][ i ] = prism_faces[ f ][ i ];
  }
}

static int mesh_cell_polyhedron ( const vo_mesh *m, int64_t cell, vo_polyhedron *p )
{
  const int b = m->corner_offsets[ cell ], n = m->corner_offsets[ cell + 1 ] - b;
  if( n == 4 )
    make_tet_topology( p );
  else if( n == 8 )
    make_cube_topology( p );
  else
    return 1;                                 /* DUNE_THROW( InvalidStateException, "Invalid GeometryType." ) */
  for( int i = 0; i < n; ++i )
    for( int k = 0; k < 3; ++k )
      p->x[ i ][ k ] = m->corners[ 3 * ( b + i ) + k ];
  return 0;
}

/* truncVolume( upwindPolygon3d( iGeometry, v ), hs ) for an intersection with nc = 3 or 4 corners c */
static int mesh_flux3 ( const double *c, int nc, const double *v, const double *hs, double *flux )
{
  double a[ 3 ], b[ 3 ], cr[ 3 ];
  for( int k = 0; k < 3; ++k )
  {
    a[ k ] = c[ 3 + k ] - c[ k ];
    b[ k ] = c[ 6 + k ] - c[ k ];
  }
  cross3( a, b, cr );
  vo_polyhedron up, cl;
  if( nc == 3 )
    make_prism_topology( &up );
  else if( nc == 4 )
    make_cube_topology( &up );
  else
    return 1;
  const int shiftedFirst = dotn( cr, v, 3 ) < 0.0;
  for( int i = 0; i < nc; ++i )
    for( int k = 0; k < 3; ++k )
    {
      const double moved = c[ 3 * i + k ] - v[ k ];
      up.x[ shiftedFirst ? i : i + nc ][ k ] = moved;
      up.x[ shiftedFirst ? i + nc : i ][ k ] = c[ 3 * i + k ];
    }
  polyhedron_clip( &up, hs, &cl );
  *flux = polyhedron_volume( &cl );
  return 0;
}

static int mesh_dim ( const vo_mesh *m ) { return m->dim == 3 ? 3 : 2; }

/* the two primitives above on explicit corner lists (entry points for the tests of the device functions) */
int vo_locate_corners ( int nc, const double *corners, const double *normal, double fraction, double *out )
{
  vo_polyhedron p;
  if( nc == 4 )
    make_tet_topology( &p );
  else if( nc == 8 )
    make_cube_topology( &p );
  else
    return 1;
  for( int i = 0; i < nc; ++i )
    for( int k = 0; k < 3; ++k )
      p.x[ i ][ k ] = corners[ 3 * i + k ];
  return locate_3d( &p, normal, fraction, out );
}

int vo_flux_corners ( int nc, const double *corners, const double *v, const double *hs, double *out )
{
  return mesh_flux3( corners, nc, v, hs, out );
}

/* R3 on a mesh: ModifiedSwartzReconstruction::applyLocal for 2D, reconstruction/modifiedswartz.hh:102-153 */
static int mesh_interface_centroid ( const vo_mesh *m, int64_t cell, const double *normal, double color, double *centroid )
{
  vo_polygon p;
  double hs[ 3 ], line[ 2 ][ 2 ];
  if( mesh_cell_polygon( m, cell, &p ) )
    return 1;
  locate_2d( &p, normal, color, hs );
  polygon_plane_cut( &p, hs, line );
  centroid[ 0 ] = ( line[ 0 ][ 0 ] + line[ 1 ][ 0 ] ) * 0.5;   /* Line::centroid, 2d/polygon.hh:183-188 */
  centroid[ 1 ] = ( line[ 0 ][ 1 ] + line[ 1 ][ 1 ] ) * 0.5;
  return 0;
}

static int mesh_swartz_local ( const vo_mesh *m, const double *u, const uint8_t *flags, int64_t cell, double *hsAll, int maxIterations )
{
  double *reconstruction = hsAll + 3 * cell;
  double normal[ 2 ] = { reconstruction[ 0 ], reconstruction[ 1 ] };
  int iterations = 0;
  double residuum = 0.0;
  do
  {
    const double oldNormal[ 2 ] = { normal[ 0 ], normal[ 1 ] };
    normal[ 0 ] = normal[ 1 ] = 0.0;
    double cEn[ 2 ];
    if( mesh_interface_centroid( m, cell, oldNormal, u[ cell ], cEn ) )
      return 1;
    for( int s = m->stencil_offsets[ cell ]; s < m->stencil_offsets[ cell + 1 ]; ++s )
    {
      const int64_t nb = m->stencil[ s ];
      if( !is_mixed( flags[ nb ] ) )
        continue;
      double cNb[ 2 ];
      if( mesh_interface_centroid( m, nb, oldNormal, u[ nb ], cNb ) )
        return 1;
      const double dd[ 2 ] = { cNb[ 0 ] - cEn[ 0 ], cNb[ 1 ] - cEn[ 1 ] };
      double centerNormal[ 2 ] = { -dd[ 1 ], dd[ 0 ] };
      if( dotn( centerNormal, oldNormal, 2 ) < 0 )
      {
        centerNormal[ 0 ] *= -1.0;
        centerNormal[ 1 ] *= -1.0;
      }
      const double nrm = normn( centerNormal, 2 );
      centerNormal[ 0 ] /= nrm;
      centerNormal[ 1 ] /= nrm;
      const double weight = 1.0;
      normal[ 0 ] += weight * centerNormal[ 0 ];
      normal[ 1 ] += weight * centerNormal[ 1 ];
    }
    if( normn( normal, 2 ) < 0.5 )
      return 0;
    const double nrm = normn( normal, 2 );
    normal[ 0 ] /= nrm;
    normal[ 1 ] /= nrm;
    if( vo_mesh_locate( m, cell, normal, u[ cell ], reconstruction ) )
      return 1;
    residuum = 1.0 - dotn( normal, oldNormal, 2 );
    ++iterations;
  }
  while( residuum > 1e-12 && iterations < maxIterations );
  return 0;
}

int vo_mesh_swartz ( const vo_mesh *m, const double *u, const uint8_t *flags, double *hs, int max_iterations )
{
  if( vo_mesh_youngs( m, u, flags, hs ) )      /* initializer()(..), modifiedswartz.hh:72 */
    return 1;
  for( int64_t c = 0; c < m->n_cells; ++c )
    if( is_mixed( flags[ c ] ) && mesh_swartz_local( m, u, flags, c, hs, max_iterations ) )
      return 1;
  return 0;
}

