// vofx 2D geometry device functions: polygon clip, plane-constant solve, swept-face flux, interface segment.
// Restates (operation for operation, see vofx_math.cuh) dune/vof/geometry/2d/{polygon,intersect,upwindpolygon}.hh
// and the Polygon overload of locateHalfSpace (geometry/algorithm.hh:68-111) for axis-aligned cells, with
// fixed-size register arrays instead of std::vector.
#pragma once

#include "vofx_math.cuh"

namespace vofx
{

  // a convex polygon with at most 8 vertices (a clipped quadrilateral has at most 5)
  struct Poly2
  {
    int n;
    double x[ 8 ], y[ 8 ];
  };

  // half-space in 2D: (nx, ny, d), fluid side nx*x + ny*y + d > 0; valid iff |n|^2 > 0.5 (geometry/halfspace.hh:104-107)
  struct Hs2
  {
    double nx, ny, d;
  };

  VOFX_INLINE bool hs_valid ( const Hs2 &h ) { return dot2( h.nx, h.ny, h.nx, h.ny ) > 0.5; }
  VOFX_INLINE double level_set ( const Hs2 &h, double x, double y ) { return dot2( h.nx, h.ny, x, y ) + h.d; }

  // signed area, 2d/polygon.hh:99-106
  VOFX_HD inline double poly_volume ( const Poly2 &p )
  {
    double sum = 0.0;
    const int n = p.n;
    for( int i = 0; i < n; ++i )
    {
      const int j = ( i + 1 == n ) ? 0 : i + 1;
      sum += ( p.y[ i ] + p.y[ j ] ) * ( p.x[ i ] - p.x[ j ] );
    }
    return sum / 2.0;
  }

  // Sutherland-Hodgman against one half-space, 2d/intersect.hh:67-99: inside iff levelSet > 0 (strict),
  // cut point = (-l1/(l0-l1))*v0 + (l0/(l0-l1))*v1 accumulated from zero
  VOFX_HD inline void poly_clip ( const Poly2 &p, const Hs2 &hs, Poly2 &out )
  {
    out.n = 0;
    if( !hs_valid( hs ) )
      return;
    const int n = p.n;
    // the level set of a vertex enters two edges: evaluated once (the reference evaluates it per edge, same value)
    double first = 0.0, l0 = 0.0;
    for( int i = 0; i < n; ++i )
    {
      const int j = ( i + 1 == n ) ? 0 : i + 1;
      if( i == 0 )
        first = l0 = level_set( hs, p.x[ 0 ], p.y[ 0 ] );
      const double l1 = ( j == 0 ) ? first : level_set( hs, p.x[ j ], p.y[ j ] );
      if( l0 > 0.0 )
      {
        out.x[ out.n ] = p.x[ i ];
        out.y[ out.n ] = p.y[ i ];
        out.n++;
      }
      if( ( l0 > 0.0 ) != ( l1 > 0.0 ) )
      {
        double a = -l1, b = l0;
        vdiv2( a, b, l0 - l1 );          // -l1 / ( l0 - l1 ), l0 / ( l0 - l1 ): one reciprocal, the IEEE quotients
        double cx = 0.0, cy = 0.0;
        cx += a * p.x[ i ]; cx += b * p.x[ j ];
        cy += a * p.y[ i ]; cy += b * p.y[ j ];
        out.x[ out.n ] = cx;
        out.y[ out.n ] = cy;
        out.n++;
      }
      l0 = l1;
    }
  }

  // getVolumeFraction, geometry/algorithm.hh:31-37
  VOFX_HD inline double volume_fraction ( const Poly2 &p, const Hs2 &hs )
  {
    Poly2 c;
    poly_clip( p, hs, c );
    return vdiv( poly_volume( c ), poly_volume( p ) );
  }

  // the same with the volume of p already known (the plane-constant solve evaluates it for ~15 planes through one cell)
  VOFX_HD inline double volume_fraction ( const Poly2 &p, double volumeOfP, const Hs2 &hs )
  {
    Poly2 c;
    poly_clip( p, hs, c );
    return vdiv( poly_volume( c ), volumeOfP );
  }

  // makePolygon of a quadrilateral cell, 2d/polygon.hh:230-241: corners 0,1,3,2; upper corner = lower + h
  VOFX_HD inline void make_cell ( double lox, double loy, double hx, double hy, Poly2 &p )
  {
    const double x1 = lox + hx, y1 = loy + hy;
    p.n = 4;
    p.x[ 0 ] = lox; p.y[ 0 ] = loy;
    p.x[ 1 ] = x1;  p.y[ 1 ] = loy;
    p.x[ 2 ] = x1;  p.y[ 2 ] = y1;
    p.x[ 3 ] = lox; p.y[ 3 ] = y1;
  }

  struct VfResidual2
  {
    const Poly2 *p;
    double nx, ny, fraction, volume;
    VOFX_HD double operator() ( double dist ) const
    {
      const Hs2 hs = { nx, ny, dist };
      return volume_fraction( *p, volume, hs ) - fraction;
    }
  };

  // locateHalfSpace( Polygon, normal, fraction ), geometry/algorithm.hh:68-111: bracket by the planes through the
  // vertices, degenerate returns, else Brent on the plane constant with absolute tolerance 1e-12
  VOFX_HD inline double locate ( const Poly2 &polygon, double nx, double ny, double fraction )
  {
    double pMin = 0.0, pMax = 0.0;
    double volMin = -DBL_MAX, volMax = DBL_MAX;
    const double cellVolume = poly_volume( polygon );
    for( int i = 0; i < polygon.n; ++i )
    {
      const double dist = -( dot2( nx, ny, polygon.x[ i ], polygon.y[ i ] ) );
      const Hs2 hs = { nx, ny, dist };
      const double volume = volume_fraction( polygon, cellVolume, hs );
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
    if( volMax == DBL_MAX )
      return pMin;
    if( volMin == -DBL_MAX )
      return pMax;
    if( volMin == volMax )
      return pMin;
    const VfResidual2 f = { &polygon, nx, ny, fraction, cellVolume };
    return brent( f, pMin, pMax, 1e-12 );
  }

  // truncVolume( upwindPolygon2d( face, v ), hs ): 2d/upwindpolygon.hh:24-30, geometry/upwindpolygon.hh:58-62.
  // face = 2*axis+side of the cell with lower corner (lox,loy); the face's corners are its lower corner and
  // that corner + h along the spanned axis; the swept region is the face translated by -v.
  VOFX_HD inline double flux ( double lox, double loy, double hx, double hy, int face, double vx, double vy, const Hs2 &hs )
  {
    const int axis = face >> 1;
    double c0x = lox, c0y = loy;
    if( face & 1 )
    {
      if( axis == 0 ) c0x += hx; else c0y += hy;
    }
    double c1x = c0x, c1y = c0y;
    if( axis == 0 ) c1y += hy; else c1x += hx;
    const double dx = c1x - c0x, dy = c1y - c0y;
    Poly2 up, cl;
    up.n = 4;
    if( dot2( -dy, dx, vx, vy ) < 0.0 )   // generalizedCrossProduct( c1 - c0 ) * v, geometry/utility.hh:82-85
    {
      up.x[ 0 ] = c0x; up.y[ 0 ] = c0y;
      up.x[ 1 ] = c1x; up.y[ 1 ] = c1y;
      up.x[ 2 ] = c1x - vx; up.y[ 2 ] = c1y - vy;
      up.x[ 3 ] = c0x - vx; up.y[ 3 ] = c0y - vy;
    }
    else
    {
      up.x[ 0 ] = c1x; up.y[ 0 ] = c1y;
      up.x[ 1 ] = c0x; up.y[ 1 ] = c0y;
      up.x[ 2 ] = c0x - vx; up.y[ 2 ] = c0y - vy;
      up.x[ 3 ] = c1x - vx; up.y[ 3 ] = c1y - vy;
    }
    poly_clip( up, hs, cl );
    return poly_volume( cl );
  }

  // intersect( Polygon, HyperPlane ) -> Line, 2d/intersect.hh:110-148.  Returns false (and a zero line) if empty.
  VOFX_HD inline bool plane_cut ( const Poly2 &p, const Hs2 &hs, double lx[ 2 ], double ly[ 2 ] )
  {
    lx[ 0 ] = lx[ 1 ] = ly[ 0 ] = ly[ 1 ] = 0.0;
    if( !hs_valid( hs ) )
      return false;
    int j = 0;
    const int n = p.n;
    for( int i = 0; i < n; ++i )
    {
      if( j == 2 )
        break;
      const int k = ( i + 1 == n ) ? 0 : i + 1;
      const double l0 = level_set( hs, p.x[ i ], p.y[ i ] );
      const double l1 = level_set( hs, p.x[ k ], p.y[ k ] );
      if( ( l0 > 0.0 && l1 < 0.0 ) || ( l0 < 0.0 && l1 > 0.0 ) )
      {
        const double a = -l1 / ( l0 - l1 );
        const double b = l0 / ( l0 - l1 );
        double cx = 0.0, cy = 0.0;
        cx += a * p.x[ i ]; cx += b * p.x[ k ];
        cy += a * p.y[ i ]; cy += b * p.y[ k ];
        lx[ j ] = cx; ly[ j ] = cy;
        j++;
      }
      else if( fabs( l0 ) < kEps )
      {
        lx[ j ] = p.x[ i ]; ly[ j ] = p.y[ i ];
        j++;
      }
    }
    if( j == 0 )
      return false;
    if( j == 1 )
    {
      lx[ 1 ] = lx[ 0 ];
      ly[ 1 ] = ly[ 0 ];
    }
    return true;
  }

  // ---- exact circle / polygon overlap (initial data and error of the rotating-circle problem) --------------------------
  // circleIntersection, interpolation/circleinterpolation.hh:23-72: the points where the segment l0-l1 meets the circle,
  // ordered along the segment
  VOFX_HD inline int circle_intersection ( double l0x, double l0y, double l1x, double l1y, double cx, double cy, double r, double *px, double *py )
  {
    int np = 0;
    const double ddx = l1x - l0x, ddy = l1y - l0y;
    const double nx = -ddy, ny = ddx;                       // generalizedCrossProduct, geometry/utility.hh:82-85
    const double p = dot2( nx, ny, l0x, l0y );
    const double d = p - nx * cx - ny * cy;
    const double nn = dot2( nx, ny, nx, ny );
    const double diskr = r * r * nn - d * d;
    if( diskr < 0 )
      return 0;
    else if( fabs( diskr ) == 0.0 )
    {
      double vx = nx * d, vy = ny * d;
      vx /= nn; vy /= nn;
      vx += cx; vy += cy;
      if( dot2( l0x - vx, l0y - vy, l1x - vx, l1y - vy ) <= 0.0 )
      {
        px[ 0 ] = vx; py[ 0 ] = vy;
        return 1;
      }
      return 0;
    }
    const double sq = sqrt( diskr );
    double v1x = nx * d + ny * sq, v1y = ny * d - nx * sq;
    v1x /= nn; v1y /= nn;
    v1x += cx; v1y += cy;
    if( dot2( l0x - v1x, l0y - v1y, l1x - v1x, l1y - v1y ) <= 0.0 )
    {
      px[ np ] = v1x; py[ np ] = v1y;
      np++;
    }
    double v2x = nx * d - ny * sq, v2y = ny * d + nx * sq;
    v2x /= nn; v2y /= nn;
    v2x += cx; v2y += cy;
    if( dot2( l0x - v2x, l0y - v2y, l1x - v2x, l1y - v2y ) <= 0.0 )
    {
      px[ np ] = v2x; py[ np ] = v2y;
      np++;
    }
    if( np == 2 )
    {
      const double ex = px[ 1 ] - px[ 0 ], ey = py[ 1 ] - py[ 0 ];
      if( dot2( -ey, ex, nx, ny ) < 0 )
      {
        const double tx = px[ 0 ], ty = py[ 0 ];
        px[ 0 ] = px[ 1 ]; py[ 0 ] = py[ 1 ];
        px[ 1 ] = tx; py[ 1 ] = ty;
      }
    }
    return np;
  }

  // intersectionVolume( polygon, center, radius ), circleinterpolation.hh:75-138: area of polygon /\ disc (inscribed polygon
  // + circular segments).  asin / sin are the platform's (CUDA's on the device): within 2 ulp of glibc's, so cut cells
  // agree with the reference to ~1e-16 absolute, uncut cells exactly.
  VOFX_HD inline double circle_polygon_volume ( const Poly2 &polygon, double cx, double cy, double radius )
  {
    double sx[ 16 ], sy[ 16 ];
    Poly2 poly;
    double qx[ 24 ], qy[ 24 ];          // up to 2 intersections + 1 vertex per edge of a polygon with <= 8 vertices
    int nSec = 0, nPoly = 0;
    int i = 0, i0 = 0;
    for( ; i < polygon.n; ++i )
      if( norm2( polygon.x[ i ] - cx, polygon.y[ i ] - cy ) > radius )
      {
        i0 = i;
        break;
      }
    if( i == polygon.n )
      return poly_volume( polygon );
    for( int k = 0; k < polygon.n; ++k )
    {
      const int e = ( k + i0 ) % polygon.n, e1 = ( e + 1 ) % polygon.n;
      double px[ 2 ], py[ 2 ];
      const int np = circle_intersection( polygon.x[ e ], polygon.y[ e ], polygon.x[ e1 ], polygon.y[ e1 ], cx, cy, radius, px, py );
      for( int q = 0; q < np; ++q )
      {
        sx[ nSec ] = px[ q ]; sy[ nSec ] = py[ q ]; nSec++;
        qx[ nPoly ] = px[ q ]; qy[ nPoly ] = py[ q ]; nPoly++;
      }
      if( norm2( polygon.x[ e1 ] - cx, polygon.y[ e1 ] - cy ) < radius )
      {
        qx[ nPoly ] = polygon.x[ e1 ]; qy[ nPoly ] = polygon.y[ e1 ]; nPoly++;
      }
    }
    double volume = 0.0;
    if( nSec > 1 )
    {
      if( nPoly > 2 )
      {
        // Polygon::volume of the collected points, 2d/polygon.hh:99-106
        double sum = 0.0;
        for( int a = 0; a < nPoly; ++a )
        {
          const int b = ( a + 1 == nPoly ) ? 0 : a + 1;
          sum += ( qy[ a ] + qy[ b ] ) * ( qx[ a ] - qx[ b ] );
        }
        volume += sum / 2.0;
      }
      for( int q = 1; q < nSec; q = q + 2 )
      {
        const int q1 = ( q + 1 ) % nSec;
        const double len = norm2( sx[ q ] - sx[ q1 ], sy[ q ] - sy[ q1 ] );
        const double alpha = 2.0 * asin( len / ( 2.0 * radius ) );
        volume += radius * radius / 2.0 * ( alpha - sin( alpha ) );
      }
    }
    (void) poly;
    return volume;
  }

  // one cell's term of l1error( gridView, reconstructions, flags, RotatingCircle< double, 2 > ), test/errors.hh:62-95;
  // exactFraction = the cell's value of circleInterpolation
  VOFX_HD inline double l1error_circle_cell ( double lox, double loy, double hx, double hy, const Hs2 &hs, bool full, double cx, double cy, double radius )
  {
    Poly2 cell;
    make_cell( lox, loy, hx, hy, cell );
    const double volume = hx * hy;                                          // geometry().volume() of the grid (include/vofx.h)
    const double exactFraction = circle_polygon_volume( cell, cx, cy, radius ) / volume;
    const double T1 = exactFraction * volume;
    const double T0 = volume - T1;
    if( !hs_valid( hs ) )
    {
      const double c = full ? 1.0 : 0.0;
      return T0 * fabs( c ) + T1 * fabs( 1.0 - c );
    }
    Poly2 one;
    poly_clip( cell, hs, one );
    const double shared = circle_polygon_volume( one, cx, cy, radius );
    return ( T1 - shared ) + ( poly_volume( one ) - shared );
  }

} // namespace vofx
