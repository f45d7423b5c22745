// TEST INFRASTRUCTURE ONLY (oracle/).
// Problems that BASELINE.json's configs name but the reference does not ship (hard part 7):
// the LeVeque deformation field.  They follow the reference's Problem concept
// (evaluate / velocityField, cf. dune/vof/test/problems/rotatingcircle.hh:67-103) so that the reference's own
// Velocity adaptor (dune/vof/velocity.hh:5-33) and Evolution operator run on them unchanged.
//
// The spatial factors use libm sin() of the face-centre coordinate; the product side precomputes the same
// per-axis factors on the host with the same libm from bit-identical coordinates (include/vofx.h, separable
// velocity descriptor) and multiplies them in the same order, so the face velocities agree bit-for-bit.
// The time factor cos(pi t/T) is evaluated with the fixed polynomial below (only + - * and comparisons), which
// the device restates, because libm's and CUDA's cos are not guaranteed to agree in the last bit.
#pragma once
#include <cmath>

#include <dune/common/fvector.hh>

namespace VofxOracle
{

  // cos(pi * s) from + - * only.  |error| about 1 ulp.  This *defines* the time factor of the LeVeque problem.
  inline double detCosPi ( double s )
  {
    if( s < 0.0 )
      s = -s;
    double r = s - 2.0 * std::floor( s * 0.5 );   // [0,2)
    if( r > 1.0 )
      r = 2.0 - r;                                // [0,1]
    double sign = 1.0;
    if( r > 0.5 )
    {
      r = 1.0 - r;                                // [0,0.5]
      sign = -1.0;
    }
    const double pi = 3.14159265358979323846;
    if( r <= 0.25 )
    {
      const double x = pi * r;
      const double z = x * x;
      double p = 1.0 / 20922789888000.0;          // 1/16!
      p = p * z - 1.0 / 87178291200.0;            // 1/14!
      p = p * z + 1.0 / 479001600.0;              // 1/12!
      p = p * z - 1.0 / 3628800.0;                // 1/10!
      p = p * z + 1.0 / 40320.0;                  // 1/8!
      p = p * z - 1.0 / 720.0;                    // 1/6!
      p = p * z + 1.0 / 24.0;                     // 1/4!
      p = p * z - 0.5;
      p = p * z + 1.0;
      return sign * p;
    }
    else
    {
      const double x = pi * ( 0.5 - r );
      const double z = x * x;
      double p = 1.0 / 355687428096000.0;         // 1/17!
      p = p * z - 1.0 / 1307674368000.0;          // 1/15!
      p = p * z + 1.0 / 6227020800.0;             // 1/13!
      p = p * z - 1.0 / 39916800.0;               // 1/11!
      p = p * z + 1.0 / 362880.0;                 // 1/9!
      p = p * z - 1.0 / 5040.0;                   // 1/7!
      p = p * z + 1.0 / 120.0;                    // 1/5!
      p = p * z - 1.0 / 6.0;
      p = p * z + 1.0;
      return sign * ( p * x );
    }
  }

  inline double sinSq ( double x ) { const double s = std::sin( M_PI * x ); return s * s; }
  inline double sinTwo ( double x ) { return std::sin( 2.0 * M_PI * x ); }

} // namespace VofxOracle


template< class ctype, int dim >
struct LeVeque {};

// 2D single vortex (LeVeque 1996): u = sin^2(pi x) sin(2 pi y) c(t), v = -sin(2 pi x) sin^2(pi y) c(t); disc R=0.15 at (0.5,0.75)
template< class ctype >
struct LeVeque< ctype, 2 >
{
  using DomainType = Dune::FieldVector< ctype, 2 >;
  using RangeType = Dune::FieldVector< ctype, 1 >;

  explicit LeVeque ( double period = 8.0 ) : period_( period ) {}

  void evaluate ( const DomainType &x, double, RangeType &u ) const
  {
    DomainType y { 0.5, 0.75 };
    y -= x;
    u = ( y.two_norm() < 0.15 ) ? 1.0 : 0.0;
  }
  void evaluate ( const DomainType &x, RangeType &u ) const { evaluate( x, 0.0, u ); }

  void velocityField ( const DomainType &x, const double t, DomainType &v ) const
  {
    const double c = VofxOracle::detCosPi( t / period_ );
    v[ 0 ] = ( VofxOracle::sinSq( x[ 0 ] ) * VofxOracle::sinTwo( x[ 1 ] ) ) * c;
    v[ 1 ] = ( ( -VofxOracle::sinTwo( x[ 0 ] ) ) * VofxOracle::sinSq( x[ 1 ] ) ) * c;
  }

  double period_;
};

// 3D deformation field (LeVeque 1996), SURVEY.md section 8(d) config C3: sphere R=0.15 at (0.35,0.35,0.35), T=3
template< class ctype >
struct LeVeque< ctype, 3 >
{
  using DomainType = Dune::FieldVector< ctype, 3 >;
  using RangeType = Dune::FieldVector< ctype, 1 >;

  explicit LeVeque ( double period = 3.0 ) : period_( period ) {}

  void evaluate ( const DomainType &x, double, RangeType &u ) const
  {
    DomainType y { 0.35, 0.35, 0.35 };
    y -= x;
    u = ( y.two_norm() < 0.15 ) ? 1.0 : 0.0;
  }
  void evaluate ( const DomainType &x, RangeType &u ) const { evaluate( x, 0.0, u ); }

  void velocityField ( const DomainType &x, const double t, DomainType &v ) const
  {
    using namespace VofxOracle;
    const double c = detCosPi( t / period_ );
    v[ 0 ] = ( ( ( 2.0 * sinSq( x[ 0 ] ) ) * sinTwo( x[ 1 ] ) ) * sinTwo( x[ 2 ] ) ) * c;
    v[ 1 ] = ( ( ( -sinTwo( x[ 0 ] ) ) * sinSq( x[ 1 ] ) ) * sinTwo( x[ 2 ] ) ) * c;
    v[ 2 ] = ( ( ( -sinTwo( x[ 0 ] ) ) * sinTwo( x[ 1 ] ) ) * sinSq( x[ 2 ] ) ) * c;
  }

  double period_;
};
