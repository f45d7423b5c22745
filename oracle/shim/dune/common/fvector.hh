// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for dune-common's FieldVector so that the
// reference's header-only operators under /root/reference/dune/vof can be instantiated unchanged.
// dune-common is a third-party dependency that is not vendored by the reference (dune.module:4).
// Accumulations are plain left-to-right loops, as in dune-common's DenseVector.
#pragma once
#include <array>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <ostream>

namespace Dune
{

  template< class K, int n >
  struct FieldVector
  {
    using value_type = K;
    using field_type = K;
    using size_type = std::size_t;
    static constexpr int dimension = n;

    // the reference relies on value-initialisation (e.g. 2d/intersect.hh:87-89 does "Coord point; point.axpy(..)")
    FieldVector () { v_.fill( K( 0 ) ); }
    FieldVector ( const K &k ) { v_.fill( k ); }
    FieldVector ( std::initializer_list< K > l )
    {
      v_.fill( K( 0 ) );
      int i = 0;
      for( const K &x : l )
        v_[ i++ ] = x;
    }

    K &operator[] ( std::size_t i ) { return v_[ i ]; }
    const K &operator[] ( std::size_t i ) const { return v_[ i ]; }
    static constexpr int size () { return n; }

    FieldVector &operator= ( const K &k ) { v_.fill( k ); return *this; }
    FieldVector &operator+= ( const FieldVector &o ) { for( int i = 0; i < n; ++i ) v_[ i ] += o.v_[ i ]; return *this; }
    FieldVector &operator-= ( const FieldVector &o ) { for( int i = 0; i < n; ++i ) v_[ i ] -= o.v_[ i ]; return *this; }
    FieldVector &operator*= ( const K &k ) { for( int i = 0; i < n; ++i ) v_[ i ] *= k; return *this; }
    FieldVector &operator/= ( const K &k ) { for( int i = 0; i < n; ++i ) v_[ i ] /= k; return *this; }
    FieldVector &axpy ( const K &a, const FieldVector &x ) { for( int i = 0; i < n; ++i ) v_[ i ] += a * x.v_[ i ]; return *this; }

    K operator* ( const FieldVector &o ) const { K s = 0; for( int i = 0; i < n; ++i ) s += v_[ i ] * o.v_[ i ]; return s; }
    K two_norm2 () const { K s = 0; for( int i = 0; i < n; ++i ) s += v_[ i ] * v_[ i ]; return s; }
    K two_norm () const { return std::sqrt( two_norm2() ); }

    bool operator== ( const FieldVector &o ) const { return v_ == o.v_; }
    bool operator!= ( const FieldVector &o ) const { return !( v_ == o.v_ ); }

    operator K () const { static_assert( n == 1, "scalar conversion only for n == 1" ); return v_[ 0 ]; }

    std::array< K, n > v_;
  };

  template< class K, int n >
  FieldVector< K, n > operator+ ( FieldVector< K, n > a, const FieldVector< K, n > &b ) { a += b; return a; }
  template< class K, int n >
  FieldVector< K, n > operator- ( FieldVector< K, n > a, const FieldVector< K, n > &b ) { a -= b; return a; }
  template< class K, int n >
  FieldVector< K, n > operator* ( const K &k, FieldVector< K, n > a ) { a *= k; return a; }
  template< class K, int n >
  FieldVector< K, n > operator* ( FieldVector< K, n > a, const K &k ) { a *= k; return a; }
  template< class K, int n >
  std::ostream &operator<< ( std::ostream &o, const FieldVector< K, n > &v )
  {
    for( int i = 0; i < n; ++i )
      o << ( i ? " " : "" ) << v[ i ];
    return o;
  }

} // namespace Dune
