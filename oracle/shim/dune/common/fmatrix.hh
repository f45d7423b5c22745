// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for dune-common's FieldMatrix.
// solve()/determinant() restate the closed forms published in dune-common 2.6 densematrix.hh
// (1x1 division, 2x2 and 3x3 Cramer with the maple-generated 3x3 determinant).  dune-common itself is
// not available in this image to diff against, so parity is UNPINNED at this one boundary
// (call site on the hot path: reconstruction/modifiedyoungs.hh:126).
#pragma once
#include "fvector.hh"

namespace Dune
{

  template< class K, int r, int c >
  struct FieldMatrix
  {
    using row_type = FieldVector< K, c >;

    FieldMatrix () {}
    FieldMatrix ( const K &k ) { for( auto &x : m_ ) x = k; }
    FieldMatrix ( std::initializer_list< row_type > l )
    {
      int i = 0;
      for( const auto &x : l )
        m_[ i++ ] = x;
    }

    row_type &operator[] ( std::size_t i ) { return m_[ i ]; }
    const row_type &operator[] ( std::size_t i ) const { return m_[ i ]; }

    FieldMatrix &operator+= ( const FieldMatrix &o ) { for( int i = 0; i < r; ++i ) m_[ i ] += o.m_[ i ]; return *this; }

    K determinant () const
    {
      static_assert( r == c, "square matrices only" );
      const FieldMatrix &M = *this;
      if constexpr ( r == 1 )
        return M[ 0 ][ 0 ];
      else if constexpr ( r == 2 )
        return M[ 0 ][ 0 ] * M[ 1 ][ 1 ] - M[ 0 ][ 1 ] * M[ 1 ][ 0 ];
      else
      {
        static_assert( r == 3, "only up to 3x3" );
        const K t4 = M[ 0 ][ 0 ] * M[ 1 ][ 1 ];
        const K t6 = M[ 0 ][ 0 ] * M[ 1 ][ 2 ];
        const K t8 = M[ 0 ][ 1 ] * M[ 1 ][ 0 ];
        const K t10 = M[ 0 ][ 2 ] * M[ 1 ][ 0 ];
        const K t12 = M[ 0 ][ 1 ] * M[ 2 ][ 0 ];
        const K t14 = M[ 0 ][ 2 ] * M[ 2 ][ 0 ];
        return ( t4 * M[ 2 ][ 2 ] - t6 * M[ 2 ][ 1 ] - t8 * M[ 2 ][ 2 ] + t10 * M[ 2 ][ 1 ] + t12 * M[ 1 ][ 2 ] - t14 * M[ 1 ][ 1 ] );
      }
    }

    template< class V >
    void solve ( V &x, const V &b ) const
    {
      static_assert( r == c, "square matrices only" );
      const FieldMatrix &M = *this;
      if constexpr ( r == 1 )
        x[ 0 ] = b[ 0 ] / M[ 0 ][ 0 ];
      else if constexpr ( r == 2 )
      {
        K detinv = M[ 0 ][ 0 ] * M[ 1 ][ 1 ] - M[ 0 ][ 1 ] * M[ 1 ][ 0 ];
        detinv = 1.0 / detinv;
        x[ 0 ] = detinv * ( M[ 1 ][ 1 ] * b[ 0 ] - M[ 0 ][ 1 ] * b[ 1 ] );
        x[ 1 ] = detinv * ( M[ 0 ][ 0 ] * b[ 1 ] - M[ 1 ][ 0 ] * b[ 0 ] );
      }
      else
      {
        const K d = determinant();
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
    }

    template< class X, class Y >
    void mv ( const X &x, Y &y ) const
    {
      for( int i = 0; i < r; ++i )
      {
        y[ i ] = 0;
        for( int j = 0; j < c; ++j )
          y[ i ] += m_[ i ][ j ] * x[ j ];
      }
    }

    std::array< row_type, r > m_;
  };

} // namespace Dune
