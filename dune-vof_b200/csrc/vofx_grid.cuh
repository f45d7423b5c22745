// vofx: device view of the (slab of the) structured grid, index/coordinate helpers, velocity tables.
#pragma once

#include "vofx_math.cuh"

namespace vofx
{

  // Storage box of one context: the owned cells plus `halo` layers on either side along the last axis.
  // Coordinates are generated from GLOBAL indices (include/vofx.h), so every rank reproduces the serial bits.
  struct Grid
  {
    int dim;
    int ns[ 3 ];        // storage cells per axis (ns[2] = 1 in 2D)
    int g0[ 3 ];        // global multi-index of storage cell (0,0,0)
    int ng[ 3 ];        // global cells per axis
    double origin[ 3 ], h[ 3 ];
    int own0, own1;     // owned layers [own0, own1) along the last axis (storage index)
    int list0, list1;   // layers whose mixed cells are reconstructed (owned +- 1, inside the domain)
    long long cells;    // ns[0]*ns[1]*ns[2]
  };

  VOFX_INLINE int last_axis ( const Grid &g ) { return g.dim - 1; }

  // a context holds fewer than 2^31 cells (vofx_create): 32-bit division, not the 64-bit software routine
  VOFX_INLINE void cell_ijk ( const Grid &g, long long c, int *ijk )
  {
    unsigned r = (unsigned) c;
    const unsigned n0 = (unsigned) g.ns[ 0 ], n1 = (unsigned) g.ns[ 1 ];
    const unsigned q0 = r / n0;
    ijk[ 0 ] = int( r - q0 * n0 );
    const unsigned q1 = q0 / n1;
    ijk[ 1 ] = int( q0 - q1 * n1 );
    ijk[ 2 ] = int( q1 );
  }

  VOFX_INLINE long long cell_index ( const Grid &g, const int *ijk )
  {
    return ijk[ 0 ] + (long long) g.ns[ 0 ] * ( ijk[ 1 ] + (long long) g.ns[ 1 ] * ijk[ 2 ] );
  }

  // a neighbour exists iff it is inside the global domain (otherwise it is a wall) and inside this context's storage
  VOFX_INLINE bool cell_exists ( const Grid &g, const int *ijk )
  {
    for( int d = 0; d < 3; ++d )
    {
      if( d >= g.dim )
        break;
      const int gi = ijk[ d ] + g.g0[ d ];
      if( gi < 0 || gi >= g.ng[ d ] || ijk[ d ] < 0 || ijk[ d ] >= g.ns[ d ] )
        return false;
    }
    return true;
  }

  VOFX_INLINE bool in_domain ( const Grid &g, const int *ijk )
  {
    for( int d = 0; d < 3; ++d )
    {
      if( d >= g.dim )
        break;
      const int gi = ijk[ d ] + g.g0[ d ];
      if( gi < 0 || gi >= g.ng[ d ] )
        return false;
    }
    return true;
  }

  VOFX_INLINE double cell_lower ( const Grid &g, int d, int i ) { return g.origin[ d ] + ( i + g.g0[ d ] ) * g.h[ d ]; }

  VOFX_INLINE double cell_volume ( const Grid &g )
  {
    double v = 1;
    for( int d = 0; d < 3; ++d )
      if( d < g.dim )
        v *= g.h[ d ];
    return v;
  }

  VOFX_INLINE double face_measure_of ( const Grid &g, int face )
  {
    double v = 1;
    for( int d = 0; d < 3; ++d )
      if( d < g.dim && d != ( face >> 1 ) )
        v *= g.h[ d ];
    return v;
  }

  VOFX_INLINE bool is_mixed ( unsigned char f ) { return f == 1 || f == 2; }   // FlagSet::isMixed, flagset.hh:71-72,78

  // ---- separable face-centre velocity (include/vofx.h) ------------------------------------------------------------
  struct Velocity
  {
    int nterms[ 3 ];
    double coef[ 3 ];
    int timeKind[ 3 ];
    double period[ 3 ];
    // tab[component][term][axis][0 centre | 1 lower face | 2 upper face] -> table indexed by the GLOBAL cell index
    const double *tab[ 3 ][ 2 ][ 3 ][ 3 ];
    // general fallback (vofx_velocity_set_faces): velocity sampled by the host functor at every face centre,
    // faces[ ( cell * 2*dim + face ) * dim + component ], frozen at the start-of-step time by the caller
    const double *faces;
    // the same for the band only (vofx_velocity_set_band_faces): bandFaces[ ( bandSlot[ cell ] * 2*dim + face ) * dim + component ],
    // bandSlot = position of the cell in the band list of the pass in flight
    const double *bandFaces;
    const int *bandSlot;
  };

  // velocity at the centre of `face` of the cell with storage multi-index ijk, as that cell's own intersection sees
  // it (Velocity::operator(), velocity.hh:15-20).  v has 3 entries; unused ones are 0.
  VOFX_HD inline void face_velocity ( const Grid &g, const Velocity &vel, const int *ijk, int face, double time, double *v )
  {
    if( vel.bandFaces )
    {
      const long long c = vel.bandSlot[ cell_index( g, ijk ) ];
      for( int a = 0; a < 3; ++a )
        v[ a ] = ( a < g.dim ) ? vel.bandFaces[ ( c * 2 * g.dim + face ) * g.dim + a ] : 0.0;
      return;
    }
    if( vel.faces )
    {
      const long long c = cell_index( g, ijk );
      for( int a = 0; a < 3; ++a )
        v[ a ] = ( a < g.dim ) ? vel.faces[ ( c * 2 * g.dim + face ) * g.dim + a ] : 0.0;
      return;
    }
    const int fa = face >> 1;
    const int kindOfFaceAxis = ( face & 1 ) ? 2 : 1;
    for( int a = 0; a < 3; ++a )
    {
      v[ a ] = 0.0;
      if( a >= g.dim || vel.nterms[ a ] == 0 )
        continue;
      double sum = 0.0;
      for( int r = 0; r < vel.nterms[ a ]; ++r )
      {
        double term = vel.tab[ a ][ r ][ 0 ][ fa == 0 ? kindOfFaceAxis : 0 ][ ijk[ 0 ] + g.g0[ 0 ] ];
        term = term * vel.tab[ a ][ r ][ 1 ][ fa == 1 ? kindOfFaceAxis : 0 ][ ijk[ 1 ] + g.g0[ 1 ] ];
        if( g.dim == 3 )
          term = term * vel.tab[ a ][ r ][ 2 ][ fa == 2 ? kindOfFaceAxis : 0 ][ ijk[ 2 ] + g.g0[ 2 ] ];
        sum = ( r == 0 ) ? term : sum + term;
      }
      double va = sum * vel.coef[ a ];
      if( vel.timeKind[ a ] == 1 )
        va = va * det_cospi( time / vel.period[ a ] );
      v[ a ] = va;
    }
  }

  // the time factors of a step are the same for every face: evaluate them once per thread ...
  VOFX_HD inline void velocity_time_factors ( const Velocity &vel, double time, double *tf )
  {
    for( int a = 0; a < 3; ++a )
      tf[ a ] = ( vel.timeKind[ a ] == 1 ) ? det_cospi( time / vel.period[ a ] ) : 1.0;
  }

  // ... and use them here: the same operations as face_velocity with det_cospi( time / period ) already evaluated
  VOFX_HD inline void face_velocity_tf ( const Grid &g, const Velocity &vel, const int *ijk, int face, const double *tf, double *v )
  {
    if( vel.bandFaces )
    {
      const long long c = vel.bandSlot[ cell_index( g, ijk ) ];
      for( int a = 0; a < 3; ++a )
        v[ a ] = ( a < g.dim ) ? vel.bandFaces[ ( c * 2 * g.dim + face ) * g.dim + a ] : 0.0;
      return;
    }
    if( vel.faces )
    {
      const long long c = cell_index( g, ijk );
      for( int a = 0; a < 3; ++a )
        v[ a ] = ( a < g.dim ) ? vel.faces[ ( c * 2 * g.dim + face ) * g.dim + a ] : 0.0;
      return;
    }
    const int fa = face >> 1;
    const int kindOfFaceAxis = ( face & 1 ) ? 2 : 1;
    for( int a = 0; a < 3; ++a )
    {
      v[ a ] = 0.0;
      if( a >= g.dim || vel.nterms[ a ] == 0 )
        continue;
      double sum = 0.0;
      for( int r = 0; r < vel.nterms[ a ]; ++r )
      {
        double term = vel.tab[ a ][ r ][ 0 ][ fa == 0 ? kindOfFaceAxis : 0 ][ ijk[ 0 ] + g.g0[ 0 ] ];
        term = term * vel.tab[ a ][ r ][ 1 ][ fa == 1 ? kindOfFaceAxis : 0 ][ ijk[ 1 ] + g.g0[ 1 ] ];
        if( g.dim == 3 )
          term = term * vel.tab[ a ][ r ][ 2 ][ fa == 2 ? kindOfFaceAxis : 0 ][ ijk[ 2 ] + g.g0[ 2 ] ];
        sum = ( r == 0 ) ? term : sum + term;
      }
      double va = sum * vel.coef[ a ];
      if( vel.timeKind[ a ] == 1 )
        va = va * tf[ a ];
      v[ a ] = va;
    }
  }

} // namespace vofx
