// vofx: launchers of the interface-extraction kernels (see vofx_launch.h).
#include "vofx_launch.h"
#include "vofx_kernels.cuh"

namespace vofx
{
  namespace launch
  {
    int scan_blocks ( long long n ) { return (int) ( ( n + kScanBlock - 1 ) / kScanBlock ); }

    // exclusive scan: mode flags (counts == nullptr): out = ordered list of the owned mixed cells; mode counts: out[i] = offset of item i
    void scan ( int dim, cudaStream_t s, Grid g, long long n, const unsigned char *flags, const int *counts, int *blockSums, long long *total, int *out, int phase )
    {
      const int blocks = scan_blocks( n );
      if( blocks == 0 )
        return;
      if( phase == 0 )
      {
        if( dim == 2 )
          k_scan_block_sums< 2 ><<< blocks, 256, 0, s >>>( g, n, flags, counts, blockSums );
        else
          k_scan_block_sums< 3 ><<< blocks, 256, 0, s >>>( g, n, flags, counts, blockSums );
        k_scan_block_offsets<<< 1, 1024, 0, s >>>( blocks, blockSums, total );
      }
      else
      {
        if( dim == 2 )
          k_scan_scatter< 2 ><<< blocks, 256, 0, s >>>( g, n, flags, counts, blockSums, out );
        else
          k_scan_scatter< 3 ><<< blocks, 256, 0, s >>>( g, n, flags, counts, blockSums, out );
      }
    }

    void interface_cut ( int dim, int blocks, cudaStream_t s, Grid g, const double *hs, const int *cells, long long nMixed, int *nverts, double *verts )
    {
      if( dim == 2 )
        k_interface_cut< 2 ><<< blocks, 128, 0, s >>>( g, hs, cells, nMixed, nverts, verts );
      else
        k_interface_cut< 3 ><<< blocks, 128, 0, s >>>( g, hs, cells, nMixed, nverts, verts );
    }

    void interface_gather ( int dim, int blocks, cudaStream_t s, long long nMixed, const int *nverts, const int *offsets, const double *verts, double *csr )
    {
      if( dim == 2 )
        k_interface_gather< 2 ><<< blocks, 128, 0, s >>>( nMixed, nverts, offsets, verts, csr );
      else
        k_interface_gather< 3 ><<< blocks, 128, 0, s >>>( nMixed, nverts, offsets, verts, csr );
    }
  } // namespace launch
} // namespace vofx
