// dune/vofx/reconstruction.hh -- mirrors dune/vof/reconstruction.hh: everything lives in dune/vofx/vofx.hh
#ifndef DUNE_VOFX_RECONSTRUCTION_HH
#define DUNE_VOFX_RECONSTRUCTION_HH
#include <dune/vofx/vofx.hh>
#endif
