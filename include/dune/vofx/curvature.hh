// dune/vofx/curvature.hh -- mirrors dune/vof/curvature.hh: everything lives in dune/vofx/vofx.hh
#ifndef DUNE_VOFX_CURVATURE_HH
#define DUNE_VOFX_CURVATURE_HH
#include <dune/vofx/vofx.hh>
#endif
