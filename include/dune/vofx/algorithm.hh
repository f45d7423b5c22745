// dune/vofx/algorithm.hh -- mirrors dune/vof/algorithm.hh: everything lives in dune/vofx/vofx.hh
#ifndef DUNE_VOFX_ALGORITHM_HH
#define DUNE_VOFX_ALGORITHM_HH
#include <dune/vofx/vofx.hh>
#endif
