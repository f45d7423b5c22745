// dune/vofx/evolution.hh -- mirrors dune/vof/evolution.hh: everything lives in dune/vofx/vofx.hh
#ifndef DUNE_VOFX_EVOLUTION_HH
#define DUNE_VOFX_EVOLUTION_HH
#include <dune/vofx/vofx.hh>
#endif
