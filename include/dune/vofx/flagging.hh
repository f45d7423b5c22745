// dune/vofx/flagging.hh -- mirrors dune/vof/flagging.hh: everything lives in dune/vofx/vofx.hh
#ifndef DUNE_VOFX_FLAGGING_HH
#define DUNE_VOFX_FLAGGING_HH
#include <dune/vofx/vofx.hh>
#endif
