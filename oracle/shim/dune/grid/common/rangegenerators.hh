// TEST INFRASTRUCTURE ONLY (oracle/): forwards to the serial Cartesian grid model.
#pragma once
#include <dune/grid/common/minigrid.hh>
