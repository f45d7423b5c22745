// TEST INFRASTRUCTURE ONLY (oracle/): name-only stand-in.
#pragma once
namespace Dune { template< int dim, class ct > struct VirtualRefinement {}; }
