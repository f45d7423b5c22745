// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for dune-common typetraits.
#pragma once
#include <type_traits>
namespace Dune
{
  template< class... > using void_t = void;
  template< class T > using real_t = T;
}
