// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for dune-common StaticPower.
#pragma once
namespace Dune
{
  template< int b, int e > struct StaticPower { static constexpr int power = b * StaticPower< b, e - 1 >::power; };
  template< int b > struct StaticPower< b, 0 > { static constexpr int power = 1; };
}
