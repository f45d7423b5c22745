// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for dune-common exceptions.
#pragma once
#include <sstream>
#include <stdexcept>
namespace Dune
{
  struct Exception : std::runtime_error { using std::runtime_error::runtime_error; };
  struct InvalidStateException : Exception { using Exception::Exception; };
  struct NotImplemented : Exception { using Exception::Exception; };
}
#define DUNE_THROW( E, m ) do { std::ostringstream dune_throw_s_; dune_throw_s_ << m; throw Dune::E( dune_throw_s_.str() ); } while( 0 )
