// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for dune-common Timer.
#pragma once
#include <chrono>
namespace Dune
{
  struct Timer
  {
    explicit Timer ( bool startNow = true ) { if( startNow ) start(); }
    void start () { t0_ = clock::now(); running_ = true; }
    double stop () { if( running_ ) { acc_ += std::chrono::duration< double >( clock::now() - t0_ ).count(); running_ = false; } return acc_; }
    double elapsed () const { return acc_ + ( running_ ? std::chrono::duration< double >( clock::now() - t0_ ).count() : 0.0 ); }
  private:
    using clock = std::chrono::steady_clock;
    clock::time_point t0_;
    double acc_ = 0.0;
    bool running_ = false;
  };
}
