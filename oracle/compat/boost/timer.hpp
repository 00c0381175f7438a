// ORACLE / TEST INFRASTRUCTURE ONLY.
// Stand-in for boost::timer (frechet/src/NNFrechet.cpp:1,97,154,198,206-212,220): elapsed() = seconds
// since construction.  Boost is absent from this image; see boost/graph/adjacency_list.hpp.
#ifndef ORACLE_COMPAT_BOOST_TIMER_HPP
#define ORACLE_COMPAT_BOOST_TIMER_HPP
#include <chrono>
namespace boost {
class timer {
public:
  timer() : start_(std::chrono::steady_clock::now()) {}
  void restart() { start_ = std::chrono::steady_clock::now(); }
  double elapsed() const {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - start_).count();
  }

private:
  std::chrono::steady_clock::time_point start_;
};
} // namespace boost
#endif
