// ORACLE / TEST INFRASTRUCTURE ONLY.  Force-included (-include) when compiling
// the reference's own pr_ompl/src/RRTConnect.cpp from /root/reference: the
// reference spells boost::shared_ptr / boost::bind / _1 / _2 (OMPL <= 1.1 era,
// pr_ompl/include/pr_ompl/RRTConnect.h:154, pr_ompl/src/RRTConnect.cpp:135-136)
// and Boost is not installed in this image.  Map them onto the std:: equivalents.
#ifndef ORACLE_BOOST_COMPAT_H
#define ORACLE_BOOST_COMPAT_H
#include <functional>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
namespace boost {
using std::bind;
using std::function;
using std::shared_ptr;
} // namespace boost
using namespace std::placeholders; // boost exposes _1, _2 at global scope
#endif
