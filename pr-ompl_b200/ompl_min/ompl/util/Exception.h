// ompl_min: minimal stand-in for the OMPL 1.x public API surface that
// personalrobotics/pr-ompl touches.  OMPL itself is a
// third-party dependency that is not vendored by the reference and not
// installed in this image; these headers exist so that planners written
// against the OMPL plugin API compile and run here.  They are interface
// plumbing only: no nearest-neighbour structure and no motion validator
// lives in this tree (the GPU ones are in ../../plugin, the CPU ones used
// for checking are under /oracle).
#ifndef OMPL_MIN_UTIL_EXCEPTION_H
#define OMPL_MIN_UTIL_EXCEPTION_H
#include <stdexcept>
#include <string>
namespace ompl {
class Exception : public std::runtime_error {
public:
  explicit Exception(const std::string &what) : std::runtime_error(what) {}
  Exception(const std::string &prefix, const std::string &what)
      : std::runtime_error(prefix + ": " + what) {}
  ~Exception() noexcept override {}
};
} // namespace ompl
#endif
