// ompl_min: ompl::base::Path.
#ifndef OMPL_MIN_BASE_PATH_H
#define OMPL_MIN_BASE_PATH_H
#include <iosfwd>
#include "ompl/base/SpaceInformation.h"
namespace ompl {
namespace base {
class Path {
public:
  explicit Path(const SpaceInformationPtr &si) : si_(si) {}
  virtual ~Path() {}
  Path(const Path &) = delete;
  Path &operator=(const Path &) = delete;
  const SpaceInformationPtr &getSpaceInformation() const { return si_; }
  template <class T> T *as() { return static_cast<T *>(this); }
  template <class T> const T *as() const { return static_cast<const T *>(this); }
  virtual double length() const = 0;
  virtual bool check() const = 0;

protected:
  SpaceInformationPtr si_;
};
typedef std::shared_ptr<Path> PathPtr;
} // namespace base
} // namespace ompl
#endif
