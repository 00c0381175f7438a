// ompl_min: ompl::base::State (opaque base; concrete layout belongs to the space).
#ifndef OMPL_MIN_BASE_STATE_H
#define OMPL_MIN_BASE_STATE_H
namespace ompl {
namespace base {
class State {
public:
  template <class T> const T *as() const { return static_cast<const T *>(this); }
  template <class T> T *as() { return static_cast<T *>(this); }
  State(const State &) = delete;
  State &operator=(const State &) = delete;

protected:
  State() {}
  virtual ~State() {}
};
} // namespace base
} // namespace ompl
#endif
