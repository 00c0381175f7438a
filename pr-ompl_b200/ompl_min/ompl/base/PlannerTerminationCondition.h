// ompl_min: ompl::base::PlannerTerminationCondition and factory helpers.
#ifndef OMPL_MIN_BASE_PLANNERTERMINATIONCONDITION_H
#define OMPL_MIN_BASE_PLANNERTERMINATIONCONDITION_H
#include <atomic>
#include <chrono>
#include <functional>
#include <memory>
namespace ompl {
namespace base {
typedef std::function<bool()> PlannerTerminationConditionFn;

class PlannerTerminationCondition {
public:
  PlannerTerminationCondition(const PlannerTerminationConditionFn &fn) : impl_(new Impl(fn)) {}
  void terminate() const { impl_->terminate_ = true; }
  bool eval() const { return impl_->terminate_ || impl_->fn_(); }
  bool operator()() const { return eval(); }
  operator bool() const { return eval(); }

private:
  struct Impl {
    explicit Impl(const PlannerTerminationConditionFn &fn) : fn_(fn), terminate_(false) {}
    PlannerTerminationConditionFn fn_;
    std::atomic<bool> terminate_;
  };
  std::shared_ptr<Impl> impl_;
};

inline PlannerTerminationCondition plannerNonTerminatingCondition() {
  return PlannerTerminationCondition([] { return false; });
}
inline PlannerTerminationCondition plannerAlwaysTerminatingCondition() {
  return PlannerTerminationCondition([] { return true; });
}
inline PlannerTerminationCondition timedPlannerTerminationCondition(double duration) {
  auto end = std::chrono::steady_clock::now() +
             std::chrono::duration_cast<std::chrono::steady_clock::duration>(
                 std::chrono::duration<double>(duration));
  return PlannerTerminationCondition([end] { return std::chrono::steady_clock::now() >= end; });
}
/// terminates after being evaluated numIterations times (upstream IterationTerminationCondition)
inline PlannerTerminationCondition iterationPlannerTerminationCondition(unsigned int numIterations) {
  auto count = std::make_shared<unsigned int>(0u);
  return PlannerTerminationCondition(
      [count, numIterations] { return ++(*count) > numIterations; });
}
inline PlannerTerminationCondition plannerOrTerminationCondition(const PlannerTerminationCondition &a,
                                                                 const PlannerTerminationCondition &b) {
  return PlannerTerminationCondition([a, b] { return a() || b(); });
}
} // namespace base
} // namespace ompl
#endif
