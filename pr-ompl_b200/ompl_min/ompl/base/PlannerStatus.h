// ompl_min: ompl::base::PlannerStatus.
#ifndef OMPL_MIN_BASE_PLANNERSTATUS_H
#define OMPL_MIN_BASE_PLANNERSTATUS_H
#include <string>
namespace ompl {
namespace base {
struct PlannerStatus {
  enum StatusType {
    UNKNOWN = 0,
    INVALID_START,
    INVALID_GOAL,
    UNRECOGNIZED_GOAL_TYPE,
    TIMEOUT,
    APPROXIMATE_SOLUTION,
    EXACT_SOLUTION,
    CRASH,
    TYPE_COUNT
  };
  PlannerStatus(StatusType status = UNKNOWN) : status_(status) {}
  PlannerStatus(bool hasSolution, bool isApproximate)
      : status_(hasSolution ? (isApproximate ? APPROXIMATE_SOLUTION : EXACT_SOLUTION) : TIMEOUT) {}
  std::string asString() const {
    static const char *names[] = {"Unknown status",         "Invalid start",
                                  "Invalid goal",           "Unrecognized goal type",
                                  "Timeout",                "Approximate solution",
                                  "Exact solution",         "Crash"};
    return status_ < TYPE_COUNT ? names[status_] : "Unknown status";
  }
  operator bool() const { return status_ == APPROXIMATE_SOLUTION || status_ == EXACT_SOLUTION; }
  operator StatusType() const { return status_; }

private:
  StatusType status_;
};
} // namespace base
} // namespace ompl
#endif
