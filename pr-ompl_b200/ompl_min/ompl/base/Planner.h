// ompl_min: ompl::base::Planner, PlannerSpecs, PlannerInputStates.
// PlannerInputStates follows upstream semantics as recalled in SURVEY.md
// Appendix B.7 (each start yielded once if in bounds and valid; nextGoal
// samples the GoalSampleableRegion up to maxSampleCount(), the ptc overload
// keeps waiting while the region could still produce samples).
#ifndef OMPL_MIN_BASE_PLANNER_H
#define OMPL_MIN_BASE_PLANNER_H
#include <chrono>
#include <thread>
#include "ompl/base/GenericParam.h"
#include "ompl/base/PlannerData.h"
#include "ompl/base/PlannerStatus.h"
#include "ompl/base/PlannerTerminationCondition.h"
#include "ompl/base/ProblemDefinition.h"
#include "ompl/base/goals/GoalSampleableRegion.h"
namespace ompl {
namespace base {
class Planner;
typedef std::shared_ptr<Planner> PlannerPtr;

class PlannerInputStates {
public:
  explicit PlannerInputStates(const Planner *planner) : planner_(planner) {
    tempState_ = nullptr;
    update();
  }
  PlannerInputStates() : planner_(nullptr) {
    tempState_ = nullptr;
    clear();
  }
  ~PlannerInputStates() { clear(); }
  void clear() {
    if (tempState_) {
      si_->freeState(tempState_);
      tempState_ = nullptr;
    }
    addedStartStates_ = 0;
    sampledGoalsCount_ = 0;
    pdef_ = nullptr;
    si_ = nullptr;
  }
  void restart() {
    addedStartStates_ = 0;
    sampledGoalsCount_ = 0;
  }
  bool update();
  bool use(const ProblemDefinitionPtr &pdef) { return use(pdef.get()); }
  bool use(const ProblemDefinition *pdef) {
    if (pdef && pdef_ != pdef) {
      clear();
      pdef_ = pdef;
      si_ = pdef->getSpaceInformation().get();
      return true;
    }
    return false;
  }
  void checkValidity() const {
    std::string error;
    if (!pdef_)
      error = "Problem definition not specified";
    else {
      if (pdef_->getStartStateCount() <= 0)
        error = "No start states specified";
      else if (!pdef_->getGoal())
        error = "No goal specified";
    }
    if (!error.empty())
      throw Exception(error);
  }
  const State *nextStart() {
    if (pdef_ == nullptr || si_ == nullptr) {
      OMPL_ERROR("PlannerInputStates: Missing space information or problem definition");
      return nullptr;
    }
    while (addedStartStates_ < pdef_->getStartStateCount()) {
      const State *st = pdef_->getStartState(addedStartStates_);
      addedStartStates_++;
      bool bounds = si_->satisfiesBounds(st);
      bool valid = bounds ? si_->isValid(st) : false;
      if (bounds && valid)
        return st;
      OMPL_WARN("Skipping invalid start state (invalid %s)", bounds ? "state" : "bounds");
    }
    return nullptr;
  }
  const State *nextGoal() { return nextGoal(plannerAlwaysTerminatingCondition()); }
  const State *nextGoal(const PlannerTerminationCondition &ptc) {
    if (pdef_ == nullptr || si_ == nullptr) {
      OMPL_ERROR("PlannerInputStates: Missing space information or problem definition");
      return nullptr;
    }
    const GoalSampleableRegion *goal =
        pdef_->getGoal()->hasType(GOAL_SAMPLEABLE_REGION)
            ? pdef_->getGoal()->as<GoalSampleableRegion>()
            : nullptr;
    if (goal) {
      bool attempt = true;
      while (attempt) {
        attempt = false;
        if (sampledGoalsCount_ < goal->maxSampleCount() && goal->canSample()) {
          if (tempState_ == nullptr)
            tempState_ = si_->allocState();
          do {
            goal->sampleGoal(tempState_);
            sampledGoalsCount_++;
            bool bounds = si_->satisfiesBounds(tempState_);
            bool valid = bounds ? si_->isValid(tempState_) : false;
            if (bounds && valid)
              return tempState_;
            OMPL_WARN("Skipping invalid goal state (invalid %s)", bounds ? "state" : "bounds");
          } while (!ptc && sampledGoalsCount_ < goal->maxSampleCount() && goal->canSample());
        }
        if (goal->couldSample() && !ptc) {
          std::this_thread::sleep_for(std::chrono::milliseconds(10));
          attempt = !ptc;
        }
      }
    }
    return nullptr;
  }
  bool haveMoreStartStates() const {
    return pdef_ && addedStartStates_ < pdef_->getStartStateCount();
  }
  bool haveMoreGoalStates() const {
    if (pdef_ && pdef_->getGoal() && pdef_->getGoal()->hasType(GOAL_SAMPLEABLE_REGION))
      return sampledGoalsCount_ < pdef_->getGoal()->as<GoalSampleableRegion>()->maxSampleCount();
    return false;
  }
  unsigned int getSeenStartStatesCount() const { return addedStartStates_; }
  unsigned int getSampledGoalsCount() const { return sampledGoalsCount_; }

private:
  const Planner *planner_;
  unsigned int addedStartStates_ = 0;
  unsigned int sampledGoalsCount_ = 0;
  State *tempState_;
  const ProblemDefinition *pdef_ = nullptr;
  const SpaceInformation *si_ = nullptr;
};

struct PlannerSpecs {
  GoalType recognizedGoal = GOAL_ANY;
  bool multithreaded = false;
  bool approximateSolutions = false;
  bool optimizingPaths = false;
  bool directed = false;
  bool provingSolutionNonExistence = false;
};

class Planner {
public:
  Planner(const SpaceInformationPtr &si, const std::string &name)
      : si_(si), pis_(this), name_(name), setup_(false) {
    if (!si_)
      throw Exception(name_, "Invalid space information instance for planner");
  }
  virtual ~Planner() {}
  Planner(const Planner &) = delete;
  Planner &operator=(const Planner &) = delete;
  template <class T> T *as() { return static_cast<T *>(this); }
  template <class T> const T *as() const { return static_cast<const T *>(this); }

  const SpaceInformationPtr &getSpaceInformation() const { return si_; }
  const ProblemDefinitionPtr &getProblemDefinition() const { return pdef_; }
  const PlannerInputStates &getPlannerInputStates() const { return pis_; }
  virtual void setProblemDefinition(const ProblemDefinitionPtr &pdef) {
    pdef_ = pdef;
    pis_.update();
  }
  virtual PlannerStatus solve(const PlannerTerminationCondition &ptc) = 0;
  PlannerStatus solve(const PlannerTerminationConditionFn &ptc, double /*checkInterval*/) {
    return solve(PlannerTerminationCondition(ptc));
  }
  PlannerStatus solve(double solveTime) {
    if (solveTime < 1.0)
      return solve(timedPlannerTerminationCondition(solveTime));
    return solve(timedPlannerTerminationCondition(solveTime));
  }
  virtual void clear() {
    pis_.clear();
    pis_.update();
  }
  virtual void getPlannerData(PlannerData &data) const { (void)data; }
  const std::string &getName() const { return name_; }
  void setName(const std::string &name) { name_ = name; }
  const PlannerSpecs &getSpecs() const { return specs_; }
  virtual void setup() {
    if (!si_->isSetup()) {
      OMPL_INFORM("%s: Space information setup was not yet called. Calling now.", getName().c_str());
      si_->setup();
    }
    if (setup_)
      OMPL_WARN("%s: Planner setup called multiple times", getName().c_str());
    else
      setup_ = true;
  }
  virtual void checkValidity() {
    if (!isSetup())
      setup();
    pis_.checkValidity();
  }
  bool isSetup() const { return setup_; }
  ParamSet &params() { return params_; }
  const ParamSet &params() const { return params_; }

protected:
  template <typename T, typename PlannerType, typename SetterType, typename GetterType>
  void declareParam(const std::string &name, const PlannerType &planner, const SetterType &setter,
                    const GetterType &getter, const std::string &rangeSuggestion = "") {
    params_.declareParam<T>(name, std::bind(setter, planner, std::placeholders::_1),
                            std::bind(getter, planner));
    if (!rangeSuggestion.empty())
      params_[name]->setRangeSuggestion(rangeSuggestion);
  }
  template <typename T, typename PlannerType, typename SetterType>
  void declareParam(const std::string &name, const PlannerType &planner, const SetterType &setter,
                    const std::string &rangeSuggestion = "") {
    params_.declareParam<T>(name, std::bind(setter, planner, std::placeholders::_1));
    if (!rangeSuggestion.empty())
      params_[name]->setRangeSuggestion(rangeSuggestion);
  }

  SpaceInformationPtr si_;
  ProblemDefinitionPtr pdef_;
  PlannerInputStates pis_;
  std::string name_;
  PlannerSpecs specs_;
  ParamSet params_;
  bool setup_;
};

inline bool PlannerInputStates::update() {
  if (!planner_)
    throw Exception("No planner set for PlannerInputStates");
  return use(planner_->getProblemDefinition());
}

typedef std::function<PlannerPtr(const SpaceInformationPtr &)> PlannerAllocator;
} // namespace base
} // namespace ompl
#endif
