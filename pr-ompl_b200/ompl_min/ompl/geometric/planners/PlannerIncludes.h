// ompl_min: umbrella header used by geometric planners.
#ifndef OMPL_MIN_GEOMETRIC_PLANNERS_PLANNERINCLUDES_H
#define OMPL_MIN_GEOMETRIC_PLANNERS_PLANNERINCLUDES_H
#include <sstream>
#include "ompl/base/Planner.h"
#include "ompl/base/PlannerData.h"
#include "ompl/base/ProblemDefinition.h"
#include "ompl/base/SpaceInformation.h"
#include "ompl/base/goals/GoalSampleableRegion.h"
#include "ompl/base/goals/GoalState.h"
#include "ompl/base/goals/GoalStates.h"
#include "ompl/geometric/PathGeometric.h"
#include "ompl/util/Console.h"
#include "ompl/util/Exception.h"
#include "ompl/util/RandomNumbers.h"
#endif
