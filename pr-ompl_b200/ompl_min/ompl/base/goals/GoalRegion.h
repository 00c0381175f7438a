#ifndef OMPL_MIN_BASE_GOALS_GOALREGION_H
#define OMPL_MIN_BASE_GOALS_GOALREGION_H
#include "ompl/base/Goal.h"
#endif
