// ompl_min: ompl::base::PlannerData / PlannerDataVertex (the subset the
// reference's getPlannerData uses: pr_ompl/src/RRTConnect.cpp:378-415).
#ifndef OMPL_MIN_BASE_PLANNERDATA_H
#define OMPL_MIN_BASE_PLANNERDATA_H
#include <algorithm>
#include <limits>
#include <map>
#include <vector>
#include "ompl/base/SpaceInformation.h"
namespace ompl {
namespace base {
class PlannerDataVertex {
public:
  PlannerDataVertex(const State *st, int tag = 0) : state_(st), tag_(tag) {}
  virtual ~PlannerDataVertex() {}
  virtual int getTag() const { return tag_; }
  virtual void setTag(int tag) { tag_ = tag; }
  virtual const State *getState() const { return state_; }
  bool operator==(const PlannerDataVertex &rhs) const { return state_ == rhs.state_; }

protected:
  const State *state_;
  int tag_;
};

class PlannerData {
public:
  static const unsigned int INVALID_INDEX;
  explicit PlannerData(const SpaceInformationPtr &si) : si_(si) {}
  virtual ~PlannerData() {}
  PlannerData(const PlannerData &) = delete;
  PlannerData &operator=(const PlannerData &) = delete;

  unsigned int addVertex(const PlannerDataVertex &st) {
    if (!st.getState())
      return INVALID_INDEX;
    auto it = stateIndexMap_.find(st.getState());
    if (it != stateIndexMap_.end())
      return it->second;
    unsigned int idx = (unsigned int)vertices_.size();
    vertices_.push_back(st);
    out_.emplace_back();
    stateIndexMap_[st.getState()] = idx;
    return idx;
  }
  unsigned int addStartVertex(const PlannerDataVertex &v) {
    unsigned int idx = addVertex(v);
    if (idx != INVALID_INDEX && !isStartVertex(idx))
      startVertexIndices_.push_back(idx);
    return idx;
  }
  unsigned int addGoalVertex(const PlannerDataVertex &v) {
    unsigned int idx = addVertex(v);
    if (idx != INVALID_INDEX && !isGoalVertex(idx))
      goalVertexIndices_.push_back(idx);
    return idx;
  }
  bool addEdge(unsigned int v1, unsigned int v2) {
    if (v1 >= vertices_.size() || v2 >= vertices_.size())
      return false;
    if (std::find(out_[v1].begin(), out_[v1].end(), v2) != out_[v1].end())
      return false;
    out_[v1].push_back(v2);
    return true;
  }
  bool addEdge(const PlannerDataVertex &v1, const PlannerDataVertex &v2) {
    unsigned int i1 = addVertex(v1), i2 = addVertex(v2);
    if (i1 == INVALID_INDEX || i2 == INVALID_INDEX)
      return false;
    return addEdge(i1, i2);
  }
  unsigned int vertexIndex(const PlannerDataVertex &v) const {
    auto it = stateIndexMap_.find(v.getState());
    return it == stateIndexMap_.end() ? INVALID_INDEX : it->second;
  }
  bool vertexExists(const PlannerDataVertex &v) const { return vertexIndex(v) != INVALID_INDEX; }
  bool edgeExists(unsigned int v1, unsigned int v2) const {
    return v1 < out_.size() && std::find(out_[v1].begin(), out_[v1].end(), v2) != out_[v1].end();
  }
  unsigned int numVertices() const { return (unsigned int)vertices_.size(); }
  unsigned int numEdges() const {
    unsigned int n = 0;
    for (auto &o : out_)
      n += (unsigned int)o.size();
    return n;
  }
  const PlannerDataVertex &getVertex(unsigned int index) const { return vertices_.at(index); }
  unsigned int getEdges(unsigned int v, std::vector<unsigned int> &edgeList) const {
    edgeList = out_.at(v);
    return (unsigned int)edgeList.size();
  }
  unsigned int numStartVertices() const { return (unsigned int)startVertexIndices_.size(); }
  unsigned int numGoalVertices() const { return (unsigned int)goalVertexIndices_.size(); }
  unsigned int getStartIndex(unsigned int i) const { return startVertexIndices_.at(i); }
  unsigned int getGoalIndex(unsigned int i) const { return goalVertexIndices_.at(i); }
  bool isStartVertex(unsigned int index) const {
    return std::find(startVertexIndices_.begin(), startVertexIndices_.end(), index) !=
           startVertexIndices_.end();
  }
  bool isGoalVertex(unsigned int index) const {
    return std::find(goalVertexIndices_.begin(), goalVertexIndices_.end(), index) !=
           goalVertexIndices_.end();
  }
  virtual void clear() {
    vertices_.clear();
    out_.clear();
    stateIndexMap_.clear();
    startVertexIndices_.clear();
    goalVertexIndices_.clear();
  }
  const SpaceInformationPtr &getSpaceInformation() const { return si_; }

protected:
  SpaceInformationPtr si_;
  std::vector<PlannerDataVertex> vertices_;
  std::vector<std::vector<unsigned int>> out_;
  std::map<const State *, unsigned int> stateIndexMap_;
  std::vector<unsigned int> startVertexIndices_;
  std::vector<unsigned int> goalVertexIndices_;
};
__attribute__((weak)) const unsigned int PlannerData::INVALID_INDEX =
    std::numeric_limits<unsigned int>::max();
} // namespace base
} // namespace ompl
#endif
