// ompl_min: ompl::NearestNeighbors<T> abstract interface (the drop-in boundary
// for the tree's NN structure; reference use: pr_ompl/include/pr_ompl/RRTConnect.h:154,
// pr_ompl/src/RRTConnect.cpp:131-136,181,209).
#ifndef OMPL_MIN_DATASTRUCTURES_NEARESTNEIGHBORS_H
#define OMPL_MIN_DATASTRUCTURES_NEARESTNEIGHBORS_H
#include <functional>
#include <vector>
namespace ompl {
template <typename _T> class NearestNeighbors {
public:
  typedef std::function<double(const _T &, const _T &)> DistanceFunction;
  NearestNeighbors() {}
  virtual ~NearestNeighbors() {}
  virtual void setDistanceFunction(const DistanceFunction &distFun) { distFun_ = distFun; }
  const DistanceFunction &getDistanceFunction() const { return distFun_; }
  virtual bool reportsSortedResults() const = 0;
  virtual void clear() = 0;
  virtual void add(const _T &data) = 0;
  virtual void add(const std::vector<_T> &data) {
    for (auto &d : data)
      add(d);
  }
  virtual bool remove(const _T &data) = 0;
  virtual _T nearest(const _T &data) const = 0;
  virtual void nearestK(const _T &data, std::size_t k, std::vector<_T> &nbh) const = 0;
  virtual void nearestR(const _T &data, double radius, std::vector<_T> &nbh) const = 0;
  virtual std::size_t size() const = 0;
  virtual void list(std::vector<_T> &data) const = 0;

protected:
  DistanceFunction distFun_;
};
} // namespace ompl
#endif
