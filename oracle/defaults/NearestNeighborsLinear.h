// ORACLE / TEST INFRASTRUCTURE ONLY -- never included by the product path.
// CPU restatement of OMPL's NearestNeighborsLinear<T> (third-party dependency
// of the reference, unpinned: pr_ompl/package.xml:12; semantics UPSTREAM-RECALLED,
// SURVEY.md Appendix B.4).  This is the NN structure the reference's
// RRTConnect gets through SelfConfig::getDefaultNearestNeighbors
// (pr_ompl/src/RRTConnect.cpp:131-134) in the oracle build.
//   nearest : linear scan in insertion order, first strict minimum wins
//             => ties resolve to the lowest insertion index.
//   nearestK: order by (distance, insertion index)  [upstream partial_sort leaves
//             tie order unspecified; the fixed rule is (distance, index)].
// Parity unpinned upstream: the reference has no tests.
#ifndef ORACLE_NEARESTNEIGHBORSLINEAR_H
#define ORACLE_NEARESTNEIGHBORSLINEAR_H
#include <algorithm>
#include <utility>
#include <vector>
#include "ompl/base/StateSpace.h"
#include "ompl/datastructures/NearestNeighbors.h"
#include "ompl/util/Exception.h"
namespace ompl {
template <typename _T> class NearestNeighborsLinear : public NearestNeighbors<_T> {
public:
  NearestNeighborsLinear() : NearestNeighbors<_T>() {}
  ~NearestNeighborsLinear() override {}
  bool reportsSortedResults() const override { return true; }
  void clear() override { data_.clear(); }
  void add(const _T &data) override { data_.push_back(data); }
  void add(const std::vector<_T> &data) override {
    data_.reserve(data_.size() + data.size());
    data_.insert(data_.end(), data.begin(), data.end());
  }
  bool remove(const _T &data) override {
    if (!data_.empty())
      for (int i = (int)data_.size() - 1; i >= 0; --i)
        if (data_[i] == data) {
          data_.erase(data_.begin() + i);
          return true;
        }
    return false;
  }
  _T nearest(const _T &data) const override {
    const std::size_t sz = data_.size();
    std::size_t pos = sz;
    double dmin = 0.0;
    for (std::size_t i = 0; i < sz; ++i) {
      double distance = NearestNeighbors<_T>::distFun_(data_[i], data);
      if (pos == sz || dmin > distance) {
        pos = i;
        dmin = distance;
      }
    }
    if (pos != sz)
      return data_[pos];
    throw Exception("No elements found in nearest neighbors data structure");
  }
  void nearestK(const _T &data, std::size_t k, std::vector<_T> &nbh) const override {
    std::vector<std::pair<double, std::size_t>> d(data_.size());
    for (std::size_t i = 0; i < data_.size(); ++i)
      d[i] = std::make_pair(NearestNeighbors<_T>::distFun_(data_[i], data), i);
    if (k > d.size())
      k = d.size();
    std::partial_sort(d.begin(), d.begin() + k, d.end());
    nbh.resize(k);
    for (std::size_t i = 0; i < k; ++i)
      nbh[i] = data_[d[i].second];
  }
  void nearestR(const _T &data, double radius, std::vector<_T> &nbh) const override {
    std::vector<std::pair<double, std::size_t>> d;
    for (std::size_t i = 0; i < data_.size(); ++i) {
      double dist = NearestNeighbors<_T>::distFun_(data_[i], data);
      if (dist <= radius)
        d.push_back(std::make_pair(dist, i));
    }
    std::sort(d.begin(), d.end());
    nbh.resize(d.size());
    for (std::size_t i = 0; i < d.size(); ++i)
      nbh[i] = data_[d[i].second];
  }
  std::size_t size() const override { return data_.size(); }
  void list(std::vector<_T> &data) const override { data = data_; }

protected:
  std::vector<_T> data_;
};
} // namespace ompl

// hook consumed by ompl_min's SelfConfig::getDefaultNearestNeighbors in the oracle build
template <typename T>
ompl::NearestNeighbors<T> *ompl_min_default_nn(const ompl::base::StateSpacePtr &) {
  return new ompl::NearestNeighborsLinear<T>();
}
#endif
