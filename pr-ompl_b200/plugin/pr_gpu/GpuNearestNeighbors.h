// pr_gpu::GpuNearestNeighbors<T> -- an ompl::NearestNeighbors<T> whose queries run on the GPU
// (prgpu_knn_*, kernel family K1).  Drop-in for the structure pr_ompl::RRTConnect obtains from
// SelfConfig::getDefaultNearestNeighbors or setNearestNeighbors<NN>()
// (pr_ompl/src/RRTConnect.cpp:131-136, pr_ompl/include/pr_ompl/RRTConnect.h:113-118): default
// constructible, stores the caller's elements without owning them, answers nearest() with the
// first strict minimum in insertion order (the order of OMPL's linear structure).
//
// The elements are opaque to OMPL's NN interface (only the distance functor looks inside), so a
// GPU structure needs the coordinates.  Two ways:
//  * default accessor: if an element is a pointer to something with a member `state` (an
//    ompl::base::State* -- the shape of every OMPL planner's Motion, pr_ompl/include/pr_ompl/RRTConnect.h:
//    129-151), the coordinates are state->as<RealVectorStateSpace::StateType>()->values.  The member is
//    reached through the deduced element type, so the reference's *protected* nested Motion class works
//    unmodified.  The dimension comes from the state space given to the constructor (what
//    SelfConfig::getDefaultNearestNeighbors passes) or, for the default-constructed structure that
//    setNearestNeighbors<NN>() creates (RRTConnect.h:113-118), from setDefaultDimension();
//  * setCoordinateAccessor(fn, dim) for any other element type.
// The distance function given through setDistanceFunction() is kept for API compatibility but never
// evaluated: distances are RealVectorStateSpace::distance, bit for bit.
#ifndef PR_GPU_GPUNEARESTNEIGHBORS_H
#define PR_GPU_GPUNEARESTNEIGHBORS_H
#include <algorithm>
#include <functional>
#include <vector>
#include <type_traits>
#include <ompl/base/spaces/RealVectorStateSpace.h>
#include <ompl/datastructures/NearestNeighbors.h>
#include "pr_gpu/Error.h"
namespace pr_gpu {
namespace detail {
template <class T, class = void> struct HasStateMember : std::false_type {};
template <class T>
struct HasStateMember<T, decltype(void(std::declval<const T &>()->state))> : std::true_type {};
template <class T>
typename std::enable_if<HasStateMember<T>::value, const double *>::type motionCoordinates(const T &e) {
  return e->state->template as<ompl::base::RealVectorStateSpace::StateType>()->values;
}
template <class T>
typename std::enable_if<!HasStateMember<T>::value, const double *>::type motionCoordinates(const T &) {
  throw ompl::Exception("GpuNearestNeighbors",
                        "element type has no `state` member: call setCoordinateAccessor(fn, dim) before add()");
}
inline int &defaultDimension() {
  static int dim = 0;
  return dim;
}
} // namespace detail
/// dimension used by default-constructed GpuNearestNeighbors (the setNearestNeighbors<NN>() path)
inline void setDefaultDimension(int dim) { detail::defaultDimension() = dim; }

template <typename _T> class GpuNearestNeighbors : public ompl::NearestNeighbors<_T> {
public:
  typedef std::function<const double *(const _T &)> CoordinateAccessor;

  GpuNearestNeighbors() : handle_(nullptr), dim_(0), device_(0) {}
  explicit GpuNearestNeighbors(const std::shared_ptr<ompl::base::StateSpace> &space)
      : handle_(nullptr), dim_(0), device_(0) {
    requireRealVectorSpace(space.get(), "GpuNearestNeighbors");
    dim_ = (int)space->getDimension();
  }
  ~GpuNearestNeighbors() override { prgpu_knn_destroy(handle_); }
  GpuNearestNeighbors(const GpuNearestNeighbors &) = delete;
  GpuNearestNeighbors &operator=(const GpuNearestNeighbors &) = delete;

  void setCoordinateAccessor(const CoordinateAccessor &fn, int dim) {
    if (!data_.empty() && dim != dim_)
      throw ompl::Exception("GpuNearestNeighbors", "dimension changed on a non-empty structure");
    coords_ = fn;
    dim_ = dim;
  }
  void setDevice(int device) { device_ = device; }
  bool reportsSortedResults() const override { return true; }

  void clear() override {
    data_.clear();
    if (handle_)
      check(prgpu_knn_clear(handle_), "GpuNearestNeighbors::clear");
  }
  void add(const _T &data) override {
    ensure();
    check(prgpu_knn_add(handle_, coords_(data), 1), "GpuNearestNeighbors::add");
    data_.push_back(data);
  }
  void add(const std::vector<_T> &data) override {
    if (data.empty())
      return;
    ensure();
    std::vector<double> buf(data.size() * dim_);
    for (std::size_t i = 0; i < data.size(); ++i)
      std::copy(coords_(data[i]), coords_(data[i]) + dim_, buf.begin() + i * dim_);
    check(prgpu_knn_add(handle_, buf.data(), data.size()), "GpuNearestNeighbors::add");
    data_.insert(data_.end(), data.begin(), data.end());
  }
  bool remove(const _T &data) override {
    // not on the planners' path (never called by the reference): erase and re-upload
    for (std::size_t i = data_.size(); i-- > 0;)
      if (data_[i] == data) {
        std::vector<_T> keep(data_);
        keep.erase(keep.begin() + i);
        clear();
        add(keep);
        return true;
      }
    return false;
  }
  _T nearest(const _T &data) const override {
    if (data_.empty())
      throw ompl::Exception("No elements found in nearest neighbors data structure");
    uint32_t idx = PRGPU_INVALID_INDEX;
    check(prgpu_knn_nearest_k(handle_, coords_(data), 1, 1, nullptr, &idx, nullptr),
          "GpuNearestNeighbors::nearest");
    return data_[idx];
  }
  void nearestK(const _T &data, std::size_t k, std::vector<_T> &nbh) const override {
    nbh.clear();
    k = std::min(k, data_.size());
    if (k == 0)
      return;
    if (k > PRGPU_MAX_K)
      throw ompl::Exception("GpuNearestNeighbors::nearestK", "k exceeds PRGPU_MAX_K");
    std::vector<uint32_t> idx(k);
    check(prgpu_knn_nearest_k(handle_, coords_(data), 1, (int)k, nullptr, idx.data(), nullptr),
          "GpuNearestNeighbors::nearestK");
    for (uint32_t i : idx)
      if (i != PRGPU_INVALID_INDEX)
        nbh.push_back(data_[i]);
  }
  void nearestR(const _T &data, double radius, std::vector<_T> &nbh) const override {
    nbh.clear();
    const std::size_t k = std::min<std::size_t>(data_.size(), PRGPU_MAX_K);
    if (k == 0)
      return;
    std::vector<uint32_t> idx(k);
    std::vector<double> dist(k);
    check(prgpu_knn_nearest_k(handle_, coords_(data), 1, (int)k, nullptr, idx.data(), dist.data()),
          "GpuNearestNeighbors::nearestR");
    for (std::size_t i = 0; i < k; ++i)
      if (idx[i] != PRGPU_INVALID_INDEX && dist[i] <= radius)
        nbh.push_back(data_[idx[i]]);
    if (nbh.size() == k && k < data_.size())
      throw ompl::Exception("GpuNearestNeighbors::nearestR", "more than PRGPU_MAX_K elements within radius");
  }
  std::size_t size() const override { return data_.size(); }
  void list(std::vector<_T> &data) const override { data = data_; }

  /// batched queries against the same tree (kernel K1 at its efficient shape)
  void nearestBatch(const double *queries, std::size_t nq, std::vector<_T> &out) const {
    out.clear();
    if (nq == 0 || data_.empty())
      return;
    std::vector<uint32_t> idx(nq);
    check(prgpu_knn_nearest_k(handle_, queries, nq, 1, nullptr, idx.data(), nullptr),
          "GpuNearestNeighbors::nearestBatch");
    for (uint32_t i : idx)
      out.push_back(data_[i]);
  }
  prgpu_knn_t *handle() const { return handle_; }

private:
  void ensure() {
    if (dim_ <= 0)
      dim_ = detail::defaultDimension();
    if (!coords_ && dim_ > 0)
      coords_ = [](const _T &e) { return detail::motionCoordinates(e); };
    if (!coords_ || dim_ <= 0)
      throw ompl::Exception("GpuNearestNeighbors", "unknown dimension: construct with the state space, call "
                                                   "pr_gpu::setDefaultDimension(dim) or setCoordinateAccessor(fn, dim)");
    if (!handle_)
      check(prgpu_knn_create(device_, dim_, 1024, &handle_), "GpuNearestNeighbors");
  }
  std::vector<_T> data_; // insertion order == device index == tie-break order
  CoordinateAccessor coords_;
  prgpu_knn_t *handle_;
  int dim_, device_;
};
} // namespace pr_gpu

// hook for ompl_min's SelfConfig::getDefaultNearestNeighbors in the product build
template <typename T>
ompl::NearestNeighbors<T> *ompl_min_default_nn(const std::shared_ptr<ompl::base::StateSpace> &space) {
  return new pr_gpu::GpuNearestNeighbors<T>(space);
}
#endif
