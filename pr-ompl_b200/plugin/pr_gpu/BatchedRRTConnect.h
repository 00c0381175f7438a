// pr_gpu::BatchedRRTConnect -- many independent RRTConnect queries against one GpuWorld, planned
// entirely on the device (prgpu_rrtc_batch_*, one thread block per query).  Each query is what one
// pr_ompl::RRTConnect instance would do with a single start, a single GoalState and a
// CounterSampler(seed, queryIdBase + i): same trees, same path.
#ifndef PR_GPU_BATCHEDRRTCONNECT_H
#define PR_GPU_BATCHEDRRTCONNECT_H
#include <string>
#include <vector>
#include <ompl/base/PlannerStatus.h>
#include <ompl/base/spaces/RealVectorBounds.h>
#include "pr_gpu/GpuWorld.h"
namespace pr_gpu {
class BatchedRRTConnect {
public:
  struct Result {
    ompl::base::PlannerStatus::StatusType status;
    unsigned iterations, startTreeSize, goalTreeSize;
  };
  BatchedRRTConnect(const GpuWorldPtr &world, const ompl::base::RealVectorBounds &bounds, unsigned nQueries)
      : world_(world), n_(nQueries), handle_(nullptr) {
    params_ = prgpu_rrtc_params();
    params_.dim = world->dof();
    for (int i = 0; i < params_.dim; ++i) {
      params_.lo[i] = bounds.low.at(i);
      params_.hi[i] = bounds.high.at(i);
    }
    params_.range = 0.0;
    params_.resolution_fraction = 0.01;
    params_.ext_first = PRGPU_EXT_EXTEND;
    params_.ext_second = PRGPU_EXT_CONNECT;
    params_.max_iters = 1000;
    params_.max_nodes = 2048;
  }
  ~BatchedRRTConnect() { prgpu_rrtc_batch_destroy(handle_); }
  BatchedRRTConnect(const BatchedRRTConnect &) = delete;
  BatchedRRTConnect &operator=(const BatchedRRTConnect &) = delete;

  void setRange(double r) { params_.range = r; }
  void setStateValidityCheckingResolution(double f) { params_.resolution_fraction = f; }
  void setExtensionTypesString(const std::string &s) {
    if (s == "con-con") { params_.ext_first = 1; params_.ext_second = 1; }
    else if (s == "con-ext") { params_.ext_first = 1; params_.ext_second = 0; }
    else if (s == "ext-con") { params_.ext_first = 0; params_.ext_second = 1; }
    else if (s == "ext-ext") { params_.ext_first = 0; params_.ext_second = 0; }
    else throw std::runtime_error("Invalid extension type string.");
  }
  void setSampleStream(std::uint64_t seed, std::uint64_t queryIdBase) {
    params_.seed = seed;
    params_.query_id_base = queryIdBase;
  }
  void setLimits(unsigned maxIterations, unsigned maxNodesPerTree) {
    params_.max_iters = maxIterations;
    params_.max_nodes = maxNodesPerTree;
  }
  /// starts/goals: nQueries x dof, row-major
  std::vector<Result> solve(const double *starts, const double *goals) {
    if (handle_)
      prgpu_rrtc_batch_destroy(handle_);
    handle_ = nullptr;
    check(prgpu_rrtc_batch_create(world_->handle(), &params_, n_, &handle_), "BatchedRRTConnect");
    check(prgpu_rrtc_batch_run(handle_, starts, goals), "BatchedRRTConnect::solve");
    std::vector<int32_t> st(n_);
    std::vector<uint32_t> it(n_), ns(n_), ng(n_);
    check(prgpu_rrtc_batch_summary(handle_, st.data(), it.data(), ns.data(), ng.data()),
          "BatchedRRTConnect::solve");
    std::vector<Result> out(n_);
    for (unsigned i = 0; i < n_; ++i)
      out[i] = Result{(ompl::base::PlannerStatus::StatusType)st[i], it[i], ns[i], ng[i]};
    return out;
  }
  /// solution path of one query as dof-sized rows (empty if unsolved)
  std::vector<std::vector<double>> path(unsigned query) const {
    const int dim = params_.dim;
    uint32_t len = 0;
    check(prgpu_rrtc_batch_fetch(handle_, query, nullptr, nullptr, nullptr, nullptr, &len, nullptr, 0),
          "BatchedRRTConnect::path");
    std::vector<double> flat((std::size_t)len * dim);
    if (len)
      check(prgpu_rrtc_batch_fetch(handle_, query, nullptr, nullptr, nullptr, nullptr, &len, flat.data(), len),
            "BatchedRRTConnect::path");
    std::vector<std::vector<double>> rows(len);
    for (uint32_t i = 0; i < len; ++i)
      rows[i].assign(flat.begin() + (std::size_t)i * dim, flat.begin() + (std::size_t)(i + 1) * dim);
    return rows;
  }

private:
  GpuWorldPtr world_;
  unsigned n_;
  prgpu_rrtc_params params_;
  prgpu_rrtc_batch_t *handle_;
};
} // namespace pr_gpu
#endif
