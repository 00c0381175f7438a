// pr_gpu::GpuWorld, GpuStateValidityChecker, GpuMotionValidator -- the synthetic world of the
// BASELINE configs behind OMPL's validity interfaces (kernel family K2).
//   GpuStateValidityChecker : ompl::base::StateValidityChecker   (isValid, RRTConnect.cpp:198)
//   GpuMotionValidator      : ompl::base::MotionValidator        (si_->checkMotion, RRTConnect.cpp:198)
// checkMotion follows DiscreteMotionValidator: nd = validSegmentCount(s1, s2), tested set
// {s2} U {j/nd}; s1 is assumed valid.
#ifndef PR_GPU_GPUWORLD_H
#define PR_GPU_GPUWORLD_H
#include <memory>
#include <vector>
#include <ompl/base/MotionValidator.h>
#include <ompl/base/SpaceInformation.h>
#include <ompl/base/StateValidityChecker.h>
#include "pr_gpu/Error.h"
namespace pr_gpu {
class GpuWorld {
public:
  GpuWorld(int device, int dof, const double *dh, int nLinkSpheres, const int32_t *sphereLink,
           const double *sphereXyzr, int nObsSpheres, const double *obsSpheres, int nObsBoxes,
           const double *obsBoxes)
      : dof_(dof), device_(device), handle_(nullptr) {
    check(prgpu_world_create(device, dof, dh, nLinkSpheres, sphereLink, sphereXyzr, nObsSpheres, obsSpheres,
                             nObsBoxes, obsBoxes, &handle_),
          "GpuWorld");
  }
  ~GpuWorld() { prgpu_world_destroy(handle_); }
  GpuWorld(const GpuWorld &) = delete;
  GpuWorld &operator=(const GpuWorld &) = delete;
  prgpu_world_t *handle() const { return handle_; }
  int dof() const { return dof_; }
  int device() const { return device_; }

private:
  int dof_, device_;
  prgpu_world_t *handle_;
};
typedef std::shared_ptr<GpuWorld> GpuWorldPtr;

inline const double *reals(const ompl::base::SpaceInformation *si, const ompl::base::State *s) {
  return si->getStateSpace()->getValueAddressAtIndex(s, 0); // contiguous for RealVectorStateSpace
}

class GpuStateValidityChecker : public ompl::base::StateValidityChecker {
public:
  GpuStateValidityChecker(const ompl::base::SpaceInformationPtr &si, const GpuWorldPtr &world)
      : ompl::base::StateValidityChecker(si), world_(world) {
    requireRealVectorSpace(si->getStateSpace().get(), "GpuStateValidityChecker");
    if ((int)si->getStateDimension() != world->dof())
      throw ompl::Exception("GpuStateValidityChecker", "state dimension differs from the world's dof");
  }
  bool isValid(const ompl::base::State *state) const override {
    uint8_t v = 0;
    check(prgpu_is_valid_batch(world_->handle(), reals(si_, state), 1, &v), "GpuStateValidityChecker::isValid");
    return v != 0;
  }
  /// many states at once (n x dof row-major)
  void isValidBatch(const double *q, std::size_t n, std::vector<uint8_t> &valid) const {
    valid.resize(n);
    check(prgpu_is_valid_batch(world_->handle(), q, n, valid.data()), "GpuStateValidityChecker::isValidBatch");
  }
  const GpuWorldPtr &world() const { return world_; }

private:
  GpuWorldPtr world_;
};

class GpuMotionValidator : public ompl::base::MotionValidator {
public:
  GpuMotionValidator(const ompl::base::SpaceInformationPtr &si, const GpuWorldPtr &world)
      : ompl::base::MotionValidator(si), world_(world) {
    requireRealVectorSpace(si->getStateSpace().get(), "GpuMotionValidator");
    if ((int)si->getStateDimension() != world->dof())
      throw ompl::Exception("GpuMotionValidator", "state dimension differs from the world's dof");
  }
  bool checkMotion(const ompl::base::State *s1, const ompl::base::State *s2) const override {
    uint8_t v = 0;
    check(prgpu_check_motion_batch(world_->handle(), reals(si_, s1), reals(si_, s2), 1,
                                   si_->getStateSpace()->getLongestValidSegmentLength(),
                                   PRGPU_MODE_OMPL_DISCRETE, 0, &v),
          "GpuMotionValidator::checkMotion");
    (v ? valid_ : invalid_)++;
    return v != 0;
  }
  /// isValid(s1) && checkMotion(s1, s2) in one launch: the goal-tree form of RRTConnect.cpp:198
  bool checkMotionWithFirst(const ompl::base::State *s1, const ompl::base::State *s2) const {
    uint8_t v = 0;
    check(prgpu_check_motion_batch(world_->handle(), reals(si_, s1), reals(si_, s2), 1,
                                   si_->getStateSpace()->getLongestValidSegmentLength(),
                                   PRGPU_MODE_OMPL_DISCRETE, 1, &v),
          "GpuMotionValidator::checkMotionWithFirst");
    (v ? valid_ : invalid_)++;
    return v != 0;
  }
  bool checkMotion(const ompl::base::State *s1, const ompl::base::State *s2,
                   std::pair<ompl::base::State *, double> &lastValid) const override {
    // all interior states j/nd and s2 are tested on the device in one batch; the first invalid
    // one decides lastValid exactly as DiscreteMotionValidator's sequential scan would
    const ompl::base::StateSpace *sp = si_->getStateSpace().get();
    const int nd = (int)sp->validSegmentCount(s1, s2);
    const unsigned dim = sp->getDimension();
    std::vector<double> q((std::size_t)std::max(nd, 1) * dim);
    ompl::base::State *t = si_->allocState();
    int n = 0;
    for (int j = 1; j < nd; ++j, ++n) {
      sp->interpolate(s1, s2, (double)j / (double)nd, t);
      std::copy(reals(si_, t), reals(si_, t) + dim, q.begin() + (std::size_t)n * dim);
    }
    std::copy(reals(si_, s2), reals(si_, s2) + dim, q.begin() + (std::size_t)n * dim);
    ++n;
    si_->freeState(t);
    std::vector<uint8_t> v(n);
    check(prgpu_is_valid_batch(world_->handle(), q.data(), n, v.data()), "GpuMotionValidator::checkMotion");
    for (int i = 0; i < n; ++i)
      if (!v[i]) {
        const int j = i + 1; // index of the first invalid state (nd for s2 itself)
        lastValid.second = (double)(std::min(j, std::max(nd, 1)) - 1) / (double)std::max(nd, 1);
        if (lastValid.first)
          sp->interpolate(s1, s2, lastValid.second, lastValid.first);
        invalid_++;
        return false;
      }
    valid_++;
    return true;
  }
  /// many edges at once (n x dof row-major each)
  void checkMotionBatch(const double *from, const double *to, std::size_t n, std::vector<uint8_t> &valid) const {
    valid.resize(n);
    check(prgpu_check_motion_batch(world_->handle(), from, to, n,
                                   si_->getStateSpace()->getLongestValidSegmentLength(),
                                   PRGPU_MODE_OMPL_DISCRETE, 0, valid.data()),
          "GpuMotionValidator::checkMotionBatch");
  }
  const GpuWorldPtr &world() const { return world_; }

private:
  GpuWorldPtr world_;
};
} // namespace pr_gpu
#endif
