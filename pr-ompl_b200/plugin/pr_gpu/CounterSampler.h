// pr_gpu::CounterSampler -- ompl::base::StateSampler drawing from the counter-based stream
// u(seed, query, iteration, dim) (csrc/sample_stream.cuh), so a host planner and the batched
// device planner (prgpu_rrtc_batch_*) see the same samples for the same (seed, query).
#ifndef PR_GPU_COUNTERSAMPLER_H
#define PR_GPU_COUNTERSAMPLER_H
#include <cstdint>
#include <ompl/base/spaces/RealVectorStateSpace.h>
namespace pr_gpu {
struct SampleCounter {
  std::uint64_t seed = 0, query = 0;
  std::uint32_t count = 0;
};
inline std::uint64_t streamMix64(std::uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
inline double streamUniform(std::uint64_t seed, std::uint64_t query, std::uint64_t it, std::uint64_t dim,
                            double low, double high) {
  const std::uint64_t g = 0x9E3779B97F4A7C15ULL;
  std::uint64_t x = streamMix64(seed + g);
  x = streamMix64(x ^ (query + g));
  x = streamMix64(x ^ (it + g));
  x = streamMix64(x ^ (dim + g));
  const double u = (double)(x >> 11) * (1.0 / 9007199254740992.0);
  volatile double span = high - low; // each operation rounded separately (no contraction)
  volatile double prod = span * u;
  return low + prod;
}
class CounterSampler : public ompl::base::StateSampler {
public:
  CounterSampler(const ompl::base::StateSpace *space, SampleCounter *counter)
      : ompl::base::StateSampler(space), c_(counter) {}
  void sampleUniform(ompl::base::State *state) override {
    const auto *sp = static_cast<const ompl::base::RealVectorStateSpace *>(space_);
    const ompl::base::RealVectorBounds &b = sp->getBounds();
    double *v = state->as<ompl::base::RealVectorStateSpace::StateType>()->values;
    for (unsigned int i = 0; i < sp->getDimension(); ++i)
      v[i] = streamUniform(c_->seed, c_->query, c_->count, i, b.low[i], b.high[i]);
    c_->count++;
  }
  void sampleUniformNear(ompl::base::State *, const ompl::base::State *, double) override {
    throw ompl::Exception("CounterSampler", "sampleUniformNear is not provided");
  }
  void sampleGaussian(ompl::base::State *, const ompl::base::State *, double) override {
    throw ompl::Exception("CounterSampler", "sampleGaussian is not provided");
  }

private:
  SampleCounter *c_;
};
} // namespace pr_gpu
#endif
