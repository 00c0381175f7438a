// ompl_min: ompl::RNG.  Upstream seeds every instance from a global seed
// generator in construction order; parity tests therefore never rely on it
// and inject a deterministic StateSampler instead.
#ifndef OMPL_MIN_UTIL_RANDOMNUMBERS_H
#define OMPL_MIN_UTIL_RANDOMNUMBERS_H
#include <cmath>
#include <cstdint>
#include <random>
namespace ompl {
class RNG {
public:
  RNG() : gen_(nextSeed()) {}
  explicit RNG(std::uint_fast32_t seed) : gen_(seed) {}
  static void setSeed(std::uint_fast32_t seed) { seedCounter() = seed; }
  static std::uint_fast32_t getSeed() { return seedCounter(); }
  double uniform01() { return std::generate_canonical<double, 53>(gen_); }
  double uniformReal(double lo, double hi) { return (hi - lo) * uniform01() + lo; }
  int uniformInt(int lo, int hi) {
    int r = (int)std::floor(uniformReal((double)lo, (double)(hi) + 1.0));
    return r > hi ? hi : r;
  }
  bool uniformBool() { return uniform01() <= 0.5; }
  double gaussian01() { return normal_(gen_); }
  double gaussian(double mean, double stddev) { return gaussian01() * stddev + mean; }

private:
  static std::uint_fast32_t &seedCounter() {
    static std::uint_fast32_t s = 1;
    return s;
  }
  static std::uint_fast32_t nextSeed() { return seedCounter()++ * 2654435761u + 12345u; }
  std::mt19937 gen_;
  std::normal_distribution<double> normal_{0.0, 1.0};
};
} // namespace ompl
#endif
