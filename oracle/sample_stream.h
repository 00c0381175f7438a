// ORACLE / TEST INFRASTRUCTURE ONLY.
// Counter-based uniform sample stream injected into the planners in place of
// OMPL's RealVectorStateSampler (whose ompl::RNG seed depends on global
// construction order).  Sample `it` of
// query `query` in dimension `dim` depends only on (seed, query, it, dim):
//   x  = mix(seed, query, it, dim)        (64-bit integer hash, below)
//   u  = (x >> 11) * 2^-53                in [0,1)
//   q  = low + (high - low) * u           two roundings, no FMA
// The product restates the same stream independently in csrc/sample_stream.cuh.
#ifndef ORACLE_SAMPLE_STREAM_H
#define ORACLE_SAMPLE_STREAM_H
#include <cstdint>
namespace oracle {
inline std::uint64_t mix64(std::uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
inline double sampleU01(std::uint64_t seed, std::uint64_t query, std::uint64_t it, std::uint64_t dim) {
  std::uint64_t x = mix64(seed + 0x9E3779B97F4A7C15ULL);
  x = mix64(x ^ (query + 0x9E3779B97F4A7C15ULL));
  x = mix64(x ^ (it + 0x9E3779B97F4A7C15ULL));
  x = mix64(x ^ (dim + 0x9E3779B97F4A7C15ULL));
  return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}
inline double sampleUniform(std::uint64_t seed, std::uint64_t query, std::uint64_t it, std::uint64_t dim,
                            double low, double high) {
  volatile double span = high - low;
  volatile double prod = span * sampleU01(seed, query, it, dim);
  return low + prod;
}
} // namespace oracle
#endif
