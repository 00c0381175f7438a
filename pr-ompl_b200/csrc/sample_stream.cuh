// Counter-based uniform sample stream used in place of OMPL's RealVectorStateSampler: sample
// `it` of query `query` in dimension `dim` depends only on (seed, query, it, dim), so many
// independent planners can draw their samples on the device with no shared RNG state, and a
// CPU run with the same stream (an injected ompl::base::StateSampler) is comparable node for
// node.
//   x = mix(mix(mix(mix(seed+g) ^ (query+g)) ^ (it+g)) ^ (dim+g)),  g = 0x9E3779B97F4A7C15
//   u = (x >> 11) * 2^-53,   q = low + (high - low) * u   (each operation rounded, no FMA)
#pragma once
#include <cstdint>
namespace prgpu {
__host__ __device__ __forceinline__ uint64_t stream_mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ double stream_u01(uint64_t seed, uint64_t query, uint64_t it,
                                                      uint64_t dim) {
  const uint64_t g = 0x9E3779B97F4A7C15ULL;
  uint64_t x = stream_mix64(seed + g);
  x = stream_mix64(x ^ (query + g));
  x = stream_mix64(x ^ (it + g));
  x = stream_mix64(x ^ (dim + g));
  return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ double stream_uniform(uint64_t seed, uint64_t query, uint64_t it, uint64_t dim,
                                                 double low, double high) {
  return __dadd_rn(low, __dmul_rn(__dsub_rn(high, low), stream_u01(seed, query, it, dim)));
}
} // namespace prgpu
