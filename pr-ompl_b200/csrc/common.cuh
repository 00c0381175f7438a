// Shared device/host helpers for the prgpu kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include "../../include/prgpu.h"

namespace prgpu {

// thread-local error text behind prgpu_last_error()
void set_error(const std::string &msg);
int fail(int code, const std::string &msg);
// cudaSetDevice + "is this an sm_100 part" check shared by every create call (capi.cu)
int select_device(int device, int *sm_count);

#define PRGPU_CUDA_TRY(expr)                                                                       \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::prgpu::fail(_e == cudaErrorMemoryAllocation ? PRGPU_ENOMEM : PRGPU_ECUDA,           \
                           std::string(#expr) + ": " + cudaGetErrorString(_e));                    \
  } while (0)

// number of kernels this library has launched in this process (prgpu_diag_launch_count): bench.py's
// gpu_launches is the difference of two readings around its timed region, not a constant
extern unsigned long long g_launch_count;
#define PRGPU_COUNT_LAUNCH(n) (::prgpu::g_launch_count += (n))

// optional CUDA-event bracket around a handle's dominant kernel (bench.py's live roofline timing)
struct KernelTimer {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  bool on = false, armed = false;
  int enable(bool want) {
    if (want && !e0) {
      PRGPU_CUDA_TRY(cudaEventCreate(&e0));
      PRGPU_CUDA_TRY(cudaEventCreate(&e1));
    }
    on = want;
    armed = false;
    return PRGPU_OK;
  }
  void begin(cudaStream_t s) {
    if (on)
      cudaEventRecord(e0, s);
  }
  void end(cudaStream_t s) {
    if (on) {
      cudaEventRecord(e1, s);
      armed = true;
    }
  }
  int read(float *ms) {
    if (!on || !armed)
      return ::prgpu::fail(PRGPU_EINVAL, "profiling not enabled or no kernel launched yet");
    PRGPU_CUDA_TRY(cudaEventSynchronize(e1));
    PRGPU_CUDA_TRY(cudaEventElapsedTime(ms, e0, e1));
    return PRGPU_OK;
  }
  void destroy() {
    if (e0)
      cudaEventDestroy(e0);
    if (e1)
      cudaEventDestroy(e1);
    e0 = e1 = nullptr;
  }
};

inline uint64_t round_up(uint64_t x, uint64_t m) { return (x + m - 1) / m * m; }

// ---- bit-exact double arithmetic (no FMA contraction), matching the x86-64 reference ----
// RealVectorStateSpace::distance: sequential sum of (a-b)^2.
template <int DIM> __device__ __forceinline__ double sqdist_exact(const double *a, const double *b) {
  double diff = __dsub_rn(a[0], b[0]);
  double s = __dmul_rn(diff, diff);
#pragma unroll
  for (int i = 1; i < DIM; ++i) {
    diff = __dsub_rn(a[i], b[i]);
    s = __dadd_rn(s, __dmul_rn(diff, diff));
  }
  return s;
}
__device__ __forceinline__ double sqdist_exact_rt(const double *a, const double *b, int dim) {
  double s = 0.0;
  for (int i = 0; i < dim; ++i) {
    double diff = __dsub_rn(a[i], b[i]);
    s = __dadd_rn(s, __dmul_rn(diff, diff));
  }
  return s;
}
// RealVectorStateSpace::interpolate (Appendix B.2): from + (to - from) * t
__device__ __forceinline__ double lerp_exact(double from, double to, double t) {
  return __dadd_rn(from, __dmul_rn(__dsub_rn(to, from), t));
}

// ---- mbarrier + 1-D bulk async copy (TMA engine; SASS: UBLKCP) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile("{\n"
               ".reg .pred p;\n"
               "WAIT_%=:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra DONE_%=;\n"
               "bra WAIT_%=;\n"
               "DONE_%=:\n"
               "}" ::"r"(smem_u32(bar)),
               "r"(parity)
               : "memory");
}
// global -> shared bulk copy, completion signalled on `bar` (bytes multiple of 16, 16B aligned)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

} // namespace prgpu
