// Measurement aids behind the C ABI (include/prgpu.h, prgpu_diag_*): the launch counter and the live
// micro-benchmark of the warp-level TF32 MMA issue rate that K1's scan is measured against.
#include <algorithm>
#include "common.cuh"

namespace prgpu {
unsigned long long g_launch_count = 0;

// 8 independent accumulator fragments per warp, mma.sync.m16n8k8 TF32 back to back: the issue peak of the
// instruction knn_collect_kernel is built on (2048 flop per warp-level MMA)
__global__ void diag_mma_tf32_kernel(float *out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 4; ++j)
      c[i][j] = 0.f;
  unsigned a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = threadIdx.x ^ 5, b1 = 11;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
                   "{%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0;
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 4; ++j)
      s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
} // namespace prgpu

using namespace prgpu;

extern "C" {
uint64_t prgpu_diag_launch_count(void) { return g_launch_count; }

int prgpu_diag_mma_tf32_peak(int device, double *tflops) {
  if (!tflops)
    return fail(PRGPU_EINVAL, "diag_mma_tf32_peak: NULL argument");
  int sm = 0;
  int rc = select_device(device, &sm);
  if (rc != PRGPU_OK)
    return rc;
  const int warps_per_block = 4, blocks = sm * 4, iters = 20000; // 16 warps per SM
  float *out = nullptr;
  PRGPU_CUDA_TRY(cudaMalloc(&out, (size_t)blocks * warps_per_block * 32 * sizeof(float)));
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  diag_mma_tf32_kernel<<<blocks, warps_per_block * 32>>>(out, 100); PRGPU_COUNT_LAUNCH(1);
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    diag_mma_tf32_kernel<<<blocks, warps_per_block * 32>>>(out, iters); PRGPU_COUNT_LAUNCH(1);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.f;
    if (e == cudaSuccess)
      e = cudaEventElapsedTime(&ms, e0, e1);
    if (e != cudaSuccess) {
      cudaFree(out);
      return fail(PRGPU_ECUDA, std::string("diag_mma_tf32_peak: ") + cudaGetErrorString(e));
    }
    const double mmas = (double)blocks * warps_per_block * iters * 8;
    best = std::max(best, mmas * 2048.0 / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return PRGPU_OK;
}
}
