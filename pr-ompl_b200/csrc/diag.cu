// Measurement aids behind the C ABI (include/prgpu.h, prgpu_diag_*): the launch counter and the live
// micro-benchmark of the warp-level TF32 MMA issue rate that K1's scan is measured against.
#include <algorithm>
#include "capi_internal.cuh"
#include "kdtree.cuh"

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

int prgpu_diag_kd_leaf_capacity(uint64_t n, int *slots, int *levels) {
  if (!slots || !levels)
    return prgpu::fail(PRGPU_EINVAL, "diag_kd_leaf_capacity: NULL argument");
  if (n < 128 || n >= 0xFFFFFFFFull)
    return prgpu::fail(PRGPU_EINVAL, "diag_kd_leaf_capacity: the index needs 128 .. 2^32-2 nodes");
  int ppl = 0, L = 0;
  prgpu::kd_leaf_capacity((uint32_t)n, &ppl, &L);
  *slots = 32 * ppl;
  *levels = L;
  return PRGPU_OK;
}
uint64_t prgpu_diag_launch_count(void) { return g_launch_count; }

int prgpu_diag_knn_scan(prgpu_knn_t *h, const double *queries, uint32_t nq, float *thresh, uint32_t cap, uint32_t *counts,
                        uint32_t *idx, float *D, float *key32, double *eps) {
  if (!h || !queries || !thresh || !counts || !idx || !D || !key32 || !eps || nq == 0 || cap == 0)
    return fail(PRGPU_EINVAL, "diag_knn_scan: NULL argument");
  PRGPU_CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  DevBuf dq, dt, dc, di, dk, d32, de;
  auto done = [&](int rc) {
    cudaStreamSynchronize(s);
    for (DevBuf *b : {&dq, &dt, &dc, &di, &dk, &d32, &de})
      b->release();
    return rc;
  };
  const size_t cells = (size_t)nq * cap;
  int rc;
  if ((rc = dq.reserve((size_t)nq * h->dim * 8)) != PRGPU_OK || (rc = dt.reserve((size_t)nq * 4)) != PRGPU_OK ||
      (rc = dc.reserve((size_t)nq * 4)) != PRGPU_OK || (rc = di.reserve(cells * 4)) != PRGPU_OK ||
      (rc = dk.reserve(cells * 4)) != PRGPU_OK || (rc = d32.reserve(cells * 4)) != PRGPU_OK ||
      (rc = de.reserve((size_t)nq * 16)) != PRGPU_OK)
    return done(rc);
  cudaMemcpyAsync(dq.p, queries, (size_t)nq * h->dim * 8, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(dt.p, thresh, (size_t)nq * 4, cudaMemcpyHostToDevice, s);
  rc = knn_diag_scan(h->dim, h->view(), dq.as<double>(), nq, dt.as<float>(), cap, dc.as<uint32_t>(), di.as<uint32_t>(),
                     dk.as<float>(), d32.as<float>(), de.as<double>(), s);
  if (rc != PRGPU_OK)
    return done(rc);
  cudaMemcpyAsync(thresh, dt.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(counts, dc.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(idx, di.p, cells * 4, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(D, dk.p, cells * 4, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(key32, d32.p, cells * 4, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(eps, de.p, (size_t)nq * 16, cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess)
    return done(fail(PRGPU_ECUDA, std::string("diag_knn_scan: ") + cudaGetErrorString(e)));
  return done(PRGPU_OK);
}

int prgpu_diag_fk_centre_error(prgpu_world_t *w, const double *states, uint64_t n, double *max_err, double *delta) {
  if (!w || (n && (!states || !max_err)))
    return fail(PRGPU_EINVAL, "diag_fk_centre_error: NULL argument");
  if (delta)
    *delta = (double)w->host.delta;
  if (n == 0)
    return PRGPU_OK;
  PRGPU_CUDA_TRY(cudaSetDevice(w->device));
  DevBuf dq, de;
  auto done = [&](int rc) {
    cudaStreamSynchronize(w->stream);
    dq.release();
    de.release();
    return rc;
  };
  int rc;
  if ((rc = dq.reserve((size_t)n * w->host.dof * 8)) != PRGPU_OK || (rc = de.reserve((size_t)n * 8)) != PRGPU_OK)
    return done(rc);
  cudaMemcpyAsync(dq.p, states, (size_t)n * w->host.dof * 8, cudaMemcpyHostToDevice, w->stream);
  if ((rc = diag_fk_centre_launch(w->host, w->dev, dq.as<double>(), n, de.as<double>(), w->stream)) != PRGPU_OK)
    return done(rc);
  cudaMemcpyAsync(max_err, de.p, (size_t)n * 8, cudaMemcpyDeviceToHost, w->stream);
  cudaError_t e = cudaStreamSynchronize(w->stream);
  if (e != cudaSuccess)
    return done(fail(PRGPU_ECUDA, std::string("diag_fk_centre_error: ") + cudaGetErrorString(e)));
  return done(PRGPU_OK);
}

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
