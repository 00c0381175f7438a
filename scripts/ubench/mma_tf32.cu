// Microbenchmark: issue rate of the warp-level mma.sync m16n8k8 TF32 on sm_100a (development aid).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float *out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  unsigned a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = threadIdx.x ^ 5, b1 = 11;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float *out; cudaMalloc(&out, 148 * 16 * 1024 * 4);
  for (int warps = 4; warps <= 32; warps *= 2) {
    int blocks = 148 * (warps >= 8 ? warps / 8 * 2 : 1), threads = warps >= 8 ? 128 : warps * 32;
    // warps per SM = blocks/148 * threads/32
    int iters = 20000;
    k<<<blocks, threads>>>(out, 10);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double mmas = (double)blocks * (threads / 32) * iters * 8;
    printf("warps/SM=%d: %.3f ms, %.1f warp-MMA/us/SM, %.1f TFLOP/s tf32, %.1f outputs/ns/SM\n", blocks / 148 * threads / 32, ms,
           mmas / 148 / (ms * 1e3), mmas * 2048 / (ms * 1e-3) / 1e12, mmas * 128 / 148 / (ms * 1e6));
  }
  return 0;
}
