// Microbenchmark: what limits the collect loop?  Variants of {LDS-fed B/C, C != D, sign OR} around
// mma.sync m16n8k8 TF32 (development aid).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], float2 b, const float4 &c) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(__float_as_uint(b.x)), "r"(__float_as_uint(b.y)), "f"(c.x),
        "f"(c.y), "f"(c.z), "f"(c.w));
}
template <int MODE, int MT>
__global__ void __launch_bounds__(128) k(uint32_t *out, int iters) {
  __shared__ float recs[256 * 8];
  __shared__ float n2s[256 * 2];
  for (int i = threadIdx.x; i < 256 * 8; i += 128) recs[i] = 0.001f * i;
  for (int i = threadIdx.x; i < 256 * 2; i += 128) n2s[i] = 1.0f + i;
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  uint32_t a[MT][4];
  for (int mt = 0; mt < MT; ++mt) for (int j = 0; j < 4; ++j) a[mt][j] = __float_as_uint(0.5f + mt + j + lane);
  const float2 *R = reinterpret_cast<const float2 *>(recs) + g * 4 + t4;
  const float4 *N2 = reinterpret_cast<const float4 *>(n2s) + t4;
  uint32_t hist = 0;
  float2 b0 = R[0]; float4 c0 = N2[0];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int nt = 0; nt < 32; ++nt) {
      float2 b = b0; float4 c = c0;
      if (MODE & 1) { b = R[nt * 32]; c = N2[nt * 4]; }
      float d[MT][4];
      if (MODE & 4) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) mma(d[mt], a[mt], b, make_float4(0.f, 0.f, 0.f, 0.f));
      } else if (MODE & 8) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const float4 cc = N2[(nt * 4 + mt * 32) % 124];
          d[mt][0] = cc.x; d[mt][1] = cc.y; d[mt][2] = cc.z; d[mt][3] = cc.w;
          asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(d[mt][0]), "+f"(d[mt][1]), "+f"(d[mt][2]), "+f"(d[mt][3])
                       : "r"(a[mt][0]), "r"(a[mt][1]), "r"(a[mt][2]), "r"(a[mt][3]), "r"(__float_as_uint(b.x)), "r"(__float_as_uint(b.y)));
        }
      } else {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) mma(d[mt], a[mt], b, c);
      }
      uint32_t m = 0;
      if (MODE & 2) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
          m |= __float_as_uint(d[mt][0]) | __float_as_uint(d[mt][1]) | __float_as_uint(d[mt][2]) | __float_as_uint(d[mt][3]);
      } else {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) m ^= __float_as_uint(d[mt][0]);
      }
      hist = __funnelshift_l(m, hist, 1);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = hist;
}
template <int MODE, int MT> void run(uint32_t *out, int bps) {
  int iters = 2000;
  k<MODE, MT><<<148 * bps, 128>>>(out, 10);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE, MT><<<148 * bps, 128>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double mmas = (double)148 * bps * 4 * iters * 32 * MT;
  printf("mode=%d MT=%d blocks/SM=%d: %.3f ms, %.1f warp-MMA/us/SM (peak ~917)\n", MODE, MT, bps, ms, mmas / 148 / (ms * 1e3));
}
int main() {
  uint32_t *out; cudaMalloc(&out, 148 * 8 * 128 * 4);
  run<1, 4>(out, 8); run<3, 4>(out, 8); run<5, 4>(out, 8); run<7, 4>(out, 8); run<9, 4>(out, 8); run<11, 4>(out, 8);
  return 0;
}
