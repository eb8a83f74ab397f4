// Microbenchmark 2: LSU cost of (a) sparse 64-byte taps that miss L1 and hit L2, (b) the sweep's store patterns.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lsu_cost lsu_cost.cu && ./lsu_cost
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ float4 ldg128(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st_na(float* p, float v) { asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void st_na4(float* p, float4 v) { asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }

constexpr int ITERS = 1000;

// MODE 0: taps from an L1-resident window (48 KB)   MODE 1: taps from an L2-resident window (32 MB, random-ish walk)
template <int MODE>
__global__ void __launch_bounds__(512) k_ld(const float* __restrict__ buf, long long nfloats, float* out, long long* cyc) {
  const int lane = threadIdx.x & 31, quad = lane >> 2, q = lane & 3, warp = threadIdx.x >> 5;
  float acc = 0.f;
  const int base = ((int)(1.6f * quad)) * 16 + q * 4;
  long long t0 = clock64();
  unsigned long long h = (blockIdx.x * 16 + warp) * 0x9E3779B97F4A7C15ull + 12345;
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      h = h * 6364136223846793005ull + 1442695040888963407ull;
      long long o = MODE == 0 ? (long long)((h >> 33) % 700) * 16 : (long long)((h >> 20) % (unsigned long long)(nfloats / 16 - 64)) * 16;
      float4 r = ldg128(buf + o + base);
      acc += (r.x + r.y) + (r.z + r.w);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// Store patterns for one warp-iteration = 8 pixels x 16 channels (512 B) into a [16][N] plane layout.
// MODE 0: 4 x STG.32 per lane (lane = (px, q), channels 4q..4q+3)        -> 4 requests, each 4 planes x 32 B
// MODE 1: 1 x STG.128 per lane (lane = (ch = lane/2, half = lane%2) 4 px)  -> 1 request, 16 planes x 32 B
// MODE 2: full lines: warp writes 4 planes x 128 B with STG.128 (needs 32 px per warp; for comparison)
template <int MODE>
__global__ void __launch_bounds__(512) k_st(float* __restrict__ vol, long long plane, long long* cyc) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long wbase = ((long long)(blockIdx.x * 16 + warp)) * (ITERS * 32);
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) {
      float* p = vol + (long long)((lane & 3) * 4) * plane + wbase + it * 8 + (lane >> 2);
#pragma unroll
      for (int i = 0; i < 4; ++i) st_na(p + i * plane, 1.0f);
    } else if (MODE == 1) {
      float* p = vol + (long long)(lane >> 1) * plane + wbase + it * 8 + (lane & 1) * 4;
      st_na4(p, make_float4(1.f, 1.f, 1.f, 1.f));
    } else {
      // 32 px x 4 planes per instruction; 4 instructions cover 16 planes; every 4th iteration only (same bytes/iter avg)
      if ((it & 3) == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float* p = vol + (long long)(i * 4 + (lane >> 3)) * plane + wbase + it * 8 + (lane & 7) * 4;
          st_na4(p, make_float4(1.f, 1.f, 1.f, 1.f));
        }
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

double avg_cycles(long long* cyc, int blocks) {
  static long long h[4096];
  cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
  double a = 0;
  for (int i = 0; i < blocks; ++i) a += h[i];
  return a / blocks;
}

int main() {
  const int blocks = 148 * 2;
  const long long nf = 8ll << 20;  // 32 MB of floats (L2-resident)
  float *buf, *out, *vol; long long* cyc;
  cudaMalloc(&buf, nf * 4); cudaMemset(buf, 0, nf * 4);
  cudaMalloc(&out, blocks * 512 * 4); cudaMalloc(&cyc, blocks * 8);
  const long long plane = (long long)blocks * 16 * ITERS * 32 + 1024;
  cudaMalloc(&vol, plane * 16 * 4);
  for (int rep = 0; rep < 2; ++rep) {
    k_ld<0><<<blocks, 512>>>(buf, nf, out, cyc); cudaDeviceSynchronize();
    if (rep) printf("taps, L1-resident window   : %6.2f clk/request/SM\n", avg_cycles(cyc, blocks) / (2.0 * 16 * ITERS * 4));
    k_ld<1><<<blocks, 512>>>(buf, nf, out, cyc); cudaDeviceSynchronize();
    if (rep) printf("taps, L2-resident (L1 miss): %6.2f clk/request/SM\n", avg_cycles(cyc, blocks) / (2.0 * 16 * ITERS * 4));
    k_st<0><<<blocks, 512>>>(vol, plane, cyc); cudaDeviceSynchronize();
    if (rep) printf("store 4 x STG.32 (4 planes x 32B each): %6.2f clk per 512 B/SM\n", avg_cycles(cyc, blocks) / (2.0 * 16 * ITERS));
    k_st<1><<<blocks, 512>>>(vol, plane, cyc); cudaDeviceSynchronize();
    if (rep) printf("store 1 x STG.128 (16 planes x 32B)   : %6.2f clk per 512 B/SM\n", avg_cycles(cyc, blocks) / (2.0 * 16 * ITERS));
    k_st<2><<<blocks, 512>>>(vol, plane, cyc); cudaDeviceSynchronize();
    if (rep) printf("store full lines STG.128 (4 x 128B)   : %6.2f clk per 512 B/SM\n", avg_cycles(cyc, blocks) / (2.0 * 16 * ITERS));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
