// Microbenchmark 3: L1-miss (L2-hit) read throughput per SM as a function of the contiguous chunk size a group of
// lanes reads at a random position of a 32 MB window.  LDG.128 per lane; chunk = 16 B x lanes-per-chunk.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ float4 ldg128(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
constexpr int ITERS = 600;
template <int LPC, int UNR, int KIND>  // KIND 0 nc random, 1 plain ld random, 2 ld.cg random, 3 nc sequential; lanes per chunk: 2 -> 32 B, 4 -> 64 B, 8 -> 128 B, 16 -> 256 B, 32 -> 512 B
__global__ void __launch_bounds__(512) k(const float* __restrict__ buf, long long nfloats, float* out, long long* cyc) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunk = lane / LPC, inl = lane % LPC;
  float acc = 0.f;
  unsigned h0 = ((blockIdx.x * 16 + warp) * 64 + chunk) * 2654435761u + 12345u;
  const unsigned cmask = (unsigned)(nfloats / (LPC * 4)) - 1;  // window is a power of two
  const unsigned gw = blockIdx.x * 16 + warp;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    float4 r[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      h0 = h0 * 1664525u + 1013904223u;
      long long o = (long long)((h0 >> 7) & cmask) * (LPC * 4);  // chunk-aligned
      if (KIND == 3) o = (long long)(((gw * (ITERS * UNR) + it * UNR + u) * 128u) & (unsigned)(nfloats - 1));
      const float* p = buf + o + inl * 4;
      if (KIND == 0 || KIND == 3) r[u] = ldg128(p);
      else if (KIND == 1) asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r[u].x), "=f"(r[u].y), "=f"(r[u].z), "=f"(r[u].w) : "l"(p));
      else asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r[u].x), "=f"(r[u].y), "=f"(r[u].z), "=f"(r[u].w) : "l"(p));
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) acc += (r[u].x + r[u].y) + (r[u].z + r[u].w);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
double avg_cycles(long long* cyc, int blocks) {
  static long long h[4096];
  cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
  double a = 0; for (int i = 0; i < blocks; ++i) a += h[i]; return a / blocks;
}
template <int LPC, int UNR, int KIND> void run(const float* buf, long long nf, float* out, long long* cyc, int blocks) {
  k<LPC, UNR, KIND><<<blocks, 512>>>(buf, nf, out, cyc); cudaDeviceSynchronize();
  k<LPC, UNR, KIND><<<blocks, 512>>>(buf, nf, out, cyc); cudaDeviceSynchronize();
  double c = avg_cycles(cyc, blocks);
  double bytes = 2.0 * 16 * ITERS * UNR * 512;
  printf("kind %d chunk %4d B, %d loads in flight/thread: %6.1f B/clk/SM  (%5.1f clk per 512-B request)  %s\n", KIND, LPC * 16, UNR, bytes / c,
         c / (2.0 * 16 * ITERS * UNR), cudaGetErrorString(cudaGetLastError()));
}
int main() {
  const int blocks = 148 * 2; const long long nf = 8ll << 20;
  float *buf, *out; long long* cyc;
  cudaMalloc(&buf, nf * 4); cudaMemset(buf, 0, nf * 4); cudaMalloc(&out, blocks * 512 * 4); cudaMalloc(&cyc, blocks * 8);
  run<2, 4, 0>(buf, nf, out, cyc, blocks); run<4, 4, 0>(buf, nf, out, cyc, blocks); run<8, 4, 0>(buf, nf, out, cyc, blocks);
  run<16, 4, 0>(buf, nf, out, cyc, blocks); run<32, 4, 0>(buf, nf, out, cyc, blocks);
  run<4, 8, 0>(buf, nf, out, cyc, blocks); run<4, 2, 0>(buf, nf, out, cyc, blocks); run<4, 1, 0>(buf, nf, out, cyc, blocks);
  run<4, 4, 2>(buf, nf, out, cyc, blocks);
  run<32, 4, 3>(buf, nf, out, cyc, blocks); run<32, 8, 3>(buf, nf, out, cyc, blocks);
  return 0;
}
