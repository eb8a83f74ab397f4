// Microbenchmark 6 (round 2): data-pipe wavefronts of a warp-wide 16-byte tap request (eight 64-byte chunks) as a
// function of (a) which bank halves the two chunks of a quarter-warp fall into and (b) whether the chunks hit L1.
// Run under ncu and read l1tex__data_pipe_lsu_wavefronts / l1tex__t_requests / sectors per launch:
//   ./tap_miss            -> prints clk/request for each (window, pattern)
//   kernels are named k_win<KB>_pat<P> through the launch order printed on stdout
// Window 44 KB: L1-resident after warm-up.  Window 32 MB: every chunk misses L1 and hits L2.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// pattern 0: quads 2j / 2j+1 opposite parity (pair-opp)  1: all even (pair-same)  2: random parity
// pattern 3: the two chunks of a quarter-warp are the two halves of ONE 128-byte line
__global__ void __launch_bounds__(512, 2) k_taps(const float* __restrict__ buf, uint32_t winPx, int pattern, int iters, float* out,
                                                 long long* cyc) {
  const int lane = threadIdx.x & 31, quad = lane >> 2, q = lane & 3, warp = threadIdx.x >> 5;
  float acc = 0.f;
  const uint32_t wid = (blockIdx.x * 16 + warp);
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    float4 r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      uint32_t h = hash32((wid * 4099u + it) * 32u + u * 8u + quad);
      uint32_t pix = h % (winPx - 2);
      if (pattern == 0) pix = (pix & ~1u) | (quad & 1u);
      else if (pattern == 1) pix = pix & ~1u;
      else if (pattern == 3) { h = hash32((wid * 4099u + it) * 32u + u * 8u + (quad >> 1)); pix = ((h % (winPx - 2)) & ~1u) | (quad & 1u); }
      r[u] = __ldg(reinterpret_cast<const float4*>(buf + (size_t)pix * 16 + q * 4));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += (r[u].x + r[u].y) + (r[u].z + r[u].w);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  const size_t bigPx = (32u << 20) / 64;  // 32 MB of 64-byte pixels
  float *buf, *out;
  long long* cyc;
  cudaMalloc(&buf, bigPx * 64);
  cudaMemset(buf, 0, bigPx * 64);
  const int grid = 296;
  cudaMalloc(&out, grid * 512 * 4);
  cudaMalloc(&cyc, grid * 8);
  long long h[296];
  const uint32_t wins[3] = {704, 2816, (uint32_t)bigPx};  // 44 KB, 176 KB (L1 capacity edge), 32 MB
  for (int w = 0; w < 3; ++w)
    for (int p = 0; p < 4; ++p) {
      const int iters = w == 2 ? 200 : 1000;
      k_taps<<<grid, 512>>>(buf, wins[w], p, iters, out, cyc);  // warm
      k_taps<<<grid, 512>>>(buf, wins[w], p, iters, out, cyc);
      cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double s = 0;
      for (int i = 0; i < grid; ++i) s += h[i];
      s /= grid;
      // per SM: 2 CTAs x 16 warps x iters x 4 requests
      printf("win %8u px  pattern %d  %.2f clk/request/SM\n", wins[w], p, s / (2.0 * 16 * iters * 4));
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
