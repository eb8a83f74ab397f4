// Microbenchmark 4: LSU cost of one warp-wide STG.32 as a function of how its 32 lanes spread over 128-byte lines.
// All stores stay inside a 32 MB window (L2-resident write-back), so DRAM bandwidth is not the limiter.
//   shape (L lines x S sectors): lanes [0, 8*S) of each line group write consecutive floats.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void st_na(float* p, float v) { asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
constexpr int ITERS = 2000;
template <int LINES>  // 32 lanes = LINES lines x (32/LINES) consecutive floats
__global__ void __launch_bounds__(512) k(float* __restrict__ buf, unsigned mask, long long* cyc) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = 32 / LINES;
  const int line = lane / per, inl = lane % per;
  unsigned h = (blockIdx.x * 16 + warp) * 2654435761u + 7u;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      h = h * 1664525u + 1013904223u;
      // each line group lands on its own random 128-byte line
      unsigned o = (((h >> 7) + line * 9973u) & mask) * 32u;
      st_na(buf + o + inl, 1.0f);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
double avg_cycles(long long* cyc, int blocks) {
  static long long h[4096];
  cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
  double a = 0; for (int i = 0; i < blocks; ++i) a += h[i]; return a / blocks;
}
template <int LINES> void run(float* buf, long long* cyc, int blocks) {
  const unsigned mask = (8u << 20) / 32 - 1;
  k<LINES><<<blocks, 512>>>(buf, mask, cyc); cudaDeviceSynchronize();
  k<LINES><<<blocks, 512>>>(buf, mask, cyc); cudaDeviceSynchronize();
  double c = avg_cycles(cyc, blocks);
  printf("STG.32: %2d lines x %d sectors per instruction: %6.2f clk/instr/SM   %s\n", LINES, 4 / LINES > 0 ? 4 / LINES : 1,
         c / (2.0 * 16 * ITERS * 4), cudaGetErrorString(cudaGetLastError()));
}
int main() {
  const int blocks = 148 * 2;
  float* buf; long long* cyc;
  cudaMalloc(&buf, (8ll << 20) * 4 + 4096); cudaMalloc(&cyc, blocks * 8);
  run<1>(buf, cyc, blocks); run<2>(buf, cyc, blocks); run<4>(buf, cyc, blocks); run<8>(buf, cyc, blocks); run<16>(buf, cyc, blocks); run<32>(buf, cyc, blocks);
  return 0;
}
