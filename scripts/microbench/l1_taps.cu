// Microbenchmark: L1 / shared-memory delivery rate for the tap patterns of the sweep kernel (run on the GPU box).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1_taps l1_taps.cu && ./l1_taps
// Every pattern reads an L1-resident 48 KB window; a warp = 8 quads, quad k samples feature pixel floor(1.6 k) + shift
// (the 1.6 px output->feature stride of the G16V shape) or a dense run.  Reports useful bytes / clk / SM.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

struct __align__(32) F8 { float v[8]; };
__device__ __forceinline__ F8 ldg256(const float* p) {
  F8 r;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ldg128(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

constexpr int WIN = 48 * 1024 / 4;  // floats
constexpr int ITERS = 2000;

// mode 0: LDG.128 sparse (quad = 64 B of px floor(1.6k))   useful 512 B / request
// mode 1: LDG.256 pair   (quad = 128 B of px floor(1.6k), +1)  useful 1024 B
// mode 2: LDG.128 dense  (512 B contiguous)
// mode 3: LDG.256 dense  (1024 B contiguous)
// mode 4: LDG.128 pair via 8 lanes per pixel (4 pixels x 128 B)  useful 512 B
// mode 5: LDS.128 sparse from shared memory (same addresses as mode 0)
// mode 7: LDG.256, 2 lanes per pixel (16 px x 64 B, stride 1.6 px)   useful 1024 B
// mode 6: LDG.128 sparse, but 2 rows (half the warp row y0, other half row y1) -> 4 px x 2 rows
template <int MODE>
__global__ void __launch_bounds__(512) k(const float* __restrict__ buf, float* out, long long* cyc) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, quad = lane >> 2, q = lane & 3, warp = threadIdx.x >> 5;
  if (MODE == 5) { for (int i = threadIdx.x; i < WIN; i += blockDim.x) sm[i] = buf[i]; __syncthreads(); }
  float acc = 0.f;
  int base;
  if (MODE == 0 || MODE == 5) base = ((int)(1.6f * quad)) * 16 + q * 4;
  else if (MODE == 1) base = ((int)(1.6f * quad)) * 16 + q * 8;
  else if (MODE == 2) base = lane * 4;
  else if (MODE == 3) base = lane * 8;
  else if (MODE == 4) base = ((int)(1.6f * (lane >> 3))) * 16 + (lane & 7) * 4;
  else if (MODE == 7) base = ((int)(1.6f * (lane >> 1))) * 16 + (lane & 1) * 8;
  else if (MODE == 10) base = quad * 8 + q * 2;         // bf16-staged: quad = 32 B of px k (stride 1 px), LDG.64
  else if (MODE == 11) base = quad * 16 + q * 4;        // fp32, stride 1 px (the measured G16V step), LDG.128
  else if (MODE == 12) base = (quad * 3) * 8 + q * 2;   // bf16-staged, stride 3 px (no two quads share a sector)
  else if (MODE == 8) base = (((int)(1.6f * quad)) & ~1) * 16 + q * 8;  // LDG.256, quad = one ALIGNED 128-B line
  else if (MODE == 9) base = (((int)(1.6f * quad)) & ~1) * 16 + q * 8;  // mode 8 + LDG.128 from the next line for odd quads
  else base = ((int)(1.6f * (quad & 3))) * 16 + q * 4 + (quad >> 2) * 514 * 16;
  const int rowspan = (MODE == 6) ? 2 : 4;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    // a different window position per warp and iteration (multiple of one pixel = 16 floats), 4 independent loads
    int o = (warp * 37 + it * 13) * 16;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int idx = (base + o + u * 16 * 29) % (WIN - rowspan * 514 * 16 > 0 ? WIN - 64 : WIN - 64);
      idx &= ~3;
      if (MODE == 8 || MODE == 9) idx = ((idx >> 5) << 5) + q * 8;
      if (MODE == 9) { idx &= ~7; F8 r = ldg256(buf + idx); acc += (r.v[0] + r.v[1]) + (r.v[2] + r.v[3]) + (r.v[4] + r.v[5]) + (r.v[6] + r.v[7]);
        if (quad & 1) { float4 r2 = ldg128(buf + idx + 32); acc += (r2.x + r2.y) + (r2.z + r2.w); } }
      else if (MODE == 1 || MODE == 3 || MODE == 7 || MODE == 8) { idx &= ~7; F8 r = ldg256(buf + idx); acc += (r.v[0] + r.v[1]) + (r.v[2] + r.v[3]) + (r.v[4] + r.v[5]) + (r.v[6] + r.v[7]); }
      else if (MODE == 10 || MODE == 12) { idx = (base + o * 0 + ((warp * 37 + it * 13) * 8) + u * 8 * 29) % (WIN - 64); idx &= ~1;
        float2 r = __ldg(reinterpret_cast<const float2*>(buf + idx)); acc += r.x + r.y; }
      else if (MODE == 5) { float4 r = *reinterpret_cast<const float4*>(sm + idx); acc += (r.x + r.y) + (r.z + r.w); }
      else { float4 r = ldg128(buf + idx); acc += (r.x + r.y) + (r.z + r.w); }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, int bytes_per_lane, const float* buf, float* out, long long* cyc, int blocks) {
  size_t smem = MODE == 5 ? WIN * 4 : 0;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<MODE><<<blocks, 512, smem>>>(buf, out, cyc);
  k<MODE><<<blocks, 512, smem>>>(buf, out, cyc);
  cudaDeviceSynchronize();
  long long h[1024];
  cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0; double avg = 0;
  for (int i = 0; i < blocks; ++i) { if (h[i] > mx) mx = h[i]; avg += h[i]; }
  avg /= blocks;
  // per SM: 2 blocks x 16 warps x ITERS x 4 requests x 32 lanes x bytes
  double bytes_sm = 2.0 * 16 * ITERS * 4 * 32 * bytes_per_lane;
  printf("%-44s avg %9.0f clk  -> %6.1f useful B/clk/SM   %5.2f clk/request  (%s)\n", name, avg, bytes_sm / avg,
         avg / (2.0 * 16 * ITERS * 4), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float *buf, *out; long long* cyc;
  int blocks = 148 * 2;
  cudaMalloc(&buf, WIN * 4 + 4096); cudaMemset(buf, 0, WIN * 4 + 4096);
  cudaMalloc(&out, blocks * 512 * 4); cudaMalloc(&cyc, blocks * 8);
  run<0>("LDG.128 sparse (quad=64B, stride 1.6px)", 16, buf, out, cyc, blocks);
  run<1>("LDG.256 pair   (quad=128B, stride 1.6px)", 32, buf, out, cyc, blocks);
  run<2>("LDG.128 dense  (512B contiguous)", 16, buf, out, cyc, blocks);
  run<3>("LDG.256 dense  (1024B contiguous)", 32, buf, out, cyc, blocks);
  run<4>("LDG.128 pair, 8 lanes/px (4px x 128B)", 16, buf, out, cyc, blocks);
  run<5>("LDS.128 sparse (smem, quad=64B)", 16, buf, out, cyc, blocks);
  run<6>("LDG.128 sparse, 4px x 2 rows", 16, buf, out, cyc, blocks);
  run<7>("LDG.256, 2 lanes/px (16px x 64B, stride 1.6px)", 32, buf, out, cyc, blocks);
  run<8>("LDG.256 aligned pair line per quad (8 lines, 1024B)", 32, buf, out, cyc, blocks);
  run<9>("mode 8 + LDG.128 next line for odd quads (1024B useful+256)", 32, buf, out, cyc, blocks);
  run<11>("LDG.128 sparse (quad=64B, stride 1px)", 16, buf, out, cyc, blocks);
  run<10>("LDG.64 bf16 chunk (quad=32B, stride 1px)", 8, buf, out, cyc, blocks);
  run<12>("LDG.64 bf16 chunk (quad=32B, stride 3px)", 8, buf, out, cyc, blocks);
  return 0;
}
