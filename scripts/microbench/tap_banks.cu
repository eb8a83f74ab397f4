// Microbenchmark 5 (round 2): what a warp-wide 16-byte tap request costs as a function of WHICH bank halves its eight
// 64-byte chunks fall into, from shared memory (LDS.128) and from an L1-resident global window (LDG.128), and what a
// TMA bulk fill of shared memory costs the LDS stream next to it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tap_banks tap_banks.cu && ./tap_banks
// A warp = 8 quads; quad k reads the 64 bytes of "pixel" pix_k (lane q reads bytes 16q..16q+15).  A 64-byte chunk
// covers 16 of the 32 banks, which half is decided by the pixel's parity.  Patterns:
//   0 pair-opp : quads 2j and 2j+1 (one quarter-warp phase) always have opposite parity          (expect 4 wavefronts)
//   1 pair-same: every quad even parity                                                         (expect 8)
//   2 random   : parity random per quad                                                         (expect ~6)
//   3 half-half: quads 0-3 even, 4-7 odd (balanced over the warp, but every phase conflicts)     (8 if phase-based)
//   4 stride1.6: pix_k = base + floor(1.6 k)  (the sweep's own pattern)
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int WIN_PX = 704;  // 44 KB window of 64-byte pixels
constexpr int ITERS = 2000;

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

__device__ __forceinline__ int pick(int pattern, uint32_t seed, int quad) {
  const uint32_t h = hash32(seed * 8u + (uint32_t)quad);
  int pix = (int)(h % (WIN_PX - 16));
  switch (pattern) {
    case 0: pix = (pix & ~1) | (quad & 1); break;
    case 1: pix = pix & ~1; break;
    case 2: break;
    case 3: pix = (pix & ~1) | ((quad >> 2) & 1); break;
    default: pix = (int)(hash32(seed) % (WIN_PX - 16)) + (int)(1.6f * quad); break;
  }
  return pix;
}

template <bool SMEM>
__global__ void __launch_bounds__(512, 2) k_taps(const float* __restrict__ buf, int pattern, float* out, long long* cyc) {
  extern __shared__ __align__(128) float sm[];
  const int lane = threadIdx.x & 31, quad = lane >> 2, q = lane & 3, warp = threadIdx.x >> 5;
  if (SMEM) {
    for (int i = threadIdx.x; i < WIN_PX * 16; i += blockDim.x) sm[i] = buf[i];
    __syncthreads();
  }
  const float* src = SMEM ? sm : buf;
  float acc = 0.f;
  // warm the L1 window
  if (!SMEM) for (int i = threadIdx.x; i < WIN_PX * 4; i += blockDim.x) acc += __ldg(reinterpret_cast<const float4*>(buf) + i).x;
  __syncthreads();
  const long long t0 = clock64();
  int pre[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) pre[u] = (pick(pattern, (uint32_t)(blockIdx.x * 16 + warp) * 4u + u, quad) % (WIN_PX / 2)) * 16 + q * 4;
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    const int sh = ((it * 26) % (WIN_PX / 2 - 16)) * 16;  // even pixel shift: parities unchanged
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float4 r;
      if (SMEM) r = *reinterpret_cast<const float4*>(src + pre[u] + sh);
      else r = __ldg(reinterpret_cast<const float4*>(src + pre[u] + sh));
      acc += (r.x + r.y) + (r.z + r.w);
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// address-generation only (same loop without the loads) to subtract
__global__ void __launch_bounds__(512, 2) k_addr(int pattern, float* out, long long* cyc) {
  const int lane = threadIdx.x & 31, quad = lane >> 2, q = lane & 3, warp = threadIdx.x >> 5;
  int acc = 0;
  const long long t0 = clock64();
  int pre[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) pre[u] = (pick(pattern, (uint32_t)(blockIdx.x * 16 + warp) * 4u + u, quad) % (WIN_PX / 2)) * 16 + q * 4;
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    const int sh = ((it * 26) % (WIN_PX / 2 - 16)) * 16;
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += (pre[u] + sh) ^ it;
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
}

// ---- TMA bulk fill next to an LDS stream -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// MODE 0: LDS stream only (15 warps, conflict-free pattern 0)   MODE 1: fill only   MODE 2: both
// The fill: warp 15 lane 0 keeps NBUF bulk copies of CHUNK bytes in flight from a large global buffer into a ring.
constexpr int CHUNK = 8192, NBUF = 4, FILL_ITERS = 1024;
template <int MODE>
__global__ void __launch_bounds__(512, 2) k_fill(const float* __restrict__ buf, const float* __restrict__ big, long long big_bytes,
                                                 float* out, long long* cyc, long long* cyc_fill) {
  extern __shared__ __align__(128) float sm[];
  float* ring = sm + WIN_PX * 16;
  __shared__ uint64_t bars[NBUF];
  __shared__ volatile int done;
  const int lane = threadIdx.x & 31, quad = lane >> 2, q = lane & 3, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < WIN_PX * 16; i += blockDim.x) sm[i] = buf[i];
  if (threadIdx.x == 0) {
    for (int i = 0; i < NBUF; ++i) mbar_init(&bars[i], 1);
    done = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  float acc = 0.f;
  const long long t0 = clock64();
  if (warp == 15) {
    if (MODE != 0 && lane == 0) {
      const char* src = reinterpret_cast<const char*>(big) + ((long long)blockIdx.x * FILL_ITERS * CHUNK) % (big_bytes - (long long)FILL_ITERS * CHUNK);
      for (int i = 0; i < NBUF; ++i) {
        mbar_expect(&bars[i], CHUNK);
        bulk_g2s(reinterpret_cast<char*>(ring) + i * CHUNK, src + (long long)i * CHUNK, CHUNK, &bars[i]);
      }
      for (int i = NBUF; i < FILL_ITERS + NBUF; ++i) {
        const int s = i % NBUF;
        mbar_wait(&bars[s], ((i / NBUF) - 1) & 1);
        if (i < FILL_ITERS) {
          mbar_expect(&bars[s], CHUNK);
          bulk_g2s(reinterpret_cast<char*>(ring) + s * CHUNK, src + (long long)i * CHUNK, CHUNK, &bars[s]);
        }
      }
      cyc_fill[blockIdx.x] = clock64() - t0;
      done = 1;
    }
  } else if (MODE != 1) {
    int pre[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) pre[u] = (pick(0, (uint32_t)(blockIdx.x * 16 + warp) * 4u + u, quad) % (WIN_PX / 2)) * 16 + q * 4;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
      const int sh = ((it * 26) % (WIN_PX / 2 - 16)) * 16;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 r = *reinterpret_cast<const float4*>(sm + pre[u] + sh);
        acc += (r.x + r.y) + (r.z + r.w);
      }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + ring[threadIdx.x];
}

static double avg_cycles(long long* cyc, int blocks) {
  static long long h[4096];
  cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
  double a = 0;
  for (int i = 0; i < blocks; ++i) a += h[i];
  return a / blocks;
}

int main() {
  const int blocks = 148 * 2;
  float *buf, *out, *big;
  long long *cyc, *cyc2;
  const long long big_bytes = 1ll << 30;
  cudaMalloc(&buf, WIN_PX * 64);
  cudaMemset(buf, 0, WIN_PX * 64);
  cudaMalloc(&big, big_bytes);
  cudaMemset(big, 0, big_bytes);
  cudaMalloc(&out, blocks * 512 * 4);
  cudaMalloc(&cyc, blocks * 8);
  cudaMalloc(&cyc2, blocks * 8);
  const char* names[5] = {"pair-opp ", "pair-same", "random   ", "half-half", "stride1.6"};
  const size_t smem = WIN_PX * 64;
  cudaFuncSetAttribute(k_taps<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const double nreq = 2.0 * 16 * ITERS * 4;  // requests per SM (2 CTAs x 16 warps)
  for (int p = 0; p < 5; ++p) {
    double a[3];
    for (int rep = 0; rep < 2; ++rep) {
      k_addr<<<blocks, 512>>>(p, out, cyc);
      cudaDeviceSynchronize();
      a[0] = avg_cycles(cyc, blocks) / nreq;
      k_taps<true><<<blocks, 512, smem>>>(buf, p, out, cyc);
      cudaDeviceSynchronize();
      a[1] = avg_cycles(cyc, blocks) / nreq;
      k_taps<false><<<blocks, 512>>>(buf, p, out, cyc);
      cudaDeviceSynchronize();
      a[2] = avg_cycles(cyc, blocks) / nreq;
    }
    printf("%s  addr-only %5.2f   LDS.128 %5.2f   LDG.128(L1-resident) %5.2f   clk/request/SM\n", names[p], a[0], a[1], a[2]);
  }
  const size_t smem2 = WIN_PX * 64 + NBUF * CHUNK;
  cudaFuncSetAttribute(k_fill<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
  cudaFuncSetAttribute(k_fill<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
  cudaFuncSetAttribute(k_fill<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
  const double nreq15 = 2.0 * 15 * ITERS * 4;
  for (int rep = 0; rep < 2; ++rep) {
    k_fill<0><<<blocks, 512, smem2>>>(buf, big, big_bytes, out, cyc, cyc2);
    cudaDeviceSynchronize();
    const double lds_only = avg_cycles(cyc, blocks);
    k_fill<1><<<blocks, 512, smem2>>>(buf, big, big_bytes, out, cyc, cyc2);
    cudaDeviceSynchronize();
    const double fill_only = avg_cycles(cyc2, blocks);
    k_fill<2><<<blocks, 512, smem2>>>(buf, big, big_bytes, out, cyc, cyc2);
    cudaDeviceSynchronize();
    const double both = avg_cycles(cyc, blocks), both_fill = avg_cycles(cyc2, blocks);
    if (rep)
      printf("LDS only %.2f clk/req/SM | bulk fill only %.1f B/clk/SM (%.0f clk per 8 KB x2 CTAs) | both: LDS stream %.0f clk total (%.2f clk/req), fill %.1f B/clk/SM\n",
             lds_only / nreq15, 2.0 * FILL_ITERS * CHUNK / fill_only, fill_only / FILL_ITERS, both, both / nreq15,
             2.0 * FILL_ITERS * CHUNK / both_fill);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
