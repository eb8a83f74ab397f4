// softargmin_tma.cuh — the scale-factor-2 soft-argmin as a persistent, TMA-pipelined kernel (sm_100a).
//
// Replaces DistanceRegressorWithFixedCandidates.forward
// (dsta_mvs/model/distance_regressor/distance_regressor.py:51-80; configs/base_model.yaml:28-34: pre_interp,
// interp_scale_factor 2).  Same arithmetic, operation for operation, as softargmin_x2_kernel (softargmin.cuh) — the
// two kernels give identical bits — but a different data path:
//   * the output is cut into tiles of kSaRows x 32 blocks of 2x2 output pixels (one warp per block row, one lane per
//     block).  A tile reads a (kSaRows+1 rows x 40 columns x D candidates) box of the cost volume; ONE
//     cp.async.bulk.tensor.4d (TMA) issued by one thread brings the box into shared memory and completes an mbarrier.
//     The box starts 4 columns (16 bytes: the innermost TMA coordinate must be 16-byte aligned,
//     scripts/microbench/tma_probe.cu) and 1 row before the tile's first full block, so it may start at (-4, -1): TMA
//     fills what lies outside the tensor with zeros, and the clamped indices never read it.
//   * the kernel is persistent: gridDim.x = SMs x resident CTAs, a CTA walks tiles blockIdx.x, + gridDim.x, ... with a
//     ring of kSaStages boxes.  The box of tile i + kSaStages is requested as soon as every thread has interpolated
//     tile i out of its stage (one __syncthreads), so it travels while the CTA runs the softmax of tile i and the
//     whole of tile i + 1: loads and arithmetic overlap inside a CTA instead of only between the CTAs of an SM (the
//     one-tile-per-CTA version of this kernel ran its waves in lockstep — load, then compute — and took 47 us for
//     32 G16V frames against 16 us of issue time).
//   * a thread reads its 4 x D costs with LDS at compile-time offsets: no per-load address arithmetic (the
//     pointer-per-load kernel spends 349 of its 1264 instructions on addresses and issues 64 scalar LDG per thread).
//   * the two output columns of a block are processed as packed pairs (FMUL2 / FFMA2 / FADD2, round-to-nearest per
//     lane): interpolation, softmax numerator argument, sums and expectation — half the FP issue slots.
#pragma once
#include <cuda.h>

#include "softargmin.cuh"
#include "sweep.cuh"  // f32x2 helpers

namespace mvsgi {

constexpr int kSaRows = 3;         // rows of 2x2 blocks per tile (one warp each); G16V: 81 block rows = 27 tiles
constexpr int kSaCols = 32;        // columns of 2x2 blocks per tile (one lane each)
constexpr int kSaLead = 4;         // box columns before the first block column of the tile: keeps the box 16-byte aligned
constexpr int kSaBoxW = 40;        // box columns: 3 unused + 33 needed, rounded up to a multiple of 16 bytes
constexpr int kSaBoxH = kSaRows + 1;
constexpr int kSaThreads = kSaRows * 32;
constexpr int kSaStages = 2;       // boxes in flight per CTA

struct SaTiles {
  int tilesX, tilesY, total;  // tiles per row of tiles, per frame column of tiles, in the whole call
  int stepX, stepY, stepB;    // gridDim.x = (stepB * tilesY + stepY) * tilesX + stepX: one stride of the walk, decomposed
};

struct SaTile {
  int tx, ty, b;
};

// the tile gridDim.x further along the walk, without a division
__device__ __forceinline__ SaTile sa_next(SaTile c, const SaTiles& tl) {
  c.tx += tl.stepX;
  c.ty += tl.stepY;
  c.b += tl.stepB;
  if (c.tx >= tl.tilesX) { c.tx -= tl.tilesX; ++c.ty; }
  if (c.ty >= tl.tilesY) { c.ty -= tl.tilesY; ++c.b; }
  return c;
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait(uint32_t mb, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SA_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SA_DONE;\n"
      "bra SA_WAIT;\n"
      "SA_DONE:\n"
      "}\n" ::"r"(mb), "r"(parity)
      : "memory");
}

template <int DREG, bool FULL, bool NC, bool AM>
__global__ void __launch_bounds__(kSaThreads, NC ? 4 : (AM ? 6 : 8)) softargmin_x2_tma_kernel(const SoftKP kp, const __grid_constant__ CUtensorMap tmap,
                                                                       const SaTiles tl) {
  // a stage holds [d][row][col] as TMA writes the box (innermost coordinate first)
  constexpr int kStageFloats = DREG * kSaBoxH * kSaBoxW;
  __shared__ __align__(128) float tiles[kSaStages][kStageFloats];
  __shared__ __align__(8) float2 sidx[DREG];  // every inverse-distance index twice: one LDS.64 gives the packed pair
  __shared__ __align__(8) unsigned long long mbar[kSaStages];
  __shared__ unsigned released[kSaStages];  // warps that have read a stage's box, counted over the whole walk

  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int D = FULL ? DREG : kp.D;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kSaStages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&mbar[s])));
      released[s] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < DREG) {
    const float id = (FULL || (int)threadIdx.x < kp.D) ? __ldg(kp.idx + threadIdx.x) : 0.f;
    sidx[threadIdx.x] = make_float2(id, id);
  }
  __syncthreads();

  // thread 0: request the box of tile c into stage s (the box holds min(DREG, D) = D candidate planes, host: boxDim[2])
  auto request = [&](const SaTile& c, int s) {
    const uint32_t mb = smem_addr(&mbar[s]);
    const uint32_t bytes = (uint32_t)D * kSaBoxH * kSaBoxW * 4;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_addr(&tiles[s][0])), "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(mb),
        "r"(c.tx * kSaCols - kSaLead), "r"(c.ty * kSaRows - 1), "r"(0), "r"(c.b)
        : "memory");
  };
  // cur: the tile being computed; ahead: the tile kSaStages strides further, whose box is requested when cur's is read
  SaTile cur;
  cur.tx = (int)blockIdx.x % tl.tilesX;
  cur.ty = ((int)blockIdx.x / tl.tilesX) % tl.tilesY;
  cur.b = (int)blockIdx.x / (tl.tilesX * tl.tilesY);
  SaTile ahead = cur;
#pragma unroll
  for (int s = 0; s < kSaStages; ++s) {
    if (threadIdx.x == 0 && ahead.b < kp.B) request(ahead, s);
    ahead = sa_next(ahead, tl);
  }

  const int hw = kp.Hout * kp.Wout;
  for (int it = 0; cur.b < kp.B; ++it, cur = sa_next(cur, tl), ahead = sa_next(ahead, tl)) {
  const int stage = it % kSaStages;
  const uint32_t parity = (uint32_t)(it / kSaStages) & 1u;
  const int tx = cur.tx, ty = cur.ty, b = cur.b;
  // first block column / row of the tile (-1 = the first output column / row on its own), and the box's first column
  const int cx = tx * kSaCols - 1, cy = ty * kSaRows - 1;
  const int bx0 = tx * kSaCols - kSaLead;

  // geometry of this thread's block while the box is in flight.  With scale 0.5 exactly (host: the x2 condition)
  // src_index() (softargmin.cuh) gives source position k + 0.25 for output 2k+1 and k + 0.75 for output 2k+2 — weights
  // (0.75, 0.25) and (0.25, 0.75), exact in binary — and position 0, weights (1, 0), for output 0 (block -1).
  const int m = cx + lane, k = cy + w;  // block column -1 .. Win-1, block row -1 .. Hin-1
  const bool live = m < kp.Win && k < kp.Hin;
  const int ra = min(max(k, 0), kp.Hin - 1), rb = min(ra + 1, kp.Hin - 1);
  const int ca = min(max(m, 0), kp.Win - 1), cb = min(ca + 1, kp.Win - 1);
  const float ly0[2] = {0.75f, k < 0 ? 1.0f : 0.25f}, ly1[2] = {0.25f, k < 0 ? 0.0f : 0.75f};
  const float lx0[2] = {0.75f, m < 0 ? 1.0f : 0.25f}, lx1[2] = {0.25f, m < 0 ? 0.0f : 0.75f};
  const bool rowok[2] = {live && k >= 0, live && 2 * k + 2 < kp.Hout};
  const bool colok[2] = {m >= 0, 2 * m + 2 < kp.Wout};
  const f32x2 LX0 = pk2(lx0[0], lx0[1]), LX1 = pk2(lx1[0], lx1[1]);
  const float* s00 = &tiles[stage][0] + (ra - cy) * kSaBoxW + (ca - bx0);
  const float* s01 = s00 + (cb - ca);
  const float* s10 = s00 + (rb - ra) * kSaBoxW;
  const float* s11 = s10 + (cb - ca);

  mbar_wait(smem_addr(&mbar[stage]), parity);

  // v[r][d]: the two output columns of output row r as a packed pair
  f32x2 v[2][DREG];
#pragma unroll
  for (int d = 0; d < DREG; ++d) {
    if (FULL || d < D) {
      constexpr int plane = kSaBoxH * kSaBoxW;
      const float v00 = s00[d * plane], v01 = s01[d * plane], v10 = s10[d * plane], v11 = s11[d * plane];
      const f32x2 top = fma2(LX0, pk2(v00, v00), mul2(LX1, pk2(v01, v01)));
      const f32x2 bot = fma2(LX0, pk2(v10, v10), mul2(LX1, pk2(v11, v11)));
#pragma unroll
      for (int r = 0; r < 2; ++r) v[r][d] = fma2(pk2(ly0[r], ly0[r]), top, mul2(pk2(ly1[r], ly1[r]), bot));
    } else {
      v[0][d] = v[1][d] = pk2(-CUDART_INF_F, -CUDART_INF_F);
    }
  }

  // this warp has its costs in registers.  The last of the tile's warps to get here requests the box of tile
  // it + kSaStages into the stage: nobody waits at a CTA barrier, the warps of a CTA drift apart freely (a warp that
  // runs ahead blocks on the full-barrier of a stage until the slowest warp has released the stage's previous box)
  __syncwarp();
  if (lane == 0) {
    __threadfence_block();
    const unsigned arrived = atomicAdd(&released[stage], 1u);
    if (arrived % kSaRows == kSaRows - 1 && ahead.b < kp.B) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      request(ahead, stage);
    }
  }

  float* nc = NC ? kp.normCosts + (long long)b * D * hw : nullptr;
  constexpr float kLog2e = 1.4426950408889634f;
  const f32x2 LOG2E = pk2(kLog2e, kLog2e);
  // both output rows of the block go through the softmax together (RJ = 2: four independent max / sum / expectation
  // chains per thread for the scheduler to interleave); with norm_costs kept the numerators of one row fill the registers
  constexpr int RJ = NC ? 1 : 2;
#pragma unroll
  for (int r0 = 0; r0 < 2; r0 += RJ) {
    if (NC && !rowok[r0]) continue;
    float mx[RJ][2];
    int am[RJ][2];
#pragma unroll
    for (int j = 0; j < RJ; ++j) {
      am[j][0] = am[j][1] = 0;
      unpk2(v[r0 + j][0], mx[j][0], mx[j][1]);
    }
#pragma unroll
    for (int d = 1; d < DREG; ++d) {
#pragma unroll
      for (int j = 0; j < RJ; ++j) {
        float x[2];
        unpk2(v[r0 + j][d], x[0], x[1]);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (AM) {
            if (x[c] > mx[j][c]) {  // strict '>' keeps the first maximum, like torch.argmax (unused slots hold -inf)
              mx[j][c] = x[c];
              am[j][c] = d;
            }
          } else {
            mx[j][c] = fmaxf(mx[j][c], x[c]);
          }
        }
      }
    }
    // softmax numerators 2^((v - max) * log2 e): one FFMA + one MUFU.EX2 each (ex2.approx.ftz, 2 ulp); the expectation
    // is accumulated on the un-normalised numerators and scaled once (distance_regressor.py:62-78)
    f32x2 NMLOG[RJ], sum[RJ], acc[RJ];
#pragma unroll
    for (int j = 0; j < RJ; ++j) {
      NMLOG[j] = pk2(-(mx[j][0] * kLog2e), -(mx[j][1] * kLog2e));
      sum[j] = acc[j] = 0ull;
    }
    f32x2 e[NC ? DREG : 1];
#pragma unroll
    for (int d = 0; d < DREG; ++d) {
      const f32x2 ID = *reinterpret_cast<const f32x2*>(&sidx[d]);
#pragma unroll
      for (int j = 0; j < RJ; ++j) {
        float a0, a1;
        unpk2(fma2(v[r0 + j][d], LOG2E, NMLOG[j]), a0, a1);  // 2^-inf = 0 for unused slots
        const f32x2 ed = pk2(ex2_ftz(a0), ex2_ftz(a1));
        if (NC) e[d] = ed;
        sum[j] = add2(sum[j], ed);
        acc[j] = fma2(ed, ID, acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < RJ; ++j) {
      const int r = r0 + j;
      float s[2], a[2];
      unpk2(sum[j], s[0], s[1]);
      unpk2(acc[j], a[0], a[1]);
      if (!rowok[r]) continue;
      const float rs[2] = {1.0f / s[0], 1.0f / s[1]};
      const int opix = (2 * k + 1 + r) * kp.Wout + (2 * m + 1);
      if (NC) {  // the two columns of a block back to back: both halves of a 32-byte sector reach L2 together
#pragma unroll
        for (int d = 0; d < DREG; ++d) {
          float e0, e1;
          unpk2(e[d], e0, e1);
          if (FULL || d < D) {
            if (colok[0]) __stcs(nc + (long long)d * hw + opix, e0 * rs[0]);
            if (colok[1]) __stcs(nc + (long long)d * hw + opix + 1, e1 * rs[1]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (!colok[c]) continue;
        kp.invDist[(long long)b * hw + opix + c] = a[c] * rs[c];
        if (AM) kp.argmax[(long long)b * hw + opix + c] = am[j][c];
      }
    }
  }
  }  // tiles
}

}  // namespace mvsgi
