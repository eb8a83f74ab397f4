// softargmin_bwd.cuh — gradient of the scale-factor-2 soft-argmin with respect to the cost volume, as a gather.
//
// Adjoint of DistanceRegressorWithFixedCandidates.forward (dsta_mvs/model/distance_regressor/distance_regressor.py:51-80;
// configs/base_model.yaml:28-34: pre_interp, interp_scale_factor 2) — what autograd runs through F.interpolate,
// F.softmax and the expectation.  The general kernel (backward.cuh) gives every output pixel a thread and scatters
// 4 atomics per candidate into a zero-filled gradient: 313 us per 16 G16V frames, 0.08 of the HBM roofline, bound by
// the L2 atomic units.  With the x2 upsample the scatter has a fixed shape: the 2x2 block of output pixels
// (2k+1, 2k+2) x (2m+1, 2m+2) reads input rows (k, k+1) x columns (m, m+1), so input pixel (y, x) collects from exactly
// the four blocks (y-1 .. y) x (x-1 .. x).  Here
//   * a thread owns a block: it interpolates the block's 4 x D costs, runs the four softmaxes, forms
//     g = p (G - sum p G) per output pixel and candidate (G = grad_inv_dist * idx + grad_norm_costs), and folds the
//     4 pixels into the 4 partial gradients of its input corners — all in registers;
//   * a lane adds its left neighbour's right-hand partials (one shuffle), a warp the lower-row partials of the warp
//     above it (one shared-memory exchange per tile), and every input pixel is written once: no atomics, no zero fill.
// A CTA of kSbRows + 1 warps produces kSbRows x 31 input pixels (the first warp and the first lane only supply halo
// partials).  Borders: block -1 puts all the weight of output row / column 0 on input row / column 0, the last block's
// two clamped taps land on the same input row / column — expressed as the per-thread weight pairs WY / WX below.
#pragma once
#include <cuda.h>

#include "softargmin.cuh"
#include "sweep.cuh"  // f32x2 helpers

namespace mvsgi {

constexpr int kSbRows = 4;                       // input rows per CTA
constexpr int kSbCols = 31;                      // input columns per CTA
constexpr int kSbThreads = (kSbRows + 1) * 32;

// the grad_norm_costs box of a CTA: 2 (kSbRows + 1) output rows x (64 columns + up to 3 of alignment lead), D planes
constexpr int kSbBoxH = 2 * (kSbRows + 1);
constexpr int kSbBoxW = 68;

__device__ __forceinline__ uint32_t sb_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// FULL: D == DREG.  GNC: 0 = no grad_norm_costs, 1 = read from global memory (two passes, the second out of L1/L2),
// 2 = the CTA's box of grad_norm_costs staged in shared memory by one cp.async.bulk.tensor.4d (TMA) that travels while
// the threads load and interpolate the costs; outside the image the box is zero-filled, which is the gradient there
template <int DREG, bool FULL, int GNC>
__global__ void __launch_bounds__(kSbThreads) softargmin_x2_backward_kernel(const SoftBwdKP kb,
                                                                            const __grid_constant__ CUtensorMap gmap) {
  const SoftKP& kp = kb.f;
  // dynamic shared memory: xch[kSbRows + 1][DREG][32] (lower-row partials of a warp, for the warp below), then for
  // GNC == 2 the box gbox[d][kSbBoxH][kSbBoxW]
  extern __shared__ __align__(128) float dyn[];
  float (*xch)[DREG][32] = reinterpret_cast<float (*)[DREG][32]>(dyn);
  float* gbox = dyn + (kSbRows + 1) * DREG * 32;
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ __align__(8) float2 sidx[DREG];        // every inverse-distance index twice (one LDS.64 = a packed pair)

  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int D = FULL ? DREG : kp.D;
  if (threadIdx.x < DREG) {
    const float id = (FULL || (int)threadIdx.x < kp.D) ? __ldg(kp.idx + threadIdx.x) : 0.f;
    sidx[threadIdx.x] = make_float2(id, id);
  }
  if (GNC == 2 && threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb_smem_addr(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int b = blockIdx.z;
  // first output column of the CTA's blocks, and how far it sits past the 16-byte aligned column the box starts at
  const int gcol0 = 2 * ((int)blockIdx.x * kSbCols - 1) + 1, glead = gcol0 & 3;
  if (GNC == 2 && threadIdx.x == 0) {
    const uint32_t mb = sb_smem_addr(&mbar);
    const uint32_t bytes = (uint32_t)D * kSbBoxH * kSbBoxW * 4;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(sb_smem_addr(gbox)), "l"(reinterpret_cast<unsigned long long>(&gmap)), "r"(mb), "r"(gcol0 - glead),
        "r"(2 * ((int)blockIdx.y * kSbRows - 1) + 1), "r"(0), "r"(b)
        : "memory");
  }
  const int k = blockIdx.y * kSbRows - 1 + w, m = blockIdx.x * kSbCols - 1 + lane;  // this thread's block
  const bool blk = k < kp.Hin && m < kp.Win;
  // forward taps and weights (softargmin_tma.cuh: exact for scale 0.5)
  const int ra = min(max(k, 0), kp.Hin - 1), rb = min(ra + 1, kp.Hin - 1);
  const int ca = min(max(m, 0), kp.Win - 1), cb = min(ca + 1, kp.Win - 1);
  const float ly0[2] = {0.75f, k < 0 ? 1.0f : 0.25f}, ly1[2] = {0.25f, k < 0 ? 0.0f : 0.75f};
  const float lx0[2] = {0.75f, m < 0 ? 1.0f : 0.25f}, lx1[2] = {0.25f, m < 0 ? 0.0f : 0.75f};
  const bool rowok[2] = {blk && k >= 0, blk && 2 * k + 2 < kp.Hout};
  const bool colok[2] = {m >= 0, 2 * m + 2 < kp.Wout};
  // adjoint weights: WY[r][i] = share of output row r that input row k + i receives (i = 0: row k, 1: row k + 1)
  float WY[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, WX[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  if (k < 0) {
    WY[1][1] = 1.0f;  // output row 0 reads input row 0 = row k + 1 only
  } else if (k == kp.Hin - 1) {
    WY[0][0] = 1.0f;  // output row 2 Hin - 1: both clamped taps are input row Hin - 1 = row k
  } else if (k < kp.Hin) {
    WY[0][0] = 0.75f; WY[0][1] = 0.25f; WY[1][0] = 0.25f; WY[1][1] = 0.75f;
  }
  if (m < 0) {
    WX[1][1] = 1.0f;
  } else if (m == kp.Win - 1) {
    WX[0][0] = 1.0f;
  } else if (m < kp.Win) {
    WX[0][0] = 0.75f; WX[0][1] = 0.25f; WX[1][0] = 0.25f; WX[1][1] = 0.75f;
  }

  const float* c00 = kp.costs + (long long)b * kp.sB + ra * kp.sY + ca * kp.sX;
  const long long dc = (cb - ca) * kp.sX, dr = (rb - ra) * kp.sY;
  const f32x2 LX0 = pk2(lx0[0], lx0[1]), LX1 = pk2(lx1[0], lx1[1]);

  // v[r][d]: the two output columns of output row r as a packed pair (same arithmetic as the forward kernels)
  f32x2 v[2][DREG];
#pragma unroll
  for (int d = 0; d < DREG; ++d) {
    if (FULL || d < D) {
      const float* pl = c00 + (long long)d * kp.sD;
      const float v00 = __ldg(pl), v01 = __ldg(pl + dc), v10 = __ldg(pl + dr), v11 = __ldg(pl + dr + dc);
      const f32x2 top = fma2(LX0, pk2(v00, v00), mul2(LX1, pk2(v01, v01)));
      const f32x2 bot = fma2(LX0, pk2(v10, v10), mul2(LX1, pk2(v11, v11)));
#pragma unroll
      for (int r = 0; r < 2; ++r) v[r][d] = fma2(pk2(ly0[r], ly0[r]), top, mul2(pk2(ly1[r], ly1[r]), bot));
    } else {
      v[0][d] = v[1][d] = pk2(-CUDART_INF_F, -CUDART_INF_F);
    }
  }

  const int hw = kp.Hout * kp.Wout;
  constexpr float kLog2e = 1.4426950408889634f;
  const f32x2 LOG2E = pk2(kLog2e, kLog2e);
  if (GNC == 2) {  // the box has had the whole interpolation to arrive (phase 0 of a barrier used once)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
        "@p bra SB_DONE;\n"
        "bra SB_WAIT;\n"
        "SB_DONE:\n"
        "}\n" ::"r"(sb_smem_addr(&mbar))
        : "memory");
  }
  // v[r][d] <- g[r][d] = p_d (G_d - sum_e p_e G_e) for the two pixels of output row r (0 for pixels outside the image)
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int opix = (2 * k + 1 + r) * kp.Wout + (2 * m + 1);
    const bool ok0 = rowok[r] && colok[0], ok1 = rowok[r] && colok[1];
    const float gi0 = (kb.gInv != nullptr && ok0) ? __ldg(kb.gInv + (long long)b * hw + opix) : 0.f;
    const float gi1 = (kb.gInv != nullptr && ok1) ? __ldg(kb.gInv + (long long)b * hw + opix + 1) : 0.f;
    const f32x2 GI = pk2(gi0, gi1);
    const float* gnc = GNC == 1 ? kb.gNc + (long long)b * D * hw + opix : nullptr;
    const float* gsm = gbox + (2 * w + r) * kSbBoxW + glead + 2 * lane;  // GNC == 2: this block's pair of row r
    auto upstream = [&](int d) {  // G_d of the two pixels
      const f32x2 ID = *reinterpret_cast<const f32x2*>(&sidx[d]);
      if (GNC == 0) return mul2(GI, ID);
      if (GNC == 2) return fma2(GI, ID, pk2(gsm[d * (kSbBoxH * kSbBoxW)], gsm[d * (kSbBoxH * kSbBoxW) + 1]));
      const float n0 = ok0 ? __ldg(gnc + (long long)d * hw) : 0.f, n1 = ok1 ? __ldg(gnc + (long long)d * hw + 1) : 0.f;
      return fma2(GI, ID, pk2(n0, n1));
    };
    float mx0, mx1;
    unpk2(v[r][0], mx0, mx1);
#pragma unroll
    for (int d = 1; d < DREG; ++d) {
      float x0, x1;
      unpk2(v[r][d], x0, x1);
      mx0 = fmaxf(mx0, x0);
      mx1 = fmaxf(mx1, x1);
    }
    const f32x2 NMLOG = pk2(-(mx0 * kLog2e), -(mx1 * kLog2e));
    f32x2 sum = 0ull, dot = 0ull;
#pragma unroll
    for (int d = 0; d < DREG; ++d) {
      float a0, a1;
      unpk2(fma2(v[r][d], LOG2E, NMLOG), a0, a1);  // 2^-inf = 0 for unused slots
      const f32x2 e = pk2(ex2_ftz(a0), ex2_ftz(a1));
      v[r][d] = e;
      sum = add2(sum, e);
      if (FULL || d < D) dot = fma2(e, upstream(d), dot);
    }
    float s0, s1;
    unpk2(sum, s0, s1);
    const f32x2 RS = pk2(1.0f / s0, 1.0f / s1);
    const f32x2 NDOT = mul2(mul2(dot, RS), pk2(-1.0f, -1.0f));  // -sum_e p_e G_e
#pragma unroll
    for (int d = 0; d < DREG; ++d)
      if (FULL || d < D) v[r][d] = mul2(mul2(v[r][d], RS), add2(upstream(d), NDOT));
  }

  // fold the 4 pixels into the partial gradients of the 4 input corners, add the left neighbour's right-hand partials
  float h0[DREG];  // towards input row k, column m (complete along x)
#pragma unroll
  for (int d = 0; d < DREG; ++d) {
    if (FULL || d < D) {
      float g00, g01, g10, g11;
      unpk2(v[0][d], g00, g01);
      unpk2(v[1][d], g10, g11);
      // t[r][j]: output row r towards input column m + j
      const float t00 = fmaf(WX[1][0], g01, WX[0][0] * g00), t01 = fmaf(WX[1][1], g01, WX[0][1] * g00);
      const float t10 = fmaf(WX[1][0], g11, WX[0][0] * g10), t11 = fmaf(WX[1][1], g11, WX[0][1] * g10);
      const float p00 = fmaf(WY[1][0], t10, WY[0][0] * t00), p01 = fmaf(WY[1][0], t11, WY[0][0] * t01);
      const float p10 = fmaf(WY[1][1], t10, WY[0][1] * t00), p11 = fmaf(WY[1][1], t11, WY[0][1] * t01);
      h0[d] = p00 + __shfl_up_sync(0xFFFFFFFFu, p01, 1);
      xch[w][d][lane] = p10 + __shfl_up_sync(0xFFFFFFFFu, p11, 1);
    }
  }
  __syncthreads();
  if (w == 0 || lane == 0 || !blk) return;  // halo blocks own no input pixel
  float* gc = kb.gradCosts + (long long)b * kb.gB + (long long)k * kb.gY + (long long)m * kb.gX;
#pragma unroll
  for (int d = 0; d < DREG; ++d)
    if (FULL || d < D) gc[(long long)d * kb.gD] = h0[d] + xch[w - 1][d][lane];
}

}  // namespace mvsgi
