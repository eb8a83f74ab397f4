// softargmin.cuh — fused bilinear upsample + softmax over candidates + expected inverse distance.
//
// Replaces DistanceRegressorWithFixedCandidates.forward
// (dsta_mvs/model/distance_regressor/distance_regressor.py:51-80): the reference materialises the upsampled
// [B, D, sH, sW] volume (F.interpolate), then softmax, then expand/mul/sum.  Here one thread (or DSPLIT lanes of one
// warp, combined with warp shuffles) owns an output pixel, interpolates its D costs straight out of the low-res
// volume, and writes inv_dist (and norm_costs / argmax only when asked).
//
// The interpolation reproduces PyTorch's CPU upsample_bilinear2d bit for bit (fma placement determined empirically,
// see oracle/mvsgi_oracle.c) so that the arg-max candidate is bit-exact against the reference.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace mvsgi {

struct SoftKP {
  int B, D, Hin, Win, Hout, Wout, interp;
  float scaleH, scaleW;
  const float* costs;
  long long sB, sD, sY, sX;
  const float* idx;
  float* invDist;
  float* normCosts;
  int* argmax;
};

__device__ __forceinline__ void src_index(int dst, float scale, int nIn, int& i0, int& i1, float& l0, float& l1) {
  float src = __fmaf_rn(scale, (float)dst + 0.5f, -0.5f);
  src = fmaxf(src, 0.0f);
  i0 = min((int)src, nIn - 1);
  i1 = min(i0 + 1, nIn - 1);
  l1 = fminf(fmaxf(__fsub_rn(src, (float)i0), 0.0f), 1.0f);
  l0 = __fsub_rn(1.0f, l1);
}

template <int DSPLIT>
__global__ void __launch_bounds__(128) softargmin_kernel(const SoftKP kp) {
  constexpr int PPW = 32 / DSPLIT;  // output pixels per warp
  constexpr int DREG = 16;          // candidates per lane (host guarantees ceil(D / DSPLIT) <= DREG)
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int part = lane / PPW, pp = lane % PPW;
  const int oxRaw = (blockIdx.x * 4 + wrp) * PPW + pp;
  const bool act = oxRaw < kp.Wout;
  const int ox = act ? oxRaw : kp.Wout - 1;
  const int oy = blockIdx.y, b = blockIdx.z;

  int x0 = ox, x1 = ox, y0 = oy, y1 = oy;
  float lx0 = 1.f, lx1 = 0.f, ly0 = 1.f, ly1 = 0.f;
  if (kp.interp) {
    src_index(ox, kp.scaleW, kp.Win, x0, x1, lx0, lx1);
    src_index(oy, kp.scaleH, kp.Hin, y0, y1, ly0, ly1);
  }
  const int dPer = (kp.D + DSPLIT - 1) / DSPLIT;
  const int dBeg = part * dPer, dEnd = min(kp.D, dBeg + dPer);
  const float* cb = kp.costs + (long long)b * kp.sB;
  const long long o00 = y0 * kp.sY + x0 * kp.sX, o01 = y0 * kp.sY + x1 * kp.sX;
  const long long o10 = y1 * kp.sY + x0 * kp.sX, o11 = y1 * kp.sY + x1 * kp.sX;

  float v[DREG];
  float m = -CUDART_INF_F;
  int am = 0x7FFFFFFF;
#pragma unroll
  for (int i = 0; i < DREG; ++i) {
    const int d = dBeg + i;
    v[i] = -CUDART_INF_F;
    if (d < dEnd) {
      const float* pl = cb + (long long)d * kp.sD;
      float val;
      if (kp.interp) {
        const float v00 = __ldg(pl + o00), v01 = __ldg(pl + o01), v10 = __ldg(pl + o10), v11 = __ldg(pl + o11);
        const float top = __fmaf_rn(lx0, v00, __fmul_rn(lx1, v01));
        const float bot = __fmaf_rn(lx0, v10, __fmul_rn(lx1, v11));
        val = __fmaf_rn(ly0, top, __fmul_rn(ly1, bot));
      } else {
        val = __ldg(pl + o00);
      }
      v[i] = val;
      if (am == 0x7FFFFFFF || val > m) {  // strict '>' keeps the first maximum
        m = val;
        am = d;
      }
    }
  }
  // combine the DSPLIT partial maxima of a pixel (first maximum wins, like torch.argmax)
#pragma unroll
  for (int s = PPW; s < 32; s <<= 1) {
    const float om = __shfl_xor_sync(0xFFFFFFFFu, m, s);
    const int oa = __shfl_xor_sync(0xFFFFFFFFu, am, s);
    if (om > m || (om == m && oa < am)) { m = om; am = oa; }
  }
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < DREG; ++i) {
    if (dBeg + i < dEnd) {
      v[i] = expf(v[i] - m);
      sum += v[i];
    }
  }
#pragma unroll
  for (int s = PPW; s < 32; s <<= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, s);
  const long long hw = (long long)kp.Hout * kp.Wout;
  const long long opix = (long long)oy * kp.Wout + ox;
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < DREG; ++i) {
    const int d = dBeg + i;
    if (d < dEnd) {
      const float p = __fdiv_rn(v[i], sum);
      if (kp.normCosts != nullptr && act) __stcs(kp.normCosts + ((long long)b * kp.D + d) * hw + opix, p);
      const float term = __fmul_rn(p, __ldg(kp.idx + d));
      acc = (i == 0) ? term : __fadd_rn(acc, term);
    }
  }
#pragma unroll
  for (int s = PPW; s < 32; s <<= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, s);
  if (act && part == 0) {
    kp.invDist[(long long)b * hw + opix] = acc;
    if (kp.argmax != nullptr) kp.argmax[(long long)b * hw + opix] = am;
  }
}

}  // namespace mvsgi
