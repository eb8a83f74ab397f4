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

// soft-argmin backward (backward.cuh: any scale / candidate count, atomics; softargmin_bwd.cuh: x2 gather)
struct SoftBwdKP {
  SoftKP f;              // forward problem (costs, strides, sizes, idx)
  const float* gInv;     // [B, 1, Hout, Wout] or nullptr
  const float* gNc;      // [B, D, Hout, Wout] or nullptr
  float* gradCosts;      // element (b, d, y, x) at gB*b + gD*d + gY*y + gX*x
  long long gB, gD, gY, gX;
};

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

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

// ---- scale-factor-2 fast path ----------------------------------------------------------------------------------------
// With F.interpolate(scale_factor=2, align_corners=False) the 2x2 block of output pixels (2k+1, 2k+2) x (2m+1, 2m+2)
// reads exactly the 2x2 input neighbourhood rows (k, k+1) x columns (m, m+1) (clamped at the border; k, m = -1 is the
// first output row / column on its own).  One thread owns such a block: 4 coalesced loads per candidate feed 4 output
// pixels (16 per pixel in the generic kernel), all D candidates of the 4 pixels stay in registers, and the softmax uses
// one reciprocal per pixel.  The interpolated values are computed with the same src_index() weights and the same fma
// placement as the generic kernel, so the arg-max candidate stays bit-exact against PyTorch's CPU upsample.
template <int DREG, bool FULL, bool NC, bool AM>  // FULL: D == DREG; NC / AM: norm_costs / argmax requested
__global__ void __launch_bounds__(128) softargmin_x2_kernel(const SoftKP kp) {
  const int m = blockIdx.x * 128 + threadIdx.x - 1;  // -1 .. Win-1
  const int k = blockIdx.y - 1;                      // -1 .. Hin-1
  const int b = blockIdx.z;
  if (m >= kp.Win) return;
  const int D = FULL ? DREG : kp.D;
  const int ra = max(k, 0), rb = min(ra + 1, kp.Hin - 1);
  const int ca = max(m, 0), cb = min(ca + 1, kp.Win - 1);
  // one frame of costs spans < 2^31 elements (host check): 32-bit element offsets from the frame base
  const float* fb = kp.costs + (long long)b * kp.sB;
  const int o00 = ra * (int)kp.sY + ca * (int)kp.sX;
  const int dx = (cb - ca) * (int)kp.sX, dy = (rb - ra) * (int)kp.sY;  // 0 at the clamped border
  const int sD = (int)kp.sD;

  // weights of the (up to) two output rows / columns of this block
  float ly0[2], ly1[2], lx0[2], lx1[2];
  bool rowok[2], colok[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int oy = 2 * k + 1 + r, ox = 2 * m + 1 + r;
    int i0, i1;
    rowok[r] = oy >= 0 && oy < kp.Hout;
    colok[r] = ox >= 0 && ox < kp.Wout;
    src_index(max(oy, 0), kp.scaleH, kp.Hin, i0, i1, ly0[r], ly1[r]);
    src_index(max(ox, 0), kp.scaleW, kp.Win, i0, i1, lx0[r], lx1[r]);
  }

  float v[4][DREG];  // [row r * 2 + column c][candidate]
  float idx[DREG];
#pragma unroll
  for (int d = 0; d < DREG; ++d) {
    idx[d] = 0.f;
#pragma unroll
    for (int px = 0; px < 4; ++px) v[px][d] = -CUDART_INF_F;
    if (FULL || d < D) {
      // non-negative 32-bit element offsets (host check) -> one IMAD.WIDE.U32 per address
      const unsigned o = (unsigned)(o00 + d * sD);
      const float v00 = __ldg(fb + o), v01 = __ldg(fb + (o + (unsigned)dx)), v10 = __ldg(fb + (o + (unsigned)dy)),
                  v11 = __ldg(fb + (o + (unsigned)(dy + dx)));
      idx[d] = __ldg(kp.idx + d);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float top = __fmaf_rn(lx0[c], v00, __fmul_rn(lx1[c], v01));
        const float bot = __fmaf_rn(lx0[c], v10, __fmul_rn(lx1[c], v11));
#pragma unroll
        for (int r = 0; r < 2; ++r) v[r * 2 + c][d] = __fmaf_rn(ly0[r], top, __fmul_rn(ly1[r], bot));
      }
    }
  }

  const int hw = kp.Hout * kp.Wout;
  float* nc = NC ? kp.normCosts + (long long)b * D * hw : nullptr;
  constexpr float kLog2e = 1.4426950408889634f;
#pragma unroll
  for (int px = 0; px < 4; ++px) {
    const int r = px >> 1, c = px & 1;
    if (!(rowok[r] && colok[c])) continue;
    const int opix = (2 * k + 1 + r) * kp.Wout + (2 * m + 1 + c);
    float mx = v[px][0];
    int am = 0;
#pragma unroll
    for (int d = 1; d < DREG; ++d) {
      if (AM) {
        if (v[px][d] > mx) {  // strict '>' keeps the first maximum, like torch.argmax (unused slots hold -inf)
          mx = v[px][d];
          am = d;
        }
      } else {
        mx = fmaxf(mx, v[px][d]);
      }
    }
    // softmax numerators 2^((v - max) * log2 e): one FFMA + one MUFU.EX2 each (ex2.approx.ftz, 2 ulp); the expectation
    // is accumulated on the un-normalised numerators and scaled once (distance_regressor.py:62-78)
    const float mlog = mx * kLog2e;
    float sum = 0.f, acc = 0.f;
#pragma unroll
    for (int d = 0; d < DREG; ++d) {
      const float e = ex2_ftz(__fmaf_rn(v[px][d], kLog2e, -mlog));  // 2^-inf = 0 for unused slots
      v[px][d] = e;
      sum += e;
      acc = __fmaf_rn(e, idx[d], acc);
    }
    const float rs = 1.0f / sum;
    if (NC) {
#pragma unroll
      for (int d = 0; d < DREG; ++d)
        if (FULL || d < D) __stcs(nc + (long long)d * hw + opix, v[px][d] * rs);
    }
    kp.invDist[(long long)b * hw + opix] = acc * rs;
    if (AM) kp.argmax[(long long)b * hw + opix] = am;
  }
}

}  // namespace mvsgi
