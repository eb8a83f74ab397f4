// backward.cuh — backward passes of the fused sweep and of the soft-argmin (training drop-in).
//
// What torch.autograd derives from the reference's forward code when training_step differentiates through
// cv_builder and dist_regressor (dsta_mvs/model/mvs_model/spherical_sweep_stereo.py:170-202):
//   sweep (std)    d vol / d f_k = (2 / n) (f_k - mean)        for cameras k in the valid set M, valid voxels only
//                  (spherical_sweep_avg.py:114-125; the term through `avg` cancels because sum_M (f_j - mean) = 0)
//   sweep (concat) the adjoint of the sampler (spherical_sweep.py:52-67)
//   sampler        grad of a sample goes to its four taps with the forward weights; taps in the zero padding are
//                  dropped by F.pad's backward (dsta_mvs/model/backports/backports.py:52-86)
//   soft-argmin    softmax backward p_d (G_d - sum_e p_e G_e), G_d = g_norm_costs_d + g_inv_dist * inv_dist_idx_d,
//                  followed by the adjoint of the bilinear upsample (distance_regressor.py:51-80)
//
// The sampler adjoint is a scatter-add.  The sweep's backward is the BWD instantiation of sweep_kernel (sweep.cuh): the
// forward's work decomposition — tap table in shared memory, built once for all frames, mask lookups, frame loop —
// with the scatter in place of the store.  It accumulates into a zero-initialised gradient image with the layout of
// the features it taps: the caller's dense channels-last grad_feats directly (2 launches: memset + kernel), or, for
// NCHW features, the staged layout ([B][CAM][CG][Hf+3][Wf+2][16]), where the padding taps need no test (their sums land
// in the frame and are never read back) and unstage_grad_kernel then writes the interior to grad_feats.  A tap's 16
// channels are four 16-byte vector reductions (red.global.add.v4.f32, one per lane of the quad that owns the voxel).
// This file keeps the un-staging kernel and the general soft-argmin backward.
#pragma once
#include "softargmin.cuh"
#include "sweep.cuh"

namespace mvsgi {

// red_add4: sweep.cuh

// gradient image in staged layout -> grad_feats [B, CAM, C, Hf, Wf] (any element strides): the inverse of
// stage_feats_kernel.  One block = one image row of one (b, cam, channel group).
__global__ void __launch_bounds__(256)
unstage_grad_kernel(const float* __restrict__ src, int n1, int C, int CG, int Hf, int Wf, float* __restrict__ dst,
                    long long sN0, long long sN1, long long sC, long long sH, long long sW) {
  __shared__ float tile[16 * kStagePitch];
  const int Wp = Wf + 2, Hp1 = Hf + 3;
  const int xBlocks = (Wf + kStagePx - 1) / kStagePx;
  const int xb = blockIdx.x % xBlocks;
  const int ncg = blockIdx.x / xBlocks;
  const int n = ncg / CG, cg = ncg - n * CG;
  const int y = blockIdx.y;
  const int x0 = xb * kStagePx;
  const int npx = min(kStagePx, Wf - x0);
  const int t = threadIdx.x;
  const float4* srow = reinterpret_cast<const float4*>(src + (((long long)ncg * Hp1 + (y + 1)) * Wp + (x0 + 1)) * 16);
  for (int idx = t; idx < npx * 4; idx += 256) {
    const int px = idx >> 2, c4 = idx & 3;
    const float4 v = __ldg(srow + idx);
    tile[(c4 * 4 + 0) * kStagePitch + px] = v.x;
    tile[(c4 * 4 + 1) * kStagePitch + px] = v.y;
    tile[(c4 * 4 + 2) * kStagePitch + px] = v.z;
    tile[(c4 * 4 + 3) * kStagePitch + px] = v.w;
  }
  __syncthreads();
  float* d = dst + (long long)(n / n1) * sN0 + (long long)(n % n1) * sN1 + (long long)y * sH;
  if (sW == 1) {  // NCHW destination: a warp writes consecutive pixels of one channel
    const int lane = t & 31, wrp = t >> 5;
    for (int c = wrp; c < 16; c += 8) {
      const int ch = cg * 16 + c;
      if (ch >= C) continue;
      for (int i = lane; i < npx; i += 32) d[(long long)ch * sC + x0 + i] = tile[c * kStagePitch + i];
    }
  } else {  // channels-last (or anything else): 16 consecutive lanes write the 16 channels of one pixel
    for (int idx = t; idx < npx * 16; idx += 256) {
      const int c = idx & 15, i = idx >> 4;
      const int ch = cg * 16 + c;
      if (ch < C) d[(long long)ch * sC + (long long)(x0 + i) * sW] = tile[c * kStagePitch + i];
    }
  }
}

// ---- soft-argmin backward ------------------------------------------------------------------------------------------
// SoftBwdKP: softargmin.cuh

__global__ void zero_strided4_kernel(float* p, int n0, int n1, int n2, int n3, long long s0, long long s1, long long s2,
                                     long long s3) {
  const long long total = (long long)n0 * n1 * n2 * n3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int i3 = (int)(i % n3);
    long long r = i / n3;
    const int i2 = (int)(r % n2);
    r /= n2;
    const int i1 = (int)(r % n1), i0 = (int)(r / n1);
    p[i0 * s0 + i1 * s1 + i2 * s2 + i3 * s3] = 0.f;
  }
}

// One thread per output pixel; the D interpolated costs are recomputed (three passes over L1-resident taps) instead
// of kept in registers, so any D <= MVSGI_MAX_CANDS works with one kernel.
__global__ void __launch_bounds__(128) softargmin_backward_kernel(const SoftBwdKP kb) {
  const SoftKP& kp = kb.f;
  const int ox = blockIdx.x * 128 + threadIdx.x;
  if (ox >= kp.Wout) return;
  const int oy = blockIdx.y, b = blockIdx.z;
  int x0 = ox, x1 = ox, y0 = oy, y1 = oy;
  float lx0 = 1.f, lx1 = 0.f, ly0 = 1.f, ly1 = 0.f;
  if (kp.interp) {
    src_index(ox, kp.scaleW, kp.Win, x0, x1, lx0, lx1);
    src_index(oy, kp.scaleH, kp.Hin, y0, y1, ly0, ly1);
  }
  const float* cb = kp.costs + (long long)b * kp.sB;
  const long long o00 = y0 * kp.sY + x0 * kp.sX, o01 = y0 * kp.sY + x1 * kp.sX;
  const long long o10 = y1 * kp.sY + x0 * kp.sX, o11 = y1 * kp.sY + x1 * kp.sX;
  auto value = [&](int d) {
    const float* pl = cb + (long long)d * kp.sD;
    if (!kp.interp) return __ldg(pl + o00);
    const float v00 = __ldg(pl + o00), v01 = __ldg(pl + o01), v10 = __ldg(pl + o10), v11 = __ldg(pl + o11);
    const float top = __fmaf_rn(lx0, v00, __fmul_rn(lx1, v01));
    const float bot = __fmaf_rn(lx0, v10, __fmul_rn(lx1, v11));
    return __fmaf_rn(ly0, top, __fmul_rn(ly1, bot));
  };
  float m = -CUDART_INF_F;
  for (int d = 0; d < kp.D; ++d) m = fmaxf(m, value(d));
  const long long hw = (long long)kp.Hout * kp.Wout;
  const long long opix = (long long)oy * kp.Wout + ox;
  const float gi = kb.gInv ? __ldg(kb.gInv + b * hw + opix) : 0.f;
  const float* gnc = kb.gNc ? kb.gNc + (long long)b * kp.D * hw + opix : nullptr;
  float sum = 0.f, dot = 0.f;
  for (int d = 0; d < kp.D; ++d) {
    const float e = expf(value(d) - m);
    const float G = fmaf(gi, __ldg(kp.idx + d), gnc ? __ldg(gnc + d * hw) : 0.f);
    sum += e;
    dot = fmaf(e, G, dot);
  }
  const float rs = 1.0f / sum;
  dot *= rs;  // sum_e p_e G_e
  float* gc = kb.gradCosts + (long long)b * kb.gB;
  const long long g00 = y0 * kb.gY + x0 * kb.gX, g01 = y0 * kb.gY + x1 * kb.gX;
  const long long g10 = y1 * kb.gY + x0 * kb.gX, g11 = y1 * kb.gY + x1 * kb.gX;
  for (int d = 0; d < kp.D; ++d) {
    const float p = expf(value(d) - m) * rs;
    const float G = fmaf(gi, __ldg(kp.idx + d), gnc ? __ldg(gnc + d * hw) : 0.f);
    const float gvv = p * (G - dot);
    float* pl = gc + (long long)d * kb.gD;
    if (kp.interp) {
      atomicAdd(pl + g00, ly0 * lx0 * gvv);
      atomicAdd(pl + g01, ly0 * lx1 * gvv);
      atomicAdd(pl + g10, ly1 * lx0 * gvv);
      atomicAdd(pl + g11, ly1 * lx1 * gvv);
    } else {
      pl[g00] = gvv;
    }
  }
}

}  // namespace mvsgi
