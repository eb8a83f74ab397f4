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
// The sampler adjoint is a scatter-add.  It accumulates into a zero-initialised gradient image with the SAME
// channels-last padded layout as the staged features ([B][CAM][CG][Hf+3][Wf+2][16]), so a tap's 16 channels are four
// 16-byte vector reductions (red.global.add.v4.f32, one per lane of the quad that owns the voxel) and the padding taps
// need no test: their sums land in the frame and are never read back.  unstage_grad_kernel then writes the interior
// to grad_feats in the caller's layout.
#pragma once
#include "softargmin.cuh"
#include "sweep.cuh"

namespace mvsgi {

MVSGI_DEV void red_add4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct SweepBwdKP {
  SweepKP f;             // forward problem (feat = staged features, mbits = packed masks; vol unused)
  const float* gradVol;  // [B, Cout, Dvol, Ho, Wo]
  float* gradStage;      // zero-initialised, staged layout
  // geometry shared by the frames (fused projection or shared grids): per (candidate, pixel, camera) the tap offset,
  // frac x, frac y and the mask lookup info, evaluated once by sweep_backward_geom_kernel instead of once per frame
  // in every quad; nullptr = evaluate in the backward kernel (per-frame grids)
  const uint4* geom;
};

// [D][P][CAM] entries: (tap offset incl. the camera's image, frac x, frac y, mask info | kMaskNone)
template <int CAM, int AGG>
__global__ void __launch_bounds__(128)
sweep_backward_geom_kernel(const SweepKP kp, const __grid_constant__ MvsgiRig rig, uint4* __restrict__ out) {
  const int pix = blockIdx.x * 128 + threadIdx.x;
  if (pix >= kp.P) return;
  const int d = blockIdx.y;
  const int h = pix / kp.Wo, w = pix - h * kp.Wo;
  float rx = 0.f, ry = 0.f, rz = 0.f;
  if (kp.gridMode == MVSGI_GRID_PROJECT) cv_ray(rig, h, w, rx, ry, rz);
#pragma unroll
  for (int cam = 0; cam < CAM; ++cam) {
    float gx, gy;
    bool gvalid = true;
    if (kp.gridMode == MVSGI_GRID_PROJECT) {
      float x, y, z;
      to_camera(rig.cams[cam], rx, ry, rz, rig.dist[d], x, y, z);
      gvalid = project(rig.cams[cam], x, y, z, gx, gy);
    } else {
      const float2 g = __ldg(reinterpret_cast<const float2*>(kp.grids + cam * kp.gsCam + d * kp.gsD + h * kp.gsY +
                                                               w * kp.gsX));
      gx = g.x;
      gy = g.y;
      if (AGG == MVSGI_AGG_STD) gvalid = __ldg(kp.gmask + cam * kp.gmCam + d * kp.gmD + h * kp.gmY + w * kp.gmX) != 0;
    }
    int off;
    float lx, ly;
    make_entry(kp, gx, gy, off, lx, ly);
    uint32_t info = kMaskNone;
    if (AGG == MVSGI_AGG_STD && gvalid) info = make_mask_info(gx, gy, kp.Hm, kp.Wm, kp.mPitch);
    out[((long long)d * kp.P + pix) * CAM + cam] =
        make_uint4((uint32_t)(off + cam * kp.fsCam), __float_as_uint(lx), __float_as_uint(ly), info);
  }
}

constexpr int kBwdThreads = 256;  // 64 quads = 64 voxels per block

// grid: (ceil(P / 64), D, B).  A quad of 4 lanes owns voxel (b, d, pixel); lane q evaluates the geometry and the mask
// test of camera q and the quad shares them with shuffles; then every lane handles its 4 channels of each group.
template <int CAM, int AGG>
__global__ void __launch_bounds__(kBwdThreads)
sweep_backward_kernel(const SweepBwdKP kb, const __grid_constant__ MvsgiRig rig) {
  const SweepKP& kp = kb.f;
  const int lane = threadIdx.x & 31, q = lane & 3;
  const int pixRaw = blockIdx.x * (kBwdThreads / 4) + (threadIdx.x >> 2);
  const bool act = pixRaw < kp.P;
  const int pix = act ? pixRaw : kp.P - 1;
  const int d = blockIdx.y, b = blockIdx.z;
  const int h = pix / kp.Wo, w = pix - h * kp.Wo;
  const int HW = kp.P;
  const unsigned qmask = 0xFu << (lane & ~3);

  // ---- lane q: camera q
  int off = kp.zeroOff;
  float lx = 0.f, ly = 0.f;
  uint32_t bit = 0;
  if (q < CAM) {
    uint32_t info = kMaskNone;
    if (kb.geom != nullptr) {
      const uint4 e = __ldg(kb.geom + ((long long)d * kp.P + pix) * CAM + q);
      off = (int)e.x;
      lx = __uint_as_float(e.y);
      ly = __uint_as_float(e.z);
      info = e.w;
    } else {
      float gx, gy;
      bool gvalid = true;
      if (kp.gridMode == MVSGI_GRID_PROJECT) {
        float rx, ry, rz, x, y, z;
        cv_ray(rig, h, w, rx, ry, rz);
        to_camera(rig.cams[q], rx, ry, rz, rig.dist[d], x, y, z);
        gvalid = project(rig.cams[q], x, y, z, gx, gy);
      } else {
        const long long bb = kp.perFrameGrids ? b : 0;
        const float2 g = __ldg(reinterpret_cast<const float2*>(kp.grids + bb * kp.gsB + q * kp.gsCam + d * kp.gsD +
                                                                 h * kp.gsY + w * kp.gsX));
        gx = g.x;
        gy = g.y;
        if (AGG == MVSGI_AGG_STD)
          gvalid = __ldg(kp.gmask + bb * kp.gmB + q * kp.gmCam + d * kp.gmD + h * kp.gmY + w * kp.gmX) != 0;
      }
      make_entry(kp, gx, gy, off, lx, ly);
      off += q * kp.fsCam;
      if (AGG == MVSGI_AGG_STD && gvalid) info = make_mask_info(gx, gy, kp.Hm, kp.Wm, kp.mPitch);
    }
    if (AGG == MVSGI_AGG_STD) {
      const uint32_t* mb = kp.mbits + (kp.perFrameMasks ? (long long)b * kp.mbB : 0) + q * kp.mbCam;
      bit = mask_eval(mask_fetch(mb, info, kp.mPitch), info);  // spherical_sweep_avg.py:92-102
    } else {
      bit = 1u;
    }
  }
  int offs[CAM];
  float lxs[CAM], lys[CAM];
  uint32_t v = 0;
#pragma unroll
  for (int cam = 0; cam < CAM; ++cam) {
    const int src = (lane & ~3) + cam;
    offs[cam] = __shfl_sync(0xFFFFFFFFu, off, src);
    lxs[cam] = __shfl_sync(0xFFFFFFFFu, lx, src);
    lys[cam] = __shfl_sync(0xFFFFFFFFu, ly, src);
    v |= __shfl_sync(0xFFFFFFFFu, bit, src) << cam;
  }
  (void)qmask;
  const int n = __popc(v);
  const bool need = act && (AGG == MVSGI_AGG_CONCAT || n > 1);  // :106-111, :125
  if (!need) return;

  const int cs = kp.Dvol * HW;
  const int cLane = q * 4;
  const float* fb = kp.feat + (long long)b * kp.fsB + cLane;
  float* gb = kb.gradStage + (long long)b * kp.fsB + cLane;
  const float* gv = kb.gradVol + ((long long)b * kp.Cout + cLane) * cs + (long long)(d + kp.dOff) * HW + pix;
  const float inv = 1.0f / (float)n;

  for (int cg = 0; cg < kp.CG; ++cg) {
    const int nst = min(4, kp.C - (cg * kCG + cLane));  // staged channels >= C are padding
    if (nst <= 0) continue;
    float g[CAM][4];
    if (AGG == MVSGI_AGG_STD) {
      float f[CAM][4];
#pragma unroll
      for (int cam = 0; cam < CAM; ++cam) {
        if ((v >> cam) & 1u) {
          const float* t = fb + (long long)cg * kp.fsCg + offs[cam];
          const float4 ta = ldg4(t), tc = ldg4(t + kCG), tb = ldg4(t + kp.fsY), td = ldg4(t + kp.fsY + kCG);
          const float hx = 1.0f - lxs[cam], hy = 1.0f - lys[cam];
          const float wa = hx * hy, wb = hx * lys[cam], wc = lxs[cam] * hy, wd = lxs[cam] * lys[cam];
          f[cam][0] = fmaf(td.x, wd, fmaf(tc.x, wc, fmaf(tb.x, wb, ta.x * wa)));
          f[cam][1] = fmaf(td.y, wd, fmaf(tc.y, wc, fmaf(tb.y, wb, ta.y * wa)));
          f[cam][2] = fmaf(td.z, wd, fmaf(tc.z, wc, fmaf(tb.z, wb, ta.z * wa)));
          f[cam][3] = fmaf(td.w, wd, fmaf(tc.w, wc, fmaf(tb.w, wb, ta.w * wa)));
        } else {
          f[cam][0] = f[cam][1] = f[cam][2] = f[cam][3] = 0.f;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float s = f[0][i];
#pragma unroll
        for (int cam = 1; cam < CAM; ++cam) s += f[cam][i];
        const float avg = s * inv;
        const float go = i < nst ? __ldg(gv + ((long long)cg * kCG + i) * cs) : 0.f;
        const float k2 = 2.0f * go * inv;
#pragma unroll
        for (int cam = 0; cam < CAM; ++cam) g[cam][i] = (f[cam][i] - avg) * k2;
      }
    } else {
#pragma unroll
      for (int cam = 0; cam < CAM; ++cam)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          g[cam][i] = i < nst ? __ldg(gv + ((long long)cam * kp.C + cg * kCG + i) * cs) : 0.f;
    }
#pragma unroll
    for (int cam = 0; cam < CAM; ++cam) {
      if (!((v >> cam) & 1u)) continue;
      if (offs[cam] == kp.zeroOff + cam * kp.fsCam) continue;  // sample entirely outside: nothing to accumulate
      float* t = gb + (long long)cg * kp.fsCg + offs[cam];
      const float hx = 1.0f - lxs[cam], hy = 1.0f - lys[cam];
      const float wa = hx * hy, wb = hx * lys[cam], wc = lxs[cam] * hy, wd = lxs[cam] * lys[cam];
      red_add4(t, g[cam][0] * wa, g[cam][1] * wa, g[cam][2] * wa, g[cam][3] * wa);
      red_add4(t + kCG, g[cam][0] * wc, g[cam][1] * wc, g[cam][2] * wc, g[cam][3] * wc);
      red_add4(t + kp.fsY, g[cam][0] * wb, g[cam][1] * wb, g[cam][2] * wb, g[cam][3] * wb);
      red_add4(t + kp.fsY + kCG, g[cam][0] * wd, g[cam][1] * wd, g[cam][2] * wd, g[cam][3] * wd);
    }
  }
}

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
