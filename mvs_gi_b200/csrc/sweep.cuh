// sweep.cuh — fused spherical sweep + cross-view aggregation for sm_100a.
//
// One persistent kernel replaces the reference's per-candidate loop of
// F.pad / 4x torch.gather / where / sum (~4.7k ATen calls per frame):
//   dsta_mvs/model/cost_volume_builder/spherical_sweep_avg.py:38-136  (std / variance, masked)
//   dsta_mvs/model/cost_volume_builder/spherical_sweep.py:38-68       (concat)
//   dsta_mvs/model/backports/backports.py:11-86                       (bilinear_grid_sample)
//
// Work decomposition
//   item  = (output tile of tileH x 32 pixels, chunk of `dch` candidates, frame b)
//   CTA   = persistent; owns a contiguous range of items ordered (tile, chunk, frame) so that
//           consecutive items share their geometry.
//   phase 1  (once per (tile, chunk) when grids are batch-shared / projected in-kernel; once per item otherwise):
//           evaluate or load the sampling grid and turn it into a *tap table* in shared memory:
//           per (candidate, camera, pixel) one int32 element offset of the top-left tap plus the four
//           bilinear weights, zero-padding already folded in (out-of-image taps get weight 0 and an
//           in-bounds address).  The std builder also resolves the mask test there into 1 bit per camera.
//   phase 2  4 lanes per output pixel, lane q owns channels [4q, 4q+4) of the current 16-channel group:
//           every tap is one 16-byte load from the channels-last feature map (4 lanes = 64 contiguous bytes
//           per pixel); the candidate loop runs innermost so the sliding per-camera footprint of the tile
//           stays in L1 from one candidate to the next; mean/variance across cameras happens in
//           registers; per-view warped volumes are never materialised; the (C, D, H, W) volume is written
//           once with streaming stores, 8 consecutive pixels = one full 32-byte sector per warp per channel.
#pragma once
#include "geometry.cuh"

namespace mvsgi {

constexpr int kTileW = 32;          // output pixels per tile row
constexpr int kLanesPerPixel = 4;   // 4 lanes x float4 = 16 channels
constexpr uint32_t kMaskNone = 0xFFFFFFFFu;

struct SweepKP {
  int B, C, D, Cout;
  int Ho, Wo, Hf, Wf, Hm, Wm;
  int tileH, dch, numChunks, tilesX, numTiles;
  long long totalItems;
  // channels-last feature view: element (b, cam, y, x, c) at feat[b*fsB + cam*fsCam + y*fsY + x*fsX + c]
  const float* feat;
  long long fsB, fsCam;
  int fsY, fsX;
  // packed masks (std): uint32 words, image (b, cam) at mbits + b*mbB + cam*mbCam, (Hm+2) rows of mPitch words
  const uint32_t* mbits;
  long long mbB, mbCam;
  int mPitch;
  int perFrameMasks;
  // supplied grids: (b, cam, d, h, w) -> 2 floats ; grid_masks -> 1 byte
  int gridMode, perFrameGrids;
  const float* grids;
  long long gsB, gsCam, gsD, gsY, gsX;
  const uint8_t* gmask;
  long long gmB, gmCam, gmD, gmY, gmX;
  float* vol;
  uint8_t* validOut;
};

// ---- tap table entry --------------------------------------------------------------------------------------------
struct TapEntry {
  int off;             // element offset of tap (y_base, x_base) inside one (b, cam) feature image
  float wa, wb, wc, wd;  // weights of (y_base,x_base), (y_base+1,x_base), (y_base,x_base+1), (y_base+1,x_base+1)
};

// One axis of bilinear_grid_sample with zero padding folded in.
//   v       continuous pixel coordinate (backports.py:41-42)
//   base    first of the two in-bounds pixels actually read
//   w0, w1  weights of pixel base and base+1  (== (x1-x), (x-x0) of backports.py:52-55 when both are in bounds)
MVSGI_DEV void axis_taps(float v, int size, int& base, float& w0, float& w1) {
  const float f0 = floorf(v);
  base = 0;
  w0 = 0.0f;
  w1 = 0.0f;
  // also rejects NaN; taps entirely in the zero padding contribute exactly 0 (backports.py:58-72)
  if (!(v >= -1.0f && v < (float)size)) return;
  const int i0 = (int)f0;
  const float hi = fsub(fadd(f0, 1.0f), v);  // (x1 - x)
  const float lo = fsub(v, f0);              // (x - x0)
  if (i0 < 0) {               // x0 = -1 is padding, x1 = 0 is pixel 0
    base = 0; w0 = lo; w1 = 0.0f;
  } else if (i0 >= size - 1) {  // x0 = size-1 is the last pixel, x1 is padding
    base = size - 2; w0 = 0.0f; w1 = hi;
  } else {
    base = i0; w0 = hi; w1 = lo;
  }
}

MVSGI_DEV TapEntry make_tap_entry(float gx, float gy, int Hf, int Wf, int fsY, int fsX) {
  int bx, by;
  float cw0, cw1, rw0, rw1;
  axis_taps(unnormalise(gx, Wf), Wf, bx, cw0, cw1);
  axis_taps(unnormalise(gy, Hf), Hf, by, rw0, rw1);
  TapEntry e;
  e.off = by * fsY + bx * fsX;
  e.wa = fmul(cw0, rw0);
  e.wb = fmul(cw0, rw1);
  e.wc = fmul(cw1, rw0);
  e.wd = fmul(cw1, rw1);
  return e;
}

// Mask test of spherical_sweep_avg.py:92-96 for 0/1 (non-negative) masks: bilinear(mask) > 0  <=>  some tap has
// weight > 0 and a set mask bit.  Encodes the padded bit position of tap (y0, x0) and which of the other three taps
// carry a non-zero weight.  kMaskNone = no tap can contribute.
MVSGI_DEV uint32_t make_mask_info(float gx, float gy, int Hm, int Wm, int mPitch) {
  const float x = unnormalise(gx, Wm), y = unnormalise(gy, Hm);
  if (!(x >= -1.0f && x < (float)Wm && y >= -1.0f && y < (float)Hm)) return kMaskNone;
  const float fx0 = floorf(x), fy0 = floorf(y);
  const float hx = fsub(fadd(fx0, 1.0f), x), lx = fsub(x, fx0);
  const float hy = fsub(fadd(fy0, 1.0f), y), ly = fsub(y, fy0);
  const int X0 = (int)fx0 + 1, Y0 = (int)fy0 + 1;  // padded coordinates
  uint32_t info = (uint32_t)(Y0 * mPitch * 32 + X0);
  if (fmul(hx, ly) > 0.0f) info |= 1u << 28;  // wb: (x0, y1)
  if (fmul(lx, hy) > 0.0f) info |= 1u << 29;  // wc: (x1, y0)
  if (fmul(lx, ly) > 0.0f) info |= 1u << 30;  // wd: (x1, y1)
  // wa = (x1-x)(y1-y) is always > 0 here
  return info;
}

MVSGI_DEV bool mask_lookup(const uint32_t* __restrict__ img, uint32_t info, int mPitch) {
  if (info == kMaskNone) return false;
  const uint32_t pos = info & 0x0FFFFFFFu;
  const uint32_t word = pos >> 5, sh = pos & 31u;
  const uint32_t r0 = __funnelshift_r(__ldg(img + word), __ldg(img + word + 1), sh);
  const uint32_t r1 = __funnelshift_r(__ldg(img + word + mPitch), __ldg(img + word + mPitch + 1), sh);
  uint32_t hit = r0 & 1u;                 // a
  hit |= (r1 & 1u) & (info >> 28);        // b
  hit |= ((r0 >> 1) & 1u) & (info >> 29); // c
  hit |= ((r1 >> 1) & 1u) & (info >> 30); // d
  return (hit & 1u) != 0;
}

MVSGI_DEV float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

MVSGI_DEV float4 bilerp(const float4 a, const float4 b, const float4 c, const float4 d, float wa, float wb, float wc,
                        float wd) {
  // backports.py:86  Ia*wa + Ib*wb + Ic*wc + Id*wd (left to right; contraction into FMAs allowed)
  float4 r;
  r.x = fmaf(d.x, wd, fmaf(c.x, wc, fmaf(b.x, wb, a.x * wa)));
  r.y = fmaf(d.y, wd, fmaf(c.y, wc, fmaf(b.y, wb, a.y * wa)));
  r.z = fmaf(d.z, wd, fmaf(c.z, wc, fmaf(b.z, wb, a.z * wa)));
  r.w = fmaf(d.w, wd, fmaf(c.w, wc, fmaf(b.w, wb, a.w * wa)));
  return r;
}

// ---- the kernel ---------------------------------------------------------------------------------------------------
template <int CAM, int AGG>
__global__ void __launch_bounds__(512, 1)
sweep_kernel(const SweepKP kp, const __grid_constant__ MvsgiRig rig) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int TP = kp.tileH * kTileW;            // pixels per tile
  const int nEnt = kp.dch * CAM * TP;          // tap-table entries
  int* tOff = reinterpret_cast<int*>(smem_raw);
  float* tW = reinterpret_cast<float*>(tOff + nEnt);       // 4 planes of nEnt
  uint32_t* tM = reinterpret_cast<uint32_t*>(tW + 4 * nEnt);  // only when masks are per frame
  unsigned char* vb = reinterpret_cast<unsigned char*>(tM + (kp.perFrameMasks ? nEnt : 0));  // dch * TP

  const int tid = threadIdx.x;
  // phase-1 mapping: one pixel per thread, candidates strided over the 4 thread groups
  const int p1 = tid % TP, g1 = tid / TP;
  // phase-2 mapping: 4 lanes per pixel
  const int p2 = tid >> 2, q = tid & 3;

  const long long i0 = kp.totalItems * blockIdx.x / gridDim.x;
  const long long i1 = kp.totalItems * (blockIdx.x + 1) / gridDim.x;
  long long curTc = -1;

  for (long long item = i0; item < i1; ++item) {
    const int b = (int)(item % kp.B);
    const long long tc = item / kp.B;
    const int chunk = (int)(tc % kp.numChunks);
    const int tile = (int)(tc / kp.numChunks);
    const int ty = tile / kp.tilesX, tx = tile % kp.tilesX;
    const int d0 = chunk * kp.dch;
    const int dcount = min(kp.dch, kp.D - d0);
    const bool newGeom = (tc != curTc) || kp.perFrameGrids;
    curTc = tc;

    // ------------------------------------------------------------------ phase 1: geometry -> tap table
    {
      const int h = ty * kp.tileH + p1 / kTileW, w = tx * kTileW + p1 % kTileW;
      const bool inb = (h < kp.Ho) && (w < kp.Wo);
      const uint32_t* mimg0 = nullptr;
      if (AGG == MVSGI_AGG_STD) mimg0 = kp.mbits + (kp.perFrameMasks ? (long long)b * kp.mbB : 0);
      if (newGeom) {
        float rx = 0.f, ry = 0.f, rz = 1.f;
        if (kp.gridMode == MVSGI_GRID_PROJECT && inb) cv_ray(rig, h, w, rx, ry, rz);
        for (int dd = g1; dd < dcount; dd += kLanesPerPixel) {
          const int d = d0 + dd;
          uint32_t vbit = 0;
#pragma unroll
          for (int cam = 0; cam < CAM; ++cam) {
            float gx = 2.0f, gy = 2.0f;  // far outside: all-zero taps
            bool gvalid = false;
            if (inb) {
              if (kp.gridMode == MVSGI_GRID_PROJECT) {
                float x, y, z;
                to_camera(rig.cams[cam], rx, ry, rz, rig.dist[d], x, y, z);
                gvalid = project(rig.cams[cam], x, y, z, gx, gy);
              } else {
                const long long go = (long long)b * kp.gsB + cam * kp.gsCam + d * kp.gsD + h * kp.gsY + w * kp.gsX;
                const float2 g = __ldg(reinterpret_cast<const float2*>(kp.grids + go));
                gx = g.x;
                gy = g.y;
                if (AGG == MVSGI_AGG_STD)
                  gvalid = __ldg(kp.gmask + (long long)b * kp.gmB + cam * kp.gmCam + d * kp.gmD + h * kp.gmY +
                                 w * kp.gmX) != 0;
              }
            }
            const int e = (dd * CAM + cam) * TP + p1;
            const TapEntry te = make_tap_entry(gx, gy, kp.Hf, kp.Wf, kp.fsY, kp.fsX);
            tOff[e] = te.off;
            tW[e] = te.wa;
            tW[nEnt + e] = te.wb;
            tW[2 * nEnt + e] = te.wc;
            tW[3 * nEnt + e] = te.wd;
            if (AGG == MVSGI_AGG_STD) {
              const uint32_t info = gvalid ? make_mask_info(gx, gy, kp.Hm, kp.Wm, kp.mPitch) : kMaskNone;
              if (kp.perFrameMasks) tM[e] = info;
              else if (mask_lookup(mimg0 + cam * kp.mbCam, info, kp.mPitch)) vbit |= 1u << cam;
            }
          }
          if (AGG == MVSGI_AGG_STD && !kp.perFrameMasks) vb[dd * TP + p1] = (unsigned char)vbit;
        }
      }
      if (AGG == MVSGI_AGG_STD && kp.perFrameMasks) {
        // tM[(dd, cam, p1)] was written by this same thread above (identical dd striding): no barrier needed
        for (int dd = g1; dd < dcount; dd += kLanesPerPixel) {
          uint32_t vbit = 0;
#pragma unroll
          for (int cam = 0; cam < CAM; ++cam)
            if (mask_lookup(mimg0 + cam * kp.mbCam, tM[(dd * CAM + cam) * TP + p1], kp.mPitch)) vbit |= 1u << cam;
          vb[dd * TP + p1] = (unsigned char)vbit;
        }
      }
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase 2: gather + aggregate + store
    {
      const int h = ty * kp.tileH + p2 / kTileW, w = tx * kTileW + p2 % kTileW;
      const bool inb = (p2 < TP) && (h < kp.Ho) && (w < kp.Wo);
      const long long planeHW = (long long)kp.Ho * kp.Wo;
      const float* fb = kp.feat + (long long)b * kp.fsB;
      for (int cg = 0; cg < kp.C; cg += 16) {
        const int c0 = cg + q * 4;
        const bool cact = inb && (c0 < kp.C);
        const float* fq = fb + c0;
        for (int dd = 0; dd < dcount; ++dd) {
          const int d = d0 + dd;
          if (AGG == MVSGI_AGG_STD) {
            const uint32_t v = cact ? vb[dd * TP + p2] : 0u;
            const int n = __popc(v);
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n > 1) {  // spherical_sweep_avg.py:106-111
              float4 f[CAM];
#pragma unroll
              for (int cam = 0; cam < CAM; ++cam) {
                f[cam] = make_float4(0.f, 0.f, 0.f, 0.f);
                if ((v >> cam) & 1u) {
                  const int e = (dd * CAM + cam) * TP + p2;
                  const float* t = fq + cam * kp.fsCam + tOff[e];
                  f[cam] = bilerp(ldg4(t), ldg4(t + kp.fsY), ldg4(t + kp.fsX), ldg4(t + kp.fsY + kp.fsX), tW[e],
                                  tW[nEnt + e], tW[2 * nEnt + e], tW[3 * nEnt + e]);
                }
              }
              const float inv = 1.0f / (float)n;
              float4 s = f[0];
#pragma unroll
              for (int cam = 1; cam < CAM; ++cam) {
                s.x += f[cam].x; s.y += f[cam].y; s.z += f[cam].z; s.w += f[cam].w;
              }
              const float4 avg = make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv);  // :114
#pragma unroll
              for (int cam = 0; cam < CAM; ++cam) {
                if ((v >> cam) & 1u) {  // :119-122 masked-out cameras contribute (avg-avg)^2 = 0
                  const float dx = f[cam].x - avg.x, dy = f[cam].y - avg.y, dz = f[cam].z - avg.z,
                              dw = f[cam].w - avg.w;
                  o.x = fmaf(dx, dx, o.x); o.y = fmaf(dy, dy, o.y); o.z = fmaf(dz, dz, o.z); o.w = fmaf(dw, dw, o.w);
                }
              }
              o.x *= inv; o.y *= inv; o.z *= inv; o.w *= inv;
            }
            if (cact) {
              float* dst = kp.vol + (((long long)b * kp.Cout + c0) * kp.D + d) * planeHW + (long long)h * kp.Wo + w;
              const long long cs = (long long)kp.D * planeHW;
              __stcs(dst, o.x);
              __stcs(dst + cs, o.y);
              __stcs(dst + 2 * cs, o.z);
              __stcs(dst + 3 * cs, o.w);
              if (kp.validOut != nullptr && c0 == 0)
                kp.validOut[((long long)b * kp.D + d) * planeHW + (long long)h * kp.Wo + w] = (n > 1) ? 1 : 0;
            }
          } else {  // concat: spherical_sweep.py:52-67, out channel = cam*C + c, masks ignored
            if (cact) {
#pragma unroll
              for (int cam = 0; cam < CAM; ++cam) {
                const int e = (dd * CAM + cam) * TP + p2;
                const float* t = fq + cam * kp.fsCam + tOff[e];
                const float4 o = bilerp(ldg4(t), ldg4(t + kp.fsY), ldg4(t + kp.fsX), ldg4(t + kp.fsY + kp.fsX),
                                        tW[e], tW[nEnt + e], tW[2 * nEnt + e], tW[3 * nEnt + e]);
                float* dst = kp.vol + (((long long)b * kp.Cout + cam * kp.C + c0) * kp.D + d) * planeHW +
                             (long long)h * kp.Wo + w;
                const long long cs = (long long)kp.D * planeHW;
                __stcs(dst, o.x);
                __stcs(dst + cs, o.y);
                __stcs(dst + 2 * cs, o.z);
                __stcs(dst + 3 * cs, o.w);
              }
            }
          }
        }
      }
    }
    __syncthreads();  // the next item's phase 1 overwrites the table
  }
}

// ---- grid dump (mvsgi_project_only_f32): exactly the (gx, gy, valid) phase 1 computes ---------------------------
__global__ void project_only_kernel(const __grid_constant__ MvsgiRig rig, float* __restrict__ grids,
                                    uint8_t* __restrict__ gmask) {
  const int Ho = rig.cv_height, Wo = rig.cv_width;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Ho * Wo) return;
  const int h = idx / Wo, w = idx % Wo;
  float rx, ry, rz;
  cv_ray(rig, h, w, rx, ry, rz);
  for (int cam = 0; cam < rig.num_cams; ++cam)
    for (int d = 0; d < rig.num_cands; ++d) {
      float x, y, z, gx, gy;
      to_camera(rig.cams[cam], rx, ry, rz, rig.dist[d], x, y, z);
      const bool valid = project(rig.cams[cam], x, y, z, gx, gy);
      const long long o = ((long long)(cam * rig.num_cands + d) * Ho + h) * Wo + w;
      grids[o * 2] = gx;
      grids[o * 2 + 1] = gy;
      gmask[o] = valid ? 1 : 0;
    }
}

// ---- staging kernels ---------------------------------------------------------------------------------------------
// feats [N, C, H, W] (arbitrary element strides) -> channels-last [N, H, W, C] fp32.  32 pixels x C channels per block
// row, transposed through shared memory so both the read (along W) and the write (along C) are coalesced.
__global__ void __launch_bounds__(256)
nhwc_stage_kernel(const float* __restrict__ src, long long sN0, int n1, long long sN1, long long sC, long long sH,
                  long long sW, int C, int H, int W, float* __restrict__ dst) {
  extern __shared__ float tile[];  // [C][33]
  const int x0 = blockIdx.x * 32, y = blockIdx.y, n = blockIdx.z;
  const float* s = src + (long long)(n / n1) * sN0 + (long long)(n % n1) * sN1 + (long long)y * sH;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  for (int c = wrp; c < C; c += 8) {
    const int x = x0 + lane;
    tile[c * 33 + lane] = (x < W) ? __ldg(s + c * sC + x * sW) : 0.0f;
  }
  __syncthreads();
  const int c4n = C >> 2;  // float4 groups per pixel
  float* d = dst + (((long long)n * H + y) * W + x0) * C;
  for (int i = threadIdx.x; i < 32 * c4n; i += 256) {
    const int x = i / c4n, c4 = i % c4n;
    if (x0 + x < W) {
      float4 v;
      v.x = tile[(c4 * 4 + 0) * 33 + x];
      v.y = tile[(c4 * 4 + 1) * 33 + x];
      v.z = tile[(c4 * 4 + 2) * 33 + x];
      v.w = tile[(c4 * 4 + 3) * 33 + x];
      *reinterpret_cast<float4*>(d + (long long)x * C + c4 * 4) = v;
    }
  }
}

// masks -> padded bit images: word j of padded row Y holds bits X = 32j..32j+31, bit set iff mask(y=Y-1, x=X-1) > 0.
template <typename T>
__global__ void __launch_bounds__(256)
pack_masks_kernel(const T* __restrict__ src, int n1, long long sN0, long long sN1, long long sH, long long sW, int Hm,
                  int Wm, int pitch, uint32_t* __restrict__ bits, int* __restrict__ flags) {
  const int n = blockIdx.z, Y = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int word = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (word >= pitch) return;
  const int X = word * 32 + lane, x = X - 1, y = Y - 1;
  bool set = false, bad = false;
  if (x >= 0 && x < Wm && y >= 0 && y < Hm) {
    const float v = (float)src[(long long)(n / n1) * sN0 + (long long)(n % n1) * sN1 + (long long)y * sH + x * sW];
    set = v > 0.0f;
    bad = !(v >= 0.0f);  // negative or NaN
  }
  const uint32_t w = __ballot_sync(0xFFFFFFFFu, set);
  const uint32_t anybad = __ballot_sync(0xFFFFFFFFu, bad);
  if (lane == 0) {
    bits[((long long)n * (Hm + 2) + Y) * pitch + word] = w;
    if (anybad && flags != nullptr) atomicOr(flags, 1);
  }
}

}  // namespace mvsgi
