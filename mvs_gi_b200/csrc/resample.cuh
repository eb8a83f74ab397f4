// resample.cuh — fisheye -> surrogate-ERP image resampling, the step before the feature extractor.
//
// Replaces CameraModelRotation.__init__/__call__ (dsta_mvs/image_sampler/camera_model_sampler.py:55-75, 98-120) as
// api/inference_pytorch.py:61-74 uses it, one camera and one image at a time, with a precomputed grid:
//   rays of the target (Equirectangular) pixels, pixel centre +0.5         planar_as_base.py:200-201, camera_models.py:321-326, 680-720
//   x_raw = R_raw_fisheye @ x_target                                        planar_as_base.py:206-209
//   raw camera point_3d_2_pixel(normalized=True) -> grid                    camera_model_sampler.py:58-64
//   invalid (raw or target) -> grid = -1, output = invalid value            :67-71, :117
//   F.grid_sample(bilinear, zeros, align_corners=False)                     :110-114, planar_as_base.py:124-133
//   interpolation='nearest': F.grid_sample(nearest, zeros, align_corners=True)   :98-120, planar_as_base.py:127-133
// Here the geometry is evaluated in-kernel from the same MvsgiRig the sweep uses (target model in the cv_* fields,
// cams[0] = raw model with rot = R_raw_fisheye, trans = 0, one candidate of distance 1), so no grid is read from HBM,
// and a batch of N images (all cameras sharing a model/orientation, or consecutive frames) is resampled in one launch.
// HBM-bound: C*Ht*Wt*4 bytes written + the touched part of C*H*W*4 read per image.
#pragma once
#include "geometry.cuh"

namespace mvsgi {

struct ResampleKP {
  int N, C, H, W, Ht, Wt;
  const void* img;     // float32, or uint8 scaled by 1/255 on load (api/inference_pytorch.py:56-57)
  long long sN, sC, sH, sW;
  float invalid;
  int nearest;         // MVSGI_RESAMPLE_NEAREST
  float* out;          // [N, C, Ht, Wt] contiguous
  uint8_t* validOut;   // [Ht, Wt] or nullptr
};

constexpr int kResampleThreads = 256;

MVSGI_DEV float ld_px(const float* p) { return __ldg(p); }
// uint8 pixels: the reference converts on the GPU with `t_imgs.float() / 255.0` (api/inference_pytorch.py:56-57), and
// PyTorch's CUDA division by a host scalar multiplies by the float32 reciprocal, so that is the arithmetic matched
// here bit for bit (tests/test_resample.py); the true quotient differs by at most 1 ulp.
MVSGI_DEV float ld_px(const uint8_t* p) { return fmul((float)__ldg(p), __uint_as_float(0x3B808081u)); }  // float32(1/255)

template <int CU, typename T>  // CU channels handled together (3 for RGB); T = element type of the raw image
__global__ void __launch_bounds__(kResampleThreads)
resample_kernel(const ResampleKP kp, const __grid_constant__ MvsgiRig rig) {
  // a block is a 32 x 8 tile of the target image: its rays fall into one compact patch of the raw image (L1 reuse
  // between the rows of the tile), and every warp still stores 32 consecutive pixels
  const int P = kp.Ht * kp.Wt;
  const int tilesX = (kp.Wt + 31) >> 5;
  const int ty = blockIdx.x / tilesX, tx = blockIdx.x - ty * tilesX;
  // float images: inside the tile a warp is an 8 x 4 patch (not a 32 x 1 strip), so its 32 rays land in a compact
  // patch of the raw image whatever the direction of the fisheye arc (about half the sectors per request; measured
  // 298 -> 271 us).  uint8 HWC images (3-byte pixels, a whole neighbourhood per sector) measure faster as strips.
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const bool patch = sizeof(T) == 4;
  const int h = ty * (kResampleThreads / 32) + (patch ? (wrp >> 2) * 4 + (lane >> 3) : wrp);
  const int w = tx * 32 + (patch ? (wrp & 3) * 8 + (lane & 7) : lane);
  if (h >= kp.Ht || w >= kp.Wt) return;
  const int pix = h * kp.Wt + w;
  float rx, ry, rz, x, y, z, gx, gy;
  cv_ray(rig, h, w, rx, ry, rz);
  to_camera(rig.cams[0], rx, ry, rz, 1.0f, x, y, z);
  const bool valid = project(rig.cams[0], x, y, z, gx, gy);
  if (kp.validOut != nullptr && blockIdx.y == 0) kp.validOut[pix] = valid ? 1 : 0;
  if (kp.nearest) {
    // F.grid_sample(mode='nearest', zeros, align_corners=True): camera_model_sampler.py:98-120, planar_as_base.py:127-133;
    // ATen: ((g + 1) / 2) * (size - 1), nearbyint (half to even), 0 outside the image (NaN compares false)
    const float xr = rintf(fmul(fdiv(fadd(gx, 1.0f), 2.0f), (float)(kp.W - 1)));
    const float yr = rintf(fmul(fdiv(fadd(gy, 1.0f), 2.0f), (float)(kp.H - 1)));
    const bool inb = valid && xr >= 0.0f && xr <= (float)(kp.W - 1) && yr >= 0.0f && yr <= (float)(kp.H - 1);
    const long long on = inb ? (long long)(int)yr * kp.sH + (long long)(int)xr * kp.sW : 0;
    for (int n = blockIdx.y; n < kp.N; n += gridDim.y) {
      const T* im = static_cast<const T*>(kp.img) + n * kp.sN + on;
      float* o = kp.out + (long long)n * kp.C * P + pix;
      for (int c = 0; c < kp.C; ++c) {
        const float v = valid ? (inb ? ld_px(im + (long long)c * kp.sC) : 0.0f) : kp.invalid;
        asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(o + (long long)c * P), "f"(v) : "memory");
      }
    }
    return;
  }
  // grid_sample, align_corners=False, zero padding
  const float fx = unnormalise(gx, kp.W), fy = unnormalise(gy, kp.H);
  const float x0f = floorf(fx), y0f = floorf(fy);
  const float lx = fsub(fx, x0f), ly = fsub(fy, y0f);
  const float hx = fsub(fadd(x0f, 1.0f), fx), hy = fsub(fadd(y0f, 1.0f), fy);
  const float wnw = fmul(hx, hy), wne = fmul(lx, hy), wsw = fmul(hx, ly), wse = fmul(lx, ly);
  // out-of-range coordinates (incl. NaN) take no tap; the int conversion is only used when in range
  const bool inr = valid && fx > -1.0f && fx < (float)kp.W && fy > -1.0f && fy < (float)kp.H;
  const int x0 = inr ? (int)x0f : 0, y0 = inr ? (int)y0f : 0;
  const bool bx0 = inr && x0 >= 0, bx1 = inr && x0 + 1 < kp.W, by0 = inr && y0 >= 0, by1 = inr && y0 + 1 < kp.H;
  const long long o00 = (long long)y0 * kp.sH + (long long)x0 * kp.sW;
  const bool t00 = bx0 && by0, t01 = bx1 && by0, t10 = bx0 && by1, t11 = bx1 && by1;
  for (int n = blockIdx.y; n < kp.N; n += gridDim.y) {
    const T* im = static_cast<const T*>(kp.img) + n * kp.sN + o00;
    float* o = kp.out + (long long)n * kp.C * P + pix;
    for (int c0 = 0; c0 < kp.C; c0 += CU) {
      // all taps of CU channels in flight before any arithmetic
      float nw[CU], ne[CU], sw[CU], se[CU];
#pragma unroll
      for (int u = 0; u < CU; ++u) {
        const bool on = valid && c0 + u < kp.C;
        const T* r0 = im + (long long)(c0 + u) * kp.sC;
        nw[u] = (on && t00) ? ld_px(r0) : 0.0f;
        ne[u] = (on && t01) ? ld_px(r0 + kp.sW) : 0.0f;
        sw[u] = (on && t10) ? ld_px(r0 + kp.sH) : 0.0f;
        se[u] = (on && t11) ? ld_px(r0 + kp.sH + kp.sW) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < CU; ++u) {
        if (c0 + u < kp.C) {
          const float v = valid ? fadd(fadd(fadd(fmul(nw[u], wnw), fmul(ne[u], wne)), fmul(sw[u], wsw)), fmul(se[u], wse))
                                : kp.invalid;
          asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(o + (long long)(c0 + u) * P), "f"(v) : "memory");
        }
      }
    }
  }
}

}  // namespace mvsgi
