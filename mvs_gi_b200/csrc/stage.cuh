// stage.cuh — streaming pre-passes of the sweep (HBM-bound, one read + one write each).
//
//   stage_feats_kernel : feats [B, CAM, C, Hf, Wf] (any element strides; NCHW from cuDNN or channels-last) ->
//                        zero-padded, channel-group-major, channels-last fp32 images [B*CAM][CG][Hf+3][Wf+2][16].
//                        The zero frame is the F.pad of the reference sampler (dsta_mvs/model/backports/backports.py:58);
//                        row Hf+2 is an extra all-zero row (target of far-outside samples); channels >= C are zero.
//   pack_masks_kernel  : masks [.., Hm, Wm] (float32 0/1 or u8/bool) -> padded bit images, word j of padded row Y holds
//                        bits X = 32j..32j+31, bit set iff mask(y = Y-1, x = X-1) > 0
//                        (the "bilinear(mask) > 0" test of spherical_sweep_avg.py:92-96 only needs those bits).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mvsgi {

constexpr int kStagePx = 128;     // padded pixels per block
constexpr int kStagePitch = 130;  // smem row pitch in floats (== 2 mod 32: conflict-free both ways)

__global__ void __launch_bounds__(256)
stage_feats_kernel(const float* __restrict__ src, long long sN0, int n1, long long sN1, long long sC, long long sH,
                   long long sW, int C, int CG, int Hf, int Wf, int xBlocks, int chanFastest,
                   float* __restrict__ dst) {
  __shared__ float tile[16 * kStagePitch];
  const int Wp = Wf + 2, Hp1 = Hf + 3;
  const int xb = blockIdx.x % xBlocks;
  const int ncg = blockIdx.x / xBlocks;
  const int n = ncg / CG, cg = ncg - n * CG;
  const int yp = blockIdx.y;
  const int xp0 = xb * kStagePx;
  const int t = threadIdx.x;
  float* drow = dst + (((long long)ncg * Hp1 + yp) * Wp + xp0) * 16;
  const int npx = min(kStagePx, Wp - xp0);

  const bool zero_row = (yp == 0) || (yp > Hf);
  if (!zero_row) {
    const float* s = src + (long long)(n / n1) * sN0 + (long long)(n % n1) * sN1 + (long long)(yp - 1) * sH;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int c, i;
      if (chanFastest) {  // channels-last source: 16 consecutive lanes read 16 consecutive channels
        const int idx = t + 256 * k;
        c = idx & 15;
        i = idx >> 4;
      } else {  // NCHW source: a warp reads 32 consecutive pixels of one channel
        c = (t >> 5) + 8 * (k & 1);
        i = (t & 31) + 32 * (k >> 1);
      }
      const int x = xp0 + i - 1, ch = cg * 16 + c;
      float v = 0.0f;
      if (x >= 0 && x < Wf && ch < C) v = __ldg(s + (long long)ch * sC + (long long)x * sW);
      tile[c * kStagePitch + i] = v;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int idx = t + 256 * k;  // float4 index inside the block's [px][16] output
    const int px = idx >> 2, c4 = idx & 3;
    if (px < npx) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!zero_row) {
        v.x = tile[(c4 * 4 + 0) * kStagePitch + px];
        v.y = tile[(c4 * 4 + 1) * kStagePitch + px];
        v.z = tile[(c4 * 4 + 2) * kStagePitch + px];
        v.w = tile[(c4 * 4 + 3) * kStagePitch + px];
      }
      *reinterpret_cast<float4*>(drow + (long long)px * 16 + c4 * 4) = v;
    }
  }
}

// One warp packs 8 consecutive words (256 mask pixels) per step with 8 independent loads per lane in flight.
// flags[0] bit0 is set when any mask value is negative or NaN (then the packed test is not equivalent to the
// reference's float test and the caller must reject the input).
template <typename T>
__global__ void __launch_bounds__(256)
pack_masks_kernel(const T* __restrict__ src, int n1, long long sN0, long long sN1, long long sH, long long sW, int Hm,
                  int Wm, int pitch, int wordsPerImage, uint32_t* __restrict__ bits, int* __restrict__ flags) {
  const int n = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * 8 + (threadIdx.x >> 5);  // warp index inside the image
  const int word0 = gw * 8;
  if (word0 >= wordsPerImage) return;
  const T* s = src + (long long)(n / n1) * sN0 + (long long)(n % n1) * sN1;
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int word = word0 + k;
    const int Y = word / pitch, wj = word - Y * pitch;
    const int x = wj * 32 + lane - 1, y = Y - 1;
    v[k] = 0.0f;
    if (word < wordsPerImage && x >= 0 && x < Wm && y >= 0 && y < Hm)
      v[k] = (float)s[(long long)y * sH + (long long)x * sW];
  }
  bool bad = false;
  uint32_t out = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    bad |= !(v[k] >= 0.0f);  // negative or NaN
    const uint32_t w = __ballot_sync(0xFFFFFFFFu, v[k] > 0.0f);
    if (lane == k) out = w;
  }
  if (lane < 8 && word0 + lane < wordsPerImage) bits[(long long)n * wordsPerImage + word0 + lane] = out;
  if (__any_sync(0xFFFFFFFFu, bad) && lane == 0 && flags != nullptr) atomicOr(flags, 1);
}

}  // namespace mvsgi
