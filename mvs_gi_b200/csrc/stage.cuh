// stage.cuh — streaming pre-passes of the sweep (HBM-bound, one read + one write each).
//
//   stage_feats_kernel : feats [B, CAM, C, Hf, Wf] (any element strides; NCHW from cuDNN or channels-last) ->
//                        zero-padded, channel-group-major, channels-last fp32 images [B*CAM][CG][Hf+3][Wf+2][16].
//                        The zero frame is the F.pad of the reference sampler (dsta_mvs/model/backports/backports.py:58);
//                        row Hf+2 is an extra all-zero row (target of far-outside samples); channels >= C are zero.
//   pack_masks_kernel  : masks [.., Hm, Wm] (float32 0/1 or u8/bool) -> zero-framed bit images, bit set iff mask > 0
//                        (the "bilinear(mask) > 0" test of spherical_sweep_avg.py:92-96 only needs those bits).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mvsgi {

constexpr int kStagePx = 544;     // padded pixels per block (one block covers a whole 514-pixel padded row)
constexpr int kStagePitch = 546;  // smem row pitch in floats (== 2 mod 8: conflict-free both ways)

// BF16 (MVSGI_FLAG_FEATS_BF16): the staged image holds bfloat16 (round to nearest even), 32 bytes per pixel; the
// sweep converts back to fp32 and accumulates in fp32 (stated-tolerance mode, SURVEY.md §8c criterion 4).
template <bool BF16>
__global__ void __launch_bounds__(256)
stage_feats_kernel(const float* __restrict__ src, long long sN0, int n1, long long sN1, long long sC, long long sH,
                   long long sW, int C, int CG, int Hf, int Wf, int xBlocks, int chanFastest,
                   float* __restrict__ dst_) {
  __shared__ float tile[16 * kStagePitch];
  const int Wp = Wf + 2, Hp1 = Hf + 3;
  const int xb = blockIdx.x % xBlocks;
  const int ncg = blockIdx.x / xBlocks;
  const int n = ncg / CG, cg = ncg - n * CG;
  const int yp = blockIdx.y;
  const int xp0 = xb * kStagePx;
  const int t = threadIdx.x;
  const long long rowOff = (((long long)ncg * Hp1 + yp) * Wp + xp0) * 16;  // elements
  float4* drow = reinterpret_cast<float4*>(dst_ + rowOff);
  uint2* drow16 = reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(dst_) + rowOff);
  const int npx = min(kStagePx, Wp - xp0);

  if ((yp == 0) || (yp > Hf)) {  // zero frame rows
    for (int idx = t; idx < npx * 4; idx += 256) {
      if (BF16) drow16[idx] = make_uint2(0u, 0u);
      else drow[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  const float* s = src + (long long)(n / n1) * sN0 + (long long)(n % n1) * sN1 + (long long)(yp - 1) * sH;
  if (chanFastest) {  // channels-last source: 16 consecutive lanes read the 16 channels of one pixel
#pragma unroll 4
    for (int idx = t; idx < npx * 16; idx += 256) {
      const int c = idx & 15, i = idx >> 4;
      const int x = xp0 + i - 1, ch = cg * 16 + c;
      float v = 0.0f;
      if (x >= 0 && x < Wf && ch < C) v = __ldg(s + (long long)ch * sC + (long long)x * sW);
      tile[c * kStagePitch + i] = v;
    }
  } else {  // NCHW source: a warp reads 32 consecutive pixels of one channel; all loads of a thread are independent
    const int lane = t & 31, wrp = t >> 5;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int c = wrp + 8 * cc, ch = cg * 16 + c;
      const float* sc = s + (long long)ch * sC;
#pragma unroll
      for (int k = 0; k < kStagePx / 32; ++k) {
        const int i = lane + 32 * k;
        const int x = xp0 + i - 1;
        float v = 0.0f;
        if (i < npx && x >= 0 && x < Wf && ch < C) v = __ldg(sc + (long long)x * sW);
        tile[c * kStagePitch + i] = v;
      }
    }
  }
  __syncthreads();
  for (int idx = t; idx < npx * 4; idx += 256) {  // float4 index inside the block's [px][16] output
    const int px = idx >> 2, c4 = idx & 3;
    float4 v;
    v.x = tile[(c4 * 4 + 0) * kStagePitch + px];
    v.y = tile[(c4 * 4 + 1) * kStagePitch + px];
    v.z = tile[(c4 * 4 + 2) * kStagePitch + px];
    v.w = tile[(c4 * 4 + 3) * kStagePitch + px];
    if (BF16) {
      const uint32_t lo = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v.x)) |
                          ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v.y)) << 16);
      const uint32_t hi = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v.z)) |
                          ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v.w)) << 16);
      drow16[idx] = make_uint2(lo, hi);
    } else {
      drow[idx] = v;
    }
  }
}

// One block packs one padded row of one mask image: every lane reads 4 consecutive mask pixels (one 16-byte load when
// VEC), turns them into a nibble, and 8 lanes OR their nibbles into a word with three shuffles.
// Bit layout: padded row Y = y + 1 (rows 0 and Hm + 1 are zero), rows stored as interleaved pairs (below); word 0 of
// a row is a zero guard, pixel x is bit
// (x & 31) of word 1 + (x >> 5), and at least one zero guard word follows, so taps x0 = -1 and x0 + 1 = Wm read zeros.
// flags[0] bit0 is set when any mask value is negative or NaN (then the packed test is not equivalent to the
// reference's float test and the caller must reject the input).
template <typename T>
struct Vec4Of;
template <>
struct Vec4Of<float> {
  using type = float4;
};
template <>
struct Vec4Of<uint8_t> {
  using type = uchar4;
};

// masks are read exactly once: keep them out of L1 and mark them evict-first in L2
__device__ __forceinline__ float4 ld_stream4(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ uchar4 ld_stream4(const uchar4* p) { return __ldcs(p); }

constexpr int kPackThreads = 128;

template <typename T, bool VEC>
__global__ void __launch_bounds__(kPackThreads)
pack_masks_kernel(const T* __restrict__ src, int n1, long long sN0, long long sN1, long long sH, long long sW, int Hm,
                  int Wm, int pitch, uint32_t* __restrict__ bits, int* __restrict__ flags) {
  const int n = blockIdx.y, Y = blockIdx.x, t = threadIdx.x, lane = t & 31;
  // row-pair interleaved: word j of padded row Y is stored as the TOP of pair Y and as the BOTTOM of pair Y - 1, so
  // the two rows a bilinear mask sample needs are one aligned 8-byte load (sweep.cuh, mask_fetch)
  uint32_t* img = bits + (long long)n * (Hm + 2) * pitch * 2;
  auto put = [&](int j, uint32_t w) {
    img[((long long)Y * pitch + j) * 2] = w;
    if (Y > 0) img[((long long)(Y - 1) * pitch + j) * 2 + 1] = w;
  };
  const int dataWords = (Wm + 31) >> 5;
  if (Y == 0 || Y == Hm + 1) {
    for (int j = t; j < pitch; j += kPackThreads) {
      put(j, 0u);
      if (Y == Hm + 1) img[((long long)Y * pitch + j) * 2 + 1] = 0u;  // bottom of the last pair: never a real row
    }
    return;
  }
  if (t == 0) put(0, 0u);
  for (int j = 1 + dataWords + t; j < pitch; j += kPackThreads) put(j, 0u);
  const T* s = src + (long long)(n / n1) * sN0 + (long long)(n % n1) * sN1 + (long long)(Y - 1) * sH;
  bool bad = false;
  constexpr int U = 4;  // row segments per thread with all loads issued before the first use (kPackThreads * 16 pixels per pass)
  for (int xb0 = (t & ~31) * 4; xb0 < dataWords * 32; xb0 += kPackThreads * 4 * U) {  // warp-uniform trip count (shuffles below)
    float v[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int x0 = xb0 + u * (kPackThreads * 4) + lane * 4;
      v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0.f;
      if (VEC) {
        if (x0 < Wm) {  // Wm % 4 == 0 on this path
          const typename Vec4Of<T>::type q = ld_stream4(reinterpret_cast<const typename Vec4Of<T>::type*>(s + x0));
          v[u][0] = (float)q.x; v[u][1] = (float)q.y; v[u][2] = (float)q.z; v[u][3] = (float)q.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (x0 + i < Wm) v[u][i] = (float)s[(long long)(x0 + i) * sW];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int x0 = xb0 + u * (kPackThreads * 4) + lane * 4;
      uint32_t nib = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        bad |= !(v[u][i] >= 0.0f);  // negative or NaN
        nib |= (v[u][i] > 0.0f ? 1u : 0u) << i;
      }
      uint32_t w = nib << (4 * (lane & 7));
      w |= __shfl_xor_sync(0xFFFFFFFFu, w, 1);
      w |= __shfl_xor_sync(0xFFFFFFFFu, w, 2);
      w |= __shfl_xor_sync(0xFFFFFFFFu, w, 4);
      if ((lane & 7) == 0 && (x0 >> 5) < dataWords) put(1 + (x0 >> 5), w);
    }
  }
  if (__any_sync(0xFFFFFFFFu, bad) && lane == 0 && flags != nullptr) atomicOr(flags, 1);
}

}  // namespace mvsgi
