// sweep_tiled.cuh — the tile-walking sweep (round 2): same arithmetic as sweep.cuh, different work decomposition.
//
// Replaces (dsta_mvs/model/...): cost_volume_builder/spherical_sweep_avg.py:38-136, spherical_sweep.py:38-68,
// backports/backports.py:11-86 — bit-identical to sweep_kernel (sweep.cuh) on the same inputs.
//
// Why a second decomposition.  sweep_kernel keeps the tap table of a CTA's pixels (80 KB) in shared memory for the
// whole launch; two CTAs leave 68 KB of L1 for the taps, each CTA's pixels are scattered octets, and only 26 % of the
// tap sectors hit L1: 3.3 GB of L2->L1 traffic per 32-frame launch, three quarters of what the L2 can deliver
// (profiles/r1m_sweep_ncu.txt), at 350 M issued instructions.  Here
//   * the geometry of the whole volume is evaluated ONCE per launch by tile_table_kernel into a global table laid
//     out as one contiguous chunk per (tile, candidate): per voxel and camera a 16-byte record (tap offset + edge
//     flags, the two column weights, frac y), plus the mask lookup words / the static validity bits.  The table
//     (20-26 MB at G16V) stays in the 126 MB L2 for all frames; a warp reads its records with three 16-byte loads
//     per candidate, one step ahead of their use, straight into registers (L1::no_allocate).  No shared memory:
//     the whole 228 KB of the SM is L1 for the taps.
//   * a tile is a compact block of output pixels (kTileWarps octets) and a CTA's warps walk its candidates in order: the
//     sampling footprint of the tile slides along the epipolar direction by a few pixels per candidate, so what a
//     candidate step needs is mostly what the previous one left in L1.  CTAs take (frame group, tile, candidate
//     range) items from an atomic queue in frame-major order, so a frame's features are read from HBM once.
//   * FB frames per pass (template): records, addresses and bilinear weights are decoded once and applied to FB
//     frames, whose taps are all in flight together.
//   * packed fp32 arithmetic (FMUL2 / FFMA2 / FADD2, two channels per instruction, round-to-nearest each): half the
//     issue slots of the scalar form for the same bits.
//   * the two quads of a quarter-warp (the 8 lanes one LDG.128 wavefront serves) always read 64-byte chunks in
//     opposite halves of the 128-byte bank line: of its two patch columns (x0, x0 + 1) every quad loads first the
//     one whose address bit 6 equals the quad's parity (scripts/microbench/tap_banks.cu: 3.9 clk per request instead
//     of 5.3, L1-resident).  Column order does not change the result bits (one commutative add joins the columns).
//   * features that are already channels-last (16 k channels, north_star (a)) are tapped in the caller's tensor: the
//     record carries clamped offsets and edge flags instead of relying on a zero frame, so nothing is staged.
#pragma once
#include "sweep.cuh"

namespace mvsgi {

#ifndef MVSGI_TILE_WARPS
#define MVSGI_TILE_WARPS 8
#endif
#ifndef MVSGI_TILE_CTAS
#define MVSGI_TILE_CTAS 3
#endif
constexpr int kTileWarps = MVSGI_TILE_WARPS;  // warps per CTA = octets per tile
constexpr int kTileVox = kTileWarps * 8;       // voxels per (tile, candidate) chunk
constexpr int kTileThreads = kTileWarps * 32;
constexpr int kTileCtas = MVSGI_TILE_CTAS;

struct TiledKP {
  int tileH, octPerRow;  // tile = tileH rows x (octPerRow * 8) pixels, tileH * octPerRow == kTileWarps
  int tilesX, numTiles;
  int stepsPerItem, itemsPerTile, frameGroups, numItems;  // item = (frame group, tile, candidate range of <= stepsPerItem)
  int chunkBytes, minfoOff, vinfoOff;  // chunk = rec[CAM][kTileVox] uint4 | minfo[kTileVox][4] u32 (per-frame masks) | vinfo[kTileVox] u32
  int fX;                              // elements between horizontally adjacent feature pixels (16 staged, C direct)
  int direct;                          // 1: kp.feat is the caller's channels-last tensor (no zero frame)
  unsigned char* table;                // [numTiles][D] chunks
  int* queue;                          // item counter, zero before the launch
};

struct F4 {  // the 4 consecutive channels a lane owns, as two packed pairs
  f32x2 lo, hi;
};
MVSGI_DEV F4 ldg_f4(const float* p) {
  const ulonglong2 u = __ldg(reinterpret_cast<const ulonglong2*>(p));
  F4 r;
  r.lo = u.x;
  r.hi = u.y;
  return r;
}
// table records: prefetched into L1 one candidate ahead (no registers held), then read from there
MVSGI_DEV uint4 ld_rec(const unsigned char* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
MVSGI_DEV uint32_t ld_word(const unsigned char* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
MVSGI_DEV void prefetch_l1(const unsigned char* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
MVSGI_DEV void st_stream_f4(float* p, f32x2 lo, f32x2 hi) {
  asm volatile("st.global.L1::no_allocate.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(lo), "l"(hi) : "memory");
}

// ---- geometry table ----------------------------------------------------------------------------------------------------
// Record of (camera, voxel): { e | ymode << 2 | xcode, w_first, w_second, ly }            backports.py:41-72
//   e        element offset (inside a frame's features, camera included, a multiple of 16) of row y0 of the patch
//            column this quad loads FIRST — the one whose 64-byte chunk sits in the bank half of the quad's parity
//   xcode    where the second column is: 0 at +fX elements, 2 at -fX, 1 the same column (image edge, its weight is 0)
//   ymode    0: rows y0 and y0 + 1.  Image edges of un-padded features: 1: row y0 is the zero padding — both taps read
//            row y0 + 1 and the y0 weight is forced to 0; 2: row y0 + 1 is the padding — both read row y0, ly := 0
//   w_first, w_second   x weights of the two columns ((1 - lx, lx) in some order; 0 for a padding column)
//   ly       y - y0
// With staged (zero-framed) features xcode is 0 or 2 and ymode 0; samples whose four taps all fall into the padding
// (and NaN coordinates) read the all-zero patch (staged) or pixel 0 with zero weights (direct).
MVSGI_DEV uint4 make_record(const SweepKP& kp, const TiledKP& tp, int cam, int j, float gx, float gy) {
  const float x = unnormalise(gx, kp.Wf), y = unnormalise(gy, kp.Hf);
  uint4 r;
  if (!(x >= -1.0f && x < (float)kp.Wf && y >= -1.0f && y < (float)kp.Hf)) {
    r.x = (uint32_t)((tp.direct ? 0 : kp.zeroOff) + cam * kp.fsCam);
    r.y = r.z = r.w = 0u;
    return r;
  }
  const float fx0 = floorf(x), fy0 = floorf(y);
  const int x0 = (int)fx0, y0 = (int)fy0;
  const float lx = fsub(x, fx0), ly = fsub(y, fy0), hx = 1.0f - lx;
  uint32_t ymode = 0, xcode;
  int row, col;
  float wf, ws;
  if (!tp.direct) {
    row = y0 + 1;
    const int c0 = x0 + 1;  // padded coordinates
    const uint32_t swap = ((uint32_t)((row * kp.fsY + c0 * tp.fX) >> 4) ^ (uint32_t)j) & 1u;
    col = c0 + (int)swap;
    xcode = swap ? 2u : 0u;
    wf = swap ? lx : hx;
    ws = swap ? hx : lx;
  } else {
    row = y0;
    if (y0 < 0) { row = 0; ymode = 1u; }
    else if (y0 >= kp.Hf - 1) { ymode = 2u; }
    if (x0 < 0) {  // only column x0 + 1 = 0 is real
      col = 0; xcode = 1u; wf = lx; ws = 0.0f;
    } else if (x0 >= kp.Wf - 1) {  // only column x0 is real
      col = x0; xcode = 1u; wf = hx; ws = 0.0f;
    } else {
      const uint32_t swap = ((uint32_t)((row * kp.fsY + x0 * tp.fX + cam * kp.fsCam) >> 4) ^ (uint32_t)j) & 1u;
      col = x0 + (int)swap;
      xcode = swap ? 2u : 0u;
      wf = swap ? lx : hx;
      ws = swap ? hx : lx;
    }
  }
  r.x = (uint32_t)(row * kp.fsY + col * tp.fX + cam * kp.fsCam) | (ymode << 2) | xcode;
  r.y = __float_as_uint(wf);
  r.z = __float_as_uint(ws);
  r.w = __float_as_uint(ly);
  return r;
}

// grid (numTiles, D), block kTileVox: thread j = voxel j of the chunk (octet j / 8 of the tile, pixel j % 8 of it).
template <int CAM, int AGG>
__global__ void __launch_bounds__(kTileVox)
tile_table_kernel(const SweepKP kp, const TiledKP tp, const __grid_constant__ MvsgiRig rig) {
  const int tile = blockIdx.x, d = blockIdx.y, j = threadIdx.x;
  const int ty = tile / tp.tilesX, tx = tile - ty * tp.tilesX;
  const int oct = j >> 3, r = oct / tp.octPerRow, o = oct - r * tp.octPerRow;
  const int h = ty * tp.tileH + r, w = (tx * tp.octPerRow + o) * 8 + (j & 7);
  const bool inside = h < kp.Ho && w < kp.Wo;
  const int hc = min(h, kp.Ho - 1), wc = min(w, kp.Wo - 1);
  unsigned char* chunk = tp.table + ((size_t)tile * kp.D + d) * tp.chunkBytes;
  uint4* rec = reinterpret_cast<uint4*>(chunk);
  uint32_t* minfo = reinterpret_cast<uint32_t*>(chunk + tp.minfoOff);
  uint32_t* vinfo = reinterpret_cast<uint32_t*>(chunk + tp.vinfoOff);

  float rx = 0.f, ry = 0.f, rz = 0.f;
  if (kp.gridMode == MVSGI_GRID_PROJECT) cv_ray(rig, hc, wc, rx, ry, rz);
  uint32_t vbits = 0;
#pragma unroll
  for (int cam = 0; cam < CAM; ++cam) {
    float gx, gy;
    bool gvalid = true;
    if (kp.gridMode == MVSGI_GRID_PROJECT) {
      float x, y, z;
      to_camera(rig.cams[cam], rx, ry, rz, rig.dist[d], x, y, z);
      gvalid = project(rig.cams[cam], x, y, z, gx, gy);
    } else {  // grids shared by all frames (batch stride 0)
      const float2 g = __ldg(reinterpret_cast<const float2*>(kp.grids + cam * kp.gsCam + d * kp.gsD + hc * kp.gsY + wc * kp.gsX));
      gx = g.x;
      gy = g.y;
      if (AGG == MVSGI_AGG_STD) gvalid = __ldg(kp.gmask + cam * kp.gmCam + d * kp.gmD + hc * kp.gmY + wc * kp.gmX) != 0;
    }
    rec[cam * kTileVox + j] = make_record(kp, tp, cam, j, gx, gy);
    if (AGG == MVSGI_AGG_STD) {
      const uint32_t mi = gvalid ? make_mask_info(gx, gy, kp.Hm, kp.Wm, kp.mPitch) : kMaskNone;
      if (kp.perFrameMasks) minfo[j * 4 + cam] = mi;
      else vbits |= mask_eval(mask_fetch(kp.mbits + cam * kp.mbCam, mi, kp.mPitch), mi) << cam;
    } else {
      vbits |= 1u << cam;
    }
  }
  if (AGG == MVSGI_AGG_STD && kp.perFrameMasks) {
#pragma unroll
    for (int cam = CAM; cam < 4; ++cam) minfo[j * 4 + cam] = kMaskNone;
  }
  vinfo[j] = vbits | (inside ? 0x100u : 0u);
}

// ---- the kernel ------------------------------------------------------------------------------------------------------
// A record decoded for this lane: element offsets of the four taps (first column rows y0 / y1, second column rows
// y0 / y1) and the four weights.                                                          backports.py:52-55
struct Taps {
  int oa, ob, oc, od;
  float w00, w01, w10, w11;
};
MVSGI_DEV Taps decode(const uint4 r, int fsY, int fX, float wf, float ws) {
  Taps t;
  const uint32_t xc = r.x & 3u, ym = (r.x >> 2) & 3u;
  const int dx = (1 - (int)xc) * fX;
  const int dy = ym ? 0 : fsY;
  float ly = __uint_as_float(r.w);
  float hy = 1.0f - ly;
  if (ym == 1u) hy = 0.0f;
  if (ym == 2u) ly = 0.0f;
  t.oa = (int)(r.x & ~15u);
  t.ob = t.oa + dy;
  t.oc = t.oa + dx;
  t.od = t.ob + dx;
  t.w00 = __fmul_rn(wf, hy);
  t.w01 = __fmul_rn(wf, ly);
  t.w10 = __fmul_rn(ws, hy);
  t.w11 = __fmul_rn(ws, ly);
  return t;
}
// One bilinear sample of this lane's 4 channels: Ia*wa + Ib*wb + Ic*wc + Id*wd (backports.py:86), summed as
// (first column) + (second column); which column comes first does not change the bits.
MVSGI_DEV F4 sample4(const float* __restrict__ fb, const Taps& t) {
  const F4 a0 = ldg_f4(fb + t.oa), b0 = ldg_f4(fb + t.ob), a1 = ldg_f4(fb + t.oc), b1 = ldg_f4(fb + t.od);
  const f32x2 W00 = pk2(t.w00, t.w00), W01 = pk2(t.w01, t.w01), W10 = pk2(t.w10, t.w10), W11 = pk2(t.w11, t.w11);
  F4 f;
  f.lo = add2(fma2(b0.lo, W01, mul2(a0.lo, W00)), fma2(b1.lo, W11, mul2(a1.lo, W10)));
  f.hi = add2(fma2(b0.hi, W01, mul2(a0.hi, W00)), fma2(b1.hi, W11, mul2(a1.hi, W10)));
  return f;
}

// Variance across the cameras of the (compile-time) set V, spherical_sweep_avg.py:114-125: the same operations in the
// same order as sweep_kernel performs for that set (absent cameras add exact zeros there).
template <int CAM, int V>
MVSGI_DEV F4 variance_uniform(const F4 (&f)[CAM]) {
  constexpr int n = ((V >> 0) & 1) + ((V >> 1) & 1) + ((V >> 2) & 1) + ((V >> 3) & 1);
  constexpr float inv = n == 2 ? 0.5f : (n == 3 ? (1.0f / 3.0f) : 0.25f);
  const f32x2 INV = pk2(inv, inv);
  F4 sm, acc, o;
  bool first = true;
#pragma unroll
  for (int cam = 0; cam < CAM; ++cam)
    if ((V >> cam) & 1) {
      if (first) sm = f[cam];
      else { sm.lo = add2(sm.lo, f[cam].lo); sm.hi = add2(sm.hi, f[cam].hi); }
      first = false;
    }
  // :114  as an fma with +0 so that ptxas cannot contract the product into the subtraction below (it does contract
  // mul.rn.f32x2 + sub.rn.f32x2 into FFMA2); sweep_kernel rounds the mean separately, like the reference's tensor ops
  const f32x2 alo = fma2(sm.lo, INV, 0ull), ahi = fma2(sm.hi, INV, 0ull);
  first = true;
#pragma unroll
  for (int cam = 0; cam < CAM; ++cam)
    if ((V >> cam) & 1) {
      const f32x2 dlo = sub2(f[cam].lo, alo), dhi = sub2(f[cam].hi, ahi);  // :119-122
      if (first) { acc.lo = mul2(dlo, dlo); acc.hi = mul2(dhi, dhi); }     // == fma(d, d, 0)
      else { acc.lo = fma2(dlo, dlo, acc.lo); acc.hi = fma2(dhi, dhi, acc.hi); }
      first = false;
    }
  o.lo = mul2(acc.lo, INV);
  o.hi = mul2(acc.hi, INV);
  return o;
}

// Everything a lane needs to store its 4 channels of one voxel of one frame.
struct OutCtx {
  float* vcg;          // this lane's first channel of the voxel (16-channel group 0)
  long long validIdx;  // index into validOut
  bool act;            // the voxel is inside the image
};
template <bool CL>
MVSGI_DEV void store4(const SweepKP& kp, float* dst, const F4& o, int nst, int csr) {
  if (CL) {
    if (nst >= 4) st_stream_f4(dst, o.lo, o.hi);
    if (kp.numPeers > 0 && nst >= 4) {
#pragma unroll 1
      for (int pr = 0; pr < kp.numPeers; ++pr) st_stream_f4(dst + kp.peerDelta[pr], o.lo, o.hi);
    }
  } else {
    float v[4];
    unpk2(o.lo, v[0], v[1]);
    unpk2(o.hi, v[2], v[3]);
    st_lane<false>(dst, v, nst, csr);
    if (kp.numPeers > 0) {
#pragma unroll 1
      for (int pr = 0; pr < kp.numPeers; ++pr) st_lane<false>(dst + kp.peerDelta[pr], v, nst, csr);
    }
  }
}

// One octet of one candidate for FB frames whose voxels all see exactly the cameras of V (warp-uniform, compile
// time): no votes, no selects, absent cameras are not touched; the taps of all FB frames are independent loads.
template <int CAM, int AGG, bool CL, int FB, int V>
MVSGI_DEV void octet_uniform(const SweepKP& kp, const TiledKP& tp, const uint4 (&rec)[CAM], const float* const (&fb)[FB],
                             const OutCtx (&oc)[FB], int q, int csr) {
  constexpr int n = ((V >> 0) & 1) + ((V >> 1) & 1) + ((V >> 2) & 1) + ((V >> 3) & 1);
  const float* fc[FB];
#pragma unroll
  for (int k = 0; k < FB; ++k) fc[k] = fb[k];
  for (int cg = 0; cg < kp.CG; ++cg) {
    F4 f[FB][CAM];
#pragma unroll
    for (int cam = 0; cam < CAM; ++cam)
      if ((V >> cam) & 1) {
        const Taps t = decode(rec[cam], kp.fsY, tp.fX, __uint_as_float(rec[cam].y), __uint_as_float(rec[cam].z));
#pragma unroll
        for (int k = 0; k < FB; ++k) f[k][cam] = sample4(fc[k], t);
      }
    const int nst = kp.C - (cg * kCG + q * 4);  // staged channels >= C are padding
#pragma unroll
    for (int k = 0; k < FB; ++k) {
      fc[k] += kp.fsCg;
      if (!oc[k].act) continue;
      float* vcg = oc[k].vcg + (CL ? (long long)cg * kCG : (long long)cg * kCG * csr);
      if (AGG == MVSGI_AGG_STD) {
        const F4 o = variance_uniform<CAM, V>(f[k]);
        store4<CL>(kp, vcg, o, nst, csr);
        if (kp.validOut != nullptr && cg == 0 && q == 0) kp.validOut[oc[k].validIdx] = (n > 1) ? 1 : 0;
      } else {  // concat: spherical_sweep.py:52-67, out channel = cam * C + c
#pragma unroll
        for (int cam = 0; cam < CAM; ++cam)
          store4<CL>(kp, vcg + (CL ? (long long)(cam * kp.C) : (long long)(cam * kp.C) * csr), f[k][cam], nst, csr);
      }
    }
  }
}

// One octet of one candidate of one frame with per-lane validity bits v (mask boundaries).
template <int CAM, bool CL>
MVSGI_DEV void octet_generic(const SweepKP& kp, const TiledKP& tp, const uint4 (&rec)[CAM], const float* fb, const OutCtx& oc,
                             uint32_t v, int q, int csr) {
  const int n = __popc(v);
  const bool need = oc.act && n > 1;  // spherical_sweep_avg.py:106-111
  for (int cg = 0; cg < kp.CG; ++cg, fb += kp.fsCg) {
    F4 f[CAM];
#pragma unroll
    for (int cam = 0; cam < CAM; ++cam) {
      const bool ld = need && ((v >> cam) & 1u);
      if (__any_sync(0xFFFFFFFFu, ld)) {
        // lanes with nothing to sample read their own (always valid) addresses with zero weights
        const Taps t = decode(rec[cam], kp.fsY, tp.fX, ld ? __uint_as_float(rec[cam].y) : 0.0f,
                              ld ? __uint_as_float(rec[cam].z) : 0.0f);
        f[cam] = sample4(fb, t);
      } else {
        f[cam].lo = f[cam].hi = 0ull;
      }
    }
    F4 o;
    o.lo = o.hi = 0ull;
    if (need) {
      const float inv = n == 2 ? 0.5f : (n == 3 ? (1.0f / 3.0f) : 0.25f);
      const f32x2 INV = pk2(inv, inv);
      F4 sm = f[0];
#pragma unroll
      for (int cam = 1; cam < CAM; ++cam) { sm.lo = add2(sm.lo, f[cam].lo); sm.hi = add2(sm.hi, f[cam].hi); }
      // :114  as an fma with +0 so that ptxas cannot contract the product into the subtraction below (it does contract
  // mul.rn.f32x2 + sub.rn.f32x2 into FFMA2); sweep_kernel rounds the mean separately, like the reference's tensor ops
  const f32x2 alo = fma2(sm.lo, INV, 0ull), ahi = fma2(sm.hi, INV, 0ull);
      F4 acc;
      acc.lo = acc.hi = 0ull;
#pragma unroll
      for (int cam = 0; cam < CAM; ++cam) {
        const bool on = (v >> cam) & 1u;  // :119-122 masked-out cameras contribute (avg - avg)^2 = 0
        const f32x2 dlo = on ? sub2(f[cam].lo, alo) : 0ull, dhi = on ? sub2(f[cam].hi, ahi) : 0ull;
        acc.lo = fma2(dlo, dlo, acc.lo);
        acc.hi = fma2(dhi, dhi, acc.hi);
      }
      o.lo = mul2(acc.lo, INV);
      o.hi = mul2(acc.hi, INV);
    }
    if (oc.act) {
      const int nst = kp.C - (cg * kCG + q * 4);
      store4<CL>(kp, oc.vcg + (CL ? (long long)cg * kCG : (long long)cg * kCG * csr), o, nst, csr);
      if (kp.validOut != nullptr && cg == 0 && q == 0) kp.validOut[oc.validIdx] = (n > 1) ? 1 : 0;
    }
  }
}

// Voxels nobody samples (fewer than two cameras valid): zeros.
template <bool CL>
MVSGI_DEV void octet_zero(const SweepKP& kp, const OutCtx& oc, int q, int csr) {
  if (!oc.act) return;
  F4 o;
  o.lo = o.hi = 0ull;
  for (int cg = 0; cg < kp.CG; ++cg) {
    const int nst = kp.C - (cg * kCG + q * 4);
    store4<CL>(kp, oc.vcg + (CL ? (long long)cg * kCG : (long long)cg * kCG * csr), o, nst, csr);
  }
  if (kp.validOut != nullptr && q == 0) kp.validOut[oc.validIdx] = 0;
}

template <int CAM, int AGG, bool CL, int FB>
__global__ void __launch_bounds__(kTileThreads, kTileCtas)
sweep_tiled_kernel(const SweepKP kp, const TiledKP tp) {
  __shared__ int sItem[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = lane & 3, q8 = lane >> 2;
  const int j = warp * 8 + q8;
  const int rowInTile = warp / tp.octPerRow;
  const int poLane = rowInTile * kp.Wo + (warp - rowInTile * tp.octPerRow) * 8 + q8;
  const int HW = kp.Ho * kp.Wo;
  const int csr = kp.Dvol * HW;
  const long long frameVol = (long long)kp.Cout * csr;
  const int perGroup = tp.numTiles * tp.itemsPerTile;

  if (tid == 0) sItem[0] = atomicAdd(tp.queue, 1);
  __syncthreads();
  for (int it = 0;; ++it) {
    const int item = sItem[it & 1];
    if (item >= tp.numItems) break;
    if (tid == 0) sItem[(it + 1) & 1] = atomicAdd(tp.queue, 1);  // its latency hides behind this item

    const int fg = item / perGroup, rem = item - fg * perGroup;
    const int tile = rem / tp.itemsPerTile, dg = rem - tile * tp.itemsPerTile;
    const int ty = tile / tp.tilesX, tx = tile - ty * tp.tilesX;
    const int d0 = dg * tp.stepsPerItem, d1 = min(kp.D, d0 + tp.stepsPerItem);
    const int org = ty * tp.tileH * kp.Wo + tx * tp.octPerRow * 8 + poLane;

    const float* fb[FB];
    float* vq[FB];
    const uint32_t* mb[FB];
    bool live[FB];
#pragma unroll
    for (int k = 0; k < FB; ++k) {
      const int b = min(fg * FB + k, kp.B - 1);
      live[k] = fg * FB + k < kp.B;  // the last group of an odd batch repeats its frame without storing it
      fb[k] = kp.feat + (long long)b * kp.fsB + q * 4;
      vq[k] = kp.vol + (long long)b * frameVol + (CL ? q * 4 : (long long)q * 4 * csr);
      mb[k] = kp.mbits + (long long)b * kp.mbB + (q < CAM ? q : 0) * kp.mbCam;
      // opaque to the compiler: otherwise it re-derives the pointers from the kernel parameters at every tap (a 64-bit
      // add + shift per load instead of one IMAD.WIDE)
      asm volatile("" : "+l"(fb[k]), "+l"(vq[k]), "+l"(mb[k]));
    }

    const unsigned char* chunk = tp.table + ((size_t)tile * kp.D + d0) * tp.chunkBytes;
    for (int d = d0; d < d1; ++d, chunk += tp.chunkBytes) {
      uint4 rec[CAM];
#pragma unroll
      for (int cam = 0; cam < CAM; ++cam) rec[cam] = ld_rec(chunk + (size_t)(cam * kTileVox + j) * 16);
      const uint32_t vin = ld_word(chunk + tp.vinfoOff + j * 4);
      uint32_t mi = kMaskNone;
      if (AGG == MVSGI_AGG_STD && kp.perFrameMasks) mi = ld_word(chunk + tp.minfoOff + (j * 4 + q) * 4);
      if (d + 1 < d1) {  // next candidate's records into L1 while this one's taps are in flight: lane q fetches camera q's line
        const unsigned char* nchunk = chunk + tp.chunkBytes;
        if (q < CAM) prefetch_l1(nchunk + (size_t)(q * kTileVox + j) * 16);
        else prefetch_l1(nchunk + tp.vinfoOff + j * 4);
        if (AGG == MVSGI_AGG_STD && kp.perFrameMasks && q == 0) prefetch_l1(nchunk + tp.minfoOff + j * 16);
      }

      const bool inside = (vin >> 8) & 1u;
      const int po = (d + kp.dOff) * HW + org;
      OutCtx oc[FB];
      uint32_t v[FB];
#pragma unroll
      for (int k = 0; k < FB; ++k) {
        oc[k].act = inside && live[k];
        // host: Cout * Dvol * Ho * Wo < 2^31, so the in-frame part of the offset is 32-bit arithmetic
        oc[k].vcg = vq[k] + (CL ? (long long)po * kp.Cout : (long long)po);
        oc[k].validIdx = (long long)min(fg * FB + k, kp.B - 1) * csr + po;
        v[k] = vin & 0xFu;
      }
      if (AGG == MVSGI_AGG_STD && kp.perFrameMasks) {  // spherical_sweep_avg.py:92-102, lane q looks up camera q
        MaskWords mw[FB];
#pragma unroll
        for (int k = 0; k < FB; ++k) mw[k] = mask_fetch(mb[k], mi, kp.mPitch);
#pragma unroll
        for (int k = 0; k < FB; ++k) {
          const uint32_t ball = __ballot_sync(0xFFFFFFFFu, mask_eval(mw[k], mi) != 0);
          v[k] = (ball >> (lane & ~3)) & ((1u << CAM) - 1u);
        }
      }

      if (AGG == MVSGI_AGG_CONCAT) {
        octet_uniform<CAM, AGG, CL, FB, (1 << CAM) - 1>(kp, tp, rec, fb, oc, q, csr);
      } else {
        // warp-uniform validity pattern shared by the FB frames (the usual case: masks are regions): specialised bodies
        const uint32_t v0 = __shfl_sync(0xFFFFFFFFu, v[0], 0);
        bool same = v[0] == v0;
#pragma unroll
        for (int k = 1; k < FB; ++k) same = same && v[k] == v0;
        const bool uni = __all_sync(0xFFFFFFFFu, same);
        if (uni && __popc(v0) < 2) {
#pragma unroll
          for (int k = 0; k < FB; ++k) octet_zero<CL>(kp, oc[k], q, csr);
        } else if (CAM == 3 && uni && v0 == 7u) {
          octet_uniform<CAM, AGG, CL, FB, (CAM == 3 ? 7 : 3)>(kp, tp, rec, fb, oc, q, csr);
        } else if (CAM == 3 && uni && v0 == 3u) {
          octet_uniform<CAM, AGG, CL, FB, 3>(kp, tp, rec, fb, oc, q, csr);
        } else if (CAM == 3 && uni && v0 == 5u) {
          octet_uniform<CAM, AGG, CL, FB, (CAM == 3 ? 5 : 3)>(kp, tp, rec, fb, oc, q, csr);
        } else if (CAM == 3 && uni && v0 == 6u) {
          octet_uniform<CAM, AGG, CL, FB, (CAM == 3 ? 6 : 3)>(kp, tp, rec, fb, oc, q, csr);
        } else if (CAM == 4 && uni && v0 == 15u) {
          octet_uniform<CAM, AGG, CL, FB, (CAM == 4 ? 15 : 3)>(kp, tp, rec, fb, oc, q, csr);
        } else if (CAM == 2 && uni && v0 == 3u) {
          octet_uniform<CAM, AGG, CL, FB, 3>(kp, tp, rec, fb, oc, q, csr);
        } else {
#pragma unroll
          for (int k = 0; k < FB; ++k)
            if (live[k]) octet_generic<CAM, CL>(kp, tp, rec, fb[k], oc[k], v[k], q, csr);
        }
      }
    }
    __syncthreads();  // publishes the next item; keeps the CTA's warps on one tile (their taps share L1 lines)
  }
}

}  // namespace mvsgi
