// sweep_tiled.cuh — the tile-pipelined sweep (round 2): same arithmetic as sweep.cuh, different work decomposition.
//
// Replaces (dsta_mvs/model/...): cost_volume_builder/spherical_sweep_avg.py:38-136, spherical_sweep.py:38-68,
// backports/backports.py:11-86 — bit-identical to sweep_kernel (sweep.cuh) on the same inputs.
//
// Why a second decomposition.  sweep_kernel keeps the tap table of a CTA's pixels (80 KB) in shared memory for the
// whole launch; two CTAs leave 68 KB of L1 for the taps and each CTA's pixels are scattered octets, so only 26 % of
// the tap sectors hit L1 (profiles/r1m_sweep_ncu.txt).  Here
//   * the geometry of the whole volume is evaluated ONCE per launch by tile_table_kernel into a global table laid
//     out as one contiguous chunk per (tile, candidate): per voxel and camera a 16-byte record, plus the mask lookup
//     words / the static validity bits.  The table (20-26 MB at G16V) stays in the 126 MB L2 for all frames.
//   * a persistent CTA takes work items (frame, tile, candidate range) from an atomic queue in frame-major order (all
//     CTAs stay within a frame or two of each other, so a frame's features are read from HBM once), and a producer
//     warp streams the item's chunks through a ring of shared-memory stages with TMA bulk copies
//     (cp.async.bulk.shared::cluster.global + mbarrier complete_tx; SASS UBLKCP / SYNCS), 3 stages x <= 11 KB.
//     ~170 KB of the SM's 228 KB stay L1.
//   * a tile is a compact 8 x 16 block of output pixels and the 16 consumer warps walk its candidates in order: the
//     sampling footprint of the tile slides along the epipolar direction by a few pixels per candidate, so what a
//     candidate step needs is mostly what the previous one left in L1.
//   * the two quads of a quarter-warp (the 8 lanes one LDG.128 wavefront serves) always read 64-byte chunks in
//     opposite halves of the 128-byte bank line: of its two patch columns (x0, x0 + 1) every quad loads first the
//     one whose address bit 6 equals the quad's parity.  Measured (scripts/microbench/tap_banks.cu): 3.9 clk per
//     request instead of 5.3 (L1-resident) — the data-pipe cost of a tap request is set by bank conflicts between the
//     two chunks of a quarter-warp, not by the number of lines.  Column order does not change the result bits (the two
//     column sums are added with one commutative fadd).
#pragma once
#include "sweep.cuh"

namespace mvsgi {

constexpr int kTileWarps = 16;                      // consumer warps per CTA = octets per tile
constexpr int kTileVox = kTileWarps * 8;             // voxels per (tile, candidate) chunk
constexpr int kTileThreads = (kTileWarps + 1) * 32;  // + the producer warp
constexpr int kTileStages = 3;

struct TiledKP {
  int tileH, octPerRow;  // tile = tileH rows x (octPerRow * 8) pixels, tileH * octPerRow == kTileWarps
  int tilesX, numTiles;
  int stepsPerItem, itemsPerTile, numItems;  // item = (frame, tile, candidate range of <= stepsPerItem)
  int chunkBytes, minfoOff, vinfoOff;        // chunk = rec[CAM][128] uint4 | minfo[128][4] u32 (per-frame masks) | vinfo[128] u32
  unsigned char* table;                      // [numTiles][D] chunks
  int* queue;                                // item counter, zero before the launch
};

// ---- mbarrier / TMA bulk copy (PTX) ----------------------------------------------------------------------------------
MVSGI_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
MVSGI_DEV void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
MVSGI_DEV void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
MVSGI_DEV void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
MVSGI_DEV void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      "WAIT_%=:\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra DONE_%=;\n"
      " bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy by the TMA unit; completes `bytes` of transaction on `bar`.  16-byte aligned, size % 16 == 0.
MVSGI_DEV void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- geometry table ----------------------------------------------------------------------------------------------------
// grid (numTiles, D), block kTileVox: thread j = voxel j of the chunk (octet j / 8 of the tile, pixel j % 8 of it).
// Record of (camera, voxel): { e0 | swap, cx0, cx1, ly }
//   e0   element offset (inside a frame's staged features, camera included) of patch column x0, a multiple of 16
//   swap (bit 0 of the same word) 1: this quad loads column x0 + 1 first, 0: column x0 first
//   cx0, cx1  x weights of the first / second column ((1 - lx, lx) or swapped), ly = y - y0     backports.py:47-55
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
    int off;
    float lx, ly;
    make_entry(kp, gx, gy, off, lx, ly);
    const uint32_t e = (uint32_t)(off + cam * kp.fsCam);
    const uint32_t swap = ((e >> 4) ^ (uint32_t)j) & 1u;
    const float hx = 1.0f - lx;
    uint4 rr;
    rr.x = e | swap;
    rr.y = __float_as_uint(swap ? lx : hx);
    rr.z = __float_as_uint(swap ? hx : lx);
    rr.w = __float_as_uint(ly);
    rec[cam * kTileVox + j] = rr;
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
// One bilinear sample of 4 channels (backports.py:41-86).  r = the (camera, voxel) record; fb / fb2 = this lane's
// pointers to rows y0 / y0 + 1 of the frame's staged features (element offset 0, channel quarter included).
// The record holds the offset of column x0 with the swap bit in bit 0: this quad loads first the column whose bank
// half matches its parity (x0 if swap == 0, else x0 + 1), then the other one.
MVSGI_DEV void tile_sample(const float* __restrict__ fb, const float* __restrict__ fb2, const uint4 r, float cx0, float cx1,
                           float f[4]) {
  const int sw = (int)(r.x & 1u);
  const int o0 = (int)r.x + 15 * sw;  // (x0 column | swap bit) -> first column:  + 16 * swap, bit 0 cleared
  const int o1 = (int)r.x - 17 * sw;  // second column minus 16 elements (the + 16 is an immediate of the load)
  const float ly = __uint_as_float(r.w), hy = 1.0f - ly;
  const float4 a0 = ldg4(fb + o0), b0 = ldg4(fb2 + o0), a1 = ldg4(fb + o1 + 16), b1 = ldg4(fb2 + o1 + 16);
  const float w00 = __fmul_rn(cx0, hy), w01 = __fmul_rn(cx0, ly), w10 = __fmul_rn(cx1, hy), w11 = __fmul_rn(cx1, ly);
  // backports.py:86 as (column sum) + (column sum); which column comes first does not change the bits
  f[0] = __fadd_rn(fmaf(b0.x, w01, __fmul_rn(a0.x, w00)), fmaf(b1.x, w11, __fmul_rn(a1.x, w10)));
  f[1] = __fadd_rn(fmaf(b0.y, w01, __fmul_rn(a0.y, w00)), fmaf(b1.y, w11, __fmul_rn(a1.y, w10)));
  f[2] = __fadd_rn(fmaf(b0.z, w01, __fmul_rn(a0.z, w00)), fmaf(b1.z, w11, __fmul_rn(a1.z, w10)));
  f[3] = __fadd_rn(fmaf(b0.w, w01, __fmul_rn(a0.w, w00)), fmaf(b1.w, w11, __fmul_rn(a1.w, w10)));
}

// Variance across the cameras of the (compile-time) set V, spherical_sweep_avg.py:114-125: the same operations in the
// same order as the generic loop below performs for that set (absent cameras add exact zeros there).
template <int CAM, int V>
MVSGI_DEV void tile_variance_uniform(const float f[CAM][4], float o[4]) {
  constexpr int n = ((V >> 0) & 1) + ((V >> 1) & 1) + ((V >> 2) & 1) + ((V >> 3) & 1);
  constexpr float inv = n == 2 ? 0.5f : (n == 3 ? (1.0f / 3.0f) : 0.25f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float sm = 0.f;
    bool first = true;
#pragma unroll
    for (int cam = 0; cam < CAM; ++cam)
      if ((V >> cam) & 1) {
        sm = first ? f[cam][i] : sm + f[cam][i];
        first = false;
      }
    const float avg = sm * inv;
    float acc = 0.f;
#pragma unroll
    for (int cam = 0; cam < CAM; ++cam)
      if ((V >> cam) & 1) {
        const float dlt = f[cam][i] - avg;
        acc = fmaf(dlt, dlt, acc);
      }
    o[i] = acc * inv;
  }
}

// One octet of one candidate.  V >= 0: every voxel of the warp sees exactly the cameras of V (warp-uniform, compile
// time): no votes, no selects, absent cameras are not touched.  V < 0: per-lane validity bits v.
template <int CAM, int AGG, bool FULLC, bool VALID, bool PEERS, bool CL, int V>
MVSGI_DEV void tile_octet(const SweepKP& kp, const uint4* __restrict__ rec, const float* fb, const float* fb2, float* vcg,
                          uint32_t v, bool act, int q, int q8, int csr, long long validIdx) {
  constexpr bool UNI = V >= 0;
  const int n = UNI ? __popc((unsigned)V) : __popc(v);
  const bool need = UNI ? true : (act && (AGG == MVSGI_AGG_CONCAT || n > 1));  // spherical_sweep_avg.py:106-111
  for (int cg = 0; cg < kp.CG; ++cg, fb += kp.fsCg, fb2 += kp.fsCg, vcg += (CL ? (long long)kCG : kCG * (long long)csr)) {
    float f[CAM][4];
#pragma unroll
    for (int cam = 0; cam < CAM; ++cam) {
      if (UNI) {
        if ((V >> cam) & 1) {
          const uint4 r = rec[cam * kTileVox];
          tile_sample(fb, fb2, r, __uint_as_float(r.y), __uint_as_float(r.z), f[cam]);
          // keep the next camera's 4 tap loads behind this camera's: 4 x 16 bytes in flight per lane is what the
          // 56-register budget (2 CTAs x 17 warps) holds without spilling
          asm volatile("" ::: "memory");
        } else {
          f[cam][0] = f[cam][1] = f[cam][2] = f[cam][3] = 0.0f;
        }
      } else {
        const bool ld = need && ((v >> cam) & 1u);
        if (__any_sync(0xFFFFFFFFu, ld)) {
          uint4 r = rec[cam * kTileVox];
          // lanes with nothing to sample read the all-zero patch (the pixel of it in this quad's bank half) with zero weights
          if (!ld) r.x = (uint32_t)(kp.zeroOff + (q8 & 1) * 16);
          tile_sample(fb, fb2, r, ld ? __uint_as_float(r.y) : 0.0f, ld ? __uint_as_float(r.z) : 0.0f, f[cam]);
        } else {
          f[cam][0] = f[cam][1] = f[cam][2] = f[cam][3] = 0.0f;
        }
      }
    }

    const int nst = FULLC ? 4 : kp.C - (cg * kCG + q * 4);
    if (AGG == MVSGI_AGG_STD) {
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      if (UNI) {
        tile_variance_uniform<CAM, (V >= 0 ? V : 0)>(f, o);
      } else if (need) {
        const float inv = n == 2 ? 0.5f : (n == 3 ? (1.0f / 3.0f) : 0.25f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float sm = f[0][i];
#pragma unroll
          for (int cam = 1; cam < CAM; ++cam) sm += f[cam][i];
          const float avg = sm * inv;  // :114
          float acc = 0.f;
#pragma unroll
          for (int cam = 0; cam < CAM; ++cam) {
            const float dlt = ((v >> cam) & 1u) ? f[cam][i] - avg : 0.0f;  // :119-122
            acc = fmaf(dlt, dlt, acc);
          }
          o[i] = acc * inv;
        }
      }
      if (act) {
        st_lane<CL>(vcg, o, nst, csr);
        if (PEERS) {
          for (int pr = 0; pr < kp.numPeers; ++pr) st_lane<CL>(vcg + kp.peerDelta[pr], o, nst, csr);
        }
        if (VALID && cg == 0 && q == 0) kp.validOut[validIdx] = (n > 1) ? 1 : 0;
      }
    } else {  // concat: spherical_sweep.py:52-67
      if (act) {
#pragma unroll
        for (int cam = 0; cam < CAM; ++cam) {
          float* dst = vcg + (CL ? (long long)(cam * kp.C) : (long long)(cam * kp.C) * csr);
          st_lane<CL>(dst, f[cam], nst, csr);
          if (PEERS) {
            for (int pr = 0; pr < kp.numPeers; ++pr) st_lane<CL>(dst + kp.peerDelta[pr], f[cam], nst, csr);
          }
        }
      }
    }
  }
}

// Voxels nobody samples (fewer than two cameras valid): zeros.
template <bool FULLC, bool VALID, bool PEERS, bool CL>
MVSGI_DEV void tile_octet_zero(const SweepKP& kp, float* vcg, bool act, int q, int csr, long long validIdx) {
  const float o[4] = {0.f, 0.f, 0.f, 0.f};
  for (int cg = 0; cg < kp.CG; ++cg, vcg += (CL ? (long long)kCG : kCG * (long long)csr)) {
    const int nst = FULLC ? 4 : kp.C - (cg * kCG + q * 4);
    if (act) {
      st_lane<CL>(vcg, o, nst, csr);
      if (PEERS) {
        for (int pr = 0; pr < kp.numPeers; ++pr) st_lane<CL>(vcg + kp.peerDelta[pr], o, nst, csr);
      }
      if (VALID && cg == 0 && q == 0) kp.validOut[validIdx] = 0;
    }
  }
}

template <int CAM, int AGG, bool FULLC, bool VALID, bool PEERS, bool CL>
__global__ void __launch_bounds__(kTileThreads, 2)
sweep_tiled_kernel(const SweepKP kp, const TiledKP tp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];  // bulk copies need 16-byte aligned destinations
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kTileStages * tp.chunkBytes);
  uint64_t* empty = full + kTileStages;
  // per stage: { frame (-1 = no more work), output offset of the tile origin at this candidate, -, - } and the frame's
  // element offsets into the staged features / the volume as two 64-bit values
  int4* info = reinterpret_cast<int4*>(empty + kTileStages);
  longlong2* infoOff = reinterpret_cast<longlong2*>(info + kTileStages);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kTileStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kTileWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == kTileWarps) {
    // ------------------------------------------------------------ producer: item queue + TMA bulk copies of the chunks
    if (lane != 0) return;
    int stage = 0, phase = 0;
    int item = atomicAdd(tp.queue, 1);
    const int perFrame = tp.numTiles * tp.itemsPerTile;
    while (item < tp.numItems) {
      const int next = atomicAdd(tp.queue, 1);  // its latency hides behind this item's copies
      const int b = item / perFrame, rem = item - b * perFrame;
      const int tile = rem / tp.itemsPerTile, dg = rem - tile * tp.itemsPerTile;
      const int ty = tile / tp.tilesX, tx = tile - ty * tp.tilesX;
      const int d0 = dg * tp.stepsPerItem, d1 = min(kp.D, d0 + tp.stepsPerItem);
      const unsigned char* src = tp.table + ((size_t)tile * kp.D + d0) * tp.chunkBytes;
      const int org = ty * tp.tileH * kp.Wo + tx * tp.octPerRow * 8;
      for (int d = d0; d < d1; ++d, src += tp.chunkBytes) {
        mbar_wait(&empty[stage], phase ^ 1);  // passes at once the first time round (fresh barrier)
        info[stage] = make_int4(b, (d + kp.dOff) * kp.Ho * kp.Wo + org, 0, 0);
        infoOff[stage] = make_longlong2((long long)b * kp.fsB, (long long)b * kp.Cout * kp.Dvol * kp.Ho * kp.Wo);
        mbar_expect_tx(&full[stage], (uint32_t)tp.chunkBytes);
        bulk_g2s(smem_raw + (size_t)stage * tp.chunkBytes, src, (uint32_t)tp.chunkBytes, &full[stage]);
        if (++stage == kTileStages) { stage = 0; phase ^= 1; }
      }
      item = next;
    }
    mbar_wait(&empty[stage], phase ^ 1);
    info[stage] = make_int4(-1, 0, 0, 0);
    mbar_arrive(&full[stage]);
    return;
  }

  // -------------------------------------------------------------- consumers: warp = one octet of the tile
  // (nothing but the ring position is carried from step to step: 56 registers hold one camera's taps in flight)
  const int q = lane & 3, q8 = lane >> 2;
  const int j = warp * 8 + q8;
  const int rowInTile = warp / tp.octPerRow;
  const int poLane = rowInTile * kp.Wo + (warp - rowInTile * tp.octPerRow) * 8 + q8;
  const int csr = kp.Dvol * kp.Ho * kp.Wo;
  int stage = 0, phase = 0;

  for (;;) {
    mbar_wait(&full[stage], phase);
    const int4 in = info[stage];
    const int b = in.x;
    if (b < 0) break;
    const longlong2 offs = infoOff[stage];
    const float* fb = kp.feat + offs.x + q * 4;  // this lane's row y0 / y0 + 1 pointers into the frame's staged features
    const float* fb2 = fb + kp.fsY;
    const unsigned char* chunk = smem_raw + (size_t)stage * tp.chunkBytes;
    const uint4* rec = reinterpret_cast<const uint4*>(chunk) + j;
    const uint32_t vin = reinterpret_cast<const uint32_t*>(chunk + tp.vinfoOff)[j];
    uint32_t v = vin & 0xFu;
    if (AGG == MVSGI_AGG_STD && kp.perFrameMasks) {  // spherical_sweep_avg.py:92-102, lane q looks up camera q
      const uint32_t mi = reinterpret_cast<const uint32_t*>(chunk + tp.minfoOff)[j * 4 + q];
      const uint32_t* mb = kp.mbits + (long long)b * kp.mbB + (q < CAM ? q : 0) * kp.mbCam;
      const uint32_t bit = mask_eval(mask_fetch(mb, mi, kp.mPitch), mi);
      const uint32_t ball = __ballot_sync(0xFFFFFFFFu, bit != 0);
      v = (ball >> (lane & ~3)) & ((1u << CAM) - 1u);
    }
    const bool act = (vin >> 8) & 1u;
    const int po = in.y + poLane;
    // host: Cout * Dvol * Ho * Wo < 2^31, so the in-frame part of the offset is 32-bit arithmetic
    float* vcg = kp.vol + offs.y + (CL ? po * kp.Cout + q * 4 : po + q * 4 * csr);
    const long long validIdx = VALID ? (long long)b * csr + po : 0;

    if (AGG == MVSGI_AGG_CONCAT) {
      tile_octet<CAM, AGG, FULLC, VALID, PEERS, CL, (1 << CAM) - 1>(kp, rec, fb, fb2, vcg, v, act, q, q8, csr, validIdx);
    } else {
      // warp-uniform validity pattern (the usual case: masks are regions): specialised bodies
      const uint32_t v0 = __shfl_sync(0xFFFFFFFFu, v, 0);
      const bool uni = __all_sync(0xFFFFFFFFu, v == v0);
      if (uni && __popc(v0) < 2) {
        tile_octet_zero<FULLC, VALID, PEERS, CL>(kp, vcg, act, q, csr, validIdx);
      } else if (CAM == 3 && uni && v0 == 7u) {
        tile_octet<CAM, AGG, FULLC, VALID, PEERS, CL, (CAM == 3 ? 7 : -1)>(kp, rec, fb, fb2, vcg, v, act, q, q8, csr, validIdx);
      } else if (CAM == 3 && uni && v0 == 3u) {
        tile_octet<CAM, AGG, FULLC, VALID, PEERS, CL, (CAM == 3 ? 3 : -1)>(kp, rec, fb, fb2, vcg, v, act, q, q8, csr, validIdx);
      } else if (CAM == 3 && uni && v0 == 5u) {
        tile_octet<CAM, AGG, FULLC, VALID, PEERS, CL, (CAM == 3 ? 5 : -1)>(kp, rec, fb, fb2, vcg, v, act, q, q8, csr, validIdx);
      } else if (CAM == 3 && uni && v0 == 6u) {
        tile_octet<CAM, AGG, FULLC, VALID, PEERS, CL, (CAM == 3 ? 6 : -1)>(kp, rec, fb, fb2, vcg, v, act, q, q8, csr, validIdx);
      } else if (CAM == 4 && uni && v0 == 15u) {
        tile_octet<CAM, AGG, FULLC, VALID, PEERS, CL, (CAM == 4 ? 15 : -1)>(kp, rec, fb, fb2, vcg, v, act, q, q8, csr, validIdx);
      } else if (CAM == 2 && uni && v0 == 3u) {
        tile_octet<CAM, AGG, FULLC, VALID, PEERS, CL, (CAM == 2 ? 3 : -1)>(kp, rec, fb, fb2, vcg, v, act, q, q8, csr, validIdx);
      } else {
        tile_octet<CAM, AGG, FULLC, VALID, PEERS, CL, -1>(kp, rec, fb, fb2, vcg, v, act, q, q8, csr, validIdx);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
    if (++stage == kTileStages) { stage = 0; phase ^= 1; }
  }
}

}  // namespace mvsgi
