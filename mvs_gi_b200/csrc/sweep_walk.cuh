// sweep_walk.cuh — the fused sweep for a channels-last volume (MVSGI_FLAG_VOL_CHANNELS_LAST): "walking" quads.
//
// Same math, same tap table and same staged inputs as sweep.cuh (spherical_sweep_avg.py:38-136,
// spherical_sweep.py:38-68, backports.py:11-86); what changes is who does which voxel.
//
// sweep.cuh lets the 8 quads of a warp take 8 consecutive pixels of one candidate so that the NCDHW stores fill
// 32-byte sectors; every voxel then fetches its own four taps and the L1 data pipe (one wavefront per 128-byte line
// and quarter-warp) is the limiter.  With a channels-last volume a voxel's 16 channels are 64
// contiguous bytes, so a quad can own a *segment* of L consecutive pixels of one candidate and walk it.  Along the
// walk the sampling position moves about one feature pixel per step (G16V: |dx| = 0.97 px, same feature row 79 % of
// the time), so the 2x2 patch of step i+1 shares a column with the patch of step i.  The quad keeps the two columns
// of its current patch in registers, coloured by the parity of the column's offset: a step that advances x by one
// re-loads exactly one column (2 loads instead of 4), a step that stays in the patch loads nothing.  No register
// moves are needed: only the two x-weights swap roles.
//
// STATUS: correct (bit-identical to sweep.cuh, tests/test_gpu_parity.py) but NOT the default.  Measured on B200 at
// G16V B=32 (profiles/r1g_walk_experiment.md): L1 data-pipe wavefronts drop from 140.5 M to 115.7 M per launch, but
// the cached columns cost 116 registers (16 warps/SM) and serialise consecutive steps of a quad on one memory round
// trip, so the kernel is latency-bound: 850-890 us against 543 us for sweep.cuh with a channels-last volume.
// L1 prefetches ahead of the walk add wavefronts and do not help.  Enabled with MVSGI_WALK=1 for experiments.
#pragma once
#include "sweep.cuh"

namespace mvsgi {

constexpr int kWalkThreads = 512;  // 128 quads, one CTA per SM

template <int CAM, int AGG, bool VALID, bool PEERS>
__global__ void __launch_bounds__(kWalkThreads, 1)
sweep_walk_kernel(const SweepKP kp, const __grid_constant__ MvsgiRig rig) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int ES = entry_words(CAM);
  const int MP = kp.maxPairs;
  const int L = kp.wL;
  uint32_t* tab = reinterpret_cast<uint32_t*>(smem_raw);               // [MP][ES]
  float* rays = reinterpret_cast<float*>(tab + (size_t)MP * ES);       // [wSegs * L][3]
  uint32_t* tMi = reinterpret_cast<uint32_t*>(rays + 3 * kp.wSegs * L);  // [CAM][MP] mask lookup info (std only)

  const int tid = threadIdx.x, q = tid & 3, quad = tid >> 2;
  constexpr int nquads = kWalkThreads >> 2;
  const int cLane = q * 4;
  const int HW = kp.Ho * kp.Wo;
  const int cs = kp.Dvol * HW;

  for (int item = blockIdx.x; item < kp.numItems; item += gridDim.x) {
    const int run = item / kp.numChunks, chunk = item - run * kp.numChunks;
    // A run is every numRuns-th segment of L pixels (row-major pixel order): cheap (masked-out) and expensive image
    // regions are spread evenly over the CTAs.
    const int nseg = run < kp.wNumSeg ? min(kp.wSegs, (kp.wNumSeg - run + kp.numRuns - 1) / kp.numRuns) : 0;
    const int npx = nseg * L;
    auto pixel_of = [&](int p) {
      const int s = p / L;
      return (run + kp.numRuns * s) * L + (p - s * L);  // >= P in the ragged last segment
    };
    const int d0 = chunk * kp.dch;
    const int dcount = min(kp.dch, kp.D - d0);
    const int ntasks = nseg * dcount;  // task t = (segment s, candidate dd): t = s * dcount + dd
    const int npairs = ntasks * L;     // record j = t * L + i

    if (kp.gridMode == MVSGI_GRID_PROJECT) {
      for (int p = tid; p < npx; p += kWalkThreads) {
        const int pix = min(pixel_of(p), kp.P - 1);
        float rx, ry, rz;
        cv_ray(rig, pix / kp.Wo, pix % kp.Wo, rx, ry, rz);
        rays[3 * p] = rx;
        rays[3 * p + 1] = ry;
        rays[3 * p + 2] = rz;
      }
    }
    __syncthreads();

    // ---------------------------------------------------------------- build: geometry -> tap table (as sweep.cuh)
    auto build = [&](int b) {
      for (int e = tid; e < CAM * npairs; e += kWalkThreads) {
        const int cam = e / npairs, j = e - cam * npairs;
        const int t = j / L, i = j - t * L;
        const int s = t / dcount, dd = t - s * dcount;
        const int p = s * L + i;
        const int d = d0 + dd;
        const int pixr = pixel_of(p);
        const int pix = min(pixr, kp.P - 1);
        const int h = pix / kp.Wo, w = pix - h * kp.Wo;
        float gx, gy;
        bool gvalid = true;
        if (kp.gridMode == MVSGI_GRID_PROJECT) {
          float x, y, z;
          to_camera(rig.cams[cam], rays[3 * p], rays[3 * p + 1], rays[3 * p + 2], rig.dist[d], x, y, z);
          gvalid = project(rig.cams[cam], x, y, z, gx, gy);
        } else {
          const long long go = (long long)b * kp.gsB + cam * kp.gsCam + d * kp.gsD + h * kp.gsY + w * kp.gsX;
          const float2 g = __ldg(reinterpret_cast<const float2*>(kp.grids + go));
          gx = g.x;
          gy = g.y;
          if (AGG == MVSGI_AGG_STD)
            gvalid = __ldg(kp.gmask + (long long)b * kp.gmB + cam * kp.gmCam + d * kp.gmD + h * kp.gmY + w * kp.gmX) != 0;
        }
        int off;
        float lx, ly;
        make_entry(kp, gx, gy, off, lx, ly);
        uint32_t* rec = tab + (size_t)j * ES;
        rec[2 + 3 * cam] = (uint32_t)(off + cam * kp.fsCam);
        rec[3 + 3 * cam] = __float_as_uint(lx);
        rec[4 + 3 * cam] = __float_as_uint(ly);
        if (AGG == MVSGI_AGG_STD)
          tMi[cam * MP + j] = gvalid ? make_mask_info(gx, gy, kp.Hm, kp.Wm, kp.mPitch) : kMaskNone;
        if (cam == 0) {
          rec[0] = pixr < kp.P ? (uint32_t)((d + kp.dOff) * HW + pix) : 0xFFFFFFFFu;  // negative: nothing to store
          if (AGG == MVSGI_AGG_CONCAT) rec[1] = ((1u << CAM) - 1u) * 0x0101u;
        }
      }
    };
    auto lookup_masks = [&](int b) {  // spherical_sweep_avg.py:92-102, one validity byte per record and frame parity
      unsigned char* vb8 = reinterpret_cast<unsigned char*>(tab + 1) + (b & 1);
      const uint32_t* mb = kp.mbits + (kp.perFrameMasks ? (long long)b * kp.mbB : 0);
      for (int j = tid; j < npairs; j += kWalkThreads) {
        uint32_t info[CAM];
        MaskWords m[CAM];
#pragma unroll
        for (int cam = 0; cam < CAM; ++cam) info[cam] = tMi[cam * MP + j];
#pragma unroll
        for (int cam = 0; cam < CAM; ++cam) m[cam] = mask_fetch(mb + cam * kp.mbCam, info[cam], kp.mPitch);
        uint32_t v = 0;
#pragma unroll
        for (int cam = 0; cam < CAM; ++cam) v |= mask_eval(m[cam], info[cam]) << cam;
        vb8[(size_t)j * ES * 4] = (unsigned char)v;
      }
    };
    if (!kp.perFrameGrids) {
      build(0);
      __syncthreads();
      if (AGG == MVSGI_AGG_STD) {
        lookup_masks(0);
        if (!kp.perFrameMasks) lookup_masks(1);
        __syncthreads();
      }
    }

    // ---------------------------------------------------------------- sweep: frames outer, quads walk segments
    for (int b = 0; b < kp.B; ++b) {
      if (kp.perFrameGrids) {
        __syncthreads();
        build(b);
        __syncthreads();
      }
      if (AGG == MVSGI_AGG_STD && kp.perFrameGrids) {
        lookup_masks(b);
        __syncthreads();
      }
      if (AGG == MVSGI_AGG_STD && kp.perFrameMasks && !kp.perFrameGrids && b + 1 < kp.B) lookup_masks(b + 1);

      const float* fbq = kp.feat + (long long)b * kp.fsB + cLane;
      float* vq = kp.vol + (long long)b * kp.Cout * cs + cLane;  // voxel (c, po) at frame + po*Cout + c
      const int fsY = kp.fsY;
      const int vsh = (b & 1) * 8;

      for (int t = quad; t < ntasks; t += nquads) {
        const uint32_t* rec0 = tab + (size_t)t * L * ES;
        for (int cg = 0; cg < kp.CG; ++cg) {
          const int nst = kp.C - (cg * kCG + cLane);  // host: C % 4 == 0, so 4 real channels or none
          const float* fcg = fbq + (long long)cg * kp.fsCg;
          // cached columns of the current 2x2 patch, per camera: E = the column whose offset/16 is even, O = odd
          int ce[CAM], co[CAM];
          float4 Et[CAM], Eb[CAM], Ot[CAM], Ob[CAM];
#pragma unroll
          for (int cam = 0; cam < CAM; ++cam) {
            ce[cam] = co[cam] = -1;
            Et[cam] = Eb[cam] = Ot[cam] = Ob[cam] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          for (int i = 0; i < L; ++i) {
            const uint32_t* rec = rec0 + i * ES;
            const int po = (int)rec[0];
            const bool act = po >= 0;
            const uint32_t v = reinterpret_cast<const unsigned char*>(rec + 1)[vsh >> 3];  // this frame's byte only: the other byte is being rewritten for the next frame
            const int n = __popc(v);
            const bool need = act && (AGG == MVSGI_AGG_CONCAT || n > 1);  // spherical_sweep_avg.py:106-111
            float wEt[CAM], wEb[CAM], wOt[CAM], wOb[CAM];
            // pass 1: all loads of this step (only the columns that changed), so they are in flight together
#pragma unroll
            for (int cam = 0; cam < CAM; ++cam) {
              const bool ld = need && ((v >> cam) & 1u);
              const int off = (int)rec[2 + 3 * cam];
              const float lx = __uint_as_float(rec[3 + 3 * cam]), ly = __uint_as_float(rec[4 + 3 * cam]);
              const bool par = (off >> 4) & 1;
              const int offE = off + (par ? kCG : 0), offO = off + (par ? 0 : kCG);
              if (ld && offE != ce[cam]) {
                Et[cam] = ldg4(fcg + offE);
                Eb[cam] = ldg4(fcg + offE + fsY);
                ce[cam] = offE;
              }
              if (ld && offO != co[cam]) {
                Ot[cam] = ldg4(fcg + offO);
                Ob[cam] = ldg4(fcg + offO + fsY);
                co[cam] = offO;
              }
              const float hx = 1.0f - lx, hy = 1.0f - ly;
              const float wE = ld ? (par ? lx : hx) : 0.0f, wO = ld ? (par ? hx : lx) : 0.0f;
              wEt[cam] = __fmul_rn(wE, hy);  // backports.py:52-55: wa|wc on the top row, wb|wd on the bottom row
              wEb[cam] = __fmul_rn(wE, ly);
              wOt[cam] = __fmul_rn(wO, hy);
              wOb[cam] = __fmul_rn(wO, ly);
            }
            // pass 2: bilinear value per camera (column sums, then their sum: independent of the colouring)
            float f[CAM][4];
#pragma unroll
            for (int cam = 0; cam < CAM; ++cam) {
              f[cam][0] = __fadd_rn(fmaf(Eb[cam].x, wEb[cam], __fmul_rn(Et[cam].x, wEt[cam])),
                                    fmaf(Ob[cam].x, wOb[cam], __fmul_rn(Ot[cam].x, wOt[cam])));
              f[cam][1] = __fadd_rn(fmaf(Eb[cam].y, wEb[cam], __fmul_rn(Et[cam].y, wEt[cam])),
                                    fmaf(Ob[cam].y, wOb[cam], __fmul_rn(Ot[cam].y, wOt[cam])));
              f[cam][2] = __fadd_rn(fmaf(Eb[cam].z, wEb[cam], __fmul_rn(Et[cam].z, wEt[cam])),
                                    fmaf(Ob[cam].z, wOb[cam], __fmul_rn(Ot[cam].z, wOt[cam])));
              f[cam][3] = __fadd_rn(fmaf(Eb[cam].w, wEb[cam], __fmul_rn(Et[cam].w, wEt[cam])),
                                    fmaf(Ob[cam].w, wOb[cam], __fmul_rn(Ot[cam].w, wOt[cam])));
            }
            float* vcg = vq + (long long)po * kp.Cout + cg * kCG;
            if (AGG == MVSGI_AGG_STD) {
              float o[4] = {0.f, 0.f, 0.f, 0.f};
              if (need) {
                const float inv = n == 2 ? 0.5f : (n == 3 ? (1.0f / 3.0f) : 0.25f);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  float s = f[0][k];
#pragma unroll
                  for (int cam = 1; cam < CAM; ++cam) s += f[cam][k];
                  const float avg = s * inv;  // :114
                  float acc = 0.f;
#pragma unroll
                  for (int cam = 0; cam < CAM; ++cam) {
                    const float dlt = ((v >> cam) & 1u) ? f[cam][k] - avg : 0.0f;  // :119-122
                    acc = fmaf(dlt, dlt, acc);
                  }
                  o[k] = acc * inv;
                }
              }
              if (act) {
                st_lane<true>(vcg, o, nst, 0);
                if (PEERS)
                  for (int pr = 0; pr < kp.numPeers; ++pr) st_lane<true>(vcg + kp.peerDelta[pr], o, nst, 0);
                if (VALID && cg == 0 && q == 0) kp.validOut[(long long)b * cs + po] = (n > 1) ? 1 : 0;
              }
            } else {  // concat: spherical_sweep.py:52-67, out channel = cam*C + c
              if (act) {
#pragma unroll
                for (int cam = 0; cam < CAM; ++cam) {
                  float* dst = vcg + cam * kp.C;
                  st_lane<true>(dst, f[cam], nst, 0);
                  if (PEERS)
                    for (int pr = 0; pr < kp.numPeers; ++pr) st_lane<true>(dst + kp.peerDelta[pr], f[cam], nst, 0);
                }
              }
            }
          }
        }
      }
      __syncthreads();  // publishes the next frame's validity bits; keeps the CTA on one frame (L2 residency)
    }
  }
}

}  // namespace mvsgi
