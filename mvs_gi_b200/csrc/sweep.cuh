// sweep.cuh — fused spherical sweep + cross-view aggregation for sm_100a.
//
// One persistent kernel replaces the reference's per-candidate loop of
// F.pad / 4x torch.gather / where / sum (~4.7k ATen calls per frame):
//   dsta_mvs/model/cost_volume_builder/spherical_sweep_avg.py:38-136  (std / variance, masked)
//   dsta_mvs/model/cost_volume_builder/spherical_sweep.py:38-68       (concat)
//   dsta_mvs/model/backports/backports.py:11-86                       (bilinear_grid_sample)
//
// Data layout (produced by stage.cuh): features are staged once per call into zero-padded, channel-group-major,
// channels-last images  [B][CAM][CG][Hf+3][Wf+2][16]  (fp32).  The one-pixel zero frame IS the reference's F.pad
// (backports.py:58), so a bilinear sample is always four unconditional taps; the extra all-zero row at the bottom gives
// far-outside samples an all-zero 2x2 patch to read.  Two horizontally adjacent pixels of a 16-channel group are
// 128 contiguous bytes.
//
// Work decomposition
//   run   = a contiguous range of output pixels in row-major (h, w) order; every CTA owns the same number of pixels.
//   item  = (run, chunk of <= dch candidates).  With the default plan there is exactly one item per CTA.
//   build = once per item (once per frame only when the caller supplies per-frame grids): evaluate the projection
//           (geometry.cuh) or read the supplied grid and turn it into a *tap table* in shared memory — per
//           (camera, candidate, pixel) the tap offset, frac x, frac y and the mask lookup info (16 bytes in all).
//   masks = once per item when masks are shared by all frames, else once per frame: a coalesced pass turns the mask
//           lookup info into one validity byte per (candidate, pixel).
//   sweep = frames OUTER, so all CTAs walk the batch together and one frame's features (12.6 MB at G16V) are read
//           from HBM once and then served by the 126 MB L2.  Inside a frame a quad of 4 lanes owns one
//           (pixel, candidate) pair; pairs are ordered pixel-group-major / candidate / pixel so that the sliding
//           footprint of ~32 pixels x all candidates stays in L1.
//           Each tap is one 16-byte load per lane (4 lanes = the 64 contiguous bytes of one staged pixel); mean /
//           variance across the cameras happens in registers; per-view warped volumes are never materialised;
//           the (C, D, H, W) volume is written once with streaming stores, 8 consecutive pixels = one full
//           32-byte sector per warp and channel.
#pragma once
#include "geometry.cuh"

namespace mvsgi {

constexpr int kCG = 16;             // channels per staged channel group
#ifdef MVSGI_SWEEP_THREADS
constexpr int kSweepThreads = MVSGI_SWEEP_THREADS;
constexpr int kSweepCtas = 1024 / MVSGI_SWEEP_THREADS;
#else
constexpr int kSweepThreads = 512;  // 128 quads per CTA
constexpr int kSweepCtas = 2;
#endif
// FB > 1 frames per pass (static masks and shared grids only): one decode of a table record, its addresses and its
// bilinear weights serves FB frames, whose taps are in flight together; fewer, fatter threads.  Measured shapes
// (profiles/r2t_fb_variants.jsonl, bit-identical volumes): 2 frames x 384 threads is the best all-round (80 registers);
// 4 frames x 256 threads (128 registers) wins for <= 3 cameras once the batch amortises the slower table build;
// 3 frames x 256 threads for 4 cameras (4 frames would spill).
__host__ __device__ constexpr int sweep_threads(int fb) { return fb == 1 ? kSweepThreads : (fb == 2 ? 384 : 256); }
__host__ __device__ constexpr int sweep_ctas(int fb) { return fb == 1 ? kSweepCtas : 2; }
inline int sweep_frames_per_pass(int cams, int batch, int chan_groups) {
  if (batch < 2) return 1;
  if (cams == 4) return batch >= 6 ? 3 : 2;
  return (batch >= 16 && chan_groups == 1) ? 4 : 2;
}
constexpr uint32_t kMaskNone = 0u;

struct SweepKP {
  int B, C, D, Cout, CG;
  int Ho, Wo, Hf, Wf, Hm, Wm;
  int P, numRuns, runLen, dch, numChunks, numItems, maxPairs, grpPx, interleave;
  int framesPerPass;  // the FB instantiation of sweep_kernel: 1, or sweep_frames_per_pass() (static masks, shared grids)
  int unitOct;  // interleaved runs are dealt in units of unitOct vertically stacked octets (1 = single octets)
  // cost-balanced runs (MvsgiSweepArgs.octet_order): run r owns octets octOrder[r * octUnits + k], k < octUnits
  // (-1 = empty slot); nullptr = the interleaved / contiguous assignment below
  const int* octOrder;
  int octUnits;
  unsigned long long* ctaTimes;  // optional [gridDim.x][4] (MvsgiSweepArgs.cta_times_out)
  // MvsgiSweepArgs.tap_table: item i's records at tapTable + i * tapItemWords (16-byte aligned); tapUse = 1: copied in
  // instead of built, 0: built and stored; nullptr = not kept
  uint32_t* tapTable;
  int tapItemWords, tapUse;
  int* itemQueue;  // USE with more items than CTAs: zeroed counter (workspace) the CTAs draw their later items from
  // backward instantiation (BWD): `vol` is the upstream gradient of the volume (read), gradFeat the zero-initialised
  // gradient image with the element layout of `feat` that the sampler's adjoint accumulates into
  float* gradFeat;
  int wL, wSegs, wNumSeg;  // sweep_walk.cuh: segment length, segments per run, segments in the image
  // staged features: element (b, cam, cg, yp, xp, c) at feat[b*fsB + cam*fsCam + cg*fsCg + yp*fsY + xp*16 + c]
  const float* feat;
  long long fsB;
  int fsCam, fsCg, fsY, zeroOff;  // host checks CAM * fsCam < 2^31
  // direct = 1: `feat` is the caller's channels-last tensor (no zero frame, nothing staged): element (b, cam, y, x, c)
  // at feat[b*fsB + cam*fsCam + y*fsY + x*fX + c], fsCg = 16; tap-table offsets then carry edge flags (make_entry_direct)
  int direct, fX;
  // packed masks (std): uint32 words, image (b, cam) at mbits + b*mbB + cam*mbCam, (Hm+2) rows of mPitch words
  const uint32_t* mbits;
  long long mbB, mbCam;
  int mPitch, perFrameMasks;
  // supplied grids: (b, cam, d, h, w) -> 2 floats ; grid_masks -> 1 byte
  int gridMode, perFrameGrids;
  const float* grids;
  long long gsB, gsCam, gsD, gsY, gsX;
  const uint8_t* gmask;
  long long gmB, gmCam, gmD, gmY, gmX;
  float* vol;
  uint8_t* validOut;
  // candidate-sharded sweep: the D candidates of this call are candidates [dOff, dOff + D) of a volume with Dvol
  // candidates; every store is repeated at vol + peerDelta[i] (elements) for i < numPeers (peer-mapped volumes)
  int dOff, Dvol, numPeers, volChannelsLast, featsBf16, volBf16;
  long long peerDelta[MVSGI_MAX_PEERS];
};

// ---- packed fp32 (FMUL2 / FFMA2 / FADD2 of sm_100: half the issue slots of the scalar form, same bits)
// packed fp32 (two lanes of one 64-bit register, IEEE round-to-nearest per lane) ---------------------------------
typedef unsigned long long f32x2;
MVSGI_DEV f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
MVSGI_DEV void unpk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
MVSGI_DEV f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
MVSGI_DEV f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
MVSGI_DEV f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
MVSGI_DEV f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
MVSGI_DEV float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// bf16-staged features (MVSGI_FLAG_FEATS_BF16): 4 channels = 8 bytes, widened to fp32 exactly (bf16 is the top half)
MVSGI_DEV float4 ldg4(const uint16_t* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u), __uint_as_float(u.y << 16),
                     __uint_as_float(u.y & 0xFFFF0000u));
}
template <bool BF16> struct StagedT { typedef float type; };
template <> struct StagedT<true> { typedef uint16_t type; };

// Mask test of spherical_sweep_avg.py:92-96 for 0/1 (non-negative) masks: bilinear(mask) > 0  <=>  some tap has
// weight > 0 and a set mask bit.  Encodes the padded bit position of tap (y0, x0) and which of the other three taps
// carry a non-zero weight.  kMaskNone (= 0: a bit of the zero frame, no other tap) = no tap can contribute, so the
// lookup needs no branch.
MVSGI_DEV uint32_t make_mask_info(float gx, float gy, int Hm, int Wm, int mPitch) {
  const float x = unnormalise(gx, Wm), y = unnormalise(gy, Hm);
  if (!(x >= -1.0f && x < (float)Wm && y >= -1.0f && y < (float)Hm)) return kMaskNone;
  const float fx0 = floorf(x), fy0 = floorf(y);
  const float hx = fsub(fadd(fx0, 1.0f), x), lx = fsub(x, fx0);
  const float hy = fsub(fadd(fy0, 1.0f), y), ly = fsub(y, fy0);
  const int x0 = (int)fx0, Y0 = (int)fy0 + 1;  // padded row; bit 32 + x of a row is pixel x (stage.cuh)
  uint32_t info = (uint32_t)(Y0 * mPitch * 32 + 32 + x0);
  if (fmul(hx, ly) > 0.0f) info |= 1u << 28;  // wb: (x0, y1)
  if (fmul(lx, hy) > 0.0f) info |= 1u << 29;  // wc: (x1, y0)
  if (fmul(lx, ly) > 0.0f) info |= 1u << 30;  // wd: (x1, y1)
  // wa = (x1-x)(y1-y) is always > 0 here
  return info;
}

struct MaskWords {
  uint32_t a, b, c, d;
};
MVSGI_DEV MaskWords mask_fetch(const uint32_t* __restrict__ img, uint32_t info, int mPitch) {
  const uint32_t word = (info & 0x0FFFFFFFu) >> 5;
  // packed rows are stored as interleaved pairs (stage.cuh): one 8-byte load = word `word` of rows Y0 and Y0 + 1
  const uint2* pairs = reinterpret_cast<const uint2*>(img);
  MaskWords m;
  const uint2 ac = __ldg(pairs + word);
  m.a = ac.x;
  m.c = ac.y;
  m.b = m.d = 0u;
  if ((info & 31u) == 31u) {  // only then does tap x0 + 1 live in the next word (1 lookup in 32)
    const uint2 bd = __ldg(pairs + word + 1);
    m.b = bd.x;
    m.d = bd.y;
  }
  return m;
}
MVSGI_DEV uint32_t mask_eval(const MaskWords& m, uint32_t info) {
  const uint32_t sh = info & 31u;
  const uint32_t r0 = __funnelshift_r(m.a, m.b, sh);
  const uint32_t r1 = __funnelshift_r(m.c, m.d, sh);
  uint32_t hit = r0;                 // a
  hit |= r1 & (info >> 28);          // b
  hit |= (r0 >> 1) & (info >> 29);   // c
  hit |= (r1 >> 1) & (info >> 30);   // d
  return hit & 1u;
}

MVSGI_DEV void st_stream(float* p, float v) {  // written once, never re-read here: keep it out of L1
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
MVSGI_DEV void st_stream4(float* p, const float* v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3])
               : "memory");
}
// The 4 consecutive channels a lane owns of one voxel.  CL (channels-last volume, MVSGI_FLAG_VOL_CHANNELS_LAST): they
// are 16 contiguous bytes, so a quad writes the 64 bytes of a 16-channel group and a warp 8 consecutive pixels = 512
// contiguous bytes.  Otherwise (NCDHW) they are `cs` elements apart and a warp writes one 32-byte sector per channel.
template <bool CL>
MVSGI_DEV void st_lane(float* p, const float* v, int nst, long long cs) {
  if (CL) {
    if (nst >= 4) st_stream4(p, v);  // host: C % 4 == 0, so a lane owns 4 real channels or none
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < nst) st_stream(p + i * cs, v[i]);
  }
}

// Tap-table entry of one (camera, candidate, pixel): element offset of tap (y0, x0) inside a staged (b, cam, cg)
// image, frac x (x - x0) and frac y.  backports.py:41-55 with the zero padding folded into the staged image; samples
// whose four taps all fall into the padding (and NaN coordinates) read the all-zero patch.
MVSGI_DEV void make_entry(const SweepKP& kp, float gx, float gy, int& off, float& lx, float& ly) {
  const float x = unnormalise(gx, kp.Wf), y = unnormalise(gy, kp.Hf);
  if (!(x >= -1.0f && x < (float)kp.Wf && y >= -1.0f && y < (float)kp.Hf)) {
    off = kp.zeroOff;
    lx = 0.0f;
    ly = 0.0f;
  } else {
    const float fx0 = floorf(x), fy0 = floorf(y);
    off = ((int)fy0 + 1) * kp.fsY + ((int)fx0 + 1) * kCG;
    lx = fsub(x, fx0);
    ly = fsub(y, fy0);
  }
}

// The same entry for features that are tapped where they are (channels-last, no zero frame; backports.py:58-72 pads and
// clamps instead).  The offset is a multiple of 16 elements, so its low 4 bits carry what the padding would have done:
//   bits 0-1  xmode  0: columns x0, x0 + 1   1: column x0 is padding — both taps read column x0 + 1, its weight hx := 0
//                                            2: column x0 + 1 is padding — both read column x0, lx := 0
//   bits 2-3  ymode  likewise for rows y0, y0 + 1 (hy := 0 / ly := 0)
// A padding tap thus reads a real pixel with weight exactly 0, which adds the same +0 the zero frame would (for finite
// features).  Samples entirely outside read pixel 0 with all weights 0.
MVSGI_DEV void make_entry_direct(const SweepKP& kp, float gx, float gy, int& off, float& lx, float& ly) {
  const float x = unnormalise(gx, kp.Wf), y = unnormalise(gy, kp.Hf);
  if (!(x >= -1.0f && x < (float)kp.Wf && y >= -1.0f && y < (float)kp.Hf)) {
    off = 1 | (1 << 2);  // xmode 1, ymode 1 with lx = ly = 0: all four weights are 0
    lx = 0.0f;
    ly = 0.0f;
    return;
  }
  const float fx0 = floorf(x), fy0 = floorf(y);
  int x0 = (int)fx0, y0 = (int)fy0, xmode = 0, ymode = 0;
  if (x0 < 0) { x0 = 0; xmode = 1; }
  else if (x0 >= kp.Wf - 1) xmode = 2;
  if (y0 < 0) { y0 = 0; ymode = 1; }
  else if (y0 >= kp.Hf - 1) ymode = 2;
  off = (y0 * kp.fsY + x0 * kp.fX) | (ymode << 2) | xmode;
  lx = fsub(x, fx0);
  ly = fsub(y, fy0);
}

// bfloat16 volume (MVSGI_FLAG_VOL_BF16, channels-last only): the lane's 4 channels are 8 contiguous bytes, rounded to
// nearest even from the fp32 result.
template <bool CL>
MVSGI_DEV void st_lane(uint16_t* p, const float* v, int nst, long long) {
  if (nst >= 4) {
    uint32_t lo, hi;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(v[1]), "f"(v[0]));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v[3]), "f"(v[2]));
    asm volatile("st.global.L1::no_allocate.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(lo), "r"(hi) : "memory");
  }
}
template <bool BF16> struct VolT { typedef float type; };
template <> struct VolT<true> { typedef uint16_t type; };

// ---- the kernel ---------------------------------------------------------------------------------------------------
// Shared-memory tap table: one record of kEntryWords(CAM) words per (candidate, pixel) pair
//   [0] output offset d*Ho*Wo + h*Wo + w   [1] validity bits of the cameras (rewritten per frame if masks are per frame)
//   [2+3c .. 4+3c] camera c: tap offset, frac x, frac y
// The record stride is odd, so the 8 records a warp reads at a time sit in 8 different banks.
__host__ __device__ constexpr int entry_words(int cam) { return (2 + 3 * cam) | 1; }

MVSGI_DEV uint32_t smem_addr32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
MVSGI_DEV void mbar_wait32(uint32_t mb, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(mb), "r"(parity)
        : "memory");
  } while (!done);
}

MVSGI_DEV unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

MVSGI_DEV void red_add4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// BWD = true: the backward pass on the same work decomposition (table build, mask lookups, frame loop) — per voxel the
// forward samples are re-gathered, d vol / d f_k = (2 / n)(f_k - mean) * grad_vol (std) or grad_vol itself (concat) is
// formed in registers and scattered to the four taps of every sampled camera with red.global.add.v4.f32
// (backports.py:76-86 backward; spherical_sweep_avg.py:114-125).  384 threads x 2 CTAs (85 registers), one frame per pass.
constexpr int kSweepBwdThreads = 384;
__host__ __device__ constexpr int sweep_nt(int fb, bool bwd) { return bwd ? kSweepBwdThreads : sweep_threads(fb); }

template <int CAM, int AGG, bool FULLC, bool VALID, bool PEERS, bool CL, bool BF16, bool DIRECT = false, bool VBF16 = false,
          int FB = 1, bool BWD = false>
__global__ void __launch_bounds__(sweep_nt(FB, BWD), BWD ? 2 : sweep_ctas(FB))
sweep_kernel(const SweepKP kp, const __grid_constant__ MvsgiRig rig) {
  constexpr int NT = sweep_nt(FB, BWD);
  static_assert(!BWD || (FB == 1 && !VALID && !PEERS && !BF16 && !VBF16), "backward: one fp32 frame per pass");
  // MvsgiSweepArgs.cta_times_out: (SM id, start, end of the first table build, end) of every CTA in globaltimer
  // nanoseconds — what ops.calibrated_octet_order() reads to find the SMs that run this kernel slower than the rest
  if (kp.ctaTimes != nullptr && threadIdx.x == 0) {
    unsigned sm;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    kp.ctaTimes[4 * blockIdx.x] = sm;
    kp.ctaTimes[4 * blockIdx.x + 1] = gtimer();
  }
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int ES = entry_words(CAM);
  const int MP = kp.maxPairs;
  uint32_t* tab = reinterpret_cast<uint32_t*>(smem_raw);             // [MP][ES]
  float* rays = reinterpret_cast<float*>(tab + (size_t)MP * ES);     // [runLen][3]
  uint32_t* tMi = reinterpret_cast<uint32_t*>(rays + 3 * kp.runLen);  // [CAM][MP] mask lookup info (std only)

  const int tid = threadIdx.x, lane = tid & 31, q = lane & 3;
  const int warp = tid >> 5;
  constexpr int nwarps = NT >> 5;
  // Tap table kept across calls (MvsgiSweepArgs.tap_table, USE): an item's records arrive by one bulk asynchronous copy
  // (cp.async.bulk global -> shared) that completes this mbarrier; its phase flips once per item.
  __shared__ __align__(8) unsigned long long tapBar;
  const bool tapUse = !BWD && kp.tapTable != nullptr && kp.tapUse != 0;
  uint32_t tapPhase = 0;
  if (tapUse) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr32(&tapBar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  const int cLane = q * 4;                       // first of the 4 channels (inside a 16-channel group) this lane owns
  const int HW = kp.Ho * kp.Wo;
  const int cs = kp.Dvol * HW;                   // channel stride of vol (host checks Cout * cs < 2^31)
  const int numOct = (kp.P + 7) >> 3;            // runs and locality groups are made of 8-pixel octets (32-byte sectors)

  // Items: CTA c takes items c, c + grid, ...; with a kept tap table (USE) and more items than CTAs the items after a CTA's
  // first come from a queue (one atomicAdd per item, issued by thread 0 when the previous item starts so that its latency
  // is hidden): the CTAs of the high-D sweep, 16 items each, ended between 296 and 488 us with the static deal
  // (profiles/r4e_cta_times_highd.txt) — any order of the items gives the same volume.
  __shared__ int nextItem;
  const bool dynItems = tapUse && kp.itemQueue != nullptr;
  // Candidate chunk of an item: rotated by the item's wave, because grid % numChunks == 0 would otherwise hand CTA c the
  // same chunk (c % numChunks) in every wave, and the chunks differ in cost (near candidates sample fewer cameras).
  const int chunkRot = kp.numChunks > 1 ? max(1, (int)gridDim.x / kp.numChunks) : 1;
  for (int item = blockIdx.x; item < kp.numItems;) {
    const int run = item / kp.numChunks;
    const int chunk = kp.numChunks > 1 ? (item - run * kp.numChunks + run / chunkRot) % kp.numChunks : 0;
    // A run is either a contiguous range of octets or (interleave) every numRuns-th octet: the strided form spreads
    // the expensive image regions (all cameras valid) and the cheap ones (masked out) evenly over the CTAs.
    int p0, npx;
    const int U = kp.unitOct, octPerRow = kp.Wo >> 3;  // U > 1: host guarantees Wo % 8 == 0 and Ho % U == 0
    const int numUnits = U > 1 ? (kp.Ho / U) * octPerRow : numOct;
    if (kp.octOrder != nullptr) {
      p0 = 0;
      npx = 8 * kp.octUnits;
    } else if (kp.interleave) {
      p0 = 0;
      npx = 8 * U * ((numUnits - run + kp.numRuns - 1) / kp.numRuns);
    } else {
      p0 = 8 * (int)((long long)numOct * run / kp.numRuns);
      npx = min(kp.P, 8 * (int)((long long)numOct * (run + 1) / kp.numRuns)) - p0;
    }
    // U stacked octets share feature rows between them (vertical neighbours sample rows 1.6 apart), and a run's units
    // are still spread over the image for load balance
    auto pixel_of = [&](int p) {
      if (kp.octOrder != nullptr) {  // an empty slot maps past the image: its records are inactive
        const int o = __ldg(kp.octOrder + run * kp.octUnits + (p >> 3));
        return o >= 0 ? 8 * o + (p & 7) : kp.P + (p & 7);
      }
      if (!kp.interleave) return p0 + p;
      const int k = p >> 3;
      if (U == 1) return 8 * (run + kp.numRuns * k) + (p & 7);
      const int u = run + kp.numRuns * (k / U), sub = k - (k / U) * U;
      const int ur = u / octPerRow, uc = u - ur * octPerRow;
      return (ur * U + sub) * kp.Wo + uc * 8 + (p & 7);
    };
    const int d0 = chunk * kp.dch;
    const int dcount = min(kp.dch, kp.D - d0);
    const int npairs = npx * dcount;
    // locality groups of about grpPx pixels: pairs are ordered (group, candidate, pixel)
    const int ngroups = (npx + kp.grpPx - 1) / kp.grpPx;
    const int gsz = 8 * ((((npx + 7) >> 3) + ngroups - 1) / ngroups);

    if (kp.gridMode == MVSGI_GRID_PROJECT && !tapUse) {
      for (int p = tid; p < npx; p += NT) {
        const int pix = min(pixel_of(p), kp.P - 1);
        float rx, ry, rz;
        cv_ray(rig, pix / kp.Wo, pix % kp.Wo, rx, ry, rz);
        rays[3 * p] = rx;
        rays[3 * p + 1] = ry;
        rays[3 * p + 2] = rz;
      }
    }
    __syncthreads();

    // ---------------------------------------------------------------- build: geometry -> tap table
    // one table entry: camera `cam` of pair j = (candidate d0 + dd, pixel slot p of the run)
    auto build_entry = [&](int b, int cam, int j, int dd, int p) {
      const int d = d0 + dd;
      const int pixr = pixel_of(p);
      const int pix = min(pixr, kp.P - 1);  // only the last octet of the image can be ragged
      float gx, gy;
      bool gvalid = true;
      if (kp.gridMode == MVSGI_GRID_PROJECT) {
        float x, y, z;
        to_camera(rig.cams[cam], rays[3 * p], rays[3 * p + 1], rays[3 * p + 2], rig.dist[d], x, y, z);
        gvalid = project(rig.cams[cam], x, y, z, gx, gy);
      } else {
        const int h = pix / kp.Wo, w = pix - h * kp.Wo;
        const long long go = (long long)b * kp.gsB + cam * kp.gsCam + d * kp.gsD + h * kp.gsY + w * kp.gsX;
        const float2 g = __ldg(reinterpret_cast<const float2*>(kp.grids + go));
        gx = g.x;
        gy = g.y;
        if (AGG == MVSGI_AGG_STD)
          gvalid = __ldg(kp.gmask + (long long)b * kp.gmB + cam * kp.gmCam + d * kp.gmD + h * kp.gmY + w * kp.gmX) != 0;
      }
      int off;
      float lx, ly;
      if (DIRECT) make_entry_direct(kp, gx, gy, off, lx, ly);
      else make_entry(kp, gx, gy, off, lx, ly);
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
    };
    auto build = [&](int b) {
      if (ngroups == 1) {
        // one locality group (the default plan): pair j = dd * npx + p.  Cameras outermost with a compile-time index
        // (the camera's parameters are then constant-bank operands), the candidate by a multiply-high instead of a
        // division: j < 2^16 and npx < 2^16, so floor(j * ceil(2^32 / npx) / 2^32) = j / npx exactly.
        const uint32_t magic = 0xFFFFFFFFu / (uint32_t)npx + 1u;
#pragma unroll
        for (int cam = 0; cam < CAM; ++cam)
          for (int j = tid; j < npairs; j += NT) {
            const int dd = (int)__umulhi((uint32_t)j, magic);
            build_entry(b, cam, j, dd, j - dd * npx);
          }
        return;
      }
      for (int e = tid; e < CAM * npairs; e += NT) {
        const int cam = e / npairs, j = e - cam * npairs;
        const int gi = j / (gsz * dcount);
        const int local = j - gi * gsz * dcount;
        const int gcount = min(gsz, npx - gi * gsz);
        const int dd = local / gcount;
        build_entry(b, cam, j, dd, gi * gsz + (local - dd * gcount));
      }
    };
    // Validity bits of every pair for frame b: consecutive threads = consecutive pixels of one camera and candidate,
    // so the four mask-word loads of a warp fall into one or two sectors.  spherical_sweep_avg.py:92-102
    auto lookup_masks = [&](int b) {
      unsigned char* vb8 = reinterpret_cast<unsigned char*>(tab + 1) + (b & 1);
      const uint32_t* mb = kp.mbits + (kp.perFrameMasks ? (long long)b * kp.mbB : 0);
      for (int j = tid; j < npairs; j += NT) {
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
        if (!kp.perFrameMasks) vb8[(size_t)j * ES * 4 + 1 - 2 * (b & 1)] = (unsigned char)v;  // shared masks: both parities
      }
    };
    if (tapUse) {
      // the records of this item as an earlier launch built them (host: shared grids, static masks).  Every thread has
      // passed the barrier that ends the previous item, so the table is free; the proxy fence orders those generic
      // reads before the asynchronous write.
      if (tid == 0) {
        const uint32_t bytes = (uint32_t)(((npairs * ES + 3) >> 2) << 4);
        const uint32_t mb = smem_addr32(&tapBar);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_addr32(tab)),
                     "l"(kp.tapTable + (size_t)item * kp.tapItemWords), "r"(bytes), "r"(mb)
                     : "memory");
        // the item after this one, drawn while the copy travels.  Every thread has read the previous nextItem before the
        // barrier above; the barrier that ends this item's last pass publishes the new one.
        if (dynItems) nextItem = (int)gridDim.x + atomicAdd(kp.itemQueue, 1);
        // (asking the next item's table into L2 here — cp.async.bulk.prefetch.L2 — measured slower: high-D 0.434 ->
        // 0.451 ms, profiles/r4g_prefetch_fb.jsonl; the 18 MB of prefetched tables compete with the taps for L2)
      }
      mbar_wait32(smem_addr32(&tapBar), tapPhase);
      tapPhase ^= 1u;
    } else if (!kp.perFrameGrids) {
      build(0);
      __syncthreads();
      if (AGG == MVSGI_AGG_STD) {  // frame 0 (and, when masks are shared, every frame: both parity bytes at once)
        lookup_masks(0);
        __syncthreads();
      }
      if (!BWD && kp.tapTable != nullptr) {  // BUILD: keep the finished records for later launches
        const uint4* src = reinterpret_cast<const uint4*>(tab);
        uint4* dst = reinterpret_cast<uint4*>(kp.tapTable + (size_t)item * kp.tapItemWords);
        const int n16 = (npairs * ES + 3) >> 2;
        for (int i = tid; i < n16; i += NT) dst[i] = src[i];
      }
    }

    if (kp.ctaTimes != nullptr && threadIdx.x == 0 && item == (int)blockIdx.x) kp.ctaTimes[4 * blockIdx.x + 2] = gtimer();
    // ---------------------------------------------------------------- sweep: frames outer, FB frames per pass
    // (FB > 1: host guarantees static masks and shared grids, so the table is the same for every frame)
    for (int b0 = 0; b0 < kp.B; b0 += FB) {
      // every pass holds exactly FB frames: the last pass of a batch that is not a multiple of FB starts early and
      // re-writes the frames it shares with the pass before it (same values); host guarantees B >= FB
      const int b = FB == 1 ? b0 : min(b0, kp.B - FB);
      if (kp.perFrameGrids) {
        __syncthreads();
        build(b);
        __syncthreads();
      }
      if (AGG == MVSGI_AGG_STD && kp.perFrameGrids) {
        lookup_masks(b);
        __syncthreads();
      }
      // Per-pass lane pointers and loop constants, pinned in registers (the empty asm keeps the compiler from
      // re-deriving them from the kernel parameters inside the loop).
      typedef typename StagedT<BF16>::type FT;  // element type of the staged features (same element offsets)
      const FT* fbq = reinterpret_cast<const FT*>(kp.feat) + (long long)b * kp.fsB + cLane;
      // NCDHW: voxel (c, po) at frame + c*cs + po; channels-last: at frame + po*Cout + c
      typedef typename VolT<VBF16>::type VT;  // element type of the volume (same element offsets)
      const long long volB = (long long)kp.Cout * cs;
      VT* vq = reinterpret_cast<VT*>(kp.vol) + (long long)b * volB + (CL ? (long long)cLane : (long long)cLane * cs);
      float* gbq = BWD ? kp.gradFeat + (long long)b * kp.fsB + cLane : nullptr;
      int fsY = kp.fsY, zeroOff = kp.zeroOff, csr = cs, fX = kp.fX;
      asm volatile("" : "+l"(fbq), "+l"(vq), "+r"(fsY), "+r"(zeroOff), "+r"(csr), "+r"(fX));
      long long fsB = kp.fsB;  // frames of a pass are fsB elements apart
      asm volatile("" : "+l"(fsB));

      const int q8 = lane >> 2;
      const int vsh = (b & 1) * 8;
      for (int jw = warp * 8; jw < npairs; jw += nwarps * 8) {  // warp-uniform trip count
        {
        const uint32_t* rec = tab + (size_t)min(jw + q8, npairs - 1) * ES;
        const int po = (int)rec[0];
        const bool act = jw + q8 < npairs && po >= 0;
        const uint32_t v = reinterpret_cast<const unsigned char*>(rec + 1)[vsh >> 3];  // this frame's byte only: the other byte is being rewritten for the next frame
        const int n = __popc(v);
        const bool need = act && (AGG == MVSGI_AGG_CONCAT || n > 1);  // spherical_sweep_avg.py:106-111
        const FT* fcg = fbq;
        VT* vcg = vq + (CL ? (long long)po * kp.Cout : (long long)po);

        // backward only: tap offsets (relative to the lane's pointer into the (frame, channel group) image) and bilinear
        // weights of camera `cam` for this voxel, exactly as the forward loop below derives them
        auto decode = [&](int cam, bool ld, int& o00, int& o01, int& o10, int& o11, float& wa, float& wb, float& wc,
                          float& wd) {
          const int offr = (int)rec[2 + 3 * cam];
          const int off = DIRECT ? (offr & ~15) : (ld ? offr : zeroOff);
          const float lx = __uint_as_float(rec[3 + 3 * cam]);
          float ly = __uint_as_float(rec[4 + 3 * cam]);
          float hx = ld ? 1.0f - lx : 0.0f, hy = 1.0f - ly;
          float lxe = ld ? lx : 0.0f;
          int dx = kCG, dy = fsY;
          if (DIRECT && ((offr & 15) != 0 || fX != kCG)) {
            const int xm = offr & 3, ym = (offr >> 2) & 3;
            if (xm == 1) hx = 0.0f;
            if (xm == 2) lxe = 0.0f;
            if (ym == 1) hy = 0.0f;
            if (ym == 2) ly = 0.0f;
            if (ym) dy = 0;
            dx = xm ? 0 : fX;
          }
          o00 = off; o01 = off + dx; o10 = off + dy; o11 = off + dy + dx;
          wa = __fmul_rn(hx, hy); wb = __fmul_rn(hx, ly); wc = __fmul_rn(lxe, hy); wd = __fmul_rn(lxe, ly);
        };

        for (int cg = 0; cg < kp.CG; ++cg, fcg += kp.fsCg, vcg += (CL ? (long long)kCG : kCG * (long long)csr)) {
          if (BWD) {
            const int nstb = FULLC ? 4 : kp.C - (cg * kCG + cLane);  // staged channels >= C are padding
            const float* gvc = reinterpret_cast<const float*>(vcg);
            float g[CAM][4];
            if (AGG == MVSGI_AGG_STD) {
              float fk[CAM][4];
#pragma unroll
              for (int cam = 0; cam < CAM; ++cam) {
                const bool ld = need && ((v >> cam) & 1u);
                fk[cam][0] = fk[cam][1] = fk[cam][2] = fk[cam][3] = 0.f;
                if (__any_sync(0xFFFFFFFFu, ld)) {
                  int o00, o01, o10, o11;
                  float wa, wb, wc, wd;
                  decode(cam, ld, o00, o01, o10, o11, wa, wb, wc, wd);
                  const float4 ta = ldg4(fcg + o00), tc = ldg4(fcg + o01), tb = ldg4(fcg + o10), td = ldg4(fcg + o11);
                  // the forward's sum: (column x0) + (column x1)
                  fk[cam][0] = __fadd_rn(__fmaf_rn(tb.x, wb, __fmul_rn(ta.x, wa)), __fmaf_rn(td.x, wd, __fmul_rn(tc.x, wc)));
                  fk[cam][1] = __fadd_rn(__fmaf_rn(tb.y, wb, __fmul_rn(ta.y, wa)), __fmaf_rn(td.y, wd, __fmul_rn(tc.y, wc)));
                  fk[cam][2] = __fadd_rn(__fmaf_rn(tb.z, wb, __fmul_rn(ta.z, wa)), __fmaf_rn(td.z, wd, __fmul_rn(tc.z, wc)));
                  fk[cam][3] = __fadd_rn(__fmaf_rn(tb.w, wb, __fmul_rn(ta.w, wa)), __fmaf_rn(td.w, wd, __fmul_rn(tc.w, wc)));
                }
              }
              const float inv = n == 2 ? 0.5f : (n == 3 ? (1.0f / 3.0f) : 0.25f);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float go = 0.f;
                if (need && i < nstb) go = CL ? __ldg(gvc + i) : __ldg(gvc + (long long)i * csr);
                float sm = fk[0][i];
#pragma unroll
                for (int cam = 1; cam < CAM; ++cam) sm += fk[cam][i];
                const float avg = sm * inv;
                const float k2 = 2.0f * go * inv;  // :114-125: the term through the mean cancels
#pragma unroll
                for (int cam = 0; cam < CAM; ++cam) g[cam][i] = (fk[cam][i] - avg) * k2;
              }
            } else {  // concat: the gradient of channel cam * C + c of the volume
#pragma unroll
              for (int cam = 0; cam < CAM; ++cam)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const long long ch = (long long)cam * kp.C + i;
                  g[cam][i] = (act && i < nstb) ? (CL ? __ldg(gvc + ch) : __ldg(gvc + ch * csr)) : 0.f;
                }
            }
            float* gcg = gbq + (long long)cg * kp.fsCg;
#pragma unroll
            for (int cam = 0; cam < CAM; ++cam) {
              const bool ld = need && ((v >> cam) & 1u);
              if (__any_sync(0xFFFFFFFFu, ld)) {
                int o00, o01, o10, o11;
                float wa, wb, wc, wd;
                decode(cam, ld, o00, o01, o10, o11, wa, wb, wc, wd);
                // a sample whose taps all lie in the padding reads the all-zero patch: nothing to accumulate
                if (ld && (DIRECT || (int)rec[2 + 3 * cam] != zeroOff + cam * kp.fsCam)) {
                  red_add4(gcg + o00, g[cam][0] * wa, g[cam][1] * wa, g[cam][2] * wa, g[cam][3] * wa);
                  red_add4(gcg + o01, g[cam][0] * wc, g[cam][1] * wc, g[cam][2] * wc, g[cam][3] * wc);
                  red_add4(gcg + o10, g[cam][0] * wb, g[cam][1] * wb, g[cam][2] * wb, g[cam][3] * wb);
                  red_add4(gcg + o11, g[cam][0] * wd, g[cam][1] * wd, g[cam][2] * wd, g[cam][3] * wd);
                }
              }
            }
            continue;
          }
          f32x2 f[FB][CAM][2];  // this lane's 4 channels of every camera's sample, as two packed pairs
#pragma unroll
          for (int cam = 0; cam < CAM; ++cam) {
            const bool ld = need && ((v >> cam) & 1u);
            if (__any_sync(0xFFFFFFFFu, ld)) {
              // lanes with nothing to sample read the all-zero patch with zero weights: no divergence, no predication
              // (DIRECT: their own, always valid, taps with zero weights)
              const int offr = (int)rec[2 + 3 * cam];
#ifdef MVSGI_ABL_L1TAPS  // ablation build: every tap inside two 16 KB windows (rows y0, y0 + 1) (L1-resident)
              const int off = (offr & 0xFF0);
#else
              const int off = DIRECT ? (offr & ~15) : (ld ? offr : zeroOff);
#endif
              const float lx = __uint_as_float(rec[3 + 3 * cam]);
              float ly = __uint_as_float(rec[4 + 3 * cam]);
              float hx = ld ? 1.0f - lx : 0.0f, hy = 1.0f - ly;
              float lxe = ld ? lx : 0.0f;
              const FT* t = fcg + off;
              const FT* t1 = t + fsY;
              float4 ta[FB], tb[FB], tc[FB], td[FB];
              if (DIRECT && ((offr & 15) != 0 || fX != kCG)) {
                // image edges: the padding tap re-reads its real neighbour with weight 0 (make_entry_direct); also the
                // general pixel pitch (C > 16)
                const int xm = offr & 3, ym = (offr >> 2) & 3;
                if (xm == 1) hx = 0.0f;
                if (xm == 2) lxe = 0.0f;
                if (ym == 1) hy = 0.0f;
                if (ym == 2) ly = 0.0f;
                if (ym) t1 = t;
                const int dx = xm ? 0 : fX;
                const FT* u = t;
                const FT* u1 = t1;
#pragma unroll
                for (int fb = 0; fb < FB; ++fb, u += fsB, u1 += fsB)
                  ta[fb] = ldg4(u), tc[fb] = ldg4(u + dx), tb[fb] = ldg4(u1), td[fb] = ldg4(u1 + dx);
              } else {
                const FT* u = t;
                const FT* u1 = t1;
#pragma unroll
                for (int fb = 0; fb < FB; ++fb, u += fsB, u1 += fsB)
                  ta[fb] = ldg4(u), tc[fb] = ldg4(u + kCG), tb[fb] = ldg4(u1), td[fb] = ldg4(u1 + kCG);
              }
              const float wa = __fmul_rn(hx, hy), wb = __fmul_rn(hx, ly), wc = __fmul_rn(lxe, hy),
                          wd = __fmul_rn(lxe, ly);  // backports.py:52-55
              // backports.py:86  Ia*wa + Ib*wb + Ic*wc + Id*wd, summed as (column x0) + (column x1), two channels per
              // instruction
              const f32x2 WA = pk2(wa, wa), WB = pk2(wb, wb), WC = pk2(wc, wc), WD = pk2(wd, wd);
#pragma unroll
              for (int fb = 0; fb < FB; ++fb) {
                f[fb][cam][0] = add2(fma2(pk2(tb[fb].x, tb[fb].y), WB, mul2(pk2(ta[fb].x, ta[fb].y), WA)),
                                     fma2(pk2(td[fb].x, td[fb].y), WD, mul2(pk2(tc[fb].x, tc[fb].y), WC)));
                f[fb][cam][1] = add2(fma2(pk2(tb[fb].z, tb[fb].w), WB, mul2(pk2(ta[fb].z, ta[fb].w), WA)),
                                     fma2(pk2(td[fb].z, td[fb].w), WD, mul2(pk2(tc[fb].z, tc[fb].w), WC)));
              }
            } else {
#pragma unroll
              for (int fb = 0; fb < FB; ++fb) f[fb][cam][0] = f[fb][cam][1] = 0ull;
            }
          }

          const int nst = FULLC ? 4 : kp.C - (cg * kCG + cLane);  // staged channels >= C are padding
#pragma unroll
          for (int fb = 0; fb < FB; ++fb) {
            VT* vfb = vcg + fb * volB;
            if (AGG == MVSGI_AGG_STD) {
              float o[4] = {0.f, 0.f, 0.f, 0.f};
              if (need) {
                const float inv = n == 2 ? 0.5f : (n == 3 ? (1.0f / 3.0f) : 0.25f);
                const f32x2 INV = pk2(inv, inv);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                  f32x2 s = f[fb][0][i];
#pragma unroll
                  for (int cam = 1; cam < CAM; ++cam) s = add2(s, f[fb][cam][i]);
                  // :114  every step rounds on its own, as the reference's separate tensor ops do.  The mean is an fma
                  // with +0: ptxas contracts mul.rn.f32x2 + sub.rn.f32x2 into one FFMA2, which would round once.
                  const f32x2 avg = fma2(s, INV, 0ull);
                  f32x2 acc = 0ull;
#pragma unroll
                  for (int cam = 0; cam < CAM; ++cam) {
                    // :119-122 masked-out cameras contribute (avg-avg)^2 = 0
                    const f32x2 dlt = ((v >> cam) & 1u) ? sub2(f[fb][cam][i], avg) : 0ull;
                    acc = fma2(dlt, dlt, acc);
                  }
                  unpk2(mul2(acc, INV), o[2 * i], o[2 * i + 1]);
                }
              }
#ifdef MVSGI_ABL_NOSTORE  // ablation build: only never-true stores (keeps the arithmetic alive)
              if (act && o[0] == 123.456f) {
#else
              if (act) {
#endif
                st_lane<CL>(vfb, o, nst, csr);
                if (PEERS) {  // the all-gather of the candidate slices, fused: same voxel into every peer's volume
                  for (int pr = 0; pr < kp.numPeers; ++pr) st_lane<CL>(vfb + kp.peerDelta[pr], o, nst, csr);
                }
                if (VALID && cg == 0 && q == 0) kp.validOut[(long long)(b + fb) * cs + po] = (n > 1) ? 1 : 0;
              }
            } else {  // concat: spherical_sweep.py:52-67, out channel = cam*C + c, masks ignored
              if (act) {
#pragma unroll
                for (int cam = 0; cam < CAM; ++cam) {
                  VT* dst = vfb + (CL ? (long long)(cam * kp.C) : (long long)(cam * kp.C) * csr);
                  float fc[4];
                  unpk2(f[fb][cam][0], fc[0], fc[1]);
                  unpk2(f[fb][cam][1], fc[2], fc[3]);
                  st_lane<CL>(dst, fc, nst, csr);
                  if (PEERS) {
                    for (int pr = 0; pr < kp.numPeers; ++pr) st_lane<CL>(dst + kp.peerDelta[pr], fc, nst, csr);
                  }
                }
              }
            }
          }
        }
        }
      }
      // Per-frame masks: each warp looks up its share of the NEXT frame's validity bits (other byte of the record) when
      // it has finished its taps of this frame, so the lookups' L2 latency hides behind the warps that are still
      // sampling (measured 551 -> 530 us against doing them at the start of the frame).  spherical_sweep_avg.py:92-102
      if (FB == 1 && AGG == MVSGI_AGG_STD && kp.perFrameMasks && !kp.perFrameGrids && b + 1 < kp.B) lookup_masks(b + 1);
      // One barrier per pass: it publishes the next frame's validity bits and keeps the CTA's warps on the same
      // frames, so a frame's features are fetched into L1 / L2 once, not once per straggler (measured: free-running
      // warps cost +30 % L2->L1 traffic).
      __syncthreads();
    }
    // (the barrier above also protects the table: the next item's build overwrites it; it also publishes nextItem,
    // which thread 0 rewrites only after the first barrier of the item that follows)
    item = dynItems ? nextItem : item + (int)gridDim.x;
  }
  if (kp.ctaTimes != nullptr && threadIdx.x == 0) kp.ctaTimes[4 * blockIdx.x + 3] = gtimer();
}

}  // namespace mvsgi
