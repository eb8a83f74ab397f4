// mvsgi_b200.cu — C-ABI entry points of libmvsgi_b200.so (see include/mvsgi_b200.h).
// Host side: argument validation, workspace carving, launch configuration.  No allocation, no sync.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "softargmin.cuh"
#include "stage.cuh"
#include "launch.h"
#ifdef MVSGI_EXPERIMENTS
#include "sweep_walk.cuh"
#endif
#include "backward.cuh"
#include "resample.cuh"

namespace mvsgi {

// ---- grid dump (mvsgi_project_only_f32): exactly the (gx, gy, valid) the build step computes -----------------------
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

// ---- octet costs (mvsgi_sweep_octet_cost): what a warp of sweep_kernel will do for an octet.  A warp owns the 8 voxels
// (octet, candidate) and loads a camera's taps when ANY of the 8 samples it (sweep.cuh: __any_sync(ld)), so the unit of
// cost is (octet, candidate, camera loaded): measured per-CTA times follow the count of those (R^2 0.69) far better
// than the count of sampled voxels (0.47), profiles/r3i_cta_times.txt.  Validity as the build step + lookup_masks
// compute it.  grid (ceil(P / 128), D); 8 consecutive lanes = one octet -> one atomic per octet and warp.
template <int CAM, int AGG>
__global__ void __launch_bounds__(128) octet_cost_kernel(const SweepKP kp, const __grid_constant__ MvsgiRig rig,
                                                         int* __restrict__ cost) {
  const int pixr = blockIdx.x * 128 + threadIdx.x, d = blockIdx.y;
  const bool act = pixr < kp.P;
  const int pix = act ? pixr : kp.P - 1;
  const int h = pix / kp.Wo, w = pix - h * kp.Wo;
  float rx = 0.f, ry = 0.f, rz = 0.f;
  if (kp.gridMode == MVSGI_GRID_PROJECT) cv_ray(rig, h, w, rx, ry, rz);
  uint32_t v = 0;  // bit cam: this voxel would sample camera cam
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
    if (AGG == MVSGI_AGG_STD) {
      const uint32_t info = gvalid ? make_mask_info(gx, gy, kp.Hm, kp.Wm, kp.mPitch) : kMaskNone;
      v |= mask_eval(mask_fetch(kp.mbits + cam * kp.mbCam, info, kp.mPitch), info) << cam;  // spherical_sweep_avg.py:92-102
    } else {
      v |= 1u << cam;
    }
  }
  if (!act || (AGG == MVSGI_AGG_STD && __popc(v) < 2)) v = 0;  // :106-111: one valid camera samples nothing
  int loaded = 0;  // cameras whose taps the octet's warp loads for this candidate
#pragma unroll
  for (int cam = 0; cam < CAM; ++cam) {
    const uint32_t any = __ballot_sync(0xFFFFFFFFu, (v >> cam) & 1u);
    loaded += ((any >> (threadIdx.x & 24)) & 0xFFu) != 0u;
  }
  if ((threadIdx.x & 7) == 0 && act) atomicAdd(cost + (pixr >> 3), 4 + 10 * loaded);
}

}  // namespace mvsgi

namespace {

std::atomic<long long> g_launches{0};

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// Planner knobs are compile-time constants in the shipped library; `make exp` (-DMVSGI_EXPERIMENTS) lets the
// environment override them for tuning runs.
#ifdef MVSGI_EXPERIMENTS
int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}
#else
constexpr int env_int(const char*, int dflt) { return dflt; }
#endif

struct SweepPlan {
  bool pack_masks;        // std builder: masks -> bit images
  bool per_frame_masks;
  bool per_frame_grids;
  int n_mask_images;      // B*CAM or CAM
  int mask_pitch, mask_words;
  int cg;                 // 16-channel groups of the staged features
  int64_t feats_bytes, mask_bytes, flag_bytes, total_bytes;
  int dch, num_chunks, num_runs, run_len, num_items, max_pairs, grid, group_px, interleave, unit_oct;
  int cta_fb;             // frames per pass of sweep_kernel (sweep.cuh, sweep_frames_per_pass)
  bool oct_order;         // runs follow MvsgiSweepArgs.octet_order
  int grid_target;        // resident CTAs of the sweep kernel
  bool one_item_per_cta;  // the default plan is one (run, all candidates) item per resident CTA, single octets
  size_t smem_bytes;
  int cout;
  bool walk;              // channels-last volume: sweep_walk.cuh (experiment builds only)
  int w_len, w_segs, w_num_seg;
  // tile-walking sweep (sweep_tiled.cuh)
  bool tiled;
  bool direct;            // features are tapped in the caller's channels-last tensor: nothing is staged
  int frames_per_pass, frame_groups;
  int tile_h, oct_per_row, tiles_x, num_tiles, steps_per_item, items_per_tile, tiled_items;
  int chunk_bytes, minfo_off, vinfo_off, tiled_grid;
  int64_t table_bytes;
  // MvsgiSweepArgs.tap_table: words per item (a multiple of 4, so every item starts 16-byte aligned) and total bytes;
  // 0 = this problem keeps no reusable table
  int64_t tap_item_words, tap_bytes;
};

int mask_pitch_words(int mask_w) { return (mask_w + 31) / 32 + 2; }  // zero guard word + data words + zero guard

int device_sms() {
  static int cached[mvsgi::kMaxDevices] = {};  // 0 = not asked yet
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 0 && dev < mvsgi::kMaxDevices && cached[dev] > 0) return cached[dev];
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return 148;
  if (dev >= 0 && dev < mvsgi::kMaxDevices) cached[dev] = sms;
  return sms;
}

int validate_and_plan(const MvsgiSweepArgs* a, SweepPlan* p) {
  if (a == nullptr) return MVSGI_E_NULL;
  if (a->struct_size != sizeof(MvsgiSweepArgs)) return MVSGI_E_VERSION;
  if (a->aggregate != MVSGI_AGG_STD && a->aggregate != MVSGI_AGG_CONCAT) return MVSGI_E_UNSUPPORTED;
  if (a->grid_mode != MVSGI_GRID_SUPPLIED && a->grid_mode != MVSGI_GRID_PROJECT) return MVSGI_E_UNSUPPORTED;
  if (a->flags & ~(MVSGI_FLAG_FEATS_BF16 | MVSGI_FLAG_STAGE_ONLY | MVSGI_FLAG_SWEEP_ONLY | MVSGI_FLAG_VOL_CHANNELS_LAST |
                   MVSGI_FLAG_CTA_TABLE | MVSGI_FLAG_STAGED_FEATS | MVSGI_FLAG_VOL_BF16))
    return MVSGI_E_UNSUPPORTED;
  // bf16 volume: channels-last only, plain forward (no validity output, no peers, fp32 features)
  if ((a->flags & MVSGI_FLAG_VOL_BF16) &&
      (!(a->flags & MVSGI_FLAG_VOL_CHANNELS_LAST) || (a->flags & MVSGI_FLAG_FEATS_BF16) || a->valid_out != nullptr ||
       a->num_peers > 0))
    return MVSGI_E_UNSUPPORTED;
  if ((a->flags & MVSGI_FLAG_STAGE_ONLY) && (a->flags & MVSGI_FLAG_SWEEP_ONLY)) return MVSGI_E_UNSUPPORTED;
  // bf16 staging is a plain-forward mode: no validity output, no peer stores
  if ((a->flags & MVSGI_FLAG_FEATS_BF16) && (a->valid_out != nullptr || a->num_peers > 0)) return MVSGI_E_UNSUPPORTED;
  if (a->batch <= 0 || a->feat_chs <= 0 || a->out_h <= 0 || a->out_w <= 0) return MVSGI_E_SHAPE;
  if (a->num_cams < 1 || a->num_cams > MVSGI_MAX_CAMS) return MVSGI_E_SHAPE;
  if (a->num_cands < 1) return MVSGI_E_SHAPE;
  if (a->feat_h < 1 || a->feat_w < 1) return MVSGI_E_SHAPE;
  if (a->feats == nullptr || a->vol == nullptr) return MVSGI_E_NULL;
  // in-image tap offsets and in-frame output offsets are kept in int32
  const int64_t cg64 = (a->feat_chs + mvsgi::kCG - 1) / mvsgi::kCG;
  if ((int64_t)a->num_cams * cg64 * (a->feat_h + 3) * (a->feat_w + 2) * mvsgi::kCG >= (1ll << 31)) return MVSGI_E_SHAPE;
  for (int i = 0; i < 5; ++i)
    if (a->feats_stride[i] < 0) return MVSGI_E_STRIDE;
  const int64_t cout64 = (int64_t)a->feat_chs * (a->aggregate == MVSGI_AGG_CONCAT ? a->num_cams : 1);
  // candidate-sharded call: this call's candidates are a slice of the volume's
  if (a->vol_num_cands < 0 || a->cand_offset < 0 || a->num_peers < 0 || a->num_peers > MVSGI_MAX_PEERS) return MVSGI_E_SHAPE;
  if (a->vol_num_cands == 0 && a->cand_offset != 0) return MVSGI_E_SHAPE;
  const int64_t dvol64 = a->vol_num_cands > 0 ? a->vol_num_cands : a->num_cands;
  if ((int64_t)a->cand_offset + a->num_cands > dvol64) return MVSGI_E_SHAPE;
  if (cout64 * dvol64 * a->out_h * a->out_w >= (1ll << 31)) return MVSGI_E_SHAPE;
  if (a->flags & MVSGI_FLAG_VOL_CHANNELS_LAST) {  // one 16-byte store per lane: 4 whole channels, aligned
    if (a->feat_chs % 4 != 0) return MVSGI_E_SHAPE;
    if ((reinterpret_cast<uintptr_t>(a->vol) & ((a->flags & MVSGI_FLAG_VOL_BF16) ? 7 : 15)) != 0) return MVSGI_E_STRIDE;
    for (int i = 0; i < a->num_peers; ++i)
      if ((reinterpret_cast<uintptr_t>(a->peer_vols[i]) & 15) != 0) return MVSGI_E_STRIDE;
  }
  if (a->num_peers > 0) {
    if (a->valid_out != nullptr) return MVSGI_E_UNSUPPORTED;
    for (int i = 0; i < a->num_peers; ++i) {
      if (a->peer_vols[i] == nullptr) return MVSGI_E_NULL;
      if (((reinterpret_cast<uintptr_t>(a->peer_vols[i]) ^ reinterpret_cast<uintptr_t>(a->vol)) & 3) != 0) return MVSGI_E_STRIDE;
    }
  }
  const bool std_agg = a->aggregate == MVSGI_AGG_STD;
  if (std_agg) {
    if (a->masks == nullptr) return MVSGI_E_NULL;
    if (a->mask_h < 1 || a->mask_w < 1) return MVSGI_E_SHAPE;
    if (a->masks_dtype != MVSGI_MASK_F32 && a->masks_dtype != MVSGI_MASK_U8 && a->masks_dtype != MVSGI_MASK_PACKED)
      return MVSGI_E_UNSUPPORTED;
    if ((int64_t)(a->mask_h + 2) * mask_pitch_words(a->mask_w) * 32 >= (1ll << 28)) return MVSGI_E_SHAPE;
    if (a->masks_dtype == MVSGI_MASK_PACKED) {  // bits from mvsgi_pack_masks, read as 8-byte row pairs
      if ((reinterpret_cast<uintptr_t>(a->masks) & 7) != 0 || a->masks_stride[0] % 2 != 0 || a->masks_stride[1] % 2 != 0)
        return MVSGI_E_STRIDE;
      if (a->masks_stride[1] < (int64_t)(a->mask_h + 2) * mask_pitch_words(a->mask_w) * 2 || a->masks_stride[0] < 0)
        return MVSGI_E_STRIDE;
    }
  }
  if (a->grid_mode == MVSGI_GRID_PROJECT) {
    if (a->rig == nullptr) return MVSGI_E_NULL;
    if (a->rig->num_cams != a->num_cams || a->rig->num_cands != a->num_cands) return MVSGI_E_SHAPE;
    if (a->rig->cv_height != a->out_h || a->rig->cv_width != a->out_w) return MVSGI_E_SHAPE;
    if (a->num_cands > MVSGI_MAX_CANDS) return MVSGI_E_SHAPE;
    for (int i = 0; i < a->num_cams; ++i)
      if (a->rig->cams[i].model < 0 || a->rig->cams[i].model > MVSGI_CAM_LINEAR_SPHERE) return MVSGI_E_UNSUPPORTED;
  } else {
    if (a->grids == nullptr) return MVSGI_E_NULL;
    if (std_agg && a->grid_masks == nullptr) return MVSGI_E_NULL;
    if ((reinterpret_cast<uintptr_t>(a->grids) & 7) != 0) return MVSGI_E_STRIDE;
    for (int i = 0; i < 5; ++i)
      if (a->grids_stride[i] % 2 != 0) return MVSGI_E_STRIDE;  // float2 loads
  }

  p->cout = std_agg ? a->feat_chs : a->feat_chs * a->num_cams;
  p->cg = (a->feat_chs + mvsgi::kCG - 1) / mvsgi::kCG;
  p->pack_masks = std_agg && a->masks_dtype != MVSGI_MASK_PACKED;
  p->per_frame_masks = std_agg && a->batch > 1 && a->masks_stride[0] != 0;
  p->per_frame_grids = a->grid_mode == MVSGI_GRID_SUPPLIED && a->batch > 1 &&
                       (a->grids_stride[0] != 0 || (std_agg && a->grid_masks_stride[0] != 0));
  p->n_mask_images = std_agg ? (p->per_frame_masks ? a->batch : 1) * a->num_cams : 0;
  p->mask_pitch = std_agg ? mask_pitch_words(a->mask_w) : 0;
  p->mask_words = std_agg ? (a->mask_h + 2) * p->mask_pitch * 2 : 0;  // row pairs, each row stored twice (stage.cuh)
  // Tile-walking sweep (sweep_tiled.cuh): experiment builds only (MVSGI_TILED=1), measured slower than the CTA-table
  // kernel on every BASELINE.json shape but the 4-camera one.
#ifdef MVSGI_EXPERIMENTS
  p->tiled = !p->per_frame_grids && !(a->flags & (MVSGI_FLAG_FEATS_BF16 | MVSGI_FLAG_CTA_TABLE)) &&
             env_int("MVSGI_TILED", 0) != 0 && env_int("MVSGI_WALK", 0) == 0;
#else
  p->tiled = false;
#endif
  // Channels-last features with whole 16-channel groups (north_star (a)) are tapped where they are: the table records
  // carry clamped offsets and edge flags instead of relying on a zero frame (sweep_tiled.cuh, make_record).
  {
    const int64_t* fs = a->feats_stride;
    const int64_t span = (int64_t)(a->num_cams - 1) * fs[1] + (int64_t)(a->feat_h - 1) * fs[3] +
                         (int64_t)(a->feat_w - 1) * fs[4] + a->feat_chs;
    p->direct = !(a->flags & (MVSGI_FLAG_STAGED_FEATS | MVSGI_FLAG_FEATS_BF16)) && fs[2] == 1 &&
                a->feat_chs % mvsgi::kCG == 0 && fs[4] >= a->feat_chs && fs[0] % 16 == 0 && fs[1] % 16 == 0 &&
                fs[3] % 16 == 0 && fs[4] % 16 == 0 && span < (1ll << 31) &&
                (reinterpret_cast<uintptr_t>(a->feats) & 127) == 0 && env_int("MVSGI_DIRECT", 1) != 0 &&
                env_int("MVSGI_WALK", 0) == 0;
  }
  p->feats_bytes = p->direct ? 0
                             : align_up((int64_t)a->batch * a->num_cams * p->cg * (a->feat_h + 3) * (a->feat_w + 2) *
                                            mvsgi::kCG * ((a->flags & MVSGI_FLAG_FEATS_BF16) ? 2 : 4), 256);
  p->mask_bytes = p->pack_masks ? align_up((int64_t)p->n_mask_images * p->mask_words * 4 + 16, 256) : 0;  // +16: funnel-shift over-read
  p->flag_bytes = 256;
  p->total_bytes = p->feats_bytes + p->mask_bytes + p->flag_bytes;

  // Work decomposition (sweep.cuh): equal runs of output pixels, one per resident CTA when the tap table fits.
  const int P = a->out_h * a->out_w;
  // Several frames per pass (sweep.cuh, FB): static masks and shared grids, so one table record serves every frame;
  // instantiated for whole channel groups tapped in place into a channels-last volume (the inference path).
  p->cta_fb = (a->batch >= 2 && !p->per_frame_masks && !p->per_frame_grids && p->direct &&
               (a->flags & MVSGI_FLAG_VOL_CHANNELS_LAST) && a->valid_out == nullptr && a->num_peers == 0 &&
               env_int("MVSGI_FB_ON", 1) != 0)
                  ? mvsgi::sweep_frames_per_pass(a->num_cams, a->batch, p->cg) : 1;
  // Kept tap table (USE): there is no table build for the fewer, fatter threads of the wide variant to be slow at, so it
  // pays from 4 (4 cameras: 3) frames on instead of 16 (6).  The table's layout does not depend on the variant.
  // Below 16 frames the variant is picked by cost: passes x frames per pass (a batch that is not a multiple re-sweeps the
  // frames its last pass shares with the one before) x the measured relative cost of a frame in that variant
  // (profiles/r4g_prefetch_fb.jsonl, profiles/r4n_small_batch_fb1.jsonl: wide 1.00, two frames 1.10, one frame 1.17).
  if (a->tap_table_mode == MVSGI_TAP_TABLE_USE && p->cta_fb >= 2 && a->batch < 16) {
    const int fbs[3] = {p->cg == 1 ? (a->num_cams == 4 ? 3 : 4) : 0, 2, 1};
    const double rel[3] = {1.00, 1.10, 1.17};
    double best = 1e30;
    for (int i = 0; i < 3; ++i) {
      const int fb = fbs[i];
      if (fb < 1 || fb > a->batch) continue;
      const double cost = (double)((a->batch + fb - 1) / fb) * fb * rel[i];
      if (cost < best - 1e-9) {
        best = cost;
        p->cta_fb = fb;
      }
    }
  }
  const int max_ctas = mvsgi::sweep_ctas(p->cta_fb);
  int ctas_per_sm = env_int("MVSGI_CTAS_PER_SM", max_ctas);
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  if (ctas_per_sm > max_ctas) ctas_per_sm = max_ctas;  // __launch_bounds__
  const int grid_target = device_sms() * ctas_per_sm;
  int dch = env_int("MVSGI_DCH", 16);
  if (dch < 1) dch = 1;
  if (dch > a->num_cands) dch = a->num_cands;
  const bool std_tab = std_agg;
  const int64_t per_pair = (int64_t)mvsgi::entry_words(a->num_cams) * 4 + (std_tab ? a->num_cams * 4 : 0);
  const int64_t per_px = (int64_t)dch * per_pair + 12;
  const int64_t budget = env_int("MVSGI_TABLE_BYTES", ctas_per_sm >= 4 ? 48 * 1024 : (ctas_per_sm == 2 ? 80 * 1024 : 160 * 1024));
  const int num_oct = (P + 7) / 8;
  // interleaved runs are dealt in units of `unit` vertically stacked octets when the image divides into them
  p->interleave = env_int("MVSGI_INTERLEAVE", 1) ? 1 : 0;
  int unit = env_int("MVSGI_UNIT_OCT", 1);
  if (unit < 1 || !p->interleave || a->out_w % 8 != 0 || a->out_h % unit != 0) unit = 1;
  p->unit_oct = unit;
  const int num_units = unit > 1 ? (a->out_h / unit) * (a->out_w / 8) : num_oct;
  // one item = (run, candidate chunk) per resident CTA
  int run_target = grid_target / ((a->num_cands + dch - 1) / dch);
  if (run_target < 1) run_target = 1;
  // Whole waves: when the tap table of 1/run_target of the image does not fit the shared-memory budget, every CTA
  // takes k equal items instead of the grid running one full wave plus a partial one (4 cameras: 2 x 6 octets per CTA
  // instead of 8 + 8 for a quarter of the CTAs).
  auto plan_runs = [&](int n_units, int oct_per_unit) {
    const int64_t cap = budget / ((int64_t)oct_per_unit * 8 * per_px);  // units per run the table can hold
    if (cap < 1) return 0;
    for (int k = 1; k <= 64; ++k) {
      const int ru = (n_units + k * run_target - 1) / (k * run_target);
      if (ru <= cap) return ru;
    }
    return (int)cap;
  };
  int run_units = plan_runs(num_units, unit);
  if (run_units < 1) {
    if (unit == 1) return MVSGI_E_SHAPE;
    unit = p->unit_oct = 1;
    run_units = plan_runs(num_oct, 1);
    if (run_units < 1) return MVSGI_E_SHAPE;
  }
  const int units = unit > 1 ? num_units : num_oct;
  p->num_runs = (units + run_units - 1) / run_units;
  p->run_len = 8 * unit * ((units + p->num_runs - 1) / p->num_runs);
  p->dch = dch;
  p->num_chunks = (a->num_cands + dch - 1) / dch;
  p->num_items = p->num_runs * p->num_chunks;
  p->max_pairs = p->run_len * dch;
  p->grid = p->num_items < grid_target ? p->num_items : grid_target;
  // order of the (candidate, pixel) records inside a run: with interleaved runs one group = the whole run, i.e.
  // candidate-major over all of the run's octets (measured against groups of 8 px: G16V -3.5 %, 4 cameras -19 %,
  // C=48 -11 %, DoubleSphere -16 %, high-D +1 %)
  p->group_px = env_int("MVSGI_GROUP_PX", p->interleave ? p->run_len : 32);
  if (p->group_px < 1) p->group_px = 1;
  p->grid_target = grid_target;
  // the plans a caller's octet_order can replace: all candidates in one chunk, single octets; one item per resident
  // CTA, or k whole waves of items (CTA c sweeps items c, c + grid, ...)
  p->one_item_per_cta = p->num_chunks == 1 && unit == 1 && !p->per_frame_grids && !p->tiled;
  p->oct_order = false;
  if (a->octet_order != nullptr) {  // cost-balanced runs: same table budget, the caller's assignment of octets to runs
    const int64_t cap = budget / (8 * per_px);
    if (!p->one_item_per_cta || a->order_runs < 1 || a->order_runs > 64 * grid_target || a->order_units < 1 ||
        a->order_units > cap || (int64_t)a->order_runs * a->order_units < num_oct ||
        (a->order_runs > grid_target && a->order_runs % grid_target != 0))
      return MVSGI_E_UNSUPPORTED;
    p->oct_order = true;
    p->num_runs = a->order_runs;
    p->run_len = 8 * a->order_units;
    p->num_items = p->num_runs;
    p->max_pairs = p->run_len * dch;
    p->grid = p->num_items < grid_target ? p->num_items : grid_target;
    p->group_px = p->run_len;
  }
  p->smem_bytes = align_up((int64_t)p->max_pairs * per_pair + (int64_t)p->run_len * 12 + 16, 16);
  p->walk = false;

  // Channels-last volume, experiment (MVSGI_WALK=1, off by default because it measures slower, see sweep_walk.cuh):
  // walking quads.  One 512-thread CTA per SM; a run is w_segs segments of
  // w_len consecutive pixels taken round-robin over the image; task = (segment, candidate), one per quad when
  // w_segs * dch == 128.
#ifdef MVSGI_EXPERIMENTS
  if ((a->flags & MVSGI_FLAG_VOL_CHANNELS_LAST) && !(a->flags & MVSGI_FLAG_FEATS_BF16) && a->num_cams >= 3 &&
      env_int("MVSGI_WALK", 0)) {
    const int sms = device_sms();
    const int nquads = mvsgi::kWalkThreads / 4;
    int segs = env_int("MVSGI_WALK_SEGS", nquads / dch);
    if (segs < 1) segs = 1;
    const int64_t wbudget = env_int("MVSGI_WALK_TABLE_BYTES", 200 * 1024);
    int64_t len = ((int64_t)P + (int64_t)sms * segs - 1) / ((int64_t)sms * segs);
    const int64_t len_cap = wbudget / ((int64_t)segs * ((int64_t)dch * per_pair + 12));
    if (len > len_cap) len = len_cap;
    if (len >= 1) {
      p->walk = true;
      p->w_len = (int)len;
      p->w_segs = segs;
      p->w_num_seg = (int)((P + len - 1) / len);
      p->num_runs = (p->w_num_seg + segs - 1) / segs;
      p->run_len = segs * (int)len;
      p->num_items = p->num_runs * p->num_chunks;
      p->max_pairs = segs * dch * (int)len;
      p->grid = p->num_items < sms ? p->num_items : sms;
      p->smem_bytes = align_up((int64_t)p->max_pairs * per_pair + (int64_t)p->run_len * 12 + 16, 16);
    }
  }
#endif

  // Tap table kept across calls (MvsgiSweepArgs.tap_table): one slot of max_pairs records per item
  p->tap_item_words = 0;
  p->tap_bytes = 0;
  if (!p->per_frame_grids && !p->per_frame_masks && !p->tiled && !p->walk) {
    p->tap_item_words = align_up((int64_t)p->max_pairs * mvsgi::entry_words(a->num_cams), 4);
    p->tap_bytes = align_up((int64_t)p->num_items * p->tap_item_words * 4, 256);
  }
  if (a->tap_table_mode != MVSGI_TAP_TABLE_OFF) {
    if (a->tap_table_mode != MVSGI_TAP_TABLE_BUILD && a->tap_table_mode != MVSGI_TAP_TABLE_USE) return MVSGI_E_UNSUPPORTED;
    if (p->tap_bytes == 0) return MVSGI_E_UNSUPPORTED;
    if (a->tap_table == nullptr) return MVSGI_E_NULL;
    if (a->tap_table_bytes < p->tap_bytes || (reinterpret_cast<uintptr_t>(a->tap_table) & 255) != 0) return MVSGI_E_WORKSPACE;
    if (p->tap_item_words * 4 >= (1ll << 20)) return MVSGI_E_SHAPE;  // one mbarrier transaction per item
  }

  p->table_bytes = 0;
#ifdef MVSGI_EXPERIMENTS
  if (p->tiled) {
    const int tw = mvsgi::tiled_warps_per_cta(), tv = tw * 8;  // octets / voxels per tile
    // tile = tile_h rows x (kTileWarps / tile_h) octets; one chunk of the geometry table per (tile, candidate)
    p->tile_h = env_int("MVSGI_TILE_H", 8);
    if (p->tile_h < 1 || tw % p->tile_h != 0) p->tile_h = tw >= 8 ? 8 : tw;
    p->oct_per_row = tw / p->tile_h;
    p->tiles_x = (a->out_w + p->oct_per_row * 8 - 1) / (p->oct_per_row * 8);
    p->num_tiles = p->tiles_x * ((a->out_h + p->tile_h - 1) / p->tile_h);
    const int rec_bytes = a->num_cams * tv * 16;
    p->minfo_off = rec_bytes;
    p->vinfo_off = rec_bytes + (p->per_frame_masks ? tv * 16 : 0);
    p->chunk_bytes = p->vinfo_off + tv * 4;
    // frames per pass: records and weights are decoded once for 2 frames whenever there are 2
    p->frames_per_pass = a->batch >= 2 ? env_int("MVSGI_FRAMES_PER_PASS", 2) : 1;
    if (p->frames_per_pass != 2) p->frames_per_pass = 1;
    p->frame_groups = (a->batch + p->frames_per_pass - 1) / p->frames_per_pass;
    const int tiled_target = device_sms() * mvsgi::tiled_ctas_per_sm();
    // enough items to keep every CTA busy and the tail short: split the candidates of a tile when frames x tiles is small
    const int64_t want = 8ll * tiled_target;
    int ipt = 1;
    while ((int64_t)p->frame_groups * p->num_tiles * ipt < want && ipt * 2 <= a->num_cands) ipt *= 2;
    p->steps_per_item = (a->num_cands + ipt - 1) / ipt;
    p->items_per_tile = (a->num_cands + p->steps_per_item - 1) / p->steps_per_item;
    const int64_t items = (int64_t)p->frame_groups * p->num_tiles * p->items_per_tile;
    if (items >= (1ll << 30)) return MVSGI_E_SHAPE;
    p->tiled_items = (int)items;
    p->tiled_grid = (int)(items < tiled_target ? items : tiled_target);
    p->table_bytes = align_up((int64_t)p->num_tiles * a->num_cands * p->chunk_bytes, 256);
    p->total_bytes += p->table_bytes;
  }
#endif
  return MVSGI_OK;
}

// Kernel parameters of the sweep (shared by the forward and the backward kernels).
void fill_kp(const MvsgiSweepArgs* a, const SweepPlan& p, float* ws_feats, uint32_t* ws_bits, mvsgi::SweepKP& kp) {
  const bool std_agg = a->aggregate == MVSGI_AGG_STD;
  const int Wp = a->feat_w + 2, Hp1 = a->feat_h + 3;
  memset(&kp, 0, sizeof(kp));
  kp.B = a->batch; kp.C = a->feat_chs; kp.D = a->num_cands; kp.Cout = p.cout; kp.CG = p.cg;
  kp.Ho = a->out_h; kp.Wo = a->out_w; kp.Hf = a->feat_h; kp.Wf = a->feat_w; kp.Hm = a->mask_h; kp.Wm = a->mask_w;
  kp.P = a->out_h * a->out_w; kp.numRuns = p.num_runs; kp.runLen = p.run_len; kp.dch = p.dch;
  kp.numChunks = p.num_chunks; kp.numItems = p.num_items; kp.maxPairs = p.max_pairs; kp.grpPx = p.group_px; kp.interleave = p.interleave; kp.unitOct = p.unit_oct;
  kp.framesPerPass = p.cta_fb;
  kp.octOrder = p.oct_order ? a->octet_order : nullptr;
  kp.octUnits = p.oct_order ? a->order_units : 0;
  kp.ctaTimes = reinterpret_cast<unsigned long long*>(a->cta_times_out);
  kp.tapTable = a->tap_table_mode != MVSGI_TAP_TABLE_OFF ? reinterpret_cast<uint32_t*>(a->tap_table) : nullptr;
  kp.tapItemWords = (int)p.tap_item_words;
  kp.tapUse = a->tap_table_mode == MVSGI_TAP_TABLE_USE ? 1 : 0;
  kp.itemQueue = nullptr;  // set by mvsgi_sweep_f32 (the counter lives in the call's workspace)
  kp.wL = p.walk ? p.w_len : 0; kp.wSegs = p.walk ? p.w_segs : 0; kp.wNumSeg = p.walk ? p.w_num_seg : 0;
  if (p.direct) {  // the caller's channels-last tensor: element (b, cam, y, x, c) at b*s0 + cam*s1 + y*s3 + x*s4 + c
    kp.feat = a->feats;
    kp.fsY = (int)a->feats_stride[3];
    kp.fsCg = mvsgi::kCG;
    kp.fsCam = (int)a->feats_stride[1];
    kp.fsB = a->feats_stride[0];
    kp.zeroOff = 0;
    kp.direct = 1;
    kp.fX = (int)a->feats_stride[4];
  } else {
    kp.feat = ws_feats;
    kp.fsY = Wp * mvsgi::kCG;
    kp.fsCg = Hp1 * kp.fsY;
    kp.fsCam = kp.fsCg * p.cg;
    kp.fsB = (long long)kp.fsCam * a->num_cams;
    kp.zeroOff = (a->feat_h + 1) * kp.fsY;  // rows Hf+1 and Hf+2 are all zero
    kp.direct = 0;
    kp.fX = mvsgi::kCG;
  }
  if (std_agg) {
    kp.mPitch = p.mask_pitch;
    kp.perFrameMasks = p.per_frame_masks ? 1 : 0;
    if (p.pack_masks) {
      kp.mbits = ws_bits;
      kp.mbCam = p.mask_words;
      kp.mbB = kp.mbCam * a->num_cams;
    } else {  // MVSGI_MASK_PACKED: the caller's bits
      kp.mbits = (const uint32_t*)a->masks;
      kp.mbCam = a->masks_stride[1];
      kp.mbB = a->masks_stride[0];
    }
  }
  kp.gridMode = a->grid_mode;
  kp.perFrameGrids = p.per_frame_grids ? 1 : 0;
  if (a->grid_mode == MVSGI_GRID_SUPPLIED) {
    kp.grids = a->grids;
    kp.gsB = a->grids_stride[0]; kp.gsCam = a->grids_stride[1]; kp.gsD = a->grids_stride[2];
    kp.gsY = a->grids_stride[3]; kp.gsX = a->grids_stride[4];
    if (std_agg) {
      kp.gmask = a->grid_masks;
      kp.gmB = a->grid_masks_stride[0]; kp.gmCam = a->grid_masks_stride[1]; kp.gmD = a->grid_masks_stride[2];
      kp.gmY = a->grid_masks_stride[3]; kp.gmX = a->grid_masks_stride[4];
    }
  }
  kp.vol = a->vol;
  kp.volChannelsLast = (a->flags & MVSGI_FLAG_VOL_CHANNELS_LAST) ? 1 : 0;
  kp.featsBf16 = (a->flags & MVSGI_FLAG_FEATS_BF16) ? 1 : 0;
  kp.volBf16 = (a->flags & MVSGI_FLAG_VOL_BF16) ? 1 : 0;
  kp.validOut = std_agg ? a->valid_out : nullptr;
  kp.dOff = a->cand_offset;
  kp.Dvol = a->vol_num_cands > 0 ? a->vol_num_cands : a->num_cands;
  kp.numPeers = a->num_peers;
  for (int i = 0; i < a->num_peers; ++i) kp.peerDelta[i] = (long long)((reinterpret_cast<intptr_t>(a->peer_vols[i]) - reinterpret_cast<intptr_t>(a->vol)) / 4);
}

cudaError_t launch_pack(const void* masks, int dtype, int n0, int64_t s0, int n1, int64_t s1, int64_t sh, int64_t sw,
                        int Hm, int Wm, uint32_t* bits, int* flags, cudaStream_t st) {
  const int pitch = mask_pitch_words(Wm);
  if (Hm + 2 > 65535 * 32767 || n0 * (int64_t)n1 > 65535) return cudaErrorInvalidValue;
  dim3 grid(Hm + 2, n0 * n1);
  const int esz = dtype == MVSGI_MASK_F32 ? 4 : 1;
  // 4 pixels per lane in one load: rows contiguous and every row start aligned to 4 elements
  const bool vec = sw == 1 && Wm % 4 == 0 && sh % 4 == 0 && s0 % 4 == 0 && s1 % 4 == 0 &&
                   (reinterpret_cast<uintptr_t>(masks) % (4 * esz)) == 0;
  if (dtype == MVSGI_MASK_F32) {
    const float* m = (const float*)masks;
    if (vec) mvsgi::pack_masks_kernel<float, true><<<grid, mvsgi::kPackThreads, 0, st>>>(m, n1, s0, s1, sh, sw, Hm, Wm, pitch, bits, flags);
    else mvsgi::pack_masks_kernel<float, false><<<grid, mvsgi::kPackThreads, 0, st>>>(m, n1, s0, s1, sh, sw, Hm, Wm, pitch, bits, flags);
  } else {
    const uint8_t* m = (const uint8_t*)masks;
    if (vec) mvsgi::pack_masks_kernel<uint8_t, true><<<grid, mvsgi::kPackThreads, 0, st>>>(m, n1, s0, s1, sh, sw, Hm, Wm, pitch, bits, flags);
    else mvsgi::pack_masks_kernel<uint8_t, false><<<grid, mvsgi::kPackThreads, 0, st>>>(m, n1, s0, s1, sh, sw, Hm, Wm, pitch, bits, flags);
  }
  g_launches++;
  return cudaGetLastError();
}

template <int DREG, bool FULL>
void launch_x2_flags(const mvsgi::SoftKP& kp, dim3 grid, cudaStream_t st) {
  const bool nc = kp.normCosts != nullptr, am = kp.argmax != nullptr;
  if (nc && am) mvsgi::softargmin_x2_kernel<DREG, FULL, true, true><<<grid, 128, 0, st>>>(kp);
  else if (nc) mvsgi::softargmin_x2_kernel<DREG, FULL, true, false><<<grid, 128, 0, st>>>(kp);
  else if (am) mvsgi::softargmin_x2_kernel<DREG, FULL, false, true><<<grid, 128, 0, st>>>(kp);
  else mvsgi::softargmin_x2_kernel<DREG, FULL, false, false><<<grid, 128, 0, st>>>(kp);
}

void launch_softargmin_x2(const mvsgi::SoftKP& kp, dim3 grid, cudaStream_t st) {
  if (kp.D == 16) launch_x2_flags<16, true>(kp, grid, st);
  else if (kp.D == 8) launch_x2_flags<8, true>(kp, grid, st);
  else if (kp.D < 8) launch_x2_flags<8, false>(kp, grid, st);
  else launch_x2_flags<16, false>(kp, grid, st);
}

#ifdef MVSGI_EXPERIMENTS
void fill_tp(const MvsgiSweepArgs* a, const SweepPlan& p, unsigned char* table, int* queue, mvsgi::TiledKP& tp) {
  tp.tileH = p.tile_h; tp.octPerRow = p.oct_per_row; tp.tilesX = p.tiles_x; tp.numTiles = p.num_tiles;
  tp.stepsPerItem = p.steps_per_item; tp.itemsPerTile = p.items_per_tile; tp.frameGroups = p.frame_groups;
  tp.numItems = p.tiled_items;
  tp.chunkBytes = p.chunk_bytes; tp.minfoOff = p.minfo_off; tp.vinfoOff = p.vinfo_off;
  tp.fX = p.direct ? (int)a->feats_stride[4] : mvsgi::kCG;
  tp.direct = p.direct ? 1 : 0;
  tp.table = table; tp.queue = queue;
}
#endif

}  // namespace

namespace {

// the forward problem as the backward sees it: no outputs, no peers, workspace laid out by the forward plan
int backward_plan(const MvsgiSweepBwdArgs* a, MvsgiSweepArgs* f, SweepPlan* p) {
  if (a == nullptr || a->fwd == nullptr) return MVSGI_E_NULL;
  if (a->struct_size != sizeof(MvsgiSweepBwdArgs)) return MVSGI_E_VERSION;
  if (a->flags != 0) return MVSGI_E_UNSUPPORTED;
  if (a->grad_vol == nullptr || a->grad_feats == nullptr) return MVSGI_E_NULL;
  *f = *a->fwd;
  f->vol = const_cast<float*>(a->grad_vol);  // only has to be non-NULL for validation; never written
  f->valid_out = nullptr;
  f->num_peers = 0;
  f->octet_order = nullptr;
  f->order_runs = f->order_units = 0;
  f->cta_times_out = nullptr;
  f->tap_table_mode = MVSGI_TAP_TABLE_OFF;  // the backward's plan is its own (one frame per pass, no balanced runs)
  // Channels-last features tapped in place (the forward's own rule) when the gradient has the same dense layout: the
  // scatter then goes straight into grad_feats.  Otherwise features are staged and the gradient image has the staged
  // layout (un-staged into grad_feats afterwards).
  f->flags = MVSGI_FLAG_CTA_TABLE;
  int rc = validate_and_plan(f, p);
  if (rc != MVSGI_OK) return rc;
  bool same = p->direct;
  const int64_t* fs = f->feats_stride;
  for (int i = 0; i < 5 && same; ++i) same = a->grad_feats_stride[i] == fs[i];
  // dense: the frames tile memory without gaps, so one memset clears the gradient
  same = same && fs[4] == f->feat_chs && fs[3] == fs[4] * f->feat_w && fs[1] == fs[3] * f->feat_h &&
         fs[0] == fs[1] * f->num_cams && (reinterpret_cast<uintptr_t>(a->grad_feats) & 15) == 0;
  if (same) return MVSGI_OK;
  f->flags = MVSGI_FLAG_CTA_TABLE | MVSGI_FLAG_STAGED_FEATS;
  return validate_and_plan(f, p);
}

}  // namespace

extern "C" {

int mvsgi_abi_version(void) { return MVSGI_ABI_VERSION; }

int64_t mvsgi_launch_counter(void) { return g_launches.load(); }

const char* mvsgi_strerror(int code) {
  switch (code) {
    case MVSGI_OK: return "ok";
    case MVSGI_E_NULL: return "required pointer is NULL";
    case MVSGI_E_SHAPE: return "bad shape (size <= 0, over a limit, feat_chs not a multiple of 4, or rig/shape mismatch)";
    case MVSGI_E_STRIDE: return "unsupported stride or alignment";
    case MVSGI_E_UNSUPPORTED: return "unsupported enum value or flag";
    case MVSGI_E_WORKSPACE: return "workspace too small or not 256-byte aligned";
    case MVSGI_E_VERSION: return "struct_size does not match this library (ABI mismatch)";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown mvsgi error";
  }
}

int32_t mvsgi_mask_pitch_words(int32_t mask_w) { return mask_pitch_words(mask_w); }

int64_t mvsgi_sweep_workspace_bytes(const MvsgiSweepArgs* a) {
  SweepPlan p;
  const int rc = validate_and_plan(a, &p);
  return rc != MVSGI_OK ? (int64_t)rc : p.total_bytes;
}

int mvsgi_sweep_balance_shape(const MvsgiSweepArgs* a, int32_t* runs_out, int32_t* units_out, int32_t* num_octets_out,
                              int32_t* ctas_out) {
  if (a == nullptr || runs_out == nullptr || units_out == nullptr || num_octets_out == nullptr || ctas_out == nullptr)
    return MVSGI_E_NULL;
  MvsgiSweepArgs b = *a;
  b.octet_order = nullptr;
  b.order_runs = b.order_units = 0;
  b.tap_table_mode = MVSGI_TAP_TABLE_OFF;
  if (b.vol == nullptr) b.vol = reinterpret_cast<float*>(static_cast<uintptr_t>(256));  // never written here: the problem may be described without outputs
  SweepPlan p;
  const int rc = validate_and_plan(&b, &p);
  if (rc != MVSGI_OK) return rc;
  if (!p.one_item_per_cta) return MVSGI_E_UNSUPPORTED;
  const int num_oct = (a->out_h * a->out_w + 7) / 8;
  // k whole waves of runs, k as in the default plan (its runs fit the table; more, smaller runs fit as well)
  const int waves = (p.num_runs + p.grid_target - 1) / p.grid_target;
  const int runs = num_oct < p.grid_target ? num_oct : waves * p.grid_target;
  *runs_out = runs;
  *units_out = (num_oct + runs - 1) / runs;
  *num_octets_out = num_oct;
  *ctas_out = runs < p.grid_target ? runs : p.grid_target;
  return MVSGI_OK;
}

int mvsgi_sweep_grid(const MvsgiSweepArgs* a) {
  if (a == nullptr) return MVSGI_E_NULL;
  MvsgiSweepArgs b = *a;
  if (b.vol == nullptr) b.vol = reinterpret_cast<float*>(static_cast<uintptr_t>(256));
  SweepPlan p;
  const int rc = validate_and_plan(&b, &p);
  if (rc != MVSGI_OK) return rc;
  return p.tiled ? p.tiled_grid : p.grid;
}

int64_t mvsgi_sweep_tap_table_bytes(const MvsgiSweepArgs* a) {
  if (a == nullptr) return MVSGI_E_NULL;
  MvsgiSweepArgs b = *a;
  b.tap_table = nullptr;
  b.tap_table_bytes = 0;
  b.tap_table_mode = MVSGI_TAP_TABLE_OFF;
  if (b.vol == nullptr) b.vol = reinterpret_cast<float*>(static_cast<uintptr_t>(256));
  SweepPlan p;
  const int rc = validate_and_plan(&b, &p);
  if (rc != MVSGI_OK) return rc;
  return p.tap_bytes;
}

int mvsgi_sweep_octet_cost(const MvsgiSweepArgs* a, int32_t* cost_out, void* stream) {
  if (a == nullptr || cost_out == nullptr) return MVSGI_E_NULL;
  MvsgiSweepArgs b = *a;
  b.tap_table_mode = MVSGI_TAP_TABLE_OFF;
  b.octet_order = nullptr;
  b.order_runs = b.order_units = 0;
  if (b.vol == nullptr) b.vol = reinterpret_cast<float*>(static_cast<uintptr_t>(256));  // never written here: the problem may be described without outputs
  SweepPlan p;
  const int rc = validate_and_plan(&b, &p);
  if (rc != MVSGI_OK) return rc;
  const bool std_agg = a->aggregate == MVSGI_AGG_STD;
  if (p.per_frame_grids || (std_agg && p.pack_masks)) return MVSGI_E_UNSUPPORTED;  // packed masks only
  mvsgi::SweepKP kp;
  fill_kp(&b, p, nullptr, nullptr, kp);
  MvsgiRig rig;
  if (a->grid_mode == MVSGI_GRID_PROJECT) rig = *a->rig;
  else memset(&rig, 0, sizeof(rig));
  cudaStream_t st = (cudaStream_t)stream;
  const int num_oct = (kp.P + 7) / 8;
  cudaError_t e = cudaMemsetAsync(cost_out, 0, (size_t)num_oct * sizeof(int32_t), st);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((kp.P + 127) / 128, a->num_cands);
  if (grid.y > 65535) return MVSGI_E_SHAPE;
#define MVSGI_COST(CAM)                                                                                      \
  if (std_agg) mvsgi::octet_cost_kernel<CAM, MVSGI_AGG_STD><<<grid, 128, 0, st>>>(kp, rig, cost_out);         \
  else mvsgi::octet_cost_kernel<CAM, MVSGI_AGG_CONCAT><<<grid, 128, 0, st>>>(kp, rig, cost_out)
  switch (a->num_cams) {
    case 1: MVSGI_COST(1); break;
    case 2: MVSGI_COST(2); break;
    case 3: MVSGI_COST(3); break;
    case 4: MVSGI_COST(4); break;
    default: return MVSGI_E_UNSUPPORTED;
  }
#undef MVSGI_COST
  g_launches++;
  return (int)cudaGetLastError();
}

int mvsgi_sweep_launch_count(const MvsgiSweepArgs* a) {
  SweepPlan p;
  const int rc = validate_and_plan(a, &p);
  if (rc != MVSGI_OK) return rc;
  const int stage = (p.direct ? 0 : 1) + (p.pack_masks ? 1 : 0);
  if (a->flags & MVSGI_FLAG_STAGE_ONLY) return stage;
  const int sweep = p.tiled ? 2 : 1;  // geometry table + sweep
  if (a->flags & MVSGI_FLAG_SWEEP_ONLY) return sweep;
  return sweep + stage;
}

int mvsgi_pack_masks(const void* masks, int32_t dtype, int32_t n_outer, int64_t stride_outer, int32_t n_inner,
                     int64_t stride_inner, int64_t stride_h, int64_t stride_w, int32_t mask_h, int32_t mask_w,
                     uint32_t* bits_out, int32_t* flags_out, void* stream) {
  if (masks == nullptr || bits_out == nullptr) return MVSGI_E_NULL;
  if (n_outer < 1 || n_inner < 1 || mask_h < 1 || mask_w < 1) return MVSGI_E_SHAPE;
  if (dtype != MVSGI_MASK_F32 && dtype != MVSGI_MASK_U8) return MVSGI_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(bits_out) & 7) != 0) return MVSGI_E_STRIDE;  // row pairs are read as 8-byte words
  return (int)launch_pack(masks, dtype, n_outer, stride_outer, n_inner, stride_inner, stride_h, stride_w, mask_h,
                          mask_w, bits_out, flags_out, (cudaStream_t)stream);
}

int mvsgi_sweep_f32(const MvsgiSweepArgs* a, void* stream) {
  SweepPlan p;
  int rc = validate_and_plan(a, &p);
  if (rc != MVSGI_OK) return rc;
  if (a->workspace == nullptr) return MVSGI_E_NULL;
  if (a->workspace_bytes < p.total_bytes || (reinterpret_cast<uintptr_t>(a->workspace) & 255) != 0)
    return MVSGI_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)a->workspace;
  float* ws_feats = (float*)ws;
  uint32_t* ws_bits = (uint32_t*)(ws + p.feats_bytes);
  int* ws_flags = (int*)(ws + p.feats_bytes + p.mask_bytes);
  const bool do_stage = !(a->flags & MVSGI_FLAG_SWEEP_ONLY);
  const bool do_sweep = !(a->flags & MVSGI_FLAG_STAGE_ONLY);
  const int Wp = a->feat_w + 2, Hp1 = a->feat_h + 3;

  // 1. features -> zero-padded channel-group-major channels-last (not when they are tapped where they are)
  if (do_stage && !p.direct) {
    const int N = a->batch * a->num_cams;
    const int x_blocks = (Wp + mvsgi::kStagePx - 1) / mvsgi::kStagePx;
    if ((int64_t)N * p.cg * x_blocks >= (1ll << 31) || Hp1 > 65535) return MVSGI_E_SHAPE;
    dim3 grid((unsigned)(N * p.cg * x_blocks), Hp1);
    const int chan_fastest = (a->feats_stride[2] == 1 && a->feats_stride[4] != 1) ? 1 : 0;
    if (a->flags & MVSGI_FLAG_FEATS_BF16)
      mvsgi::stage_feats_kernel<true><<<grid, 256, 0, st>>>(a->feats, a->feats_stride[0], a->num_cams,
                                                            a->feats_stride[1], a->feats_stride[2], a->feats_stride[3],
                                                            a->feats_stride[4], a->feat_chs, p.cg, a->feat_h,
                                                            a->feat_w, x_blocks, chan_fastest, ws_feats);
    else
      mvsgi::stage_feats_kernel<false><<<grid, 256, 0, st>>>(a->feats, a->feats_stride[0], a->num_cams,
                                                             a->feats_stride[1], a->feats_stride[2], a->feats_stride[3],
                                                             a->feats_stride[4], a->feat_chs, p.cg, a->feat_h,
                                                             a->feat_w, x_blocks, chan_fastest, ws_feats);
    g_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }

  // 2. masks -> bits
  if (p.pack_masks && do_stage) {
    cudaError_t e = cudaMemsetAsync(ws_flags, 0, sizeof(int), st);  // bit 0: a mask value was negative or NaN
    if (e != cudaSuccess) return (int)e;
    e = launch_pack(a->masks, a->masks_dtype, p.per_frame_masks ? a->batch : 1, a->masks_stride[0],
                                a->num_cams, a->masks_stride[1], a->masks_stride[2], a->masks_stride[3], a->mask_h,
                                a->mask_w, ws_bits, ws_flags, st);
    if (e != cudaSuccess) return (int)e;
  }

  // 3. the sweep
  if (!do_sweep) return MVSGI_OK;
  mvsgi::SweepKP kp;
  fill_kp(a, p, ws_feats, ws_bits, kp);

  MvsgiRig rig;
  if (a->grid_mode == MVSGI_GRID_PROJECT) rig = *a->rig;
  else memset(&rig, 0, sizeof(rig));

  cudaError_t e;
#ifdef MVSGI_EXPERIMENTS
  if (p.tiled) {
    // geometry table (once per call, shared by all frames), then the item queue counter, then the sweep
    int* queue = ws_flags + 32;
    mvsgi::TiledKP tp;
    fill_tp(a, p, (unsigned char*)(ws + p.feats_bytes + p.mask_bytes + p.flag_bytes), queue, tp);
    e = mvsgi::launch_tile_table(a->aggregate, a->num_cams, kp, tp, rig, st);
    if (e != cudaSuccess) return (int)e;
    g_launches++;
    e = cudaMemsetAsync(queue, 0, sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
    e = mvsgi::launch_sweep_tiled(a->aggregate, a->num_cams, kp, tp, p.tiled_grid, p.frames_per_pass, st);
    g_launches++;
    return (int)e;
  }
  if (p.walk) {
    e = mvsgi::launch_sweep_walk(a->aggregate, a->num_cams, kp, rig, p.grid, p.smem_bytes, st);
    g_launches++;
    return (int)e;
  }
#endif
  // kept tap table and several items per CTA: the items after the first are drawn from a queue (sweep.cuh); its counter
  // is 4 bytes of the call's workspace, zeroed on the stream (a memset node when captured into a CUDA graph)
  if (a->tap_table_mode == MVSGI_TAP_TABLE_USE && p.num_items > p.grid && !p.oct_order) {
    kp.itemQueue = ws_flags + 32;
    e = cudaMemsetAsync(kp.itemQueue, 0, sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
  }
  e = mvsgi::launch_sweep_cta(a->aggregate, a->num_cams, kp, rig, p.grid, p.smem_bytes, st);
  g_launches++;
  return (int)e;
}

int mvsgi_project_only_f32(const MvsgiProjectArgs* a, void* stream) {
  if (a == nullptr || a->rig == nullptr || a->grids == nullptr || a->grid_masks == nullptr) return MVSGI_E_NULL;
  if (a->struct_size != sizeof(MvsgiProjectArgs)) return MVSGI_E_VERSION;
  const MvsgiRig& r = *a->rig;
  if (r.num_cams < 1 || r.num_cams > MVSGI_MAX_CAMS || r.num_cands < 1 || r.num_cands > MVSGI_MAX_CANDS ||
      r.cv_width < 1 || r.cv_height < 1)
    return MVSGI_E_SHAPE;
  const int n = r.cv_width * r.cv_height;
  mvsgi::project_only_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(r, a->grids, a->grid_masks);
  g_launches++;
  return (int)cudaGetLastError();
}

int mvsgi_resample_f32(const MvsgiResampleArgs* a, void* stream) {
  if (a == nullptr || a->rig == nullptr || a->img == nullptr || a->out == nullptr) return MVSGI_E_NULL;
  if (a->struct_size != sizeof(MvsgiResampleArgs)) return MVSGI_E_VERSION;
  if (a->flags & ~(MVSGI_RESAMPLE_IMG_U8 | MVSGI_RESAMPLE_NEAREST)) return MVSGI_E_UNSUPPORTED;
  const MvsgiRig& r = *a->rig;
  if (r.num_cams != 1 || r.num_cands != 1 || r.dist[0] != 1.0f) return MVSGI_E_SHAPE;
  if (r.cams[0].model < 0 || r.cams[0].model > MVSGI_CAM_LINEAR_SPHERE) return MVSGI_E_UNSUPPORTED;
  // an Equirectangular raw image is sampled with border padding by the reference (camera_models.py:633); only the
  // zero padding of the fisheye models is implemented
  if (r.cams[0].model == MVSGI_CAM_EQUIRECT) return MVSGI_E_UNSUPPORTED;
  if (a->batch < 1 || a->channels < 1 || a->in_h < 1 || a->in_w < 1 || r.cv_width < 1 || r.cv_height < 1)
    return MVSGI_E_SHAPE;
  if (a->in_h != r.cams[0].height || a->in_w != r.cams[0].width) return MVSGI_E_SHAPE;  // check_input_shape :92-96
  if ((int64_t)r.cv_width * r.cv_height >= (1ll << 31)) return MVSGI_E_SHAPE;
  mvsgi::ResampleKP kp;
  kp.N = a->batch; kp.C = a->channels; kp.H = a->in_h; kp.W = a->in_w; kp.Ht = r.cv_height; kp.Wt = r.cv_width;
  kp.img = a->img;
  kp.sN = a->img_stride[0]; kp.sC = a->img_stride[1]; kp.sH = a->img_stride[2]; kp.sW = a->img_stride[3];
  kp.invalid = a->invalid_value;
  kp.nearest = (a->flags & MVSGI_RESAMPLE_NEAREST) ? 1 : 0;
  kp.out = a->out;
  kp.validOut = a->valid_out;
  // the geometry is evaluated once per thread and reused for its images: split the batch over blockIdx.y only as far as
  // needed to fill the GPU
  const int blocks = ((kp.Wt + 31) / 32) * ((kp.Ht + mvsgi::kResampleThreads / 32 - 1) / (mvsgi::kResampleThreads / 32));
  int gy = 1;
  while (blocks * gy < device_sms() * 16 && gy * 2 <= a->batch && gy < 64) gy *= 2;
  const int gy_env = env_int("MVSGI_RESAMPLE_GY", 0);
  if (gy_env > 0) gy = gy_env < a->batch ? gy_env : a->batch;
  dim3 grid(blocks, gy);
  const bool u8 = (a->flags & MVSGI_RESAMPLE_IMG_U8) != 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (a->channels % 3 == 0) {
    if (u8) mvsgi::resample_kernel<3, uint8_t><<<grid, mvsgi::kResampleThreads, 0, st>>>(kp, r);
    else mvsgi::resample_kernel<3, float><<<grid, mvsgi::kResampleThreads, 0, st>>>(kp, r);
  } else if (a->channels % 4 == 0) {
    if (u8) mvsgi::resample_kernel<4, uint8_t><<<grid, mvsgi::kResampleThreads, 0, st>>>(kp, r);
    else mvsgi::resample_kernel<4, float><<<grid, mvsgi::kResampleThreads, 0, st>>>(kp, r);
  } else {
    if (u8) mvsgi::resample_kernel<1, uint8_t><<<grid, mvsgi::kResampleThreads, 0, st>>>(kp, r);
    else mvsgi::resample_kernel<1, float><<<grid, mvsgi::kResampleThreads, 0, st>>>(kp, r);
  }
  g_launches++;
  return (int)cudaGetLastError();
}

// ---- backward passes -------------------------------------------------------------------------------------------------

int64_t mvsgi_sweep_backward_workspace_bytes(const MvsgiSweepBwdArgs* a) {
  MvsgiSweepArgs f;
  SweepPlan p;
  const int rc = backward_plan(a, &f, &p);
  if (rc != MVSGI_OK) return (int64_t)rc;
  return p.direct ? p.total_bytes : p.total_bytes + p.feats_bytes;
}

int mvsgi_sweep_backward_f32(const MvsgiSweepBwdArgs* a, void* stream) {
  MvsgiSweepArgs f;
  SweepPlan p;
  int rc = backward_plan(a, &f, &p);
  if (rc != MVSGI_OK) return rc;
  if (a->workspace == nullptr) return MVSGI_E_NULL;
  if (p.tiled || p.walk) return MVSGI_E_UNSUPPORTED;  // experiment plans have no backward
  if (a->workspace_bytes < p.total_bytes + (p.direct ? 0 : p.feats_bytes) ||
      (reinterpret_cast<uintptr_t>(a->workspace) & 255) != 0)
    return MVSGI_E_WORKSPACE;
  if (f.batch > 65535 || f.num_cands > 65535 || f.feat_h > 65535) return MVSGI_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)a->workspace;
  if (p.direct) {
    // channels-last features and gradient: pack the masks (if they are not packed already), clear grad_feats, and run
    // the backward instantiation of the sweep kernel on the caller's tensors — no staged copies
    f.workspace = ws;
    f.workspace_bytes = p.total_bytes;
    f.flags |= MVSGI_FLAG_STAGE_ONLY;
    rc = mvsgi_sweep_f32(&f, stream);
    if (rc != MVSGI_OK) return rc;
    const size_t gbytes = (size_t)f.batch * f.feats_stride[0] * sizeof(float);
    cudaError_t e = cudaMemsetAsync(a->grad_feats, 0, gbytes, st);
    if (e != cudaSuccess) return (int)e;
    mvsgi::SweepKP kp;
    fill_kp(&f, p, (float*)ws, (uint32_t*)(ws + p.feats_bytes), kp);
    kp.vol = const_cast<float*>(a->grad_vol);
    kp.volChannelsLast = 0;
    kp.gradFeat = a->grad_feats;
    kp.framesPerPass = 1;
    MvsgiRig rig;
    if (f.grid_mode == MVSGI_GRID_PROJECT) rig = *f.rig;
    else memset(&rig, 0, sizeof(rig));
    e = mvsgi::launch_sweep_bwd(f.aggregate, f.num_cams, kp, rig, p.grid, p.smem_bytes, st);
    g_launches++;
    return (int)e;
  }
  // 1. stage feats + pack masks exactly as the forward does
  f.workspace = ws;
  f.workspace_bytes = p.total_bytes;
  f.flags |= MVSGI_FLAG_STAGE_ONLY;  // on top of the plan's flags (staged features, no geometry table)
  rc = mvsgi_sweep_f32(&f, stream);
  if (rc != MVSGI_OK) return rc;
  // 2. zero the staged gradient image
  float* gstage = (float*)(ws + p.total_bytes);
  cudaError_t e = cudaMemsetAsync(gstage, 0, (size_t)p.feats_bytes, st);
  if (e != cudaSuccess) return (int)e;
  // 3. scatter: the backward instantiation of the sweep kernel (sweep.cuh, BWD) on the staged copies
  mvsgi::SweepKP kp;
  fill_kp(&f, p, (float*)ws, (uint32_t*)(ws + p.feats_bytes), kp);
  kp.vol = const_cast<float*>(a->grad_vol);
  kp.volChannelsLast = 0;
  kp.gradFeat = gstage;
  kp.framesPerPass = 1;
  MvsgiRig rig;
  if (f.grid_mode == MVSGI_GRID_PROJECT) rig = *f.rig;
  else memset(&rig, 0, sizeof(rig));
  e = mvsgi::launch_sweep_bwd(f.aggregate, f.num_cams, kp, rig, p.grid, p.smem_bytes, st);
  g_launches++;
  if (e != cudaSuccess) return (int)e;
  // 4. staged gradient -> grad_feats
  const int N = f.batch * f.num_cams;
  const int x_blocks = (f.feat_w + mvsgi::kStagePx - 1) / mvsgi::kStagePx;
  if ((int64_t)N * p.cg * x_blocks >= (1ll << 31)) return MVSGI_E_SHAPE;
  dim3 grid((unsigned)(N * p.cg * x_blocks), f.feat_h);
  mvsgi::unstage_grad_kernel<<<grid, 256, 0, st>>>(gstage, f.num_cams, f.feat_chs, p.cg, f.feat_h, f.feat_w,
                                                   a->grad_feats, a->grad_feats_stride[0], a->grad_feats_stride[1],
                                                   a->grad_feats_stride[2], a->grad_feats_stride[3],
                                                   a->grad_feats_stride[4]);
  g_launches++;
  return (int)cudaGetLastError();
}

constexpr int kGatherRows = 4;  // softargmin_bwd.cuh: kSbRows

int mvsgi_softargmin_backward_f32(const MvsgiSoftArgminBwdArgs* a, void* stream) {
  if (a == nullptr || a->fwd == nullptr) return MVSGI_E_NULL;
  if (a->struct_size != sizeof(MvsgiSoftArgminBwdArgs)) return MVSGI_E_VERSION;
  if (a->flags != 0) return MVSGI_E_UNSUPPORTED;
  const MvsgiSoftArgminArgs* f = a->fwd;
  if (f->struct_size != sizeof(MvsgiSoftArgminArgs)) return MVSGI_E_VERSION;
  if (f->costs == nullptr || f->inv_dist_idx == nullptr || a->grad_costs == nullptr) return MVSGI_E_NULL;
  if (f->batch < 1 || f->num_cands < 1 || f->num_cands > MVSGI_MAX_CANDS || f->in_h < 1 || f->in_w < 1 ||
      f->out_h < 1 || f->out_w < 1 || f->batch > 65535 || f->out_h > 65535)
    return MVSGI_E_SHAPE;
  mvsgi::SoftBwdKP kb;
  mvsgi::SoftKP& kp = kb.f;
  kp.B = f->batch; kp.D = f->num_cands; kp.Hin = f->in_h; kp.Win = f->in_w; kp.Hout = f->out_h; kp.Wout = f->out_w;
  kp.interp = (f->out_h != f->in_h || f->out_w != f->in_w) ? 1 : 0;
  kp.scaleH = f->src_scale_h; kp.scaleW = f->src_scale_w;
  kp.costs = f->costs;
  kp.sB = f->costs_stride[0]; kp.sD = f->costs_stride[1]; kp.sY = f->costs_stride[2]; kp.sX = f->costs_stride[3];
  kp.idx = f->inv_dist_idx; kp.invDist = nullptr; kp.normCosts = nullptr; kp.argmax = nullptr;
  kb.gInv = a->grad_inv_dist; kb.gNc = a->grad_norm_costs; kb.gradCosts = a->grad_costs;
  kb.gB = a->grad_costs_stride[0]; kb.gD = a->grad_costs_stride[1]; kb.gY = a->grad_costs_stride[2];
  kb.gX = a->grad_costs_stride[3];
  cudaStream_t st = (cudaStream_t)stream;
  const bool x2 = kp.interp && f->out_h == 2 * f->in_h && f->out_w == 2 * f->in_w && f->src_scale_h == 0.5f &&
                  f->src_scale_w == 0.5f && f->num_cands <= 16 && (int64_t)f->out_h * f->out_w < (1ll << 31) &&
                  f->costs_stride[1] >= 0 && f->costs_stride[2] >= 0 && f->costs_stride[3] >= 0 &&
                  (f->in_h + kGatherRows - 1) / kGatherRows <= 65535 && !env_int("MVSGI_SOFTARGMIN_BWD_ATOMIC", 0);
  if (x2) {  // configs/base_model.yaml:28-34: every input pixel gathers its four 2x2 blocks (softargmin_bwd.cuh)
    mvsgi::launch_softargmin_x2_backward(kb, st);
    g_launches++;
    return (int)cudaGetLastError();
  }
  if (kp.interp) {  // the upsample's adjoint accumulates with atomics
    const long long total = (long long)kp.B * kp.D * kp.Hin * kp.Win;
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    mvsgi::zero_strided4_kernel<<<blocks, 256, 0, st>>>(a->grad_costs, kp.B, kp.D, kp.Hin, kp.Win, kb.gB, kb.gD, kb.gY,
                                                        kb.gX);
    g_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  dim3 grid((kp.Wout + 127) / 128, kp.Hout, kp.B);
  mvsgi::softargmin_backward_kernel<<<grid, 128, 0, st>>>(kb);
  g_launches++;
  return (int)cudaGetLastError();
}

// ---- peer-mapped volumes (explicit allocation helpers; see the header) -------------------------------------------
int mvsgi_p2p_alloc(int64_t bytes, void** ptr_out) {
  if (ptr_out == nullptr) return MVSGI_E_NULL;
  if (bytes <= 0) return MVSGI_E_SHAPE;
  return (int)cudaMalloc(ptr_out, (size_t)bytes);
}

int mvsgi_p2p_free(void* ptr) { return ptr ? (int)cudaFree(ptr) : MVSGI_E_NULL; }

int mvsgi_p2p_export(const void* ptr, uint8_t handle_out[MVSGI_P2P_HANDLE_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == MVSGI_P2P_HANDLE_BYTES, "IPC handle size");
  if (ptr == nullptr || handle_out == nullptr) return MVSGI_E_NULL;
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(ptr));
  if (e != cudaSuccess) return (int)e;
  memcpy(handle_out, &h, sizeof(h));
  return MVSGI_OK;
}

int mvsgi_p2p_open(const uint8_t handle[MVSGI_P2P_HANDLE_BYTES], void** ptr_out) {
  if (handle == nullptr || ptr_out == nullptr) return MVSGI_E_NULL;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  return (int)cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess);
}

int mvsgi_p2p_close(void* ptr) { return ptr ? (int)cudaIpcCloseMemHandle(ptr) : MVSGI_E_NULL; }

int mvsgi_softargmin_f32(const MvsgiSoftArgminArgs* a, void* stream) {
  if (a == nullptr) return MVSGI_E_NULL;
  if (a->struct_size != sizeof(MvsgiSoftArgminArgs)) return MVSGI_E_VERSION;
  if (a->costs == nullptr || a->inv_dist_idx == nullptr || a->inv_dist == nullptr) return MVSGI_E_NULL;
  if (a->batch < 1 || a->num_cands < 1 || a->num_cands > MVSGI_MAX_CANDS || a->in_h < 1 || a->in_w < 1 ||
      a->out_h < 1 || a->out_w < 1 || a->batch > 65535 || a->out_h > 65535)
    return MVSGI_E_SHAPE;
  mvsgi::SoftKP kp;
  kp.B = a->batch; kp.D = a->num_cands; kp.Hin = a->in_h; kp.Win = a->in_w; kp.Hout = a->out_h; kp.Wout = a->out_w;
  kp.interp = (a->out_h != a->in_h || a->out_w != a->in_w) ? 1 : 0;
  kp.scaleH = a->src_scale_h; kp.scaleW = a->src_scale_w;
  kp.costs = a->costs;
  kp.sB = a->costs_stride[0]; kp.sD = a->costs_stride[1]; kp.sY = a->costs_stride[2]; kp.sX = a->costs_stride[3];
  kp.idx = a->inv_dist_idx; kp.invDist = a->inv_dist; kp.normCosts = a->norm_costs; kp.argmax = a->argmax;
  cudaStream_t st = (cudaStream_t)stream;
  const bool x2 = kp.interp && a->out_h == 2 * a->in_h && a->out_w == 2 * a->in_w && a->src_scale_h == 0.5f &&
                  a->src_scale_w == 0.5f && a->num_cands <= 16 && a->in_h < 65535 && (int64_t)a->out_h * a->out_w < (1ll << 31) &&
                  a->costs_stride[1] >= 0 && a->costs_stride[2] >= 0 && a->costs_stride[3] >= 0 &&
                  a->costs_stride[1] * a->num_cands + a->costs_stride[2] * a->in_h + a->costs_stride[3] * a->in_w < (1ll << 31) &&
                  !env_int("MVSGI_SOFTARGMIN_GENERIC", 0);
  if (x2) {  // configs/base_model.yaml:28-34 (interp_scale_factor 2, pre_interp): one thread per 2x2 output block
    // input box staged by TMA when the volume fits a tensor map (softargmin_tma.cuh), else pointer-per-load: same bits
    if (env_int("MVSGI_SOFTARGMIN_TMA", 1) == 0 || !mvsgi::launch_softargmin_x2_tma(kp, st)) {
      dim3 grid((a->in_w + 1 + 127) / 128, a->in_h + 1, a->batch);
      launch_softargmin_x2(kp, grid, st);
    }
  } else {
    int split = a->num_cands <= 16 ? 1 : (a->num_cands <= 32 ? 2 : 4);
    const int forced = env_int("MVSGI_SOFTARGMIN_DSPLIT", 0);
    if ((forced == 1 || forced == 2 || forced == 4) && (a->num_cands + (forced > 0 ? forced : 1) - 1) / (forced > 0 ? forced : 1) <= 16)
      split = forced;
    const int ppb = 4 * (32 / split);  // output pixels per 128-thread block
    dim3 grid((a->out_w + ppb - 1) / ppb, a->out_h, a->batch);
    if (split == 1) mvsgi::softargmin_kernel<1><<<grid, 128, 0, st>>>(kp);
    else if (split == 2) mvsgi::softargmin_kernel<2><<<grid, 128, 0, st>>>(kp);
    else mvsgi::softargmin_kernel<4><<<grid, 128, 0, st>>>(kp);
  }
  g_launches++;
  return (int)cudaGetLastError();
}

}  // extern "C"
