/*
 * mvsgi_b200.h — C ABI of libmvsgi_b200.so
 *
 * B200-native (sm_100a) replacement for the spherical-sweep / cost-volume /
 * soft-argmin hot path of castacks/mvs_gi.  Plain pointers and sizes only: no
 * torch types cross this boundary.  Every entry point cites the reference
 * interface it replaces (paths relative to the reference checkout).
 *
 * Ownership: the caller allocates every input, output and workspace buffer.
 * The compute entry points never allocate, free or retain device memory;
 * pointers are borrowed for the duration of the enqueue.  (The only calls that
 * allocate are the explicit mvsgi_p2p_* helpers at the end of this file.)  Calibration (MvsgiRig) is read
 * from HOST memory at enqueue time and passed to the kernels by value.
 *
 * Threading/streams: re-entrant; calls from several host threads and streams do not
 * interact.  The only process-wide state is an atomic launch counter
 * (mvsgi_launch_counter, diagnostics) and write-once caches of device attributes
 * (SM count, resident CTAs and opt-in shared-memory size per kernel); none of it
 * changes what a call computes.  Work is enqueued on
 * the stream that is passed in (a cudaStream_t cast to void*); no call
 * synchronises the device, so all of them are safe under CUDA-graph capture.
 *
 * Errors: 0 = enqueued.  Negative = MVSGI_E_* argument error (nothing was
 * enqueued).  Positive = cudaError_t from the launch.  No exceptions, no exit.
 * There is no CPU fallback anywhere behind this ABI.
 */
#ifndef MVSGI_B200_H_
#define MVSGI_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVSGI_ABI_VERSION 13
#define MVSGI_MAX_CAMS 4
#define MVSGI_MAX_CANDS 64
#define MVSGI_MAX_PEERS 7 /* extra destinations of a candidate-sharded sweep (8 GPUs of one box) */

/* ---- error codes ------------------------------------------------------- */
#define MVSGI_OK 0
#define MVSGI_E_NULL (-1)        /* required pointer is NULL                 */
#define MVSGI_E_SHAPE (-2)       /* a size is <= 0 or exceeds a limit        */
#define MVSGI_E_STRIDE (-3)      /* unsupported stride / alignment           */
#define MVSGI_E_UNSUPPORTED (-4) /* enum value / combination not implemented */
#define MVSGI_E_WORKSPACE (-5)   /* workspace too small or misaligned        */
#define MVSGI_E_VERSION (-6)     /* struct_size does not match this ABI      */

/* ---- enums ------------------------------------------------------------- */
/* Camera models of dsta_mvs/mvs_utils/camera_models.py that the sweep can
 * project into (Equirectangular :577, DoubleSphere :388, LinearSphere :1427). */
enum { MVSGI_CAM_EQUIRECT = 0, MVSGI_CAM_DOUBLE_SPHERE = 1, MVSGI_CAM_LINEAR_SPHERE = 2 };
/* Where the sampling grid comes from. */
enum {
  MVSGI_GRID_SUPPLIED = 0, /* caller passes grids/grid_masks (reference interface) */
  MVSGI_GRID_PROJECT = 1   /* evaluated in-kernel from MvsgiRig                    */
};
/* Cross-view aggregation = which cv_builder (configs/cv_builder/{std,concat}_cvbuilder.yaml). */
enum {
  MVSGI_AGG_STD = 0,   /* SphericalSweepStdMasked.sweep  spherical_sweep_avg.py:38-136 */
  MVSGI_AGG_CONCAT = 1 /* SphericalSweep.sweep           spherical_sweep.py:38-68      */
};
/* Element type of the masks tensor.  MVSGI_MASK_PACKED: `masks` already holds what mvsgi_pack_masks produced for
 * these images (uint32 words, 8-byte aligned, at least 16 readable bytes after the last image): masks_stride[0] =
 * words between the image sets of consecutive frames (0 = one set shared by all frames), masks_stride[1] = words
 * between the cameras of a set (>= (mask_h + 2) * pitch_words * 2), masks_stride[2..3] unused.  The sweep then
 * launches no packing kernel and reads no float masks: a caller whose masks are static (inference,
 * api/inference_pytorch.py keeps them on the device) packs them once and pays no mask traffic per call. */
enum { MVSGI_MASK_F32 = 0, MVSGI_MASK_U8 = 1, MVSGI_MASK_PACKED = 2 };
/* flags for MvsgiSweepArgs.flags */
/* Stage the features as bfloat16 (round to nearest even; taps widened back to fp32, all arithmetic in fp32): halves
 * the bytes every tap moves through L1/L2.  Stated-tolerance mode: max|vol - fp32 vol| <= 1e-2 * max|vol|, validity
 * mask unchanged (SURVEY.md §8c criterion 4).  Not available with valid_out or peer_vols; the backward ignores it. */
#define MVSGI_FLAG_FEATS_BF16 1u
#define MVSGI_FLAG_STAGE_ONLY 2u /* only fill the workspace (channels-last feats, packed masks)          */
#define MVSGI_FLAG_SWEEP_ONLY 4u /* workspace already holds what STAGE_ONLY produced for these args      */
/* vol (and every peer_vols[i]) is laid out [B, Dvol, Ho, Wo, Cout] in memory — the torch.channels_last_3d strides of
 * a [B, Cout, Dvol, Ho, Wo] tensor, what cuDNN's NDHWC Conv3d (post_vol, spherical_sweep_avg.py:30-36,165;
 * common_modules.py:82-115) consumes without a transpose.  Needs feat_chs % 4 == 0 and 16-byte aligned volumes.
 * SURVEY.md §8 row f2. */
#define MVSGI_FLAG_VOL_CHANNELS_LAST 8u
/* Use the CTA-resident-table sweep kernel (sweep.cuh) even where the tile-walking one (sweep_tiled.cuh, the default
 * whenever the geometry is shared by all frames and the features are fp32) applies.  Both kernels produce
 * bit-identical volumes; the flag exists so that tests and bench.py can hold one against the other. */
#define MVSGI_FLAG_CTA_TABLE 16u
/* Stage the features into the zero-framed workspace copy even when they could be tapped where they are.  Features
 * are tapped in place (no staging pass, no workspace for them) when they are channels-last — feats_stride[2] == 1 —
 * with feat_chs a multiple of 16, the other strides multiples of 16 elements and a 128-byte aligned base; any other
 * layout (NCHW from cuDNN, odd channel counts) is transposed into the staged copy first.  Same volume bits either way
 * for finite features; the flag exists for tests and bench.py. */
#define MVSGI_FLAG_STAGED_FEATS 32u
/* The volume is written as bfloat16 (round to nearest even of the fp32 result; `vol` then points to 2-byte elements):
 * half the bytes the sweep writes and post_vol's Conv3d re-reads, and the layout + dtype cuDNN's bf16 tensor-core
 * convolution takes directly (spherical_sweep_avg.py:30-36,165, common_modules.py:82-115).
 * Stated-tolerance mode: |vol_bf16 - vol| <= 2^-8 * |vol| elementwise, zero pattern unchanged.  Needs
 * MVSGI_FLAG_VOL_CHANNELS_LAST, feat_chs % 4 == 0 and 8-byte aligned volumes; not available with valid_out, peer_vols
 * or MVSGI_FLAG_FEATS_BF16; the backward ignores it. */
#define MVSGI_FLAG_VOL_BF16 64u

/* ---- calibration ------------------------------------------------------- */
/* One camera the sweep projects into.  rot/trans are the INVERSE pose
 * (R^T, -R^T t) of T_cv_cam, i.e. what FTensor.inverse()
 * (dsta_mvs/mvs_utils/ftensor/ftensor.py:454-476) hands to
 * transform_3D_points_torch (dsta_mvs/support/dataset/torch_cuda_sweep.py:385-408).
 * All scalars are float32 exactly as the reference casts them.
 *   p[] by model:
 *   EQUIRECT       : lon0, lon1, lon_span_pixel, lat0, lat_range(=lat1-lat0)
 *                    (camera_models.py:640-652, :741-764)
 *   DOUBLE_SPHERE  : xi, alpha, one_minus_alpha, fx, fy, cx, cy, neg_w2, half_fov
 *                    (camera_models.py:492-529, :400-408)
 *   LINEAR_SPHERE  : fov_rad            (camera_models.py:1506-1548)
 */
typedef struct MvsgiCamera {
  int32_t model;
  int32_t width;  /* ss.W used by the normalisation */
  int32_t height; /* ss.H */
  int32_t reserved;
  float rot[9];   /* row-major 3x3 */
  float trans[3];
  float p[12];
} MvsgiCamera;

/* The cost-volume camera (always Equirectangular in the reference:
 * configs/dataset/cam_model_cv.yaml) + rig + candidate distances.
 * Rays follow Equirectangular.pixel_2_ray (camera_models.py:680-720) scaled by
 * RayMaker.make_rays_from_candidates (torch_cuda_sweep.py:186-215). */
typedef struct MvsgiRig {
  int32_t num_cams;
  int32_t num_cands;
  int32_t cv_width;
  int32_t cv_height;
  float cv_lon0;
  float cv_lon_span_pixel;
  float cv_lat0;
  float cv_lat_range;
  MvsgiCamera cams[MVSGI_MAX_CAMS];
  float dist[MVSGI_MAX_CANDS]; /* FixedDistLabelTable.dist()  dist_label_table.py:104-122 */
} MvsgiRig;

/* ---- sweep + aggregation ----------------------------------------------- */
/* Replaces SphericalSweepStdMasked.sweep(feats, grids, grid_masks, masks)
 * (spherical_sweep_avg.py:38-136) and SphericalSweep.sweep(feats, grids, masks)
 * (spherical_sweep.py:38-68), including their sampler bilinear_grid_sample
 * (dsta_mvs/model/backports/backports.py:11-86).
 * Strides are in ELEMENTS of the tensor's own dtype.  A batch stride of 0 on
 * masks / grids / grid_masks means "shared by all frames" and lets the kernel
 * evaluate the geometry once per tile instead of once per frame. */
typedef struct MvsgiSweepArgs {
  uint32_t struct_size; /* = sizeof(MvsgiSweepArgs) */
  int32_t aggregate;    /* MVSGI_AGG_*  */
  int32_t grid_mode;    /* MVSGI_GRID_* */
  uint32_t flags;
  int32_t batch, num_cams, feat_chs, num_cands;
  int32_t feat_h, feat_w; /* Hi, Wi of feats            */
  int32_t mask_h, mask_w; /* H, W of masks (std only)   */
  int32_t out_h, out_w;   /* Ho, Wo of the cost volume  */

  const float* feats;      /* [B, CAM, C, Hf, Wf] fp32 */
  int64_t feats_stride[5]; /* B, CAM, C, H, W */

  const void* masks;       /* [B, CAM, 1, Hm, Wm] fp32 0/1 (or u8); NULL for concat */
  int32_t masks_dtype;     /* MVSGI_MASK_* */
  int32_t reserved0;
  int64_t masks_stride[4]; /* B, CAM, H, W */

  const float* grids;      /* [B, CAM, D, Ho, Wo, 2]; last dim contiguous; NULL in PROJECT mode */
  int64_t grids_stride[5]; /* B, CAM, D, H, W */
  const uint8_t* grid_masks; /* [B, CAM, D, Ho, Wo, 1] bool; NULL in PROJECT mode / concat */
  int64_t grid_masks_stride[5];

  const MvsgiRig* rig;     /* HOST pointer; required in PROJECT mode */

  float* vol;              /* [B, Cout, D, Ho, Wo] contiguous; Cout = C (std) or CAM*C (concat) */
  uint8_t* valid_out;      /* optional [B, D, Ho, Wo]: 1 where >1 camera valid (std); may be NULL */

  void* workspace;         /* >= mvsgi_sweep_workspace_bytes(args), 256-byte aligned */
  int64_t workspace_bytes;

  /* Candidate-sharded sweep (one process per GPU, BASELINE.json config 5): this call sweeps num_cands candidates
   * that are candidates [cand_offset, cand_offset + num_cands) of a volume with vol_num_cands candidates, i.e.
   * vol is [B, Cout, vol_num_cands, Ho, Wo].  vol_num_cands == 0 means "not sharded" (= num_cands, offset 0).
   * Every voxel is additionally stored to the same position of each peer_vols[i] (i < num_peers): device pointers
   * to the other GPUs' volumes mapped into this process (mvsgi_p2p_open), so the all-gather of the candidate
   * slices happens inside the sweep, store by store over NVLink, instead of as a collective afterwards.
   * valid_out must be NULL when num_peers > 0. */
  int32_t cand_offset;
  int32_t vol_num_cands;
  int32_t num_peers;
  int32_t reserved1;
  float* peer_vols[MVSGI_MAX_PEERS];

  /* Cost-balanced work assignment (optional; static calibration and static masks).  By default every resident CTA
   * sweeps an equal share of 8-pixel octets dealt round-robin over the image, which balances the CTAs only
   * statistically: at G16V the per-CTA cost (cameras actually sampled per voxel) spreads +-15 % around its mean.
   * octet_order is a DEVICE table [order_runs][order_units] of octet indices (pixel >> 3, row-major output pixels),
   * -1 = empty slot, in which every octet of the image appears exactly once: run r sweeps the octets of row r.
   * Shape from mvsgi_sweep_balance_shape, costs from mvsgi_sweep_octet_cost; the assignment itself (longest-
   * processing-time-first) is host code (mvs_gi_b200/ops.py).  The volume is bit-identical with any valid table.
   * NULL = the default assignment.  Measured (profiles/r3j_balance.jsonl, profiles/r3i_cta_times.txt): the CTAs of
   * the G16V sweep finish between 332 and 405 us with the round-robin deal and within 345-353 us (one straggler at
   * 375) with the table; the drop-in modules and bench.py use it by default (balance_runs, --balance). */
  const int32_t* octet_order;
  int32_t order_runs, order_units;

  /* Optional DEVICE buffer of mvsgi_sweep_grid(args) x 4 uint64: per CTA of the sweep kernel its SM id, start time, end of
   * its first table build and end time (globaltimer ns).  With octet_order, CTA i sweeps run i, so the times show which
   * runs — and which SMs: the CTA-to-SM placement of a full one-wave grid repeats from launch to launch — finish late;
   * ops.calibrated_octet_order() re-deals the octets with them.  NULL = not recorded (no cost). */
  uint64_t* cta_times_out;

  /* Tap table kept across calls (optional; static calibration and static masks — the inference configuration).  The
   * sweep kernel turns the rig's geometry into a shared-memory table of (output offset, validity bits, per camera: tap
   * offset and the two bilinear fractions) records before it sweeps the frames; that table depends only on the rig, the
   * masks, the shapes / strides of feats and vol and the work assignment — not on the features — so it plays the role
   * of the grids the reference's dataset precomputes once per rig (multi_view_camera_model_dataset.py:519-557) and hands
   * to every forward (spherical_sweep_avg.py:138-167).  tap_table is a DEVICE buffer of mvsgi_sweep_tap_table_bytes(args)
   * bytes, 256-byte aligned:
   *   MVSGI_TAP_TABLE_BUILD  the kernel builds its tables as usual and also stores them to tap_table;
   *   MVSGI_TAP_TABLE_USE    the kernel copies each work item's table from tap_table into shared memory with one bulk
   *                          asynchronous copy (cp.async.bulk + mbarrier) instead of evaluating the geometry.
   * A table is valid for calls whose arguments equal the BUILD call's except feats / vol / workspace pointer VALUES
   * (same strides, masks contents, rig, octet_order contents, batch); the volume is bit-identical to MVSGI_TAP_TABLE_OFF.
   * Pays when the batch does not amortise the build (profiles/r4a_tap_table.jsonl): one G16V frame 35 -> 19 us, 64
   * candidates at 160x640 x 2 frames 0.71 -> 0.49 ms, 32 G16V frames 0.359 -> 0.341 ms.  Size: 4 * ((2 + 3 * cams) | 1) bytes per
   * (candidate, pixel) record (44 at 3 cameras) — 18 MB at G16V, 288 MB at high-D; ask mvsgi_sweep_tap_table_bytes. */
  void* tap_table;
  int64_t tap_table_bytes;
  int32_t tap_table_mode; /* MVSGI_TAP_TABLE_* */
  int32_t reserved2;
} MvsgiSweepArgs;

#define MVSGI_TAP_TABLE_OFF 0
#define MVSGI_TAP_TABLE_BUILD 1
#define MVSGI_TAP_TABLE_USE 2

/* Bytes of MvsgiSweepArgs.tap_table for this problem; 0 when the problem keeps no reusable table (per-frame grids or
 * per-frame masks: the table changes with every frame); negative = error.  args->tap_table* are ignored. */
int64_t mvsgi_sweep_tap_table_bytes(const MvsgiSweepArgs* args);

/* Scratch the sweep needs for this problem (channels-last staging of feats when
 * they are not already channels-last — see MVSGI_FLAG_STAGED_FEATS —, bit-packed masks,
 * the per-launch geometry table).  Negative = error. */
int64_t mvsgi_sweep_workspace_bytes(const MvsgiSweepArgs* args);

/* Enqueue staging (if needed) + the fused sweep/aggregate kernel. */
int mvsgi_sweep_f32(const MvsgiSweepArgs* args, void* cuda_stream);

/* Shape of the octet_order table this problem accepts: runs = resident CTAs of the sweep (fewer for tiny images; k times
 * as many when the tap table of 1/CTAs of the image does not fit one pass: CTA c then sweeps runs c, c + ctas, ...,
 * ctas = the CTAs the sweep is launched with once the table is attached), units = octets per run, num_octets = ceil(out_h * out_w / 8).  MVSGI_E_UNSUPPORTED when the
 * candidates are swept in chunks (more than 16) or the grids are per-frame: leave octet_order NULL then.
 * args->octet_order is ignored; args->vol and args->workspace may be NULL (nothing is written). */
int mvsgi_sweep_balance_shape(const MvsgiSweepArgs* args, int32_t* runs_out, int32_t* units_out, int32_t* num_octets_out,
                              int32_t* ctas_out);

/* CTAs the sweep kernel of this problem is launched with (rows of cta_times_out); negative = error. */
int mvsgi_sweep_grid(const MvsgiSweepArgs* args);

/* Relative cost of every octet for frame 0 of this problem, written (not accumulated) to cost_out[num_octets] (DEVICE):
 * per candidate 4 + 10 * (cameras the octet's warp loads: those that at least one of its 8 voxels samples, a voxel
 * sampling its valid cameras when more than one is valid; all cameras for the concat builder).  std builder: masks must
 * be MVSGI_MASK_PACKED (mvsgi_pack_masks).  Two launches (zero fill + cost kernel) on the given stream; args->vol and
 * args->workspace may be NULL. */
int mvsgi_sweep_octet_cost(const MvsgiSweepArgs* args, int32_t* cost_out, void* cuda_stream);

/* Number of kernels the last-described problem launches (for gpu_launches
 * accounting in bench.py): staging transposes + mask pack + sweep. */
int mvsgi_sweep_launch_count(const MvsgiSweepArgs* args);

/* ---- grid evaluation only (debug / oracle hand-off) --------------------- */
/* Replaces the CPU precompute MultiViewCameraModelDataset.create_grids_from_rays
 * (dsta_mvs/support/dataset/multi_view_camera_model_dataset.py:530-557,474-518):
 * writes the kernel's own grids so that the oracle can consume exactly what the
 * fused kernel samples with. */
typedef struct MvsgiProjectArgs {
  uint32_t struct_size;
  int32_t reserved;
  const MvsgiRig* rig;  /* HOST pointer */
  float* grids;         /* [CAM, D, Ho, Wo, 2] contiguous */
  uint8_t* grid_masks;  /* [CAM, D, Ho, Wo] contiguous (1 = valid) */
} MvsgiProjectArgs;
int mvsgi_project_only_f32(const MvsgiProjectArgs* args, void* cuda_stream);

/* ---- mask packing -------------------------------------------------------- */
/* Packs masks (>0 -> 1) into the padded bit layout the sweep samples.  Per image: (Hm+2) row PAIRS x pitch_words x 2
 * uint32; pair Y holds padded rows Y (even word of each couple) and Y+1 (odd word), so the two mask rows of a bilinear
 * sample are one 8-byte load; pixel (y, x) is bit (x & 31) of word 1 + (x >> 5) of padded row y + 1 (row 0, row Hm+1,
 * word 0 and the words after the data are zero guards).  bits_out holds n_images * (Hm+2) * pitch_words * 2 words and
 * must be 8-byte aligned.
 * flags_out[0] gets bit0 set if any mask value is negative or NaN (then the
 * "bilinear(mask) > 0" test of spherical_sweep_avg.py:92-96 is not equivalent
 * to the packed test and the caller should reject the input). */
int32_t mvsgi_mask_pitch_words(int32_t mask_w);
int mvsgi_pack_masks(const void* masks, int32_t masks_dtype, int32_t n_images, int64_t image_stride_outer,
                     int32_t n_inner, int64_t image_stride_inner, int64_t stride_h, int64_t stride_w,
                     int32_t mask_h, int32_t mask_w, uint32_t* bits_out, int32_t* flags_out, void* cuda_stream);

/* ---- soft-argmin distance regression ------------------------------------ */
/* Replaces DistanceRegressorWithFixedCandidates.forward
 * (dsta_mvs/model/distance_regressor/distance_regressor.py:51-80):
 * costs[:,0] -> optional bilinear upsample (F.interpolate, align_corners=False)
 * -> softmax over D -> sum_d p_d * inv_dist_idx_d. */
typedef struct MvsgiSoftArgminArgs {
  uint32_t struct_size;
  uint32_t flags;
  int32_t batch, num_cands;
  int32_t in_h, in_w;
  int32_t out_h, out_w;  /* == in_h,in_w when no interpolation */
  float src_scale_h;     /* source-index scale of F.interpolate = 1/scale_factor */
  float src_scale_w;
  const float* costs;      /* element (b, d, y, x) of costs[:,0] */
  int64_t costs_stride[4]; /* B, D, H, W */
  const float* inv_dist_idx; /* device [D] = bf / dist  (distance_regressor.py:29-32) */
  float* inv_dist;         /* [B, 1, out_h, out_w] contiguous */
  float* norm_costs;       /* [B, D, out_h, out_w] contiguous, or NULL (inference) */
  int32_t* argmax;         /* [B, out_h, out_w] first-max candidate index, or NULL */
} MvsgiSoftArgminArgs;
int mvsgi_softargmin_f32(const MvsgiSoftArgminArgs* args, void* cuda_stream);

/* ---- backward passes (training drop-in) ------------------------------------ */
/* What torch.autograd derives from the reference's forward code when
 * training_step differentiates through cv_builder and dist_regressor
 * (dsta_mvs/model/mvs_model/spherical_sweep_stereo.py:170-202,
 *  spherical_sweep_stereo_2024.py:97-143).  Only `feats` / `costs` receive
 * gradients; grids, masks and the candidate list are data, as in the reference
 * (dataset tensors, no requires_grad). */

/* d sum(grad_vol * vol) / d feats of mvsgi_sweep_f32 for the problem `fwd` describes.
 * Replaces the autograd graph of SphericalSweepStdMasked.sweep (spherical_sweep_avg.py:38-136) /
 * SphericalSweep.sweep (spherical_sweep.py:38-68) incl. the sampler's gather backward (backports.py:76-86).
 * Of `fwd`, the fields vol, valid_out, peer_vols, num_peers, flags, workspace and workspace_bytes are ignored
 * (cand_offset / vol_num_cands are honoured: grad_vol is then the full [B, Cout, vol_num_cands, Ho, Wo] tensor).
 * The backward runs the sweep kernel's own work decomposition with a scatter (vector atomics) in place of the store.
 * Dense channels-last feats (whole 16-channel groups, the layout the forward taps in place) with grad_feats of the
 * same strides: grad_feats is cleared and accumulated into directly (memset + one kernel).  Otherwise feats are staged
 * as in the forward, the gradient is accumulated in a staged image and un-staged into grad_feats (any strides).
 * grad_feats is overwritten, not accumulated.  Summation order is not deterministic (fp32 atomics). */
typedef struct MvsgiSweepBwdArgs {
  uint32_t struct_size; /* = sizeof(MvsgiSweepBwdArgs) */
  uint32_t flags;       /* 0 */
  const MvsgiSweepArgs* fwd; /* HOST pointer */
  const float* grad_vol;     /* [B, Cout, Dvol, Ho, Wo] contiguous */
  float* grad_feats;         /* [B, CAM, C, Hf, Wf] */
  int64_t grad_feats_stride[5]; /* B, CAM, C, H, W */
  void* workspace;           /* >= mvsgi_sweep_backward_workspace_bytes(args), 256-byte aligned */
  int64_t workspace_bytes;
} MvsgiSweepBwdArgs;
int64_t mvsgi_sweep_backward_workspace_bytes(const MvsgiSweepBwdArgs* args);
int mvsgi_sweep_backward_f32(const MvsgiSweepBwdArgs* args, void* cuda_stream);

/* Backward of mvsgi_softargmin_f32 wrt costs[:, 0] (distance_regressor.py:51-80: F.interpolate, F.softmax,
 * sum(norm_costs * inv_dist_idx)).  Of `fwd`, inv_dist / norm_costs / argmax are ignored.  Either upstream gradient
 * may be NULL (that output was not used).  grad_costs is overwritten; element (b, d, y, x) at the given strides. */
typedef struct MvsgiSoftArgminBwdArgs {
  uint32_t struct_size;
  uint32_t flags;
  const MvsgiSoftArgminArgs* fwd; /* HOST pointer */
  const float* grad_inv_dist;     /* [B, 1, out_h, out_w] contiguous, or NULL */
  const float* grad_norm_costs;   /* [B, D, out_h, out_w] contiguous, or NULL */
  float* grad_costs;
  int64_t grad_costs_stride[4];   /* B, D, H, W */
} MvsgiSoftArgminBwdArgs;
int mvsgi_softargmin_backward_f32(const MvsgiSoftArgminBwdArgs* args, void* cuda_stream);

/* ---- fisheye -> surrogate-ERP image resampling (the step before the feature extractor) -------------------------- */
/* Replaces CameraModelRotation (dsta_mvs/image_sampler/camera_model_sampler.py:12-120) as
 * InferencePytorch.preprocess_imgs drives it (api/inference_pytorch.py:61-74): every pixel of the target
 * (Equirectangular surrogate) image is a ray, rotated into the raw camera by R_raw_fisheye, projected with the raw
 * camera model and bilinearly sampled (F.grid_sample, zeros padding, align_corners=False); pixels whose ray the raw
 * camera cannot see get invalid_value.  The geometry is evaluated in-kernel from `rig`:
 *   cv_width/cv_height/cv_lon0/cv_lon_span_pixel/cv_lat0/cv_lat_range = the TARGET Equirectangular model,
 *   num_cams = 1, cams[0] = the RAW camera model with rot = R_raw_fisheye (row-major) and trans = 0,
 *   num_cands = 1, dist[0] = 1.
 * All `batch` images share that geometry (same camera model and orientation, e.g. consecutive frames). */
#define MVSGI_RESAMPLE_IMG_U8 1u /* img holds uint8 pixels; each tap is scaled to [0,1] on load exactly as the CUDA
                                  * `t_imgs.float() / 255.0` of preprocess_imgs (api/inference_pytorch.py:56-57)
                                  * does it: v * float32(1/255) */
#define MVSGI_RESAMPLE_NEAREST 2u /* interpolation='nearest' of the reference sampler (camera_model_sampler.py:98-120):
                                   * F.grid_sample(mode='nearest', zeros padding, align_corners=True —
                                   * planar_as_base.py:127-133): the pixel at nearbyint(((g + 1) / 2) * (size - 1)),
                                   * 0 outside the image.  The 'blended' mode (:122-163) is three calls and one
                                   * elementwise combination on the caller's side (mvs_gi_b200/resample.py) */
typedef struct MvsgiResampleArgs {
  uint32_t struct_size;
  uint32_t flags;          /* any of MVSGI_RESAMPLE_IMG_U8, MVSGI_RESAMPLE_NEAREST */
  int32_t batch, channels;
  int32_t in_h, in_w;      /* raw image size; must equal rig->cams[0].height/width */
  const void* img;         /* float32 (or uint8) element (n, c, y, x) at the given ELEMENT strides: an NCHW tensor, or the
                            * permuted view of the HWC camera buffers (api/inference_pytorch.py:55) without a copy */
  int64_t img_stride[4];   /* N, C, H, W */
  const MvsgiRig* rig;     /* HOST pointer */
  float invalid_value;     /* default_invalid_value of the reference sampler (0) */
  int32_t reserved;
  float* out;              /* [batch, channels, cv_height, cv_width] contiguous */
  uint8_t* valid_out;      /* optional [cv_height, cv_width]: 1 where the raw camera sees the ray */
} MvsgiResampleArgs;
int mvsgi_resample_f32(const MvsgiResampleArgs* args, void* cuda_stream);

/* ---- peer-mapped volumes for the candidate-sharded sweep ------------------- */
/* One process per GPU.  Each rank allocates its copy of the volume with mvsgi_p2p_alloc (cudaMalloc on the current
 * device), exports a 64-byte handle (cudaIpcGetMemHandle) that the host code sends to the other ranks
 * (torch.distributed all_gather), and opens the other ranks' handles (cudaIpcOpenMemHandle, peer access over
 * NVLink) to obtain the pointers it passes as MvsgiSweepArgs.peer_vols.  Returns 0 or a cudaError_t. */
#define MVSGI_P2P_HANDLE_BYTES 64
int mvsgi_p2p_alloc(int64_t bytes, void** ptr_out);
int mvsgi_p2p_free(void* ptr);
int mvsgi_p2p_export(const void* ptr, uint8_t handle_out[MVSGI_P2P_HANDLE_BYTES]);
int mvsgi_p2p_open(const uint8_t handle[MVSGI_P2P_HANDLE_BYTES], void** ptr_out);
int mvsgi_p2p_close(void* ptr);

/* ---- misc ---------------------------------------------------------------- */
const char* mvsgi_strerror(int code);
int mvsgi_abi_version(void);
/* Kernel launches issued through this library by the calling process so far. */
int64_t mvsgi_launch_counter(void);

#ifdef __cplusplus
}
#endif
#endif /* MVSGI_B200_H_ */
