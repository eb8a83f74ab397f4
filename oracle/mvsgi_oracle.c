/*
 * mvsgi_oracle.c — CPU restatement of the mvs_gi sweep / cost-volume /
 * soft-argmin path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library, and only as the checker or
 * the timed CPU baseline.  Nothing under mvs_gi_b200/ imports it.
 *
 * Parity status: the reference has no golden vectors or unit tests for this
 * path, so this restatement is pinned against outputs of the
 * reference's own modules, run from /root/reference by oracle/gen_golden.py
 * and committed under tests/golden/ (tests/test_oracle_golden.py).
 *
 * Everything is fp32 evaluated in the reference's operation order; build with
 * -ffp-contract=off so the compiler does not fuse multiply-adds the reference
 * (eager PyTorch, one rounding per op) does not have.  The one place PyTorch
 * itself fuses (the bilinear upsample of F.interpolate) uses fmaf explicitly.
 *
 * Paths below are relative to the reference checkout.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/mvsgi_b200.h" /* MvsgiRig / MvsgiCamera POD only */

#if defined(_OPENMP)
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* dsta_mvs/model/backports/backports.py:11-86  bilinear_grid_sample           */
/* (zero padding, align_corners=False).  One sample point, all channels.      */
/* im is [C, H, W] with element strides (sc, sy, sx).                         */
typedef struct {
  long x0c, x1c, y0c, y1c; /* indices into the zero-padded image, clamped (:65-72) */
  float wa, wb, wc, wd;    /* :52-55 */
} orc_taps;

static inline void orc_make_taps(float gx, float gy, int H, int W, orc_taps* t) {
  /* :41-42 (align_corners=False):  x = ((x + 1) * w - 1) / 2 */
  float x = ((gx + 1.0f) * (float)W - 1.0f) / 2.0f;
  float y = ((gy + 1.0f) * (float)H - 1.0f) / 2.0f;
  /* :47-50 */
  float fx0 = floorf(x), fy0 = floorf(y);
  long x0 = (long)fx0, y0 = (long)fy0;
  long x1 = x0 + 1, y1 = y0 + 1;
  /* :52-55 — long - float promotes the long to float32 */
  float dx1 = (float)x1 - x, dx0 = x - (float)x0;
  float dy1 = (float)y1 - y, dy0 = y - (float)y0;
  t->wa = dx1 * dy1;
  t->wb = dx1 * dy0;
  t->wc = dx0 * dy1;
  t->wd = dx0 * dy0;
  /* :58-72 pad by one, shift, clamp into the padded extent */
  long pw = W + 2, ph = H + 2;
  x0 += 1; x1 += 1; y0 += 1; y1 += 1;
  if (x0 < 0) x0 = 0; if (x0 > pw - 1) x0 = pw - 1;
  if (x1 < 0) x1 = 0; if (x1 > pw - 1) x1 = pw - 1;
  if (y0 < 0) y0 = 0; if (y0 > ph - 1) y0 = ph - 1;
  if (y1 < 0) y1 = 0; if (y1 > ph - 1) y1 = ph - 1;
  t->x0c = x0; t->x1c = x1; t->y0c = y0; t->y1c = y1;
}

/* value of the zero-padded image at padded coords (:58, :76-84) */
static inline float orc_padded(const float* im, int H, int W, int64_t sy, int64_t sx, long yp, long xp) {
  if (yp < 1 || yp > H || xp < 1 || xp > W) return 0.0f;
  return im[(yp - 1) * sy + (xp - 1) * sx];
}

static inline float orc_sample1(const float* im, int H, int W, int64_t sy, int64_t sx, const orc_taps* t) {
  float Ia = orc_padded(im, H, W, sy, sx, t->y0c, t->x0c);
  float Ib = orc_padded(im, H, W, sy, sx, t->y1c, t->x0c);
  float Ic = orc_padded(im, H, W, sy, sx, t->y0c, t->x1c);
  float Id = orc_padded(im, H, W, sy, sx, t->y1c, t->x1c);
  /* :86  Ia * wa + Ib * wb + Ic * wc + Id * wd, left to right */
  return ((Ia * t->wa + Ib * t->wb) + Ic * t->wc) + Id * t->wd;
}

/* Stand-alone sampler: im [N,C,H,W] contiguous, grid [N,Hg,Wg,2], out [N,C,Hg,Wg]. */
void orc_bilinear_grid_sample(const float* im, int N, int C, int H, int W, const float* grid, int Hg, int Wg,
                              float* out) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < N; ++n)
    for (int i = 0; i < Hg * Wg; ++i) {
      orc_taps t;
      const float* g = grid + ((int64_t)n * Hg * Wg + i) * 2;
      orc_make_taps(g[0], g[1], H, W, &t);
      for (int c = 0; c < C; ++c)
        out[((int64_t)n * C + c) * Hg * Wg + i] = orc_sample1(im + ((int64_t)n * C + c) * H * W, H, W, W, 1, &t);
    }
}

/* ------------------------------------------------------------------------- */
/* dsta_mvs/model/cost_volume_builder/spherical_sweep_avg.py:38-136           */
/* SphericalSweepStdMasked.sweep.  All tensors contiguous fp32 / uint8:        */
/*   feats [B,CAM,C,Hf,Wf]  masks [B,CAM,1,Hm,Wm]  grids [B,CAM,D,Ho,Wo,2]     */
/*   grid_masks [B,CAM,D,Ho,Wo]  ->  vol [B,C,D,Ho,Wo], valid [B,D,Ho,Wo]      */
/* A batch stride of 0 may be passed for masks/grids/grid_masks (shared).     */
void orc_sweep_std(const float* feats, const float* masks, const float* grids, const uint8_t* grid_masks, int B,
                   int CAM, int C, int D, int Hf, int Wf, int Hm, int Wm, int Ho, int Wo, int64_t masks_bstride,
                   int64_t grids_bstride, int64_t gm_bstride, float* vol, uint8_t* valid_out) {
  const int64_t HW = (int64_t)Ho * Wo;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < D; ++i) { /* the reference's `for i in range(X)` :59 */
      float* f = (float*)malloc(sizeof(float) * (size_t)CAM * C);
      for (int64_t px = 0; px < HW; ++px) {
        int m[MVSGI_MAX_CAMS * 4];
        float summed = 0.0f; /* :106 torch.sum(sampled_mask.to(float), dim=1) */
        for (int cam = 0; cam < CAM; ++cam) {
          const float* g = grids + b * grids_bstride + (((int64_t)cam * D + i) * HW + px) * 2;
          orc_taps tf, tm;
          /* :73-77 features sampled at feature resolution */
          orc_make_taps(g[0], g[1], Hf, Wf, &tf);
          for (int c = 0; c < C; ++c)
            f[cam * C + c] = orc_sample1(feats + (((int64_t)b * CAM + cam) * C + c) * Hf * Wf, Hf, Wf, Wf, 1, &tf);
          /* :92-96 masks sampled with the same normalised grid at mask resolution, > 0 */
          orc_make_taps(g[0], g[1], Hm, Wm, &tm);
          float sm = orc_sample1(masks + b * masks_bstride + (int64_t)cam * Hm * Wm, Hm, Wm, Wm, 1, &tm);
          int mk = sm > 0.0f;
          /* :60-62,:102 AND with the grid mask */
          mk = mk && (grid_masks[b * gm_bstride + ((int64_t)cam * D + i) * HW + px] != 0);
          m[cam] = mk;
          summed += mk ? 1.0f : 0.0f;
        }
        int valid = summed > 1.0f;               /* :108 */
        float cnt = valid ? summed : 1.0f;       /* :111 */
        if (valid_out) valid_out[((int64_t)b * D + i) * HW + px] = (uint8_t)valid;
        for (int c = 0; c < C; ++c) {
          /* :114 avg = sum(sampled_feat * mask) / cam_count   (sequential over cams) */
          float s = 0.0f;
          for (int cam = 0; cam < CAM; ++cam) {
            float term = f[cam * C + c] * (m[cam] ? 1.0f : 0.0f);
            s = (cam == 0) ? term : s + term;
          }
          float avg = s / cnt;
          /* :119-122 sampled_feat = where(mask, feat, avg); std2 = sum((f-avg)**2)/cnt */
          float v = 0.0f;
          for (int cam = 0; cam < CAM; ++cam) {
            float fv = m[cam] ? f[cam * C + c] : avg;
            float d = fv - avg;
            float sq = d * d;
            v = (cam == 0) ? sq : v + sq;
          }
          v = v / cnt;
          /* :125 */
          vol[(((int64_t)b * C + c) * D + i) * HW + px] = valid ? v : 0.0f;
        }
      }
      free(f);
    }
}

/* dsta_mvs/model/cost_volume_builder/spherical_sweep.py:38-68  SphericalSweep.sweep */
/* vol[b, cam*C + c, d, h, w] = bilinear(feats[b,cam,c], grids[b,cam,d,h,w])  (:52-67) */
void orc_sweep_concat(const float* feats, const float* grids, int B, int CAM, int C, int D, int Hf, int Wf, int Ho,
                      int Wo, int64_t grids_bstride, float* vol) {
  const int64_t HW = (int64_t)Ho * Wo;
#pragma omp parallel for collapse(3) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int cam = 0; cam < CAM; ++cam)
      for (int i = 0; i < D; ++i)
        for (int64_t px = 0; px < HW; ++px) {
          const float* g = grids + b * grids_bstride + (((int64_t)cam * D + i) * HW + px) * 2;
          orc_taps t;
          orc_make_taps(g[0], g[1], Hf, Wf, &t);
          for (int c = 0; c < C; ++c)
            vol[((((int64_t)b * CAM + cam) * C + c) * D + i) * HW + px] =
                orc_sample1(feats + (((int64_t)b * CAM + cam) * C + c) * Hf * Wf, Hf, Wf, Wf, 1, &t);
        }
}

/* ------------------------------------------------------------------------- */
/* dsta_mvs/model/distance_regressor/distance_regressor.py:51-80              */
/* The upsample is PyTorch's F.interpolate(mode='bilinear', align_corners=    */
/* False, scale_factor=s) — third-party arithmetic (torch, not pinned by the  */
/* reference beyond NGC 22.11; restated from ATen UpSampleKernel.cpp):        */
/*   src = fma(scale, dst+0.5, -0.5) clamped >= 0, i0=(int)src, i1=min(i0+1,n-1), */
/*   l1 = src-i0, l0 = 1-l1,                                                  */
/*   out = fma(ly0, fma(lx0,v00, lx1*v01), ly1*fma(lx0,v10, lx1*v11))         */
/* The fma placement was determined empirically against torch 2.11 CPU        */
/* (bit-exact on 2x16x80x320 random input; see oracle/gen_golden.py).         */
static inline void orc_src_index(int dst, float scale, int n_in, int* i0, int* i1, float* l0, float* l1) {
  float src = fmaf(scale, (float)dst + 0.5f, -0.5f); /* torch contracts this (bit-exact for s=2,3) */
  if (src < 0.0f) src = 0.0f;
  int a = (int)src;
  if (a > n_in - 1) a = n_in - 1;
  int b = a + 1 < n_in ? a + 1 : n_in - 1;
  float lam = src - (float)a;
  if (lam < 0.0f) lam = 0.0f;
  if (lam > 1.0f) lam = 1.0f;
  *i0 = a; *i1 = b; *l1 = lam; *l0 = 1.0f - lam;
}

static inline float orc_upsampled(const float* plane, int64_t sy, int64_t sx, int y0, int y1, float ly0, float ly1,
                                  int x0, int x1, float lx0, float lx1) {
  float v00 = plane[y0 * sy + x0 * sx], v01 = plane[y0 * sy + x1 * sx];
  float v10 = plane[y1 * sy + x0 * sx], v11 = plane[y1 * sy + x1 * sx];
  float top = fmaf(lx0, v00, lx1 * v01);
  float bot = fmaf(lx0, v10, lx1 * v11);
  return fmaf(ly0, top, ly1 * bot);
}

/* costs: element (b,d,y,x) at b*sb + d*sd + y*sy + x*sx  (channel 0 of [B,1,D,H,W], :52)
 * interp != 0: out_h x out_w differs from in_h x in_w and the upsample runs (:54-60)
 * inv_dist [B,1,Ho,Wo]; norm_costs [B,D,Ho,Wo] (may be NULL); argmax [B,Ho,Wo] (may be NULL) */
void orc_softargmin(const float* costs, int64_t sb, int64_t sd, int64_t sy, int64_t sx, int B, int D, int Hin, int Win,
                    int Hout, int Wout, int interp, float scale_h, float scale_w, const float* inv_dist_idx,
                    float* inv_dist, float* norm_costs, int32_t* argmax) {
  const int64_t HWo = (int64_t)Hout * Wout;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int oy = 0; oy < Hout; ++oy) {
      float* v = (float*)malloc(sizeof(float) * (size_t)D);
      int y0 = oy, y1 = oy;
      float ly0 = 1.0f, ly1 = 0.0f;
      if (interp) orc_src_index(oy, scale_h, Hin, &y0, &y1, &ly0, &ly1);
      for (int ox = 0; ox < Wout; ++ox) {
        int x0 = ox, x1 = ox;
        float lx0 = 1.0f, lx1 = 0.0f;
        if (interp) orc_src_index(ox, scale_w, Win, &x0, &x1, &lx0, &lx1);
        float mx = -INFINITY;
        int am = 0;
        for (int d = 0; d < D; ++d) {
          const float* plane = costs + b * sb + d * sd;
          v[d] = interp ? orc_upsampled(plane, sy, sx, y0, y1, ly0, ly1, x0, x1, lx0, lx1) : plane[oy * sy + ox * sx];
          if (v[d] > mx) { mx = v[d]; am = d; }
        }
        /* :62 softmax over D: exp(x - max) / sum  (ATen SoftMaxKernel.cpp vec_softmax) */
        float sum = 0.0f;
        for (int d = 0; d < D; ++d) { v[d] = expf(v[d] - mx); sum += v[d]; }
        /* :74-78 sum_d norm_costs * inv_dist_idx, sequential over d */
        float acc = 0.0f;
        for (int d = 0; d < D; ++d) {
          float p = v[d] / sum;
          if (norm_costs) norm_costs[((int64_t)b * D + d) * HWo + (int64_t)oy * Wout + ox] = p;
          float term = p * inv_dist_idx[d];
          acc = (d == 0) ? term : acc + term;
        }
        inv_dist[(int64_t)b * HWo + (int64_t)oy * Wout + ox] = acc;
        if (argmax) argmax[(int64_t)b * HWo + (int64_t)oy * Wout + ox] = am;
      }
      free(v);
    }
}

/* ------------------------------------------------------------------------- */
/* Backward passes: what torch.autograd derives from the  */
/* reference's forward code when training_step differentiates through it       */
/* (dsta_mvs/model/mvs_model/spherical_sweep_stereo.py:170-202).  Restated op   */
/* by op from the forward graph; sums run in a fixed order here, whereas        */
/* autograd's scatter (index_add of torch.gather's backward) and the CUDA       */
/* kernel's atomics do not, so parity is 1e-5 * max|ref|, not bit-exact.        */

/* adjoint of orc_sample1 for one tap set: adds g * w to the four taps that lie inside the image; taps in the
 * zero padding receive their share too in autograd, but F.pad's backward drops it (backports.py:58). */
static inline void orc_scatter1(float* gim, int H, int W, int64_t sy, int64_t sx, const orc_taps* t, float g) {
  if (t->y0c >= 1 && t->y0c <= H && t->x0c >= 1 && t->x0c <= W) gim[(t->y0c - 1) * sy + (t->x0c - 1) * sx] += g * t->wa;
  if (t->y1c >= 1 && t->y1c <= H && t->x0c >= 1 && t->x0c <= W) gim[(t->y1c - 1) * sy + (t->x0c - 1) * sx] += g * t->wb;
  if (t->y0c >= 1 && t->y0c <= H && t->x1c >= 1 && t->x1c <= W) gim[(t->y0c - 1) * sy + (t->x1c - 1) * sx] += g * t->wc;
  if (t->y1c >= 1 && t->y1c <= H && t->x1c >= 1 && t->x1c <= W) gim[(t->y1c - 1) * sy + (t->x1c - 1) * sx] += g * t->wd;
}

/* d sum(grad_vol * vol) / d feats for SphericalSweepStdMasked.sweep (spherical_sweep_avg.py:38-136).
 * grad_vol [B,C,D,Ho,Wo]; grad_feats [B,CAM,C,Hf,Wf] is overwritten. */
void orc_sweep_std_backward(const float* feats, const float* masks, const float* grids, const uint8_t* grid_masks,
                            int B, int CAM, int C, int D, int Hf, int Wf, int Hm, int Wm, int Ho, int Wo,
                            int64_t masks_bstride, int64_t grids_bstride, int64_t gm_bstride, const float* grad_vol,
                            float* grad_feats) {
  const int64_t HW = (int64_t)Ho * Wo;
  memset(grad_feats, 0, sizeof(float) * (size_t)B * CAM * C * Hf * Wf);
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int kc = 0; kc < CAM; ++kc) { /* one thread owns grad_feats[b, kc]: no races */
      float* f = (float*)malloc(sizeof(float) * (size_t)CAM * C);
      for (int i = 0; i < D; ++i)
        for (int64_t px = 0; px < HW; ++px) {
          int m[MVSGI_MAX_CAMS * 4];
          orc_taps tk;
          memset(&tk, 0, sizeof(tk));
          float summed = 0.0f;
          for (int cam = 0; cam < CAM; ++cam) {
            const float* g = grids + b * grids_bstride + (((int64_t)cam * D + i) * HW + px) * 2;
            orc_taps tf, tm;
            orc_make_taps(g[0], g[1], Hf, Wf, &tf);
            if (cam == kc) tk = tf;
            for (int c = 0; c < C; ++c)
              f[cam * C + c] = orc_sample1(feats + (((int64_t)b * CAM + cam) * C + c) * Hf * Wf, Hf, Wf, Wf, 1, &tf);
            orc_make_taps(g[0], g[1], Hm, Wm, &tm);
            float sm = orc_sample1(masks + b * masks_bstride + (int64_t)cam * Hm * Wm, Hm, Wm, Wm, 1, &tm);
            int mk = (sm > 0.0f) && (grid_masks[b * gm_bstride + ((int64_t)cam * D + i) * HW + px] != 0);
            m[cam] = mk;
            summed += mk ? 1.0f : 0.0f;
          }
          if (!(summed > 1.0f) || !m[kc]) continue; /* :125 where(valid, ., 0); :119 where(mask, feat, avg) */
          float cnt = summed;
          for (int c = 0; c < C; ++c) {
            float s = 0.0f;
            for (int cam = 0; cam < CAM; ++cam) s += f[cam * C + c] * (m[cam] ? 1.0f : 0.0f);
            float avg = s / cnt;
            float gv = grad_vol[(((int64_t)b * C + c) * D + i) * HW + px] / cnt; /* :122 (.)/cam_count */
            /* :122 d/d(f'_j - avg) = 2 (f'_j - avg) gv ; through avg (:114) every camera gets -(sum_j ...)/cnt */
            float gavg = 0.0f;
            for (int cam = 0; cam < CAM; ++cam)
              if (m[cam]) gavg -= 2.0f * (f[cam * C + c] - avg) * gv;
            float gk = 2.0f * (f[kc * C + c] - avg) * gv + gavg / cnt;
            orc_scatter1(grad_feats + (((int64_t)b * CAM + kc) * C + c) * Hf * Wf, Hf, Wf, Wf, 1, &tk, gk);
          }
        }
      free(f);
    }
}

/* adjoint of SphericalSweep.sweep (spherical_sweep.py:38-68): grad_vol [B,CAM*C,D,Ho,Wo] -> grad_feats */
void orc_sweep_concat_backward(const float* grids, int B, int CAM, int C, int D, int Hf, int Wf, int Ho, int Wo,
                               int64_t grids_bstride, const float* grad_vol, float* grad_feats) {
  const int64_t HW = (int64_t)Ho * Wo;
  memset(grad_feats, 0, sizeof(float) * (size_t)B * CAM * C * Hf * Wf);
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int cam = 0; cam < CAM; ++cam)
      for (int i = 0; i < D; ++i)
        for (int64_t px = 0; px < HW; ++px) {
          const float* g = grids + b * grids_bstride + (((int64_t)cam * D + i) * HW + px) * 2;
          orc_taps t;
          orc_make_taps(g[0], g[1], Hf, Wf, &t);
          for (int c = 0; c < C; ++c)
            orc_scatter1(grad_feats + (((int64_t)b * CAM + cam) * C + c) * Hf * Wf, Hf, Wf, Wf, 1, &t,
                         grad_vol[((((int64_t)b * CAM + cam) * C + c) * D + i) * HW + px]);
        }
}

/* Backward of orc_softargmin (distance_regressor.py:51-80): softmax backward (p * (G - sum p G)) followed by the
 * adjoint of the bilinear upsample.  g_inv [B,1,Ho,Wo] and g_nc [B,D,Ho,Wo] may each be NULL (no gradient through that
 * output).  grad_costs [B,D,Hin,Win] (channel 0 of the costs tensor) is overwritten. */
void orc_softargmin_backward(const float* costs, int64_t sb, int64_t sd, int64_t sy, int64_t sx, int B, int D, int Hin,
                             int Win, int Hout, int Wout, int interp, float scale_h, float scale_w,
                             const float* inv_dist_idx, const float* g_inv, const float* g_nc, float* grad_costs) {
  const int64_t HWo = (int64_t)Hout * Wout, HWi = (int64_t)Hin * Win;
  memset(grad_costs, 0, sizeof(float) * (size_t)B * D * HWi);
#pragma omp parallel for schedule(static)
  for (int b = 0; b < B; ++b) {
    float* v = (float*)malloc(sizeof(float) * (size_t)D);
    for (int oy = 0; oy < Hout; ++oy) {
      int y0 = oy, y1 = oy;
      float ly0 = 1.0f, ly1 = 0.0f;
      if (interp) orc_src_index(oy, scale_h, Hin, &y0, &y1, &ly0, &ly1);
      for (int ox = 0; ox < Wout; ++ox) {
        int x0 = ox, x1 = ox;
        float lx0 = 1.0f, lx1 = 0.0f;
        if (interp) orc_src_index(ox, scale_w, Win, &x0, &x1, &lx0, &lx1);
        float mx = -INFINITY;
        for (int d = 0; d < D; ++d) {
          const float* plane = costs + b * sb + d * sd;
          v[d] = interp ? orc_upsampled(plane, sy, sx, y0, y1, ly0, ly1, x0, x1, lx0, lx1) : plane[oy * sy + ox * sx];
          if (v[d] > mx) mx = v[d];
        }
        float sum = 0.0f;
        for (int d = 0; d < D; ++d) { v[d] = expf(v[d] - mx); sum += v[d]; }
        const int64_t o = (int64_t)oy * Wout + ox;
        const float gi = g_inv ? g_inv[(int64_t)b * HWo + o] : 0.0f;
        float dot = 0.0f;
        for (int d = 0; d < D; ++d) {
          v[d] = v[d] / sum;
          float G = gi * inv_dist_idx[d] + (g_nc ? g_nc[((int64_t)b * D + d) * HWo + o] : 0.0f);
          dot += v[d] * G;
        }
        for (int d = 0; d < D; ++d) {
          float G = gi * inv_dist_idx[d] + (g_nc ? g_nc[((int64_t)b * D + d) * HWo + o] : 0.0f);
          float gv = v[d] * (G - dot);
          float* gp = grad_costs + ((int64_t)b * D + d) * HWi;
          if (interp) {
            gp[(int64_t)y0 * Win + x0] += ly0 * lx0 * gv;
            gp[(int64_t)y0 * Win + x1] += ly0 * lx1 * gv;
            gp[(int64_t)y1 * Win + x0] += ly1 * lx0 * gv;
            gp[(int64_t)y1 * Win + x1] += ly1 * lx1 * gv;
          } else {
            gp[(int64_t)oy * Win + ox] += gv;
          }
        }
      }
    }
    free(v);
  }
}

/* ------------------------------------------------------------------------- */
/* Grid precompute the fused kernel absorbs.           */
/* A.1 Equirectangular.pixel_2_ray  dsta_mvs/mvs_utils/camera_models.py:680-720 */
static inline void orc_cv_ray(const MvsgiRig* rig, int h, int w, float ray[3]) {
  float px = (float)w + 0.5f, py = (float)h + 0.5f;
  float lon = px / (float)rig->cv_width * rig->cv_lon_span_pixel + rig->cv_lon0;   /* :708 */
  float lat = py / (float)rig->cv_height * rig->cv_lat_range + rig->cv_lat0;
  float c = cosf(lat);
  ray[0] = c * sinf(lon); /* :713-717 */
  ray[1] = sinf(lat);
  ray[2] = c * cosf(lon);
}

/* One camera model: point (x,y,z) in the camera frame -> normalised grid + valid. */
static inline void orc_project(const MvsgiCamera* cam, float x, float y, float z, float* gx, float* gy, int* valid) {
  const float* p = cam->p;
  if (cam->model == MVSGI_CAM_EQUIRECT) {
    /* Equirectangular.point_3d_2_pixel  camera_models.py:741-764 */
    float lon0 = p[0], lon1 = p[1], lon_span = p[2], lat0 = p[3], lat_range = p[4];
    float r = sqrtf(z * z + x * x);                 /* linalg.norm((z,x)) :745 */
    float lat = atan2f(y, r);
    float lon = atan2f(x, z);
    float py = (lat - lat0) / lat_range;
    /* radian.py:61-62 shift_according_2_range */
    float two_pi = (float)(2.0 * M_PI);
    lon = lon + two_pi * ((lon < lon0 ? 1.0f : 0.0f) - (lon > lon1 ? 1.0f : 0.0f));
    float px = (lon - lon0) / lon_span;
    *valid = (px >= 0.0f) && (px <= 1.0f) && (py >= 0.0f) && (py <= 1.0f); /* :759 */
    *gx = px * 2.0f - 1.0f;
    *gy = py * 2.0f - 1.0f;
  } else if (cam->model == MVSGI_CAM_DOUBLE_SPHERE) {
    /* DoubleSphere.point_3d_2_pixel  camera_models.py:492-529 */
    float xi = p[0], alpha = p[1], oma = p[2], fx = p[3], fy = p[4], cx = p[5], cy = p[6], neg_w2 = p[7],
          half_fov = p[8];
    float x2 = x * x, y2 = y * y, z2 = z * z;
    float d1 = sqrtf(x2 + y2 + z2);
    float xd = xi * d1 + z;
    float d2 = sqrtf(x2 + y2 + xd * xd);
    float t = alpha * d2 + oma * xd;
    float px = fx / t * x + cx;
    float py = fy / t * y + cy;
    *gx = px / (float)cam->width * 2.0f - 1.0f;  /* :515-516 */
    *gy = py / (float)cam->height * 2.0f - 1.0f;
    float a = atan2f(sqrtf(x2 + y2), z);         /* :525 x2y2z_2_z_angle */
    *valid = (z > neg_w2 * d1) && (a <= half_fov);
  } else {
    /* LinearSphere.point_3d_2_pixel  camera_models.py:1506-1548 */
    float fov = p[0];
    float W = (float)cam->width, H = (float)cam->height;
    float R = sqrtf(x * x + y * y + z * z);      /* (x2+y2+z2)**0.5 */
    float ph = acosf(z / R);
    float th = atan2f(y, x);
    float r = ph * W / fov;
    float u = r * cosf(th) + W / 2.0f;
    float v = r * sinf(th) + H / 2.0f;
    *valid = (u >= 0.0f) && (u < W) && (v >= 0.0f) && (v < H);
    *gx = u / W * 2.0f - 1.0f;
    *gy = v / H * 2.0f - 1.0f;
  }
}

/* create_grids_from_rays  multi_view_camera_model_dataset.py:530-557 (+ :474-518):
 * grids [CAM,D,Ho,Wo,2], grid_masks [CAM,D,Ho,Wo] */
void orc_make_grids(const MvsgiRig* rig, float* grids, uint8_t* grid_masks) {
  const int Ho = rig->cv_height, Wo = rig->cv_width, D = rig->num_cands;
#pragma omp parallel for collapse(2) schedule(static)
  for (int cam = 0; cam < rig->num_cams; ++cam)
    for (int d = 0; d < D; ++d)
      for (int h = 0; h < Ho; ++h)
        for (int w = 0; w < Wo; ++w) {
          const MvsgiCamera* c = &rig->cams[cam];
          float ray[3];
          orc_cv_ray(rig, h, w, ray);
          /* torch_cuda_sweep.py:206-209  rays * candidates */
          float X = ray[0] * rig->dist[d], Y = ray[1] * rig->dist[d], Z = ray[2] * rig->dist[d];
          /* torch_cuda_sweep.py:398-403  R @ p + t  (inverse pose, ftensor.py:470-476) */
          float x = (c->rot[0] * X + c->rot[1] * Y) + c->rot[2] * Z + c->trans[0];
          float y = (c->rot[3] * X + c->rot[4] * Y) + c->rot[5] * Z + c->trans[1];
          float z = (c->rot[6] * X + c->rot[7] * Y) + c->rot[8] * Z + c->trans[2];
          float gx, gy;
          int valid;
          orc_project(c, x, y, z, &gx, &gy, &valid);
          int64_t o = (((int64_t)cam * D + d) * Ho + h) * Wo + w;
          grids[o * 2 + 0] = gx;
          grids[o * 2 + 1] = gy;
          grid_masks[o] = (uint8_t)valid;
        }
}

/* float64 restatement of the same geometry: the yardstick for "how far may the
 * kernel's fp32 grid be from the truth" (SURVEY.md §7 H2 / §8c criterion 3b). */
void orc_make_grids_f64(const MvsgiRig* rig, double* grids, uint8_t* grid_masks) {
  const int Ho = rig->cv_height, Wo = rig->cv_width, D = rig->num_cands;
#pragma omp parallel for collapse(2) schedule(static)
  for (int cam = 0; cam < rig->num_cams; ++cam)
    for (int d = 0; d < D; ++d)
      for (int h = 0; h < Ho; ++h)
        for (int w = 0; w < Wo; ++w) {
          const MvsgiCamera* c = &rig->cams[cam];
          const float* p = c->p;
          double lon = (w + 0.5) / Wo * (double)rig->cv_lon_span_pixel + (double)rig->cv_lon0;
          double lat = (h + 0.5) / Ho * (double)rig->cv_lat_range + (double)rig->cv_lat0;
          double cl = cos(lat), dist = rig->dist[d];
          double X = cl * sin(lon) * dist, Y = sin(lat) * dist, Z = cl * cos(lon) * dist;
          double x = c->rot[0] * X + c->rot[1] * Y + c->rot[2] * Z + c->trans[0];
          double y = c->rot[3] * X + c->rot[4] * Y + c->rot[5] * Z + c->trans[1];
          double z = c->rot[6] * X + c->rot[7] * Y + c->rot[8] * Z + c->trans[2];
          double gx, gy;
          int valid;
          if (c->model == MVSGI_CAM_EQUIRECT) {
            double la = atan2(y, sqrt(z * z + x * x)), lo = atan2(x, z);
            double py = (la - p[3]) / p[4];
            lo += 2.0 * M_PI * ((lo < p[0] ? 1.0 : 0.0) - (lo > p[1] ? 1.0 : 0.0));
            double px = (lo - p[0]) / p[2];
            valid = px >= 0 && px <= 1 && py >= 0 && py <= 1;
            gx = px * 2 - 1; gy = py * 2 - 1;
          } else if (c->model == MVSGI_CAM_DOUBLE_SPHERE) {
            double d1 = sqrt(x * x + y * y + z * z), xd = p[0] * d1 + z;
            double d2 = sqrt(x * x + y * y + xd * xd);
            double t = p[1] * d2 + p[2] * xd;
            gx = (p[3] / t * x + p[5]) / c->width * 2 - 1;
            gy = (p[4] / t * y + p[6]) / c->height * 2 - 1;
            valid = (z > p[7] * d1) && (atan2(sqrt(x * x + y * y), z) <= p[8]);
          } else {
            double R = sqrt(x * x + y * y + z * z), ph = acos(z / R), th = atan2(y, x);
            double r = ph * c->width / p[0];
            double u = r * cos(th) + c->width / 2.0, v = r * sin(th) + c->height / 2.0;
            valid = u >= 0 && u < c->width && v >= 0 && v < c->height;
            gx = u / c->width * 2 - 1; gy = v / c->height * 2 - 1;
          }
          int64_t o = (((int64_t)cam * D + d) * Ho + h) * Wo + w;
          grids[o * 2 + 0] = gx;
          grids[o * 2 + 1] = gy;
          grid_masks[o] = (uint8_t)valid;
        }
}

/* ------------------------------------------------------------------------- */
/* Fisheye -> surrogate-ERP resampling.
 * CameraModelRotation  dsta_mvs/image_sampler/camera_model_sampler.py:55-75 (grid), :98-120 (__call__):
 *   rays of the target pixels (Equirectangular.pixel_2_ray at pixel centres, planar_as_base.py:200-201),
 *   x_raw = R_raw_fisheye @ x (planar_as_base.py:206-209), raw.point_3d_2_pixel(normalized=True),
 *   invalid -> grid = -1 and the output pixel = invalid value (:67-71, :117),
 *   F.grid_sample(bilinear, zeros, align_corners=False) (:110-114).
 * rig: cv_* = target model, cams[0] = raw model with rot = R_raw_fisheye, trans = 0, dist[0] = 1.
 * img [N,C,H,W] contiguous -> out [N,C,Ht,Wt], valid [Ht,Wt].
 * grid_in/valid_in (optional, [Ht,Wt,2] / [Ht,Wt]): sample with this grid instead of evaluating the geometry here —
 * how the tests hand the oracle the CUDA kernel's own grid (SURVEY.md section 8c criterion 3b). */
void orc_resample(const MvsgiRig* rig, const float* img, int N, int C, int H, int W, float invalid, float* out,
                  uint8_t* valid_out, const float* grid_in, const uint8_t* valid_in) {
  const int Ht = rig->cv_height, Wt = rig->cv_width;
  const MvsgiCamera* c = &rig->cams[0];
#pragma omp parallel for schedule(static)
  for (int h = 0; h < Ht; ++h)
    for (int w = 0; w < Wt; ++w) {
      float ray[3], gx, gy;
      int valid;
      orc_cv_ray(rig, h, w, ray);
      float x = (c->rot[0] * ray[0] + c->rot[1] * ray[1]) + c->rot[2] * ray[2];
      float y = (c->rot[3] * ray[0] + c->rot[4] * ray[1]) + c->rot[5] * ray[2];
      float z = (c->rot[6] * ray[0] + c->rot[7] * ray[1]) + c->rot[8] * ray[2];
      orc_project(c, x, y, z, &gx, &gy, &valid);
      if (grid_in) {
        gx = grid_in[((int64_t)h * Wt + w) * 2];
        gy = grid_in[((int64_t)h * Wt + w) * 2 + 1];
        valid = valid_in[(int64_t)h * Wt + w] != 0;
      }
      if (valid_out) valid_out[(int64_t)h * Wt + w] = (uint8_t)valid;
      /* grid_sample: unnormalise (align_corners=False), 4 taps, out-of-bounds taps contribute 0 */
      float fx = ((gx + 1.0f) * (float)W - 1.0f) * 0.5f, fy = ((gy + 1.0f) * (float)H - 1.0f) * 0.5f;
      float x0f = floorf(fx), y0f = floorf(fy);
      float wnw = (x0f + 1.0f - fx) * (y0f + 1.0f - fy), wne = (fx - x0f) * (y0f + 1.0f - fy);
      float wsw = (x0f + 1.0f - fx) * (fy - y0f), wse = (fx - x0f) * (fy - y0f);
      int inr = valid && fx > -1.0f && fx < (float)W && fy > -1.0f && fy < (float)H;
      long x0 = inr ? (long)x0f : 0, y0 = inr ? (long)y0f : 0;
      for (int n = 0; n < N; ++n)
        for (int ch = 0; ch < C; ++ch) {
          const float* im = img + ((int64_t)n * C + ch) * H * W;
          float v = invalid;
          if (valid) {
            float nw = (inr && x0 >= 0 && y0 >= 0) ? im[y0 * W + x0] : 0.0f;
            float ne = (inr && x0 + 1 < W && y0 >= 0) ? im[y0 * W + x0 + 1] : 0.0f;
            float sw = (inr && x0 >= 0 && y0 + 1 < H) ? im[(y0 + 1) * W + x0] : 0.0f;
            float se = (inr && x0 + 1 < W && y0 + 1 < H) ? im[(y0 + 1) * W + x0 + 1] : 0.0f;
            v = ((nw * wnw + ne * wne) + sw * wsw) + se * wse;
          }
          out[(((int64_t)n * C + ch) * Ht + h) * Wt + w] = v;
        }
    }
}

/* CameraModelRotation.__call__(interpolation='nearest') (camera_model_sampler.py:98-120): F.grid_sample(mode='nearest',
 * zeros padding, align_corners=True — planar_as_base.py:127-133).  ATen: unnormalise ((g + 1) / 2) * (size - 1), round
 * half to even (nearbyint), a sample outside the image is 0; invalid rays get `invalid`. */
void orc_resample_nearest(const MvsgiRig* rig, const float* img, int N, int C, int H, int W, float invalid, float* out,
                          uint8_t* valid_out, const float* grid_in, const uint8_t* valid_in) {
  const int Ht = rig->cv_height, Wt = rig->cv_width;
  const MvsgiCamera* c = &rig->cams[0];
#pragma omp parallel for schedule(static)
  for (int h = 0; h < Ht; ++h)
    for (int w = 0; w < Wt; ++w) {
      float ray[3], gx, gy;
      int valid;
      orc_cv_ray(rig, h, w, ray);
      float x = (c->rot[0] * ray[0] + c->rot[1] * ray[1]) + c->rot[2] * ray[2];
      float y = (c->rot[3] * ray[0] + c->rot[4] * ray[1]) + c->rot[5] * ray[2];
      float z = (c->rot[6] * ray[0] + c->rot[7] * ray[1]) + c->rot[8] * ray[2];
      orc_project(c, x, y, z, &gx, &gy, &valid);
      if (grid_in) {
        gx = grid_in[((int64_t)h * Wt + w) * 2];
        gy = grid_in[((int64_t)h * Wt + w) * 2 + 1];
        valid = valid_in[(int64_t)h * Wt + w] != 0;
      }
      if (valid_out) valid_out[(int64_t)h * Wt + w] = (uint8_t)valid;
      float fx = nearbyintf(((gx + 1.0f) / 2.0f) * (float)(W - 1)), fy = nearbyintf(((gy + 1.0f) / 2.0f) * (float)(H - 1));
      int inb = valid && fx >= 0.0f && fx <= (float)(W - 1) && fy >= 0.0f && fy <= (float)(H - 1);
      long xn = inb ? (long)fx : 0, yn = inb ? (long)fy : 0;
      for (int n = 0; n < N; ++n)
        for (int ch = 0; ch < C; ++ch) {
          const float* im = img + ((int64_t)n * C + ch) * H * W;
          out[(((int64_t)n * C + ch) * Ht + h) * Wt + w] = valid ? (inb ? im[yn * W + xn] : 0.0f) : invalid;
        }
    }
}

int orc_num_threads(void) {
#if defined(_OPENMP)
  return omp_get_max_threads();
#else
  return 1;
#endif
}
