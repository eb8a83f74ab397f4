// geometry.cuh — in-kernel evaluation of the sampling grid.
//
// What the reference precomputes on the CPU at dataset construction and then
// streams from HBM for every frame (grids [B,CAM,D,Ho,Wo,2] + grid_masks) is
// evaluated here from ~700 bytes of calibration held in kernel parameters.
//
// All arithmetic is fp32 in the reference's operation order, written with the
// round-to-nearest intrinsics so nvcc cannot contract a*b+c into an FMA where
// eager PyTorch rounds twice.  Reference (paths relative to the checkout):
//   dsta_mvs/mvs_utils/camera_models.py:680-720   Equirectangular.pixel_2_ray
//   dsta_mvs/support/dataset/torch_cuda_sweep.py:186-215, 385-408   rays * candidates, R p + t
//   dsta_mvs/mvs_utils/camera_models.py:722-772   Equirectangular.point_3d_2_pixel
//   dsta_mvs/mvs_utils/camera_models.py:475-534   DoubleSphere.point_3d_2_pixel
//   dsta_mvs/mvs_utils/camera_models.py:1489-1548 LinearSphere.point_3d_2_pixel
//   dsta_mvs/mvs_utils/radian.py:42-66            shift_according_2_range
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mvsgi_b200.h"

namespace mvsgi {

#define MVSGI_DEV __device__ __forceinline__

MVSGI_DEV float fmul(float a, float b) { return __fmul_rn(a, b); }
MVSGI_DEV float fadd(float a, float b) { return __fadd_rn(a, b); }
MVSGI_DEV float fsub(float a, float b) { return __fsub_rn(a, b); }
MVSGI_DEV float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// A.1  unit ray of cost-volume pixel (h, w), pixel centre at +0.5.
MVSGI_DEV void cv_ray(const MvsgiRig& rig, int h, int w, float& rx, float& ry, float& rz) {
  const float px = (float)w + 0.5f, py = (float)h + 0.5f;
  const float lon = fadd(fmul(fdiv(px, (float)rig.cv_width), rig.cv_lon_span_pixel), rig.cv_lon0);
  const float lat = fadd(fmul(fdiv(py, (float)rig.cv_height), rig.cv_lat_range), rig.cv_lat0);
  float sl, cl, so, co;
  sincosf(lat, &sl, &cl);
  sincosf(lon, &so, &co);
  rx = fmul(cl, so);
  ry = sl;
  rz = fmul(cl, co);
}

// A.2  scale by the candidate distance and move into the camera frame: p' = R_i p + t_i.
MVSGI_DEV void to_camera(const MvsgiCamera& c, float rx, float ry, float rz, float dist, float& x, float& y,
                         float& z) {
  const float X = fmul(rx, dist), Y = fmul(ry, dist), Z = fmul(rz, dist);
  x = fadd(fadd(fadd(fmul(c.rot[0], X), fmul(c.rot[1], Y)), fmul(c.rot[2], Z)), c.trans[0]);
  y = fadd(fadd(fadd(fmul(c.rot[3], X), fmul(c.rot[4], Y)), fmul(c.rot[5], Z)), c.trans[1]);
  z = fadd(fadd(fadd(fmul(c.rot[6], X), fmul(c.rot[7], Y)), fmul(c.rot[8], Z)), c.trans[2]);
}

// A.3-A.5  camera-frame point -> normalised grid coordinate in [-1,1] + validity bit.
MVSGI_DEV bool project(const MvsgiCamera& c, float x, float y, float z, float& gx, float& gy) {
  const float* p = c.p;
  if (c.model == MVSGI_CAM_EQUIRECT) {
    const float lon0 = p[0], lon1 = p[1], lon_span = p[2], lat0 = p[3], lat_range = p[4];
    const float r = sqrtf(fadd(fmul(z, z), fmul(x, x)));
    const float lat = atan2f(y, r);
    float lon = atan2f(x, z);
    const float py = fdiv(fsub(lat, lat0), lat_range);
    const float two_pi = 6.283185307179586f;
    const float shift = (lon < lon0 ? 1.0f : 0.0f) - (lon > lon1 ? 1.0f : 0.0f);
    lon = fadd(lon, fmul(two_pi, shift));
    const float px = fdiv(fsub(lon, lon0), lon_span);
    gx = fsub(fmul(px, 2.0f), 1.0f);
    gy = fsub(fmul(py, 2.0f), 1.0f);
    return (px >= 0.0f) && (px <= 1.0f) && (py >= 0.0f) && (py <= 1.0f);
  } else if (c.model == MVSGI_CAM_DOUBLE_SPHERE) {
    const float xi = p[0], alpha = p[1], oma = p[2], fx = p[3], fy = p[4], cx = p[5], cy = p[6], neg_w2 = p[7],
                half_fov = p[8];
    const float x2 = fmul(x, x), y2 = fmul(y, y), z2 = fmul(z, z);
    const float x2y2 = fadd(x2, y2);
    const float d1 = sqrtf(fadd(x2y2, z2));
    const float xd = fadd(fmul(xi, d1), z);
    const float d2 = sqrtf(fadd(x2y2, fmul(xd, xd)));
    const float t = fadd(fmul(alpha, d2), fmul(oma, xd));
    const float px = fadd(fmul(fdiv(fx, t), x), cx);
    const float py = fadd(fmul(fdiv(fy, t), y), cy);
    gx = fsub(fmul(fdiv(px, (float)c.width), 2.0f), 1.0f);
    gy = fsub(fmul(fdiv(py, (float)c.height), 2.0f), 1.0f);
    const float a = atan2f(sqrtf(x2y2), z);
    return (z > fmul(neg_w2, d1)) && (a <= half_fov);
  } else {
    const float fov = p[0];
    const float W = (float)c.width, H = (float)c.height;
    const float R = sqrtf(fadd(fadd(fmul(x, x), fmul(y, y)), fmul(z, z)));
    const float ph = acosf(fdiv(z, R));
    const float th = atan2f(y, x);
    const float r = fdiv(fmul(ph, W), fov);
    float st, ct;
    sincosf(th, &st, &ct);
    const float u = fadd(fmul(r, ct), fdiv(W, 2.0f));
    const float v = fadd(fmul(r, st), fdiv(H, 2.0f));
    gx = fsub(fmul(fdiv(u, W), 2.0f), 1.0f);
    gy = fsub(fmul(fdiv(v, H), 2.0f), 1.0f);
    return (u >= 0.0f) && (u < W) && (v >= 0.0f) && (v < H);
  }
}

// A.6  normalised grid -> continuous pixel coordinate (backports.py:41-42, align_corners=False).
MVSGI_DEV float unnormalise(float g, int size) {
  return fmul(fsub(fmul(fadd(g, 1.0f), (float)size), 1.0f), 0.5f);
}

}  // namespace mvsgi
