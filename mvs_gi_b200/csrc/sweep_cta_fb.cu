// sweep_cta_fb.cu — launchers of the several-frames-per-pass instantiations of sweep_kernel (sweep.cuh, FB > 1):
// static masks, shared grids, whole 16-channel groups tapped in place, channels-last volume (fp32 or bf16), no validity
// output, no peers — the inference path.  validate_and_plan (mvsgi_b200.cu) decides; sweep_frames_per_pass() picks FB.
#include "launch.h"

namespace mvsgi {
namespace {

template <int CAM, int AGG, bool VBF16, int FB>
cudaError_t launch_fb(const SweepKP& kp, const MvsgiRig& rig, int grid, size_t smem, cudaStream_t st) {
  auto kern = sweep_kernel<CAM, AGG, true, false, false, true, false, true, VBF16, FB>;
  static DynSmemSet smem_set;  // per instantiation
  cudaError_t e = ensure_dyn_smem(kern, smem, smem_set);
  if (e != cudaSuccess) return e;
  kern<<<grid, sweep_threads(FB), smem, st>>>(kp, rig);
  return cudaGetLastError();
}

template <int CAM, int AGG>
cudaError_t dispatch_fb(const SweepKP& kp, const MvsgiRig& rig, int grid, size_t smem, cudaStream_t st) {
  constexpr int FBW = CAM == 4 ? 3 : 4;  // the wide variant of this camera count (sweep_frames_per_pass)
  if (kp.framesPerPass == FBW)
    return kp.volBf16 ? launch_fb<CAM, AGG, true, FBW>(kp, rig, grid, smem, st)
                      : launch_fb<CAM, AGG, false, FBW>(kp, rig, grid, smem, st);
  if (kp.framesPerPass == 2)
    return kp.volBf16 ? launch_fb<CAM, AGG, true, 2>(kp, rig, grid, smem, st)
                      : launch_fb<CAM, AGG, false, 2>(kp, rig, grid, smem, st);
  return cudaErrorInvalidValue;
}

template <int AGG>
cudaError_t dispatch_fb_cams(int cams, const SweepKP& kp, const MvsgiRig& rig, int grid, size_t smem, cudaStream_t st) {
  switch (cams) {
#ifndef MVSGI_BENCH_ONLY
    case 1: return dispatch_fb<1, AGG>(kp, rig, grid, smem, st);
    case 2: return dispatch_fb<2, AGG>(kp, rig, grid, smem, st);
#endif
    case 3: return dispatch_fb<3, AGG>(kp, rig, grid, smem, st);
    case 4: return dispatch_fb<4, AGG>(kp, rig, grid, smem, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace

cudaError_t launch_sweep_cta_fb(int aggregate, int cams, const SweepKP& kp, const MvsgiRig& rig, int grid, size_t smem,
                                cudaStream_t st) {
#ifdef MVSGI_BENCH_ONLY
  if (aggregate != MVSGI_AGG_STD) return cudaErrorNotSupported;
  return dispatch_fb_cams<MVSGI_AGG_STD>(cams, kp, rig, grid, smem, st);
#else
  return aggregate == MVSGI_AGG_STD ? dispatch_fb_cams<MVSGI_AGG_STD>(cams, kp, rig, grid, smem, st)
                                    : dispatch_fb_cams<MVSGI_AGG_CONCAT>(cams, kp, rig, grid, smem, st);
#endif
}

}  // namespace mvsgi
