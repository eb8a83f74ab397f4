// sweep_bwd_concat.cu — backward of the concat sweep (the sampler's adjoint alone).
#include "sweep_bwd.inc"

namespace mvsgi {

cudaError_t launch_sweep_bwd_concat(int cams, const SweepKP& kp, const MvsgiRig& rig, int grid, size_t smem, cudaStream_t st) {
  return dispatch_bwd_cams<MVSGI_AGG_CONCAT>(cams, kp, rig, grid, smem, st);
}

}  // namespace mvsgi
