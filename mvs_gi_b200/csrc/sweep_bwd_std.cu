// sweep_bwd_std.cu — backward of the std (masked variance) sweep + the entry both aggregations are reached through.
#include "sweep_bwd.inc"

namespace mvsgi {

cudaError_t launch_sweep_bwd_concat(int cams, const SweepKP& kp, const MvsgiRig& rig, int grid, size_t smem, cudaStream_t st);

// kp as for the forward (fill_kp) with vol = the upstream gradient [B, Cout, Dvol, Ho, Wo] (NCDHW) and gradFeat = the
// zero-initialised gradient image in the layout of kp.feat
cudaError_t launch_sweep_bwd(int aggregate, int cams, const SweepKP& kp, const MvsgiRig& rig, int grid, size_t smem,
                             cudaStream_t st) {
  return aggregate == MVSGI_AGG_STD ? dispatch_bwd_cams<MVSGI_AGG_STD>(cams, kp, rig, grid, smem, st)
                                    : launch_sweep_bwd_concat(cams, kp, rig, grid, smem, st);
}

}  // namespace mvsgi
