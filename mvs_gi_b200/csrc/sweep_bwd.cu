// sweep_bwd.cu — launchers of the backward instantiation of the CTA-resident-table sweep (sweep.cuh, BWD = true).
#pragma nv_diag_suppress 128  // BWD instantiations: the forward half of the loop body is unreachable after `continue`
#include "launch.h"

namespace mvsgi {
namespace {

template <int CAM, int AGG, bool FULLC, bool DIRECT>
cudaError_t launch_bwd(const SweepKP& kp, const MvsgiRig& rig, int grid, size_t smem, cudaStream_t st) {
  auto kern = sweep_kernel<CAM, AGG, FULLC, false, false, false, false, DIRECT, false, 1, true>;
  static DynSmemSet smem_set;  // per instantiation
  cudaError_t e = ensure_dyn_smem(kern, smem, smem_set);
  if (e != cudaSuccess) return e;
  kern<<<grid, kSweepBwdThreads, smem, st>>>(kp, rig);
  return cudaGetLastError();
}

template <int CAM, int AGG>
cudaError_t dispatch_bwd(const SweepKP& kp, const MvsgiRig& rig, int grid, size_t smem, cudaStream_t st) {
  const bool fullc = kp.C % kCG == 0;
  if (kp.direct) return fullc ? launch_bwd<CAM, AGG, true, true>(kp, rig, grid, smem, st) : cudaErrorNotSupported;
  return fullc ? launch_bwd<CAM, AGG, true, false>(kp, rig, grid, smem, st)
               : launch_bwd<CAM, AGG, false, false>(kp, rig, grid, smem, st);
}

template <int AGG>
cudaError_t dispatch_bwd_cams(int cams, const SweepKP& kp, const MvsgiRig& rig, int grid, size_t smem, cudaStream_t st) {
  switch (cams) {
    case 1: return dispatch_bwd<1, AGG>(kp, rig, grid, smem, st);
    case 2: return dispatch_bwd<2, AGG>(kp, rig, grid, smem, st);
    case 3: return dispatch_bwd<3, AGG>(kp, rig, grid, smem, st);
    default: return dispatch_bwd<4, AGG>(kp, rig, grid, smem, st);
  }
}

}  // namespace

// kp as for the forward (fill_kp) with vol = the upstream gradient [B, Cout, Dvol, Ho, Wo] (NCDHW) and gradFeat = the
// zero-initialised gradient image in the layout of kp.feat
cudaError_t launch_sweep_bwd(int aggregate, int cams, const SweepKP& kp, const MvsgiRig& rig, int grid, size_t smem,
                             cudaStream_t st) {
  return aggregate == MVSGI_AGG_STD ? dispatch_bwd_cams<MVSGI_AGG_STD>(cams, kp, rig, grid, smem, st)
                                    : dispatch_bwd_cams<MVSGI_AGG_CONCAT>(cams, kp, rig, grid, smem, st);
}

}  // namespace mvsgi
