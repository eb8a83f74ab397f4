// sweep_cta.cu — launchers of the CTA-resident-table sweep (sweep.cuh).
#include "launch.h"
#ifdef MVSGI_EXPERIMENTS
#include "sweep_walk.cuh"
#endif

namespace mvsgi {
namespace {

template <int CAM, int AGG, bool FULLC, bool VALID, bool PEERS, bool CL, bool BF16 = false, bool DIRECT = false,
          bool VBF16 = false>
cudaError_t launch_sweep_cl(const SweepKP& kp, const MvsgiRig& rig, int grid, size_t smem, cudaStream_t st) {
  auto kern = sweep_kernel<CAM, AGG, FULLC, VALID, PEERS, CL, BF16, DIRECT, VBF16>;
  static DynSmemSet smem_set;  // per instantiation
  cudaError_t e = ensure_dyn_smem(kern, smem, smem_set);
  if (e != cudaSuccess) return e;
  kern<<<grid, kSweepThreads, smem, st>>>(kp, rig);
  return cudaGetLastError();
}

template <int CAM, int AGG, bool FULLC, bool VALID, bool PEERS = false>
cudaError_t launch_sweep(const SweepKP& kp, const MvsgiRig& rig, int grid, size_t smem, cudaStream_t st) {
  if (FULLC && kp.direct)  // channels-last features tapped in place (always whole 16-channel groups)
    return kp.volChannelsLast ? launch_sweep_cl<CAM, AGG, FULLC, VALID, PEERS, true, false, FULLC>(kp, rig, grid, smem, st)
                              : launch_sweep_cl<CAM, AGG, FULLC, VALID, PEERS, false, false, FULLC>(kp, rig, grid, smem, st);
  return kp.volChannelsLast ? launch_sweep_cl<CAM, AGG, FULLC, VALID, PEERS, true>(kp, rig, grid, smem, st)
                            : launch_sweep_cl<CAM, AGG, FULLC, VALID, PEERS, false>(kp, rig, grid, smem, st);
}

template <int CAM, int AGG>
cudaError_t dispatch_flags(const SweepKP& kp, const MvsgiRig& rig, int grid, size_t smem, cudaStream_t st) {
  const bool fullc = kp.C % kCG == 0;
  const bool valid = kp.validOut != nullptr;
  if (kp.volBf16) {  // validate_and_plan: channels-last volume, no validity output, no peers, fp32 features
    if (fullc && kp.direct) return launch_sweep_cl<CAM, AGG, true, false, false, true, false, true, true>(kp, rig, grid, smem, st);
    return fullc ? launch_sweep_cl<CAM, AGG, true, false, false, true, false, false, true>(kp, rig, grid, smem, st)
                 : launch_sweep_cl<CAM, AGG, false, false, false, true, false, false, true>(kp, rig, grid, smem, st);
  }
  if (kp.featsBf16) {  // validate_and_plan: no validity output, no peers in this mode
    if (kp.volChannelsLast)
      return fullc ? launch_sweep_cl<CAM, AGG, true, false, false, true, true>(kp, rig, grid, smem, st)
                   : launch_sweep_cl<CAM, AGG, false, false, false, true, true>(kp, rig, grid, smem, st);
    return fullc ? launch_sweep_cl<CAM, AGG, true, false, false, false, true>(kp, rig, grid, smem, st)
                 : launch_sweep_cl<CAM, AGG, false, false, false, false, true>(kp, rig, grid, smem, st);
  }
  if (kp.numPeers > 0)
    return fullc ? launch_sweep<CAM, AGG, true, false, true>(kp, rig, grid, smem, st)
                 : launch_sweep<CAM, AGG, false, false, true>(kp, rig, grid, smem, st);
  if (AGG == MVSGI_AGG_STD && valid)
    return fullc ? launch_sweep<CAM, AGG, true, true>(kp, rig, grid, smem, st)
                 : launch_sweep<CAM, AGG, false, true>(kp, rig, grid, smem, st);
  return fullc ? launch_sweep<CAM, AGG, true, false>(kp, rig, grid, smem, st)
               : launch_sweep<CAM, AGG, false, false>(kp, rig, grid, smem, st);
}

#ifdef MVSGI_EXPERIMENTS
template <int CAM, int AGG, bool VALID, bool PEERS>
cudaError_t launch_walk(const SweepKP& kp, const MvsgiRig& rig, int grid, size_t smem, cudaStream_t st) {
  auto kern = sweep_walk_kernel<CAM, AGG, VALID, PEERS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, kWalkThreads, smem, st>>>(kp, rig);
  return cudaGetLastError();
}

template <int CAM, int AGG>
cudaError_t dispatch_walk(const SweepKP& kp, const MvsgiRig& rig, int grid, size_t smem, cudaStream_t st) {
  if (kp.numPeers > 0) return launch_walk<CAM, AGG, false, true>(kp, rig, grid, smem, st);
  if (AGG == MVSGI_AGG_STD && kp.validOut != nullptr) return launch_walk<CAM, AGG, true, false>(kp, rig, grid, smem, st);
  return launch_walk<CAM, AGG, false, false>(kp, rig, grid, smem, st);
}

#endif

template <int AGG>
cudaError_t dispatch_cams(int cams, const SweepKP& kp, const MvsgiRig& rig, int grid, size_t smem,
                          cudaStream_t st) {
#ifdef MVSGI_EXPERIMENTS
  if (kp.wL > 0) return cams == 3 ? dispatch_walk<3, AGG>(kp, rig, grid, smem, st) : dispatch_walk<4, AGG>(kp, rig, grid, smem, st);
#endif
#ifdef MVSGI_BENCH_ONLY  // quick tuning builds: only the std builder with 3 or 4 cameras
  if (AGG != MVSGI_AGG_STD || cams < 3) return cudaErrorNotSupported;
  return cams == 3 ? dispatch_flags<3, MVSGI_AGG_STD>(kp, rig, grid, smem, st)
                   : dispatch_flags<4, MVSGI_AGG_STD>(kp, rig, grid, smem, st);
#else
  switch (cams) {
    case 1: return dispatch_flags<1, AGG>(kp, rig, grid, smem, st);
    case 2: return dispatch_flags<2, AGG>(kp, rig, grid, smem, st);
    case 3: return dispatch_flags<3, AGG>(kp, rig, grid, smem, st);
    default: return dispatch_flags<4, AGG>(kp, rig, grid, smem, st);
  }
#endif
}


}  // namespace

#ifdef MVSGI_EXPERIMENTS
cudaError_t launch_sweep_walk(int aggregate, int cams, const SweepKP& kp, const MvsgiRig& rig, int grid, size_t smem,
                              cudaStream_t st) {
  return aggregate == MVSGI_AGG_STD ? dispatch_cams<MVSGI_AGG_STD>(cams, kp, rig, grid, smem, st)
                                    : dispatch_cams<MVSGI_AGG_CONCAT>(cams, kp, rig, grid, smem, st);
}
#endif

cudaError_t launch_sweep_cta(int aggregate, int cams, const SweepKP& kp, const MvsgiRig& rig, int grid, size_t smem,
                             cudaStream_t st) {
  if (kp.framesPerPass > 1) return launch_sweep_cta_fb(aggregate, cams, kp, rig, grid, smem, st);  // sweep_cta_fb.cu
  return aggregate == MVSGI_AGG_STD ? dispatch_cams<MVSGI_AGG_STD>(cams, kp, rig, grid, smem, st)
                                    : dispatch_cams<MVSGI_AGG_CONCAT>(cams, kp, rig, grid, smem, st);
}

}  // namespace mvsgi
