// launch.h — internal interface between the translation units of libmvsgi_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include "sweep.cuh"
#ifdef MVSGI_EXPERIMENTS
#include "sweep_tiled.cuh"
#endif

namespace mvsgi {

// cudaFuncAttributeMaxDynamicSharedMemorySize only ever has to grow, and is per device: remember what was set
constexpr int kMaxDevices = 32;
struct DynSmemSet {
  size_t bytes[kMaxDevices] = {};
};
template <typename K>
inline cudaError_t ensure_dyn_smem(K kern, size_t smem, DynSmemSet& set) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < kMaxDevices && smem <= set.bytes[dev]) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess && dev >= 0 && dev < kMaxDevices) set.bytes[dev] = smem;
  return e;
}

// sweep_cta.cu: the CTA-resident-table sweep (sweep.cuh) — per-frame grids, bf16-staged features, any shape
cudaError_t launch_sweep_cta(int aggregate, int cams, const SweepKP& kp, const MvsgiRig& rig, int grid, size_t smem,
                             cudaStream_t st);
// sweep_cta_fb.cu: the same kernel with several frames per pass (kp.framesPerPass > 1)
cudaError_t launch_sweep_cta_fb(int aggregate, int cams, const SweepKP& kp, const MvsgiRig& rig, int grid, size_t smem,
                                cudaStream_t st);
// sweep_bwd_std.cu / sweep_bwd_concat.cu: the backward instantiation of the same kernel (sweep.cuh, BWD): scatter into kp.gradFeat
cudaError_t launch_sweep_bwd(int aggregate, int cams, const SweepKP& kp, const MvsgiRig& rig, int grid, size_t smem,
                             cudaStream_t st);
// softargmin_tma.cu: scale-factor-2 soft-argmin with the input box staged by TMA; false = not applicable, nothing launched
struct SoftKP;
bool launch_softargmin_x2_tma(const SoftKP& kp, cudaStream_t st);
// softargmin_bwd.cu: gradient of the scale-factor-2 soft-argmin as a gather (no atomics, no zero fill); D <= 16
struct SoftBwdKP;
void launch_softargmin_x2_backward(const SoftBwdKP& kb, cudaStream_t st);
#ifdef MVSGI_EXPERIMENTS
cudaError_t launch_sweep_walk(int aggregate, int cams, const SweepKP& kp, const MvsgiRig& rig, int grid, size_t smem,
                              cudaStream_t st);
#endif

#ifdef MVSGI_EXPERIMENTS
// sweep_tiled.cu: shape of the tile-walking kernel (compile-time constants of that translation unit)
int tiled_warps_per_cta();
int tiled_ctas_per_sm();
// sweep_tiled.cu: geometry table + tile-walking sweep (sweep_tiled.cuh); 2 launches
cudaError_t launch_tile_table(int aggregate, int cams, const SweepKP& kp, const TiledKP& tp, const MvsgiRig& rig,
                              cudaStream_t st);
cudaError_t launch_sweep_tiled(int aggregate, int cams, const SweepKP& kp, const TiledKP& tp, int grid, int frames_per_pass,
                               cudaStream_t st);
#endif

}  // namespace mvsgi
