// sweep_tiled.cu — launchers of the tile-walking sweep (sweep_tiled.cuh).
#include "launch.h"

namespace mvsgi {
namespace {

template <int CAM, int AGG>
cudaError_t table_cam(const SweepKP& kp, const TiledKP& tp, const MvsgiRig& rig, cudaStream_t st) {
  dim3 grid(tp.numTiles, kp.D);
  tile_table_kernel<CAM, AGG><<<grid, kTileVox, 0, st>>>(kp, tp, rig);
  return cudaGetLastError();
}

template <int AGG>
cudaError_t table_agg(int cams, const SweepKP& kp, const TiledKP& tp, const MvsgiRig& rig, cudaStream_t st) {
  switch (cams) {
    case 1: return table_cam<1, AGG>(kp, tp, rig, st);
    case 2: return table_cam<2, AGG>(kp, tp, rig, st);
    case 3: return table_cam<3, AGG>(kp, tp, rig, st);
    default: return table_cam<4, AGG>(kp, tp, rig, st);
  }
}

template <int CAM, int AGG, bool CL, int FB>
cudaError_t launch_k(const SweepKP& kp, const TiledKP& tp, int grid, cudaStream_t st) {
  auto kern = sweep_tiled_kernel<CAM, AGG, CL, FB>;
  // no shared memory to speak of: give the whole SM to L1 (the taps live there)
  static bool carveout_set[kMaxDevices] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < kMaxDevices && !carveout_set[dev]) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
    if (e != cudaSuccess) return e;
    carveout_set[dev] = true;
  }
  kern<<<grid, kTileThreads, 0, st>>>(kp, tp);
  return cudaGetLastError();
}

template <int CAM, int AGG>
cudaError_t launch_flags(const SweepKP& kp, const TiledKP& tp, int grid, int fb, cudaStream_t st) {
#ifdef MVSGI_TILE_ONLY_BENCH  // experiment builds: only the bench shape (3 cameras, std, channels-last volume)
  if (CAM != 3 || AGG != MVSGI_AGG_STD || !kp.volChannelsLast) return cudaErrorNotSupported;
  return fb == 2 ? launch_k<3, MVSGI_AGG_STD, true, 2>(kp, tp, grid, st) : launch_k<3, MVSGI_AGG_STD, true, 1>(kp, tp, grid, st);
#else
  if (kp.volChannelsLast)
    return fb == 2 ? launch_k<CAM, AGG, true, 2>(kp, tp, grid, st) : launch_k<CAM, AGG, true, 1>(kp, tp, grid, st);
  return fb == 2 ? launch_k<CAM, AGG, false, 2>(kp, tp, grid, st) : launch_k<CAM, AGG, false, 1>(kp, tp, grid, st);
#endif
}

template <int AGG>
cudaError_t launch_agg(int cams, const SweepKP& kp, const TiledKP& tp, int grid, int fb, cudaStream_t st) {
  switch (cams) {
    case 1: return launch_flags<1, AGG>(kp, tp, grid, fb, st);
    case 2: return launch_flags<2, AGG>(kp, tp, grid, fb, st);
    case 3: return launch_flags<3, AGG>(kp, tp, grid, fb, st);
    default: return launch_flags<4, AGG>(kp, tp, grid, fb, st);
  }
}

}  // namespace

int tiled_warps_per_cta() { return kTileWarps; }
int tiled_ctas_per_sm() { return kTileCtas; }

cudaError_t launch_tile_table(int aggregate, int cams, const SweepKP& kp, const TiledKP& tp, const MvsgiRig& rig,
                              cudaStream_t st) {
  return aggregate == MVSGI_AGG_STD ? table_agg<MVSGI_AGG_STD>(cams, kp, tp, rig, st)
                                    : table_agg<MVSGI_AGG_CONCAT>(cams, kp, tp, rig, st);
}

cudaError_t launch_sweep_tiled(int aggregate, int cams, const SweepKP& kp, const TiledKP& tp, int grid, int frames_per_pass,
                               cudaStream_t st) {
  return aggregate == MVSGI_AGG_STD ? launch_agg<MVSGI_AGG_STD>(cams, kp, tp, grid, frames_per_pass, st)
                                    : launch_agg<MVSGI_AGG_CONCAT>(cams, kp, tp, grid, frames_per_pass, st);
}

}  // namespace mvsgi
