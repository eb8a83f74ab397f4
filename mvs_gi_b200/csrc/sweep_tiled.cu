// sweep_tiled.cu — launchers of the tile-pipelined sweep (sweep_tiled.cuh).
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

template <int CAM, int AGG, bool FULLC, bool VALID, bool PEERS, bool CL>
cudaError_t launch_k(const SweepKP& kp, const TiledKP& tp, int grid, size_t smem, cudaStream_t st) {
  auto kern = sweep_tiled_kernel<CAM, AGG, FULLC, VALID, PEERS, CL>;
  static DynSmemSet smem_set;  // per instantiation
  cudaError_t e = ensure_dyn_smem(kern, smem, smem_set);
  if (e != cudaSuccess) return e;
  kern<<<grid, kTileThreads, smem, st>>>(kp, tp);
  return cudaGetLastError();
}

template <int CAM, int AGG, bool FULLC, bool VALID, bool PEERS>
cudaError_t launch_cl(const SweepKP& kp, const TiledKP& tp, int grid, size_t smem, cudaStream_t st) {
  return kp.volChannelsLast ? launch_k<CAM, AGG, FULLC, VALID, PEERS, true>(kp, tp, grid, smem, st)
                            : launch_k<CAM, AGG, FULLC, VALID, PEERS, false>(kp, tp, grid, smem, st);
}

template <int CAM, int AGG>
cudaError_t launch_flags(const SweepKP& kp, const TiledKP& tp, int grid, size_t smem, cudaStream_t st) {
  const bool fullc = kp.C % kCG == 0;
  if (kp.numPeers > 0)
    return fullc ? launch_cl<CAM, AGG, true, false, true>(kp, tp, grid, smem, st)
                 : launch_cl<CAM, AGG, false, false, true>(kp, tp, grid, smem, st);
  if (AGG == MVSGI_AGG_STD && kp.validOut != nullptr)
    return fullc ? launch_cl<CAM, AGG, true, true, false>(kp, tp, grid, smem, st)
                 : launch_cl<CAM, AGG, false, true, false>(kp, tp, grid, smem, st);
  return fullc ? launch_cl<CAM, AGG, true, false, false>(kp, tp, grid, smem, st)
               : launch_cl<CAM, AGG, false, false, false>(kp, tp, grid, smem, st);
}

template <int AGG>
cudaError_t launch_agg(int cams, const SweepKP& kp, const TiledKP& tp, int grid, size_t smem, cudaStream_t st) {
  switch (cams) {
    case 1: return launch_flags<1, AGG>(kp, tp, grid, smem, st);
    case 2: return launch_flags<2, AGG>(kp, tp, grid, smem, st);
    case 3: return launch_flags<3, AGG>(kp, tp, grid, smem, st);
    default: return launch_flags<4, AGG>(kp, tp, grid, smem, st);
  }
}

}  // namespace

cudaError_t launch_tile_table(int aggregate, int cams, const SweepKP& kp, const TiledKP& tp, const MvsgiRig& rig,
                              cudaStream_t st) {
  return aggregate == MVSGI_AGG_STD ? table_agg<MVSGI_AGG_STD>(cams, kp, tp, rig, st)
                                    : table_agg<MVSGI_AGG_CONCAT>(cams, kp, tp, rig, st);
}

cudaError_t launch_sweep_tiled(int aggregate, int cams, const SweepKP& kp, const TiledKP& tp, int grid, size_t smem,
                               cudaStream_t st) {
  return aggregate == MVSGI_AGG_STD ? launch_agg<MVSGI_AGG_STD>(cams, kp, tp, grid, smem, st)
                                    : launch_agg<MVSGI_AGG_CONCAT>(cams, kp, tp, grid, smem, st);
}

}  // namespace mvsgi
