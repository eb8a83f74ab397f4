// softargmin_tma.cu — host side of the TMA-staged soft-argmin (softargmin_tma.cuh): tensor map + launch.
#include "launch.h"
#include "softargmin_tma.cuh"
#include "tma_host.h"

namespace mvsgi {
namespace {

// resident CTAs per SM of one instantiation x SM count of the current device, looked up once
template <int DREG, bool FULL, bool NC, bool AM>
int persistent_ctas() {
  static const int n = [] {
    int dev = 0, sms = 0, per = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // the ring of every resident CTA lives in shared memory: ask for the largest carve-out before counting them
    cudaFuncSetAttribute(softargmin_x2_tma_kernel<DREG, FULL, NC, AM>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, softargmin_x2_tma_kernel<DREG, FULL, NC, AM>, kSaThreads, 0);
    return sms * (per > 0 ? per : 1);
  }();
  return n;
}

template <int DREG, bool FULL, bool NC, bool AM>
void launch_one(const SoftKP& kp, const CUtensorMap& tm, SaTiles tl, cudaStream_t st) {
  const int grid = tl.total < persistent_ctas<DREG, FULL, NC, AM>() ? tl.total : persistent_ctas<DREG, FULL, NC, AM>();
  tl.stepX = grid % tl.tilesX;
  tl.stepY = (grid / tl.tilesX) % tl.tilesY;
  tl.stepB = grid / (tl.tilesX * tl.tilesY);
  softargmin_x2_tma_kernel<DREG, FULL, NC, AM><<<grid, kSaThreads, 0, st>>>(kp, tm, tl);
}

template <int DREG, bool FULL>
void launch_flags(const SoftKP& kp, const CUtensorMap& tm, const SaTiles& tl, cudaStream_t st) {
  const bool nc = kp.normCosts != nullptr, am = kp.argmax != nullptr;
  if (nc && am) launch_one<DREG, FULL, true, true>(kp, tm, tl, st);
  else if (nc) launch_one<DREG, FULL, true, false>(kp, tm, tl, st);
  else if (am) launch_one<DREG, FULL, false, true>(kp, tm, tl, st);
  else launch_one<DREG, FULL, false, false>(kp, tm, tl, st);
}

}  // namespace

// Launches the TMA-staged kernel when the cost volume can be described by a tensor map (unit column stride, 16-byte
// aligned base and row / plane / frame pitches); returns false — nothing launched — otherwise, and the caller uses
// softargmin_x2_kernel (same bits).
bool launch_softargmin_x2_tma(const SoftKP& kp, cudaStream_t st) {
  if (kp.sX != 1 || kp.D > 16 || kp.Win < 2 || kp.Hin < 2) return false;
  if ((reinterpret_cast<uintptr_t>(kp.costs) & 15) != 0) return false;
  const long long sB = kp.B > 1 ? kp.sB : (long long)kp.D * kp.sD;  // the frame pitch is irrelevant for one frame
  if (kp.sY % 4 != 0 || kp.sD % 4 != 0 || sB % 4 != 0 || kp.sY < kp.Win || kp.sD <= 0 || sB <= 0) return false;
  const int dreg = kp.D <= 8 ? 8 : 16;
  CUtensorMap tm;
  if (!make_map_f32_4d(&tm, kp.costs, kp.Win, kp.Hin, kp.D, kp.B, (cuuint64_t)kp.sY * 4, (cuuint64_t)kp.sD * 4,
                       (cuuint64_t)sB * 4, kSaBoxW, kSaBoxH, kp.D))
    return false;
  SaTiles tl = {};
  tl.tilesX = (kp.Win + 1 + kSaCols - 1) / kSaCols;
  tl.tilesY = (kp.Hin + 1 + kSaRows - 1) / kSaRows;
  const long long total = (long long)tl.tilesX * tl.tilesY * kp.B;
  if (total > (1ll << 30)) return false;
  tl.total = (int)total;
  if (dreg == 16) {
    if (kp.D == 16) launch_flags<16, true>(kp, tm, tl, st);
    else launch_flags<16, false>(kp, tm, tl, st);
  } else {
    if (kp.D == 8) launch_flags<8, true>(kp, tm, tl, st);
    else launch_flags<8, false>(kp, tm, tl, st);
  }
  return true;
}

}  // namespace mvsgi
