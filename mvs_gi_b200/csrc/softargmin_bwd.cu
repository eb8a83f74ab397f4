// softargmin_bwd.cu — launch of the gather-form soft-argmin backward (softargmin_bwd.cuh).
#include "launch.h"
#include "softargmin_bwd.cuh"
#include "tma_host.h"

namespace mvsgi {
namespace {

template <int DREG, bool FULL, int GNC>
void launch_mode(const SoftBwdKP& kb, const CUtensorMap& gm, dim3 grid, cudaStream_t st) {
  size_t smem = (size_t)(kSbRows + 1) * DREG * 32 * 4;  // the exchange buffer
  if (GNC == 2) {
    smem += (size_t)DREG * kSbBoxH * kSbBoxW * 4;  // + the grad_norm_costs box: past the 48 KB default
    static DynSmemSet set;
    if (ensure_dyn_smem(softargmin_x2_backward_kernel<DREG, FULL, GNC>, smem, set) != cudaSuccess) return;
  }
  softargmin_x2_backward_kernel<DREG, FULL, GNC><<<grid, kSbThreads, smem, st>>>(kb, gm);
}

template <int DREG, bool FULL>
void launch_gnc(const SoftBwdKP& kb, dim3 grid, cudaStream_t st) {
  const SoftKP& kp = kb.f;
  CUtensorMap gm = {};
  if (kb.gNc == nullptr) return launch_mode<DREG, FULL, 0>(kb, gm, grid, st);
  // grad_norm_costs is [B, D, Hout, Wout] contiguous: a tensor map needs a 16-byte aligned base and row pitch
  const cuuint64_t row = (cuuint64_t)kp.Wout * 4, plane = row * kp.Hout, frame = plane * kp.D;
  if ((reinterpret_cast<uintptr_t>(kb.gNc) & 15) == 0 && kp.Wout % 4 == 0 &&
      make_map_f32_4d(&gm, kb.gNc, kp.Wout, kp.Hout, kp.D, kp.B, row, plane, frame, kSbBoxW, kSbBoxH, kp.D))
    return launch_mode<DREG, FULL, 2>(kb, gm, grid, st);
  launch_mode<DREG, FULL, 1>(kb, gm, grid, st);
}

}  // namespace

// The caller has checked the x2 conditions (scale 0.5 both ways, D <= 16, int32-sized offsets); grad_costs is written
// once per element, so it needs no zero fill.
void launch_softargmin_x2_backward(const SoftBwdKP& kb, cudaStream_t st) {
  const SoftKP& kp = kb.f;
  dim3 grid((kp.Win + kSbCols - 1) / kSbCols, (kp.Hin + kSbRows - 1) / kSbRows, kp.B);
  if (kp.D == 16) launch_gnc<16, true>(kb, grid, st);
  else if (kp.D == 8) launch_gnc<8, true>(kb, grid, st);
  else if (kp.D < 8) launch_gnc<8, false>(kb, grid, st);
  else launch_gnc<16, false>(kb, grid, st);
}

}  // namespace mvsgi
