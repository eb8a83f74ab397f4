// softargmin_bwd.cu — launch of the gather-form soft-argmin backward (softargmin_bwd.cuh).
#include "launch.h"
#include "softargmin_bwd.cuh"

namespace mvsgi {
namespace {

template <int DREG, bool FULL>
void launch_gnc(const SoftBwdKP& kb, dim3 grid, cudaStream_t st) {
  if (kb.gNc != nullptr) softargmin_x2_backward_kernel<DREG, FULL, true><<<grid, kSbThreads, 0, st>>>(kb);
  else softargmin_x2_backward_kernel<DREG, FULL, false><<<grid, kSbThreads, 0, st>>>(kb);
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
