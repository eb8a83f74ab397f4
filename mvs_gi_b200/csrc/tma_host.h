// tma_host.h — cuTensorMapEncodeTiled through the runtime's driver entry point table (no link-time dependency on
// libcuda), shared by the translation units that build tensor maps.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace mvsgi {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// float32 tensor [n3][n2][n1][n0] with unit innermost stride and byte pitches p1..p3, box {b0, b1, b2, 1}, zero fill
// outside the tensor; false when the driver refuses the description (misaligned base or pitches, ...)
inline bool make_map_f32_4d(CUtensorMap* tm, const float* base, cuuint64_t n0, cuuint64_t n1, cuuint64_t n2, cuuint64_t n3,
                            cuuint64_t p1, cuuint64_t p2, cuuint64_t p3, cuuint32_t b0, cuuint32_t b1, cuuint32_t b2) {
  EncodeTiledFn enc = encode_tiled();
  if (enc == nullptr) return false;
  const cuuint64_t gdim[4] = {n0, n1, n2, n3};
  const cuuint64_t gstr[3] = {p1, p2, p3};
  const cuuint32_t box[4] = {b0, b1, b2, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace mvsgi
