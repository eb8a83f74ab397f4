// Probe: which cuTensorMapEncodeTiled / cp.async.bulk.tensor configurations run on this GPU (build: nvcc -arch=sm_100a).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap tm, float* out, int nfloats, int c0, int c1, int c2, int c3) {
  extern __shared__ __align__(128) float tile[];
  __shared__ __align__(8) unsigned long long mbar;
  const unsigned mb = smem_u32(&mbar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(nfloats * 4) : "memory");
    if (RANK == 2)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(smem_u32(tile)), "l"((unsigned long long)&tm), "r"(mb), "r"(c0), "r"(c1) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                   ::"r"(smem_u32(tile)), "l"((unsigned long long)&tm), "r"(mb), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(mb) : "memory");
  for (int i = threadIdx.x; i < nfloats; i += blockDim.x) out[i] = tile[i];
}

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  printf("entry point: err %d query %d ptr %p\n", (int)e, (int)q, p);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  const int W = 320, H = 80, D = 16, B = 2;
  std::vector<float> h((size_t)W * H * D * B);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *out;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&out, 1 << 20);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  struct Case { int rank, bw, bh, bd, c0, c1; const char* name; };
  Case cases[] = {{2, 32, 4, 1, 0, 0, "2d 32x4 at 0,0"},       {2, 40, 5, 1, -4, -1, "2d 40x5 at -4,-1"},
                  {4, 40, 5, 16, -4, -1, "4d 40x5x16 at -4,-1"}, {4, 40, 5, 16, 284, 79, "4d 40x5x16 at 284,79"},
                  {4, 40, 5, 8, 28, 3, "4d 40x5x8 at 28,3"},    {4, 40, 5, 16, 316, 76, "4d 40x5x16 at 316,76"},
                  {2, 36, 5, 1, -1, 0, "2d 36x5 at -1,0 (unaligned column)"}};
  for (const Case& c : cases) {
    CUtensorMap tm;
    cuuint64_t gdim[4] = {W, H, D, B};
    cuuint64_t gstr[3] = {W * 4ull, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * D * 4};
    cuuint32_t box[4] = {(cuuint32_t)c.bw, (cuuint32_t)c.bh, (cuuint32_t)c.bd, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (c.rank == 2) { gdim[1] = (cuuint64_t)H * D * B; }
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, c.rank, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int nf = c.bw * c.bh * (c.rank == 4 ? c.bd : 1);
    cudaMemset(out, 0, nf * 4);
    if (r == CUDA_SUCCESS) {
      if (c.rank == 2) probe<2><<<1, 128, nf * 4>>>(tm, out, nf, c.c0, c.c1, 0, 0);
      else probe<4><<<1, 128, nf * 4>>>(tm, out, nf, c.c0, c.c1, 0, 1);
    }
    cudaError_t ke = cudaDeviceSynchronize();
    std::vector<float> o(nf);
    cudaMemcpy(o.data(), out, nf * 4, cudaMemcpyDeviceToHost);
    // expected element [d][row][col] = value at (b=1 for 4d, d, c1+row, c0+col)
    int bad = 0;
    for (int dd = 0; dd < (c.rank == 4 ? c.bd : 1); ++dd)
      for (int rr = 0; rr < c.bh; ++rr)
        for (int cc = 0; cc < c.bw; ++cc) {
          const int x = c.c0 + cc, y = c.c1 + rr;
          float exp = 0.f;
          if (c.rank == 4) { if (x >= 0 && x < W && y >= 0 && y < H) exp = (float)((((size_t)1 * D + dd) * H + y) * W + x); }
          else { if (x >= 0 && x < W && y >= 0) exp = (float)((size_t)y * W + x); }
          if (o[(dd * c.bh + rr) * c.bw + cc] != exp) ++bad;
        }
    printf("%-28s encode %d kernel %d (%s) mismatches %d\n", c.name, (int)r, (int)ke, cudaGetErrorString(ke), bad);
    if (ke != cudaSuccess) { printf("sticky error, stopping\n"); break; }
  }
  return 0;
}
