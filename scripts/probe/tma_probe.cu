// Stand-alone probe: FP64 TMA tile load with halo / OOB fill (debug aid, not part of the library).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap tm, const CUtensorMap* tmg, int useg, int c0, int c1, int c2,
                  int n, double* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  double* buf = (double*)sm;
  uint64_t* bar = (uint64_t*)(sm + ((n * 8 + 127) / 128) * 128);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const CUtensorMap* p = useg ? tmg : &tm;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(n * 8) : "memory");
    if (RANK == 4)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                   ::"r"(s32(buf)), "l"(p), "r"(c0), "r"(c1), "r"(c2), "r"(0), "r"(s32(bar)) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(s32(buf)), "l"(p), "r"(c0), "r"(c1), "r"(c2), "r"(s32(bar)) : "memory");
  }
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(s32(bar)), "r"(0) : "memory");
  } while (!ok);
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = buf[i];
}
int main(int argc, char** argv) {
  int rank = atoi(argv[1]), c0 = atoi(argv[2]), c1 = atoi(argv[3]), useg = atoi(argv[4]);
  int nn0 = 41, nn1 = 13, nz = 13, PX = 44, PY = 14, bx = 44, by = 18;
  if (argc > 5) bx = atoi(argv[5]);
  size_t tot = (size_t)PX * PY * nz;
  std::vector<double> h(tot);
  for (size_t i = 0; i < tot; ++i) h[i] = (double)i;
  double *d, *out;
  cudaMalloc(&d, tot * 8); cudaMalloc(&out, bx * by * 8);
  cudaMemcpy(d, h.data(), tot * 8, cudaMemcpyHostToDevice);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  PFN enc = (PFN)fp;
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)nn0, (cuuint64_t)nn1, (cuuint64_t)nz, 1};
  cuuint64_t str[3] = {(cuuint64_t)PX * 8, (cuuint64_t)PX * PY * 8, (cuuint64_t)tot * 8};
  cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)by, 1, 1}, es[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, rank, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("rank %d c0 %d c1 %d useg %d bx %d: encode rc=%d; ", rank, c0, c1, useg, bx, (int)r);
  if (r) { printf("\n"); return 0; }
  CUtensorMap* tmg; cudaMalloc(&tmg, sizeof(tm)); cudaMemcpy(tmg, &tm, sizeof(tm), cudaMemcpyHostToDevice);
  int n = bx * by;
  size_t smem = ((n * 8 + 127) / 128) * 128 + 64;
  if (rank == 4) k<4><<<1, 128, smem>>>(tm, tmg, useg, c0, c1, 3, n, out);
  else k<3><<<1, 128, smem>>>(tm, tmg, useg, c0, c1, 3, n, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s; ", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<double> o(n);
    cudaMemcpy(o.data(), out, n * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < by; ++y) for (int x = 0; x < bx; ++x) {
      int gx = c0 + x, gy = c1 + y;
      double ex = (gx < 0 || gx >= nn0 || gy < 0 || gy >= nn1) ? 0.0 : (double)((size_t)3 * PX * PY + (size_t)gy * PX + gx);
      if (o[y * bx + x] != ex) ++bad;
    }
    printf("mismatches %d of %d", bad, n);
  }
  printf("\n");
  return 0;
}
