// standalone probe of the 4-D TMA box load used by kernels_cf.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap map, float* out, int n, int x, int y, int z, int b) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* dst = (float*)smem;
  uint64_t* bar = (uint64_t*)(smem + 4096);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(&map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z), "r"(b) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = dst[i];
}
int main(int argc, char** argv) {
  int W = 64, H = 32, D = 16, B = 2, bx = argc > 1 ? atoi(argv[1]) : 36, by = argc > 2 ? atoi(argv[2]) : 10;
  size_t N = (size_t)W * H * D * B;
  std::vector<float> h(N);
  for (size_t i = 0; i < N; ++i) h[i] = (float)i;
  float *d, *o;
  cudaMalloc(&d, N * 4); cudaMalloc(&o, 4096);
  cudaMemcpy(d, h.data(), N * 4, cudaMemcpyHostToDevice);
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap m;
  cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * D * 4};
  cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)by, 1, 1}, es[4] = {1, 1, 1, 1};
  CUresult r = ((Fn)fnp)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d q=%d\n", (int)r, (int)q);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192);
  probe<<<1, 128, 8192>>>(m, o, bx * by, argc > 3 ? atoi(argv[3]) : -2, argc > 4 ? atoi(argv[4]) : -1, 3, 1);
  cudaError_t e = cudaDeviceSynchronize();
  printf("sync: %s\n", cudaGetErrorString(e));
  std::vector<float> res(bx * by);
  cudaMemcpy(res.data(), o, bx * by * 4, cudaMemcpyDeviceToHost);
  for (int yy = 0; yy < 3; ++yy) { for (int xx = 0; xx < 6; ++xx) printf("%10.0f ", res[yy * bx + xx]); printf("\n"); }
  printf("expect row1: 0 0 %d %d ...\n", (1 * D + 3) * H * W, (1 * D + 3) * H * W + 1);
  return 0;
}
