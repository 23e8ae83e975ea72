// Microbenchmark 3: the same random table gather through three request paths -- LDG (the L1TEX LSU path, one
// wavefront per distinct 128-byte line of a warp request), texture fetches on linear memory (tex1Dfetch: the TEX path
// of the same unit, quad based) and LDG with 32 lanes sharing lines -- to see whether the exact-table gathers of the
// reference-order kernels can leave the 1 lane/cycle/SM rate tools/gather_probe2.cu measures.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/gather_probe3 tools/gather_probe3.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ float sum(float v) { return v; }
__device__ __forceinline__ float sum(float2 v) { return v.x + v.y; }
__device__ __forceinline__ float sum(float4 v) { return v.x + v.y + v.z + v.w; }

template <class T, int U>
__global__ void __launch_bounds__(256) k_ldg(const T* __restrict__ tab, const unsigned* __restrict__ idx, long n, float* __restrict__ out) {
  float acc = 0.f;
  for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g * U < n; g += (long)gridDim.x * blockDim.x) {
    unsigned e[U];
#pragma unroll
    for (int u = 0; u < U; ++u) e[u] = idx[g + (long)u * (n / U)];
    T v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldg(tab + e[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) acc += sum(v[u]);
  }
  if (acc == 123.456f) out[0] = acc;
}
template <class T, int U>
__global__ void __launch_bounds__(256) k_tex(cudaTextureObject_t tex, const unsigned* __restrict__ idx, long n, float* __restrict__ out) {
  float acc = 0.f;
  for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g * U < n; g += (long)gridDim.x * blockDim.x) {
    unsigned e[U];
#pragma unroll
    for (int u = 0; u < U; ++u) e[u] = idx[g + (long)u * (n / U)];
    T v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = tex1Dfetch<T>(tex, (int)e[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) acc += sum(v[u]);
  }
  if (acc == 123.456f) out[0] = acc;
}

template <class F>
void timeit(const char* name, long n, F launch) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 2; ++i) launch();
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) launch();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
  printf("%-34s %.3f ms  %.2f SM-cycles per gather-lane (%.1f G gathers/s) %s\n", name, ms, ms * 1e-3 * 1.965e9 * 148 / (double)n,
         n / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

template <class T>
cudaTextureObject_t make_tex(const void* p, size_t bytes) {
  cudaResourceDesc rd{}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = const_cast<void*>(p);
  rd.res.linear.desc = cudaCreateChannelDesc<T>(); rd.res.linear.sizeInBytes = bytes;
  cudaTextureDesc td{}; td.readMode = cudaReadModeElementType; td.filterMode = cudaFilterModePoint; td.addressMode[0] = cudaAddressModeClamp;
  cudaTextureObject_t t = 0;
  cudaError_t e = cudaCreateTextureObject(&t, &rd, &td, nullptr);
  if (e != cudaSuccess) printf("cudaCreateTextureObject: %s\n", cudaGetErrorString(e));
  return t;
}

int main(int argc, char** argv) {
  const long n = 64l << 20;
  const long window = argc > 1 ? atol(argv[1]) : 1800000l;
  const int local = argc > 2 ? atoi(argv[2]) : 0;      // > 0: lane l of a warp draws its index within `local` entries of the warp's base
  char* tab; unsigned* idx; float* out;
  cudaMalloc(&tab, window * 16); cudaMemset(tab, 0, window * 16);
  cudaMalloc(&idx, n * sizeof(unsigned)); cudaMalloc(&out, 4);
  std::vector<unsigned> h(n);
  unsigned long long s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
  unsigned base = 0;
  for (long i = 0; i < n; ++i) {
    if (local > 0) { if ((i & 31) == 0) base = (unsigned)(rnd() % (unsigned long long)(window - local)); h[i] = base + (unsigned)(rnd() % (unsigned long long)local); }
    else h[i] = (unsigned)(rnd() % (unsigned long long)window);
  }
  cudaMemcpy(idx, h.data(), n * sizeof(unsigned), cudaMemcpyHostToDevice);
  int maxlin = 0; cudaDeviceGetAttribute(&maxlin, cudaDevAttrMaxTexture1DLinearWidth, 0);
  printf("window %ld entries, %ld gathers per launch, lane locality %d entries, maxTexture1DLinear %d\n", window, n, local, maxlin);
  cudaTextureObject_t t1 = make_tex<float>(tab, window * 4), t2 = make_tex<float2>(tab, window * 8), t4 = make_tex<float4>(tab, window * 16);
  for (int occ : {4, 8}) {
    const int blocks = 148 * occ;
    printf("-- %d warps per SM\n", occ * 8);
    timeit("LDG.32  U4", n, [&] { k_ldg<float, 4><<<blocks, 256>>>((const float*)tab, idx, n, out); });
    timeit("LDG.64  U4", n, [&] { k_ldg<float2, 4><<<blocks, 256>>>((const float2*)tab, idx, n, out); });
    timeit("LDG.128 U4", n, [&] { k_ldg<float4, 4><<<blocks, 256>>>((const float4*)tab, idx, n, out); });
    timeit("TEX float  U4", n, [&] { k_tex<float, 4><<<blocks, 256>>>(t1, idx, n, out); });
    timeit("TEX float2 U4", n, [&] { k_tex<float2, 4><<<blocks, 256>>>(t2, idx, n, out); });
    timeit("TEX float4 U4", n, [&] { k_tex<float4, 4><<<blocks, 256>>>(t4, idx, n, out); });
    timeit("TEX float2 U8", n, [&] { k_tex<float2, 8><<<blocks, 256>>>(t2, idx, n, out); });
    timeit("TEX float4 U8", n, [&] { k_tex<float4, 8><<<blocks, 256>>>(t4, idx, n, out); });
  }
  return 0;
}
