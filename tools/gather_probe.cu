// Microbenchmark: cost of the exact-table gathers (one 8/16-byte entry per lane from an L2-resident table).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/gather_probe tools/gather_probe.cu
// Variants: load flavour (ldg / cg / no_allocate), split of one warp-wide gather into S predicated
// sub-gathers, entries per thread in flight.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

template <int MODE> __device__ __forceinline__ float2 load2(const float4* t, unsigned e) {
  const float2* p = reinterpret_cast<const float2*>(t + e);
  if (MODE == 0) return __ldg(p);
  if (MODE == 1) return __ldcg(p);
  if (MODE == 2) { float2 v; asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p)); return v; }
  if (MODE == 3) return __ldcs(p);
  return *p;
}

// each thread: U independent gathers per iteration; indices from a field (coalesced read), like the stencil kernels
template <int MODE, int SPLIT, int U>
__global__ void __launch_bounds__(256) k_gather(const float4* __restrict__ tab, const unsigned* __restrict__ idx, long n, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  float acc = 0.f;
  for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g * U < n; g += (long)gridDim.x * blockDim.x) {
    unsigned e[U];
    if (U == 4) { uint4 t = *reinterpret_cast<const uint4*>(idx + g * 4); e[0] = t.x; e[1] = t.y; e[2] = t.z; e[3] = t.w; }
    else { for (int u = 0; u < U; ++u) e[u] = idx[g * U + u]; }
    float2 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (SPLIT == 1) v[u] = load2<MODE>(tab, e[u]);
      else {
        v[u] = make_float2(0.f, 0.f);
#pragma unroll
        for (int s = 0; s < SPLIT; ++s)
          if ((lane % SPLIT) == s) v[u] = load2<MODE>(tab, e[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y;
  }
  if (acc == 123.456f) out[0] = acc;
}

template <int MODE, int SPLIT, int U>
float run(const float4* tab, const unsigned* idx, long n, float* out, const char* name) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int blocks = 148 * 8;
  for (int i = 0; i < 2; ++i) k_gather<MODE, SPLIT, U><<<blocks, 256>>>(tab, idx, n, out);
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) k_gather<MODE, SPLIT, U><<<blocks, 256>>>(tab, idx, n, out);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
  const double per = ms * 1e-3 * 1.965e9 * 148 / (double)n;
  printf("%-28s mode %d split %2d U %d : %.3f ms  %.2f SM-cycles per gather-lane (%.1f G gathers/s) %s\n", name, MODE, SPLIT, U, ms, per, n / ms * 1e-6,
         cudaGetErrorString(cudaGetLastError()));
  return ms;
}

int main(int argc, char** argv) {
  const long n = 64l << 20;                       // gathers per launch
  const long window = argc > 1 ? atol(argv[1]) : (1800000l);   // distinct entries (16 B each): 1.8 M = operating window
  float4* tab; unsigned* idx; float* out;
  cudaMalloc(&tab, window * sizeof(float4)); cudaMemset(tab, 0, window * sizeof(float4));
  cudaMalloc(&idx, n * sizeof(unsigned)); cudaMalloc(&out, 4);
  std::vector<unsigned> h(n);
  unsigned long long s = 88172645463325252ull;
  for (long i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (unsigned)(s % (unsigned long long)window); }
  cudaMemcpy(idx, h.data(), n * sizeof(unsigned), cudaMemcpyHostToDevice);
  printf("window %ld entries (%.1f MB), %ld gathers\n", window, window * 16e-6, n);
  run<0, 1, 1>(tab, idx, n, out, "ldg");
  run<0, 1, 4>(tab, idx, n, out, "ldg");
  run<1, 1, 4>(tab, idx, n, out, "ldcg");
  run<2, 1, 4>(tab, idx, n, out, "nc.no_allocate");
  run<3, 1, 4>(tab, idx, n, out, "ldcs");
  run<0, 2, 4>(tab, idx, n, out, "ldg split");
  run<0, 4, 4>(tab, idx, n, out, "ldg split");
  run<0, 8, 4>(tab, idx, n, out, "ldg split");
  run<1, 4, 4>(tab, idx, n, out, "ldcg split");
  run<2, 4, 4>(tab, idx, n, out, "no_alloc split");
  return 0;
}
