// Small utility kernels: permeability de-normalisation, well-table reordering / dense scatter.
#include "pvt_ref.cuh"

namespace {

// DataSummary.nonormalize, log branch (data_processing/data_processing_utils.py:1098-1106):
//   exp( log(max/min) * ((x - lo)/(hi - lo)) + log(min) )
__global__ void __launch_bounds__(256) k_denorm_log(int64_t n, const float* __restrict__ x, float lr, float lmin,
                                                    float lo, float span, float* __restrict__ out) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const float u = __fdiv_rn(__fsub_rn(x[g], lo), span);
  float y = expf(__fadd_rn(__fmul_rn(lr, u), lmin));
  if (isnan(y) || isinf(y)) y = 0.f;   // NaN/Inf -> 0 (data_processing_utils.py:1122-1125)
  out[g] = y;
}

// sorted-by-cell well table -> caller's well order
__global__ void k_unsort_wells(const WellDev* __restrict__ wells, int nw, int64_t total,
                               const float* __restrict__ sorted, float* __restrict__ out) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  const int64_t b = g / nw;
  const int w = (int)(g % nw);
  out[b * nw + wells[w].orig] = sorted[g];
}

// tf.scatter_nd of per-connection values onto the (zeroed) grid; duplicates sum
// (welldata_processor.py:205-223).
__global__ void k_scatter_wells(const WellDev* __restrict__ wells, int nw, int64_t N, int64_t total,
                                const float* __restrict__ sorted, float* __restrict__ dense) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  const int64_t b = g / nw;
  const int w = (int)(g % nw);
  atomicAdd(&dense[b * N + wells[w].cell], sorted[g]);
}

// compares the shared-rsqrt sqrt/div sequences of pvt_ref.cuh against the IEEE intrinsics
__global__ void __launch_bounds__(256) k_selftest_rounding(int64_t n, uint64_t seed, unsigned long long* __restrict__ bad) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  // splitmix64 -> two log-uniform operands
  auto mix = [](uint64_t z) { z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
                              z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); };
  const uint64_t h1 = mix(seed + 2 * (uint64_t)g), h2 = mix(seed + 2 * (uint64_t)g + 1);
  // rs in [1e-10, 1e9]: exponent uniform, mantissa uniform
  const float u1 = (float)(h1 >> 40) * (1.0f / 16777216.0f), u2 = (float)(h2 >> 40) * (1.0f / 16777216.0f);
  float rs = exp2f(-33.2f + 63.1f * u1) * (1.0f + (float)(h1 & 0x7fffff) * (1.0f / 8388608.0f));
  rs = fmaxf(rs, SRM_EPS);
  float a = exp2f(-40.0f + 50.0f * u2) * (1.0f + (float)(h2 & 0x7fffff) * (1.0f / 8388608.0f));
  if (h2 & (1ull << 39)) a = -a;
  const SrmSqrtRcp sr = srm_sqrt_rcp(rs);
  const float s_ref = __fsqrt_rn(rs);
  const float q = srm_div_by(a, sr.s, sr.y);
  const float q_ref = __fdiv_rn(a, s_ref);
  const float q2 = srm_div_by(q, sr.s, sr.y);          // chained, as in the second-derivative path
  const float q2_ref = __fdiv_rn(q_ref, s_ref);
  if (__float_as_uint(sr.s) != __float_as_uint(s_ref)) atomicAdd(&bad[0], 1ull);
  if (__float_as_uint(q) != __float_as_uint(q_ref)) atomicAdd(&bad[1], 1ull);
  if (__float_as_uint(q2) != __float_as_uint(q2_ref)) atomicAdd(&bad[2], 1ull);
}

}  // namespace

int srm_launch_selftest_rounding(int64_t n, uint64_t seed, int64_t* bad_host, cudaStream_t s) {
  unsigned long long* d = nullptr;
  SRM_CUDA_CHECK(cudaMalloc((void**)&d, 3 * sizeof(unsigned long long)));
  SRM_CUDA_CHECK(cudaMemsetAsync(d, 0, 3 * sizeof(unsigned long long), s));
  if (n > 0) k_selftest_rounding<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, seed, d);
  cudaError_t e = cudaGetLastError();
  unsigned long long h[3] = {0, 0, 0};
  if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(d);
  if (e != cudaSuccess) { srm_set_error("srm_selftest_rounding: %s", cudaGetErrorString(e)); return SRM_ERR_CUDA; }
  for (int i = 0; i < 3; ++i) bad_host[i] = (int64_t)h[i];
  return SRM_OK;
}

int srm_launch_denorm_log(int64_t n, const float* x, float kmin, float kmax, float lo, float hi, float* out, cudaStream_t s) {
  if (n == 0) return SRM_OK;
  const float lr = logf(kmax / kmin), lmin = logf(kmin);
  k_denorm_log<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, x, lr, lmin, lo, hi - lo, out);
  SRM_CUDA_CHECK(cudaGetLastError());
  return SRM_OK;
}

int srm_launch_unsort_wells(const SrmHandle* h, int32_t B, const float* sorted, float* out, cudaStream_t s) {
  const int64_t total = (int64_t)B * h->dev.n_wells;
  if (total == 0) return SRM_OK;
  k_unsort_wells<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(h->d_wells, h->dev.n_wells, total, sorted, out);
  SRM_CUDA_CHECK(cudaGetLastError());
  return SRM_OK;
}

int srm_launch_scatter_wells(const SrmHandle* h, int32_t B, const float* sorted, float* dense, cudaStream_t s) {
  const int64_t total = (int64_t)B * h->dev.n_wells;
  if (total == 0) return SRM_OK;
  k_scatter_wells<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(h->d_wells, h->dev.n_wells, h->dev.N, total, sorted, dense);
  SRM_CUDA_CHECK(cudaGetLastError());
  return SRM_OK;
}
