// Small utility kernels: permeability de-normalisation, well-table reordering / dense scatter.
#include "srm_internal.cuh"

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

}  // namespace

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
