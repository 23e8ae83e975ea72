// Helpers shared by the reference-order and closed-form kernel families.
#pragma once
#include "srm_internal.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// block reduction of NV doubles; result valid in thread 0
// ------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double* smem /* [NV*32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_down_sync(0xffffffffu, v[q], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) smem[q * 32 + warp] = v[q];
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      double x = (lane < nw) ? smem[q * 32 + lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      v[q] = x;
    }
  }
}


// a / b, IEEE.  div.rn sends a == 0 to its slow path (a subroutine call with some thirty instructions); the
// truncation brackets vanish identically (physics_loss.py:171, 419-441) and saturation differences are often exactly
// zero, so the zero numerator is answered first: (+-0) / b = +-0 with the sign product, for any finite non-zero b.
__device__ __forceinline__ float div_z(float a, float b) {
  const float ab = fabsf(b);
  if (a == 0.f && ab > 0.f && ab <= 3.402823466e38f) return __fmul_rn(a, copysignf(1.0f, b));
  return __fdiv_rn(a, b);
}

// first well (sorted by cell) with cell >= c: binary search inside the cell's layer
__device__ __forceinline__ int well_lower_bound(const SrmDev& P, int c) {
  const int k = c / (P.H * P.W);
  int lo = P.layer_ptr[k], hi = P.layer_ptr[k + 1];
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (P.wells[mid].cell < c) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// mbc_b = (-sum q) - sum mb_cells ; SSE_mbc ; terms/counts        physics_loss.py:193,800-832
__global__ void k_finalize_fwd(const __grid_constant__ SrmDev P, int32_t B, double* __restrict__ sse,
                               const double* __restrict__ mb_sum, const double* __restrict__ q_sum,
                               float* __restrict__ mbc, float* __restrict__ terms_out) {
  __shared__ double red[32];
  double v[1] = {0.0};
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float m = __fsub_rn(-(float)q_sum[b], (float)mb_sum[b]);
    mbc[b] = m;
    v[0] += (double)m * (double)m;
  }
  block_reduce<1>(v, red);
  if (threadIdx.x == 0) {
    sse[SRM_TERM_MBC] = v[0];
    const double n = (double)B * (double)P.N;
    for (int t = 0; t < SRM_N_TERMS; ++t) {
      terms_out[t] = (float)sse[t];
      double cnt = 0.0;
      if (t == SRM_TERM_DOM || t == SRM_TERM_IBC || t == SRM_TERM_TDE) cnt = n;
      if (t == SRM_TERM_MBC) cnt = n;      // the reference counts mbc with the ic field's shape (physics_loss.py:830)
      terms_out[SRM_N_TERMS + t] = (float)cnt;
    }
  }
}

__global__ void k_finalize_adj(int32_t B, const double* __restrict__ a1, const double* __restrict__ a2,
                               float* __restrict__ gdt1, float* __restrict__ gdt2) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    gdt1[b] = (float)a1[b];
    gdt2[b] = (float)a2[b];
  }
}


}  // namespace
