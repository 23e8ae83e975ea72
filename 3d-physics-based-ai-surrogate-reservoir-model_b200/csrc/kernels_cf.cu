// SRM_NUMERICS_CLOSED_FORM: the order-1 polyharmonic interpolant in closed form.
//
//   f(x) = sum_i w_i |x - c_i| + v0 x + v1          (polyhm_splines.py:138-146 with phi(r)=sqrt(r))
// is exactly piecewise linear between knots, so per interval k it is f0[k] + slope[k]*(x - x0[k]).
// Tables are derived on the host in fp64 from the fp32 (w, v) the caller solved.
#include <cmath>
#include <cstring>
#include <vector>
#include "srm_internal.cuh"

int srm_build_closed_form(SrmHandle* h, const SrmConfig* cfg) {
  const int n = cfg->n_knots, P = cfg->n_props;
  std::vector<SrmClosedForm> host(1);
  SrmClosedForm& T = host[0];
  std::memset(&T, 0, sizeof(T));
  // interval k: k=0 is x < c0 (anchor c0), k>=1 is [c_{k-1}, c_k) (anchor c_{k-1}), k=n is x >= c_{n-1}
  for (int k = 0; k <= n; ++k) {
    const double xa = (k == 0) ? (double)cfg->knots[0] : (double)cfg->knots[k - 1];
    T.x0[k] = (float)xa;
    for (int q = 0; q < P; ++q) {
      const float* w = cfg->spline_w + (size_t)q * n;
      const double v0 = cfg->spline_v[2 * q], v1 = cfg->spline_v[2 * q + 1];
      double f = v0 * xa + v1, sl = v0;
      for (int i = 0; i < n; ++i) {
        const double ci = cfg->knots[i];
        f += (double)w[i] * std::fabs(xa - ci);
        // slope inside interval k: knots with index < k are to the left (sign +), the rest to the right
        sl += (i < k) ? (double)w[i] : -(double)w[i];
      }
      T.f0[q][k] = (float)f;
      T.slope[q][k] = (float)sl;
    }
  }
  cudaError_t e = cudaMalloc((void**)&h->d_cf, sizeof(SrmClosedForm));
  if (e == cudaSuccess) e = cudaMemcpy(h->d_cf, &T, sizeof(SrmClosedForm), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { srm_set_error("closed-form table upload: %s", cudaGetErrorString(e)); return SRM_ERR_CUDA; }
  return SRM_OK;
}

int srm_launch_pvt_eval_cf(const SrmHandle*, int64_t, const float*, float*, float*, cudaStream_t) {
  srm_set_error("SRM_NUMERICS_CLOSED_FORM kernels are not built yet");
  return SRM_ERR_INVALID;
}
int srm_forward_cf(SrmHandle*, int32_t, int32_t, const float*, const int32_t*, const float*, const float*, const float*,
                   const float*, const float*, float*, float*, const SrmWs&, bool, cudaStream_t) {
  srm_set_error("SRM_NUMERICS_CLOSED_FORM kernels are not built yet");
  return SRM_ERR_INVALID;
}
int srm_backward_cf(SrmHandle*, int32_t, int32_t, const float*, const int32_t*, const float*, const float*, const float*,
                    const float*, const float*, const float*, float*, float*, float*, float*, const SrmWs&, cudaStream_t) {
  srm_set_error("SRM_NUMERICS_CLOSED_FORM kernels are not built yet");
  return SRM_ERR_INVALID;
}
