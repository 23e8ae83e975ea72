// WellRatesPressure.compute_rates_and_bhp evaluated sparsely at the connection cells, templated on
// the gas-mobility functor (reference-order spline or closed form).
#pragma once
#include "srm_internal.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// wells: forward-mode dual numbers carry d/dp of the connection-cell pressure through the
// min/max/clip/divide_no_nan chain with TensorFlow's gradient conventions.
// ------------------------------------------------------------------------------------------
struct Dual { float v, d; };
__device__ __forceinline__ Dual dmk(float v, float d = 0.f) { Dual r; r.v = v; r.d = d; return r; }
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return dmk(__fadd_rn(a.v, b.v), a.d + b.d); }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return dmk(__fsub_rn(a.v, b.v), a.d - b.d); }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return dmk(__fmul_rn(a.v, b.v), a.d * b.v + a.v * b.d); }
__device__ __forceinline__ Dual operator/(Dual a, Dual b) {
  const float q = __fdiv_rn(a.v, b.v);
  return dmk(q, (a.d - q * b.d) / b.v);
}
// tf.math.divide_no_nan
__device__ __forceinline__ Dual ddnn(Dual a, Dual b) { return (b.v == 0.f) ? dmk(0.f, 0.f) : a / b; }
// tf.minimum / tf.maximum: ties route the gradient to the first argument
__device__ __forceinline__ Dual dmin(Dual a, Dual b) { return (a.v <= b.v) ? a : b; }
__device__ __forceinline__ Dual dmax(Dual a, Dual b) { return (a.v >= b.v) ? a : b; }
// tf.clip_by_value(t, lo, hi)
// value = max(min(t,hi),lo) (the kernel's cwiseMin/cwiseMax); gradient per _ClipByValueGrad:
// to t where lo <= t <= hi, to lo where t < lo, to hi where t > hi
__device__ __forceinline__ Dual dclip(Dual t, Dual lo, Dual hi) {
  const bool below = t.v < lo.v, above = t.v > hi.v;
  return dmk(fmaxf(fminf(t.v, hi.v), lo.v), ((!below && !above) ? t.d : 0.f) + (below ? lo.d : 0.f) + (above ? hi.d : 0.f));
}

// compute_blocking_integral_and_factor, DG branch             well_rate_bhp_Subclassed.py:840-960
template <class Mob>
__device__ Dual blocking_integral(const SrmDev& P, const Mob& mob, Dual p, Dual pwf, Dual mg_n1) {
  const int n = P.n_int;
  const Dual delta = (pwf - p) / dmk((float)n);          // tf.linspace: delta = (stop-start)/n
  Dual sum = dmk(0.f), mg_prev = mg_n1, pa = p;
  for (int i = 0; i < n; ++i) {
    const Dual pb = (i + 1 < n) ? p + delta * dmk((float)(i + 1)) : pwf;   // ends are exact
    const Dual mg1 = mob(P, pb);                  // Sg1 = Sg_max -> same krg (:912)
    const Dual dp = pa - pb;
    sum = sum + dmk(0.5f) * (mg_prev + mg1) * dp;          // :920
    mg_prev = mg1;
    pa = pb;
  }
  return sum;
}

template <class Mob>
__global__ void __launch_bounds__(128) k_wells(const __grid_constant__ SrmDev P, const Mob mob, int32_t B, int32_t R,
                                                   const float* __restrict__ kx, const int32_t* __restrict__ sample_real,
                                                   const float* __restrict__ pfield, const float* __restrict__ t_days,
                                                   float* __restrict__ qw, float* __restrict__ pwfw,
                                                   float* __restrict__ dqdp) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int nw = P.n_wells;
  if (g >= (int64_t)B * nw) return;
  const int b = (int)(g / nw), w = (int)(g % nw);
  const WellDev wd = P.wells[w];
  const int r = srm_real_of(sample_real, b, B, R);
  const float k = kx[(int64_t)r * P.N + wd.cell];
  const Dual p = dmk(pfield[(int64_t)b * P.N + wd.cell], 1.0f);
  // shut-in mask: 1 unless shut_start <= t <= shut_stop          welldata_processor.py:349-354
  const float t = t_days[b];
  const float open = (t >= wd.shut_start && t <= wd.shut_stop) ? 0.f : 1.f;
  // Peaceman                                                  well_rate_bhp_Subclassed.py:782-788
  const float ky = __fmul_rn(P.kx_ky, k);
  const float ryx = __fdiv_rn(ky, k), rxy = __fdiv_rn(k, ky);
  const float num = sqrtf(__fadd_rn(__fmul_rn(sqrtf(ryx), __fmul_rn(P.dx, P.dx)),
                                    __fmul_rn(sqrtf(rxy), __fmul_rn(P.dy, P.dy))));
  const float den = __fadd_rn(powf(ryx, 0.25f), powf(rxy, 0.25f));
  const float ro = __fdiv_rn(__fmul_rn(0.28f, num), den);
  const float two_pi = 6.283185307179586f;
  float ck = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(two_pi, wd.hc), k), P.dz), P.C);
  ck = __fdiv_rn(ck, logf(__fdiv_rn(ro, wd.rw)));
  const Dual Ck = dmk(__fmul_rn(open, ck));
  const Dual mg = mob(P, p);
  const Dual pmin = dmk(wd.pwf_min), qt = dmk(wd.q_target), zero = dmk(0.f), tiny = dmk(1e-12f);
  // _compute_phase_rates                                       :963-1007
  auto phase_rates = [&](Dual pwf_) {
    Dual ig = dmk(1.f);
    if (P.use_blk) ig = blocking_integral(P, mob, p, pwf_, mg);
    const Dual dp = (p - pwf_) + tiny;                                         // :987
    const Dual blk = P.use_blk ? ddnn(ig, mg * dp) : ig;                       // :991
    const Dual qg_max2 = Ck * blk * mg * dp;                                   // :997
    return dmax(dmin(qt, qg_max2), zero);                                      // :1001
  };
  Dual pwf;
  if (!P.bhp_iterative) {
    // ---- _non_iterative_method                               :614-724
    Dual ig_max = dmk(1.f);
    if (P.use_blk) ig_max = blocking_integral(P, mob, p, pmin, mg);
    const Dual dp_max = (p - pmin) + tiny;                                     // :650
    const Dual blk_max = P.use_blk ? ddnn(ig_max, mg * dp_max) : ig_max;       // :654-657
    const Dual ckb = Ck * blk_max;                                             // well_id == 1
    const Dual qg_max = ckb * mg * dp_max;                                     // :662
    const Dual qg_opt = dmax(dmin(qt, qg_max), zero);                          // :666
    const Dual lam = dclip(ddnn(qg_opt, ckb * mg), zero, blk_max);             // :699
    const Dual dp_opt = lam * dp_max;                                          // :721
    pwf = dclip(p - dp_opt, pmin, p);                                          // :723
  } else {
    // ---- _iterative_method                                   :515-612
    // Newton-Raphson on the bottom-hole pressure, d/dp carried through every step (tf.while_loop's gradient).  The
    // reference iterates the whole batch while ANY connection misses its target; a connection with |qg - q_target| <= tol
    // has qg == q_target (the minimum picked the target: zero gradient), so its step is the identity in value and
    // derivative -- stopping it on its own error gives the same numbers.
    const Dual eps = dmk(14.7f);                                               // :540
    pwf = pmin + dmk(0.5f) * (p - pmin);                                       // :537
    for (int it = 0; it < P.bhp_max_iters; ++it) {
      const Dual qg_it = phase_rates(pwf);                                     // :566-569 (and the cond's evaluation, :549-553)
      if (!(fabsf(__fsub_rn(qg_it.v, qt.v)) > P.bhp_tol)) break;
      const Dual qg_plus = phase_rates(pwf + eps);                             // :571-574
      const Dual dq = (qg_plus - qg_it) / eps;                                 // :576
      const Dual pwf_new = pwf - (qg_it - qt) / (dq + tiny);                   // :584
      pwf = dclip(pwf_new, pmin, p);                                           // :586
    }
  }
  const Dual qg = phase_rates(pwf);
  qw[g] = qg.v;
  pwfw[g] = pwf.v;
  dqdp[g] = qg.d;
}


}  // namespace
