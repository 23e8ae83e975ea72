// The element-wise glue either side of the physics kernels, both time levels in one pass:
//   HardLayer.call                      Hard_Layer_Subclassed.py:219-242 (via CompleteTrainableModule.call,
//                                       complete_trainable_module.py:142-176)
//   per-sample mean of the dt field     physics_loss.py:102,122
// and their cotangents (what tape.gradient delivers to the networks' outputs and to the layer's kernel_exponent).
//
//   p_l[b,c]  = init_value - alpha_t(tn_l[b]) ^ expo[c] * y_l[b,c],   alpha_t(t) = (t - t_lo) / (t_hi - t_lo)
//   dt_l[b]   = mean_c dtf_l[b,c]
//
// HBM bound: forward reads y0, y1, dtf1, dtf2 and writes p0, p1 (24 B per cell-timestep); backward reads gp0, gp1,
// y0, y1 and writes gy0, gy1, gdtf1, gdtf2 (32 B).  alpha = 2^(e * log2 alpha_t): log2 alpha_t once per sample in
// fp64, the product split into an fp32 head for ex2 and a first-order tail (~2 ulp against pow()).
#include <math_constants.h>
#include <algorithm>
#include "common.cuh"

namespace {

constexpr int GT = 256;

struct GlueLevel { double l2; float lnat; float at; };      // log2(alpha_t), ln(alpha_t) (0 where alpha_t <= 0), alpha_t

__device__ __forceinline__ GlueLevel glue_level(float tn, float t_lo, float t_hi) {
  GlueLevel g;
  g.at = __fdiv_rn(__fsub_rn(tn, t_lo), __fsub_rn(t_hi, t_lo));
  g.l2 = log2((double)g.at);
  g.lnat = (g.at > 0.f) ? (float)log((double)g.at) : 0.f;     // tf pow gradient: log of the safe base
  return g;
}
// d(alpha_t ^ e) / d alpha_t = e * alpha_t^(e-1), as tf.pow's gradient forms it (a = alpha_t ^ e)
__device__ __forceinline__ float glue_dpow(const GlueLevel& g, float a, float e) {
  if (g.at > 0.f) return __fdiv_rn(e * a, g.at);
  if (g.at == 0.f) return (e > 1.f) ? 0.f : ((e == 1.f) ? 1.f : CUDART_INF_F);
  return e * powf(g.at, e - 1.f);
}
// alpha_t ^ e
__device__ __forceinline__ float glue_pow(const GlueLevel& g, float e) {
  if (g.at == 0.f) return (e > 0.f) ? 0.f : ((e == 0.f) ? 1.f : CUDART_INF_F);
  const double pr = (double)e * g.l2;
  const float hi = (float)pr;
  const float lo = (float)(pr - (double)hi);
  return exp2f(hi) * fmaf(lo, 0.69314718f, 1.0f);
}

__global__ void __launch_bounds__(GT) k_glue_fwd(int64_t N, float init_value, float t_lo, float t_hi,
                                                 const float* __restrict__ expo, const float* __restrict__ tn0,
                                                 const float* __restrict__ tn1, const float* __restrict__ y0,
                                                 const float* __restrict__ y1, const float* __restrict__ dtf1,
                                                 const float* __restrict__ dtf2, float* __restrict__ p0,
                                                 float* __restrict__ p1, double* __restrict__ dsum) {
  __shared__ double red[2 * 32];
  const int b = blockIdx.y;
  const GlueLevel g0 = glue_level(tn0[b], t_lo, t_hi), g1 = glue_level(tn1[b], t_lo, t_hi);
  const int64_t base = (int64_t)b * N;
  float s1 = 0.f, s2 = 0.f;
  const bool al16 = ((reinterpret_cast<uintptr_t>(expo) | reinterpret_cast<uintptr_t>(y0) | reinterpret_cast<uintptr_t>(y1) | reinterpret_cast<uintptr_t>(p0) |
                      reinterpret_cast<uintptr_t>(p1) | reinterpret_cast<uintptr_t>(dtf1) | reinterpret_cast<uintptr_t>(dtf2)) & 15u) == 0;
  for (int64_t c = ((int64_t)blockIdx.x * GT + threadIdx.x) * 4; c < N; c += (int64_t)gridDim.x * GT * 4) {
    if (c + 3 < N && (N & 3) == 0 && al16) {
      const float4 e = expo ? __ldg(reinterpret_cast<const float4*>(expo + c)) : make_float4(1.f, 1.f, 1.f, 1.f);
      const float4 a = __ldcs(reinterpret_cast<const float4*>(y0 + base + c));
      const float4 bb = __ldcs(reinterpret_cast<const float4*>(y1 + base + c));
      float4 o0, o1;
      o0.x = init_value - glue_pow(g0, e.x) * a.x; o0.y = init_value - glue_pow(g0, e.y) * a.y;
      o0.z = init_value - glue_pow(g0, e.z) * a.z; o0.w = init_value - glue_pow(g0, e.w) * a.w;
      o1.x = init_value - glue_pow(g1, e.x) * bb.x; o1.y = init_value - glue_pow(g1, e.y) * bb.y;
      o1.z = init_value - glue_pow(g1, e.z) * bb.z; o1.w = init_value - glue_pow(g1, e.w) * bb.w;
      __stcs(reinterpret_cast<float4*>(p0 + base + c), o0);
      __stcs(reinterpret_cast<float4*>(p1 + base + c), o1);
      if (dtf1) { const float4 d = __ldcs(reinterpret_cast<const float4*>(dtf1 + base + c)); s1 += (d.x + d.y) + (d.z + d.w); }
      if (dtf2) { const float4 d = __ldcs(reinterpret_cast<const float4*>(dtf2 + base + c)); s2 += (d.x + d.y) + (d.z + d.w); }
    } else {
      for (int64_t q = c; q < min(c + 4, N); ++q) {
        const float e = expo ? expo[q] : 1.f;
        p0[base + q] = init_value - glue_pow(g0, e) * y0[base + q];
        p1[base + q] = init_value - glue_pow(g1, e) * y1[base + q];
        if (dtf1) s1 += dtf1[base + q];
        if (dtf2) s2 += dtf2[base + q];
      }
    }
  }
  if (dsum) {
    double v[2] = {(double)s1, (double)s2};
    block_reduce<2>(v, red);
    if (threadIdx.x == 0) { atomicAdd(&dsum[2 * b], v[0]); atomicAdd(&dsum[2 * b + 1], v[1]); }
  }
}

__global__ void k_glue_mean(int32_t B, int64_t N, const double* __restrict__ dsum, float* __restrict__ dt1, float* __restrict__ dt2) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (dt1) dt1[b] = (float)(dsum[2 * b] / (double)N);
  if (dt2) dt2[b] = (float)(dsum[2 * b + 1] / (double)N);
}

// one thread: 4 cells x SB consecutive samples; the kernel_exponent cotangent is summed over the samples in registers,
// then one atomic per cell and sample group
constexpr int SB = 8;
__global__ void __launch_bounds__(GT) k_glue_bwd(int32_t B, int64_t N, float t_lo, float t_hi, const float* __restrict__ expo,
                                                 const float* __restrict__ tn0, const float* __restrict__ tn1,
                                                 const float* __restrict__ y0, const float* __restrict__ y1,
                                                 const float* __restrict__ gp0, const float* __restrict__ gp1,
                                                 const float* __restrict__ gdt1, const float* __restrict__ gdt2,
                                                 float* __restrict__ gy0, float* __restrict__ gy1, float* __restrict__ gexpo,
                                                 float* __restrict__ gdtf1, float* __restrict__ gdtf2,
                                                 float* __restrict__ gtn0, float* __restrict__ gtn1) {
  __shared__ float tred[2][GT / 32];
  const int64_t c = ((int64_t)blockIdx.x * GT + threadIdx.x) * 4;
  const bool want_t = gtn0 != nullptr || gtn1 != nullptr;      // block-uniform: every thread stays for the reduction
  if (c >= N && !want_t) return;
  const int b0 = blockIdx.y * SB, b1 = min(b0 + SB, B);
  const bool vec = (c + 3 < N) && (N & 3) == 0 && ((reinterpret_cast<uintptr_t>(y0) | reinterpret_cast<uintptr_t>(y1) | reinterpret_cast<uintptr_t>(gp0) |
                    reinterpret_cast<uintptr_t>(gp1) | reinterpret_cast<uintptr_t>(gy0) | reinterpret_cast<uintptr_t>(gy1) |
                    reinterpret_cast<uintptr_t>(gdtf1) | reinterpret_cast<uintptr_t>(gdtf2)) & 15u) == 0;
  const int nc = c >= N ? 0 : (vec ? 4 : (int)min((int64_t)4, N - c));
  const float inv_span = 1.0f / (t_hi - t_lo);
  float e[4] = {1.f, 1.f, 1.f, 1.f}, ge[4] = {0.f, 0.f, 0.f, 0.f};
  if (expo) { _Pragma("unroll") for (int q = 0; q < 4; ++q) if (q < nc) e[q] = expo[c + q]; }
  const float invN = 1.0f / (float)N;
  for (int b = b0; b < b1; ++b) {
    const GlueLevel g0 = glue_level(tn0[b], t_lo, t_hi), g1 = glue_level(tn1[b], t_lo, t_hi);
    const int64_t o = (int64_t)b * N + c;
    float ya[4], yb[4], ga[4], gb[4], ra[4], rb[4];
    if (vec) {
      const float4 t0 = __ldcs(reinterpret_cast<const float4*>(y0 + o)), t1 = __ldcs(reinterpret_cast<const float4*>(y1 + o));
      const float4 t2 = __ldcs(reinterpret_cast<const float4*>(gp0 + o)), t3 = __ldcs(reinterpret_cast<const float4*>(gp1 + o));
      ya[0] = t0.x; ya[1] = t0.y; ya[2] = t0.z; ya[3] = t0.w; yb[0] = t1.x; yb[1] = t1.y; yb[2] = t1.z; yb[3] = t1.w;
      ga[0] = t2.x; ga[1] = t2.y; ga[2] = t2.z; ga[3] = t2.w; gb[0] = t3.x; gb[1] = t3.y; gb[2] = t3.z; gb[3] = t3.w;
    } else {
      _Pragma("unroll") for (int q = 0; q < 4; ++q) if (q < nc) { ya[q] = y0[o + q]; yb[q] = y1[o + q]; ga[q] = gp0[o + q]; gb[q] = gp1[o + q]; }
    }
    float ta = 0.f, tb = 0.f;       // d/d tn_l of -alpha_t^e * y = -y * e * alpha_t^(e-1) / (t_hi - t_lo), summed over this thread's cells
    _Pragma("unroll") for (int q = 0; q < 4; ++q) if (q < nc) {
      const float a0 = glue_pow(g0, e[q]), a1 = glue_pow(g1, e[q]);
      ra[q] = -a0 * ga[q];
      rb[q] = -a1 * gb[q];
      // d/d expo of -alpha_t^e * y = -y * alpha * ln(alpha_t)
      ge[q] = fmaf(ra[q] * ya[q], g0.lnat, fmaf(rb[q] * yb[q], g1.lnat, ge[q]));
      if (want_t) {
        ta = fmaf(-ga[q] * ya[q], glue_dpow(g0, a0, e[q]), ta);
        tb = fmaf(-gb[q] * yb[q], glue_dpow(g1, a1, e[q]), tb);
      }
    }
    if (want_t) {        // Hard_Layer_Subclassed.py:219-228: the layer's time input is differentiable (physics_loss.py:105-111)
      _Pragma("unroll") for (int o = 16; o > 0; o >>= 1) { ta += __shfl_down_sync(0xffffffffu, ta, o); tb += __shfl_down_sync(0xffffffffu, tb, o); }
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      __syncthreads();
      if (lane == 0) { tred[0][warp] = ta; tred[1][warp] = tb; }
      __syncthreads();
      if (threadIdx.x < 2) {
        float t = 0.f;
        _Pragma("unroll") for (int w = 0; w < GT / 32; ++w) t += tred[threadIdx.x][w];
        float* dst = threadIdx.x == 0 ? gtn0 : gtn1;
        if (dst) atomicAdd(&dst[b], t * inv_span);
      }
    }
    if (nc == 0) continue;
    const float d1 = gdt1 ? gdt1[b] * invN : 0.f, d2 = gdt2 ? gdt2[b] * invN : 0.f;
    if (vec) {
      __stcs(reinterpret_cast<float4*>(gy0 + o), make_float4(ra[0], ra[1], ra[2], ra[3]));
      __stcs(reinterpret_cast<float4*>(gy1 + o), make_float4(rb[0], rb[1], rb[2], rb[3]));
      if (gdtf1) __stcs(reinterpret_cast<float4*>(gdtf1 + o), make_float4(d1, d1, d1, d1));
      if (gdtf2) __stcs(reinterpret_cast<float4*>(gdtf2 + o), make_float4(d2, d2, d2, d2));
    } else {
      _Pragma("unroll") for (int q = 0; q < 4; ++q) if (q < nc) {
        gy0[o + q] = ra[q]; gy1[o + q] = rb[q];
        if (gdtf1) gdtf1[o + q] = d1;
        if (gdtf2) gdtf2[o + q] = d2;
      }
    }
  }
  if (gexpo) { _Pragma("unroll") for (int q = 0; q < 4; ++q) if (q < nc) atomicAdd(&gexpo[c + q], ge[q]); }
}

}  // namespace

extern "C" size_t srm_glue_workspace_bytes(int32_t B) { return B > 0 ? (size_t)B * 2 * sizeof(double) : 0; }

extern "C" int srm_glue_forward(const SrmHandle* h, int32_t B, float init_value, float t_lo, float t_hi, const float* expo,
                                const float* tn0, const float* tn1, const float* y0, const float* y1, const float* dtf1,
                                const float* dtf2, float* p0, float* p1, float* dt1, float* dt2, void* workspace,
                                size_t workspace_bytes, void* stream) {
  if (!h || B < 1 || B > 65535 || !tn0 || !tn1 || !y0 || !y1 || !p0 || !p1 || !(t_hi > t_lo)) { srm_set_error("srm_glue_forward: bad argument"); return SRM_ERR_INVALID; }
  if ((dtf1 != nullptr) != (dt1 != nullptr) || (dtf2 != nullptr) != (dt2 != nullptr)) { srm_set_error("srm_glue_forward: dtf_l and dt_l go together"); return SRM_ERR_INVALID; }
  const bool means = dtf1 || dtf2;
  if (means && (!workspace || workspace_bytes < srm_glue_workspace_bytes(B))) { srm_set_error("srm_glue_forward: workspace %zu < required %zu bytes", workspace_bytes, srm_glue_workspace_bytes(B)); return SRM_ERR_WORKSPACE; }
  SRM_CUDA_CHECK(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t N = h->dev.N;
  double* dsum = means ? (double*)workspace : nullptr;
  if (means) SRM_CUDA_CHECK(cudaMemsetAsync(dsum, 0, srm_glue_workspace_bytes(B), s));
  const unsigned gx = (unsigned)std::min<int64_t>((N / 4 + GT - 1) / GT + 1, 8 * 148);
  k_glue_fwd<<<dim3(gx, (unsigned)B), GT, 0, s>>>(N, init_value, t_lo, t_hi, expo, tn0, tn1, y0, y1, dtf1, dtf2, p0, p1, dsum);
  SRM_CUDA_CHECK(cudaGetLastError());
  if (means) {
    k_glue_mean<<<(unsigned)((B + 127) / 128), 128, 0, s>>>(B, N, dsum, dt1, dt2);
    SRM_CUDA_CHECK(cudaGetLastError());
  }
  return SRM_OK;
}

extern "C" int srm_glue_backward(const SrmHandle* h, int32_t B, float init_value, float t_lo, float t_hi, const float* expo,
                                 const float* tn0, const float* tn1, const float* y0, const float* y1, const float* gp0,
                                 const float* gp1, const float* gdt1, const float* gdt2, float* gy0, float* gy1,
                                 float* gexpo, float* gdtf1, float* gdtf2, float* gtn0, float* gtn1, void* stream) {
  (void)init_value;
  if (!h || B < 1 || B > 65535 || !tn0 || !tn1 || !y0 || !y1 || !gp0 || !gp1 || !gy0 || !gy1 || !(t_hi > t_lo)) { srm_set_error("srm_glue_backward: bad argument"); return SRM_ERR_INVALID; }
  if ((gdtf1 != nullptr) != (gdt1 != nullptr) || (gdtf2 != nullptr) != (gdt2 != nullptr)) { srm_set_error("srm_glue_backward: gdtf_l and gdt_l go together"); return SRM_ERR_INVALID; }
  SRM_CUDA_CHECK(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t N = h->dev.N;
  if (gexpo) SRM_CUDA_CHECK(cudaMemsetAsync(gexpo, 0, sizeof(float) * (size_t)N, s));
  if (gtn0) SRM_CUDA_CHECK(cudaMemsetAsync(gtn0, 0, sizeof(float) * (size_t)B, s));
  if (gtn1) SRM_CUDA_CHECK(cudaMemsetAsync(gtn1, 0, sizeof(float) * (size_t)B, s));
  const dim3 grid((unsigned)((N / 4 + GT) / GT), (unsigned)((B + SB - 1) / SB));
  k_glue_bwd<<<grid, GT, 0, s>>>(B, N, t_lo, t_hi, expo, tn0, tn1, y0, y1, gp0, gp1, gdt1, gdt2, gy0, gy1, gexpo, gdtf1, gdtf2, gtn0, gtn1);
  SRM_CUDA_CHECK(cudaGetLastError());
  return SRM_OK;
}

// ------------------------------------------------------------------------------------------------
// BatchGenerator.__getitem__ on a device-resident data set: tf.gather(x_all, batch_inds, axis=0)
// (training.py:110-143) without the per-step host-to-device conversion of the whole data set.  Byte work: bit-exact.
// ------------------------------------------------------------------------------------------------
namespace {

// one CTA per (row, 16 KB segment of the row): 16-byte vectors when rows are 16-byte multiples and both bases aligned
template <class V>
__global__ void __launch_bounds__(256) k_gather_rows(const V* __restrict__ src, const int32_t* __restrict__ idx, int64_t n_rows,
                                                     int64_t row_vecs, V* __restrict__ dst) {
  const int64_t r = blockIdx.y;
  const int32_t s = idx[r];
  V* out = dst + r * row_vecs;
  const bool ok = s >= 0 && (int64_t)s < n_rows;             // out-of-range index: zero row (tf.gather on a GPU)
  const V* in = src + (int64_t)(ok ? s : 0) * row_vecs;
  V zero;
  memset(&zero, 0, sizeof(V));
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < row_vecs; v += (int64_t)gridDim.x * blockDim.x)
    out[v] = ok ? in[v] : zero;
}


// ---- feature tensor: time shift and permeability channel -------------------------------------------------------
// x is (B, cells, C) with the channels innermost (the reference's (B,D,H,W,5) features).  One pass reads x once and
// writes x_n1 = x with t_norm += dn[b] (physics_loss.py:105-110) and, if asked, the de-normalised permeability of
// channel kc (DataSummary.nonormalize, log branch: data_processing_utils.py:1098-1106, same arithmetic as
// k_denorm_log).  The reference does this with a zeros_like, a strided assignment, an add and a strided slice.
template <int CT, int VEC>
__global__ void __launch_bounds__(256) k_features(int64_t per_sample, int32_t C_rt, int32_t tc, int32_t kc,
                                                  const float* __restrict__ x, const float* __restrict__ dn,
                                                  float lr, float lmin, float lo, float span,
                                                  float* __restrict__ x1, float* __restrict__ kx) {
  const int C = CT > 0 ? CT : C_rt;
  const int b = blockIdx.y;
  const float d = dn ? dn[b] : 0.f;
  const float* xs = x + (int64_t)b * per_sample;
  float* x1s = x1 ? x1 + (int64_t)b * per_sample : nullptr;
  float* kxs = kx ? kx + (int64_t)b * (per_sample / C) : nullptr;
  const int64_t nv = per_sample / VEC;                      // VEC = 4: per_sample % 4 == 0 and aligned (launcher)
#pragma unroll 4
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.x * blockDim.x) {
    float e[VEC];
    if (VEC == 4) {
      const float4 q = reinterpret_cast<const float4*>(xs)[v];
      e[0] = q.x; e[VEC > 1 ? 1 : 0] = q.y; e[VEC > 2 ? 2 : 0] = q.z; e[VEC > 3 ? 3 : 0] = q.w;
    } else {
      e[0] = xs[v];
    }
    const int64_t e0 = v * VEC;
    int ch = (int)(e0 % C);
    int64_t cell = e0 / C;
    float o[VEC];
    // at most one element of the vector is the permeability channel when C > VEC: pick it with selects and evaluate the
    // exponential once per thread (per-element branches made every warp run the division and expf VEC times)
    float kval = 0.f;
    int64_t kcell = -1;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      o[j] = (ch == tc) ? __fadd_rn(e[j], d) : e[j];
      if (ch == kc) { kval = e[j]; kcell = cell; }
      if (++ch == C) { ch = 0; ++cell; }
    }
    if (kxs && C > VEC) {
      const float u = __fdiv_rn(__fsub_rn(kval, lo), span);
      float y = expf(__fadd_rn(__fmul_rn(lr, u), lmin));
      if (isnan(y) || isinf(y)) y = 0.f;
      if (kcell >= 0) kxs[kcell] = y;
    } else if (kxs) {                 // fewer channels than vector lanes: several permeability elements per vector
      ch = (int)(e0 % C); cell = e0 / C;
      for (int j = 0; j < VEC; ++j) {
        if (ch == kc) {
          const float u = __fdiv_rn(__fsub_rn(e[j], lo), span);
          float y = expf(__fadd_rn(__fmul_rn(lr, u), lmin));
          if (isnan(y) || isinf(y)) y = 0.f;
          kxs[cell] = y;
        }
        if (++ch == C) { ch = 0; ++cell; }
      }
    }
    if (x1s) {
      if (VEC == 4) reinterpret_cast<float4*>(x1s)[v] = make_float4(o[0], o[VEC > 1 ? 1 : 0], o[VEC > 2 ? 2 : 0], o[VEC > 3 ? 3 : 0]);
      else x1s[v] = o[0];
    }
  }
}

// cotangent of dn: gdn[b] = sum over cells of gx1[b, cell, tc]   (the cotangent of x passes through unchanged)
template <int CT, int VEC>
__global__ void __launch_bounds__(256) k_features_bwd(int64_t per_sample, int32_t C_rt, int32_t tc,
                                                      const float* __restrict__ gx1, float* __restrict__ gdn) {
  __shared__ double red[32];
  const int C = CT > 0 ? CT : C_rt;
  const int b = blockIdx.y;
  const float* gs = gx1 + (int64_t)b * per_sample;
  const int64_t nv = per_sample / VEC;
  double acc = 0.0;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.x * blockDim.x) {
    float e[VEC];
    if (VEC == 4) {
      const float4 q = reinterpret_cast<const float4*>(gs)[v];
      e[0] = q.x; e[VEC > 1 ? 1 : 0] = q.y; e[VEC > 2 ? 2 : 0] = q.z; e[VEC > 3 ? 3 : 0] = q.w;
    } else {
      e[0] = gs[v];
    }
    int ch = (int)((v * VEC) % C);
    float part = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      if (ch == tc) part += e[j];
      if (++ch == C) ch = 0;
    }
    acc += (double)part;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (warp == 0) {
    double t = lane < (blockDim.x >> 5) ? red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (lane == 0) atomicAdd(&gdn[b], (float)t);
  }
}

}  // namespace

extern "C" int srm_features_forward(int32_t device, const float* x, const float* dn, int32_t B, int64_t cells, int32_t C,
                                    int32_t t_channel, int32_t k_channel, float kmin, float kmax, float lo, float hi,
                                    float* x1_out, float* kx_out, void* stream) {
  const int64_t per = cells * C;
  if (B < 1 || B > 65535 || cells < 1 || C < 1 || !x || (!x1_out && !kx_out) || (x1_out && !dn) || t_channel >= C || k_channel >= C) {
    srm_set_error("srm_features_forward: bad argument (1 <= B <= 65535, channels < C, x1_out needs dn)");
    return SRM_ERR_INVALID;
  }
  const bool vec = per % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(x1_out)) & 15u) == 0;
  if (kx_out && !(kmin > 0.f && kmax > kmin && hi > lo)) { srm_set_error("srm_features_forward: bad normalisation range"); return SRM_ERR_INVALID; }
  SRM_CUDA_CHECK(cudaSetDevice(device));
  cudaStream_t s = (cudaStream_t)stream;
  const float lr = kx_out ? logf(kmax / kmin) : 0.f, lmin = kx_out ? logf(kmin) : 0.f;
  const int64_t nv = vec ? per / 4 : per;
  // eight vectors per thread: CTAs that move 4 KB each spend their time being launched (measured 3.5 TB/s)
  const dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>((nv + 256 * 8 - 1) / (256 * 8), 1024)), (unsigned)B);
  const int kc = kx_out ? k_channel : -1, tc = x1_out ? t_channel : -1;
  if (C == 5 && vec) k_features<5, 4><<<grid, 256, 0, s>>>(per, C, tc, kc, x, dn, lr, lmin, lo, hi - lo, x1_out, kx_out);
  else if (vec) k_features<0, 4><<<grid, 256, 0, s>>>(per, C, tc, kc, x, dn, lr, lmin, lo, hi - lo, x1_out, kx_out);
  else k_features<0, 1><<<grid, 256, 0, s>>>(per, C, tc, kc, x, dn, lr, lmin, lo, hi - lo, x1_out, kx_out);   // ragged rows (39 x 39 x 5)
  SRM_CUDA_CHECK(cudaGetLastError());
  return SRM_OK;
}

extern "C" int srm_features_backward(int32_t device, const float* gx1, int32_t B, int64_t cells, int32_t C, int32_t t_channel,
                                     float* gdn, void* stream) {
  const int64_t per = cells * C;
  if (B < 1 || B > 65535 || cells < 1 || C < 1 || !gx1 || !gdn || t_channel < 0 || t_channel >= C) {
    srm_set_error("srm_features_backward: bad argument");
    return SRM_ERR_INVALID;
  }
  const bool vec = per % 4 == 0 && (reinterpret_cast<uintptr_t>(gx1) & 15u) == 0;
  SRM_CUDA_CHECK(cudaSetDevice(device));
  cudaStream_t s = (cudaStream_t)stream;
  SRM_CUDA_CHECK(cudaMemsetAsync(gdn, 0, sizeof(float) * B, s));
  const int64_t nv = vec ? per / 4 : per;
  // a few CTAs per sample: enough loads in flight when B is small and the grid is large
  const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>((nv + 256 * 64 - 1) / (256 * 64), 64));
  const dim3 grid(gx, (unsigned)B);
  if (C == 5 && vec) k_features_bwd<5, 4><<<grid, 256, 0, s>>>(per, C, t_channel, gx1, gdn);
  else if (vec) k_features_bwd<0, 4><<<grid, 256, 0, s>>>(per, C, t_channel, gx1, gdn);
  else k_features_bwd<0, 1><<<grid, 256, 0, s>>>(per, C, t_channel, gx1, gdn);
  SRM_CUDA_CHECK(cudaGetLastError());
  return SRM_OK;
}

// ---- feature construction on the device -------------------------------------------------------
// weave_tensors (data_processing_utils.py:90-223) on [permx (K, cells), time (T), x, y, z (cells)] with
// flatten_first_axes and the channel flip, fused with DataSummary.normalize ('lnk-linear-scaling', :1031-1042):
//   out[(k*T + t)][cell][0..4] = norm(z), norm(y), norm(x), norm(t), normlog(permx)
// A thread owns CPT consecutive cells of one realisation: the three coordinate channels and the permeability channel are
// normalised once and written for every time point (20 bytes per cell and sample, written as whole 16-byte vectors).
namespace {
struct WeaveStats { float mn[5], inv[5]; float lo, span; };      // channel order [z, y, x, t, k]; k: ln(min), 1/ln(max/min)
__device__ __forceinline__ float norm_lin(float v, float mn, float den, float span, float lo) {
  const float r = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(v, mn), den), span), lo);          // (((v-min)/(max-min))*(hi-lo))+lo
  return (isnan(r) || isinf(r)) ? 0.f : r;
}
template <int CPT>
__global__ void __launch_bounds__(256) k_weave(int32_t K, int32_t T, int64_t cells, const float* __restrict__ permx,
                                               const float* __restrict__ time, const float* __restrict__ xg, const float* __restrict__ yg,
                                               const float* __restrict__ zg, const float* __restrict__ st /* [5][2] min, max */,
                                               float lo, float hi, float* __restrict__ out) {
  const int64_t c0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * CPT;
  const int k = blockIdx.y;
  if (c0 >= cells) return;
  const float span = hi - lo;
  float v[CPT][5];
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
    const int64_t c = c0 + i;
    if (c < cells) {
      v[i][0] = norm_lin(zg[c], st[0], __fsub_rn(st[1], st[0]), span, lo);
      v[i][1] = norm_lin(yg[c], st[2], __fsub_rn(st[3], st[2]), span, lo);
      v[i][2] = norm_lin(xg[c], st[4], __fsub_rn(st[5], st[4]), span, lo);
      // ((log(k/min)/log(max/min))*(hi-lo))+lo
      const float r = __fadd_rn(__fmul_rn(__fdiv_rn(logf(__fdiv_rn(permx[(int64_t)k * cells + c], st[8])), logf(__fdiv_rn(st[9], st[8]))), span), lo);
      v[i][4] = (isnan(r) || isinf(r)) ? 0.f : r;
    }
  }
  const float tden = __fsub_rn(st[7], st[6]);
  for (int t = 0; t < T; ++t) {
    const float tn = norm_lin(time[t], st[6], tden, span, lo);
    float* o = out + (((int64_t)k * T + t) * cells + c0) * 5;
    if (CPT == 4 && c0 + 4 <= cells) {          // 20 floats = five 16-byte stores (rows are 16-byte aligned: 4 cells x 20 B)
      float w[20];
#pragma unroll
      for (int i = 0; i < 4; ++i) { w[5 * i] = v[i][0]; w[5 * i + 1] = v[i][1]; w[5 * i + 2] = v[i][2]; w[5 * i + 3] = tn; w[5 * i + 4] = v[i][4]; }
#pragma unroll
      for (int q = 0; q < 5; ++q) __stcs(reinterpret_cast<float4*>(o) + q, make_float4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]));
    } else {
#pragma unroll
      for (int i = 0; i < CPT; ++i)
        if (c0 + i < cells) { float* oc = o + 5 * i; oc[0] = v[i][0]; oc[1] = v[i][1]; oc[2] = v[i][2]; oc[3] = tn; oc[4] = v[i][4]; }
    }
  }
}
}  // namespace

extern "C" int srm_weave_features(int32_t device, int32_t K, int32_t T, int64_t cells, const float* permx, const float* time,
                                  const float* xg, const float* yg, const float* zg, const float* stats, float lo, float hi,
                                  float* out, void* stream) {
  if (K < 1 || K > 65535 || T < 1 || cells < 1 || !permx || !time || !xg || !yg || !zg || !stats || !out || !(hi > lo)) {
    srm_set_error("srm_weave_features: bad argument (1 <= K <= 65535, T >= 1, hi > lo, no NULL pointers)");
    return SRM_ERR_INVALID;
  }
  SRM_CUDA_CHECK(cudaSetDevice(device));
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec = cells % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
  if (vec) {
    const dim3 grid((unsigned)((cells / 4 + 255) / 256), (unsigned)K);
    k_weave<4><<<grid, 256, 0, s>>>(K, T, cells, permx, time, xg, yg, zg, stats, lo, hi, out);
  } else {
    const dim3 grid((unsigned)((cells + 255) / 256), (unsigned)K);
    k_weave<1><<<grid, 256, 0, s>>>(K, T, cells, permx, time, xg, yg, zg, stats, lo, hi, out);
  }
  SRM_CUDA_CHECK(cudaGetLastError());
  return SRM_OK;
}

extern "C" int srm_gather_rows(int32_t device, const void* src, const int32_t* idx, int64_t n_idx, int64_t n_rows,
                               int64_t row_bytes, void* dst, void* stream) {
  if (n_idx < 0 || n_rows < 1 || row_bytes < 1 || (n_idx > 0 && (!src || !idx || !dst)) || n_idx > 65535) {
    srm_set_error("srm_gather_rows: bad argument (n_idx <= 65535)");
    return SRM_ERR_INVALID;
  }
  if (n_idx == 0) return SRM_OK;
  SRM_CUDA_CHECK(cudaSetDevice(device));
  cudaStream_t s = (cudaStream_t)stream;
  const bool a16 = row_bytes % 16 == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0;
  const bool a4 = row_bytes % 4 == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 3u) == 0;
  const int64_t vecs = a16 ? row_bytes / 16 : (a4 ? row_bytes / 4 : row_bytes);
  const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>((vecs + 1023) / 1024, 1024));
  const dim3 grid(gx, (unsigned)n_idx);
  if (a16) k_gather_rows<uint4><<<grid, 256, 0, s>>>((const uint4*)src, idx, n_rows, vecs, (uint4*)dst);
  else if (a4) k_gather_rows<uint32_t><<<grid, 256, 0, s>>>((const uint32_t*)src, idx, n_rows, vecs, (uint32_t*)dst);
  else k_gather_rows<unsigned char><<<grid, 256, 0, s>>>((const unsigned char*)src, idx, n_rows, vecs, (unsigned char*)dst);
  SRM_CUDA_CHECK(cudaGetLastError());
  return SRM_OK;
}
