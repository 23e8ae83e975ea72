// Gas-condensate (two-phase gas-oil, fluid_type 'GC') path: physics_error_gas_oil and its gradient.
//
//   k_relperm         RelativePermeability.compute_krog_krgo          relative_permeability.py:49-75
//   k_stage_gc        PVTLayer (7 properties) at both time levels,     PVT_Layer_Subclassed.py:146-216
//                     mobility products, relative permeabilities       physics_loss.py:336-391,457
//   k_wells_gc        WellRatesPressure, GC branch                     well_rate_bhp_Subclassed.py:614-724,727-837,963-1034
//   k_resid_fwd_gc    residual, truncation term, material balance      physics_loss.py:419-441,465-665
//   k_resid_adj_gc    hand-derived adjoint (tape.gradient, physics_loss.py:849-859)
//   k_ibc_adj_gc      inner-boundary (well-cell) part of the adjoint
//
// Numerics: the reference's fp32 op order, explicitly rounded (no FMA contraction) in every forward
// quantity, as in kernels_ref.cu; tf.pow with the integer Corey exponents is the left-to-right product the
// oracle pins.  Two pipelines with the same per-cell arithmetic:
//   staged   k_stage_gc writes the PVT packs and per-cell products to the workspace (SRM_GC_NFIELDS fields; from the
//            exact per-pressure table where it covers the pressure, else the 37-term spline of 7 properties), the
//            residual kernels march z-columns and re-read the fields of six neighbours through L1/L2
//   fused    gc_fused.cuh (table over the whole clamp range): tile + halo in shared memory, gathers in-kernel, no
//            staged fields
#include <math_constants.h>
#include <cstring>
#include <cstdlib>
#include "ref_fused.cuh"
#include "well_tile.cuh"

namespace {

constexpr int kThreads = 256;

// staged fields (index into ws.gc, each [B*N])
enum {
  F_A0 = 0, F_B0, F_RS0, F_RV0, F_DA0, F_DB0, F_DRS0, F_DRV0,        // n0: invBg, invBo, Rs, Rv and d/dp (unmasked)
  F_D2A0, F_D2B0, F_D2RS0, F_D2RV0,                                  // n0: d2/dp2 (masked by the clamp)      [adjoint]
  F_MGG, F_MOO, F_MGO, F_MOG, F_KRG, F_KRO,                          // n1: neighbour-visible
  F_A1, F_B1, F_R1, F_V1,                                            // n1: invBg, invBo, Rs*invBo, Rv*invBg
  F_DMG, F_DMO, F_DA1, F_DB1, F_DR1, F_DV1, F_DKRG, F_DKRO,          // n1 derivatives (masked)               [adjoint]
  F_COUNT
};
static_assert(F_COUNT <= SRM_GC_NFIELDS, "workspace carve");

// ---- relative permeability ------------------------------------------------------------------
// non-integer exponents: out of line, so the residual kernels do not carry several inlined copies of powf
__device__ __noinline__ float pow_general(float x, float n) { return powf(x, n); }
// x^N as the left-to-right product x*x*...*x (N >= 1), straight-line
template <int N>
__device__ __forceinline__ float powi(float x) {
  float y = x;
#pragma unroll
  for (int i = 1; i < N; ++i) y = __fmul_rn(y, x);
  return y;
}
__device__ __forceinline__ float powi_rt(float x, int ni) {     // ni >= 1, uniform over the grid
  switch (ni) {
    case 1: return x;
    case 2: return powi<2>(x);
    case 3: return powi<3>(x);
    case 4: return powi<4>(x);
    case 5: return powi<5>(x);
    case 6: return powi<6>(x);
    default: break;
  }
  float y = powi<6>(x);
  for (int i = 6; i < ni; ++i) y = __fmul_rn(y, x);
  return y;
}
__device__ __forceinline__ float pow_pinned(float x, float n, int ni) {
  return ni > 0 ? powi_rt(x, ni) : pow_general(x, n);
}
// d/dx of pow_pinned as the product rule delivers it: n * x^(n-1)
__device__ __forceinline__ float dpow_pinned(float x, float n, int ni) {
  if (ni > 0) return (float)ni * (ni > 1 ? powi_rt(x, ni - 1) : 1.f);
  return n * pow_general(x, n - 1.f);
}
// relative_permeability.py:58-73; derivatives follow TF's routing: tf.where picks a branch, tf.minimum /
// tf.maximum pass the gradient to the first argument on ties.
// NOG / NG > 0: the integer Corey exponents as compile-time constants (the fused kernels are instantiated for the
// reference's defaults, nog = 3 and ng = 6: the left-to-right products are then straight-line code instead of a switch
// per call); 0: the handle's run-time exponents
// den: the two saturation denominators as DivC (ref_fused.cuh: div.rn's own arithmetic with the reciprocal hoisted out of
// the cell loop, same bits), or null for the IEEE intrinsic
template <int NOG = 0, int NG = 0>
__device__ __forceinline__ void corey(const SrmDev& P, float sg, float& krog, float& krgo, float& dkrog, float& dkrgo,
                                      const DivC* den = nullptr) {
  const float so = __fsub_rn(__fsub_rn(1.0f, sg), P.swmin);                              // :58
  const float xo = den ? div_c(__fsub_rn(so, P.sorg), den[0]) : __fdiv_rn(__fsub_rn(so, P.sorg), P.kr_den_o);
  const float xg = den ? div_c(__fsub_rn(sg, P.sgc), den[1]) : __fdiv_rn(__fsub_rn(sg, P.sgc), P.kr_den_g);
  const float po = NOG > 0 ? powi<NOG ? NOG : 1>(xo) : pow_pinned(xo, P.nog, P.nog_i);
  const float pg = NG > 0 ? powi<NG ? NG : 1>(xg) : pow_pinned(xg, P.ng, P.ng_i);
  float ko = __fmul_rn(P.kro_somax, po);                                                 // :59
  float kg = __fmul_rn(P.krg_sorg, pg);                                                  // :60
  const float dpo = NOG > 0 ? (float)NOG * (NOG > 1 ? powi<(NOG > 1) ? NOG - 1 : 1>(xo) : 1.f) : dpow_pinned(xo, P.nog, P.nog_i);
  const float dpg = NG > 0 ? (float)NG * (NG > 1 ? powi<(NG > 1) ? NG - 1 : 1>(xg) : 1.f) : dpow_pinned(xg, P.ng, P.ng_i);
  float dko = P.kro_somax * dpo * (-1.0f / P.kr_den_o);
  float dkg = P.krg_sorg * dpg * (1.0f / P.kr_den_g);
  if (so <= P.kr_so_zero) { ko = 0.f; dko = 0.f; }                                       // :67
  if (sg > P.kr_sg_full) { kg = P.krg_swmin; dkg = 0.f; }                                // :68
  if (!(ko <= P.kro_somax)) { ko = P.kro_somax; dko = 0.f; }                             // :71 tf.minimum
  if (!(ko >= 0.f)) { ko = 0.f; dko = 0.f; }                                             //     tf.maximum
  if (!(kg <= P.krg_swmin)) { kg = P.krg_swmin; dkg = 0.f; }                             // :72
  if (!(kg >= 0.f)) { kg = 0.f; dkg = 0.f; }
  krog = ko; krgo = kg; dkrog = dko; dkrgo = dkg;
}

__global__ void __launch_bounds__(kThreads) k_relperm(const __grid_constant__ SrmDev P, int64_t n, const float* __restrict__ sg,
                                                      float* __restrict__ krog, float* __restrict__ krgo,
                                                      float* __restrict__ dkrog, float* __restrict__ dkrgo) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  float ko, kg, dko, dkg;
  corey(P, sg[g], ko, kg, dko, dkg);
  if (krog) krog[g] = ko;
  if (krgo) krgo[g] = kg;
  if (dkrog) dkrog[g] = dko;
  if (dkrgo) dkrgo[g] = dkg;
}

// ---- stage -------------------------------------------------------------------------------------
// Everything the GC residual and its adjoint need from the PVT layer, as two packs of 12 floats, each a
// pure function of ONE clamped fp32 pressure (derivatives w.r.t. the clamped input; the consumer applies
// the clamp's gradient mask):
//   pack0 (time level n)  : invBg, invBo, Rs, Rv | d/dp of the four | d2/dp2 of the four
//   pack1 (level n+1)     : Mgg, Moo, Mgo, Mog | invBg, invBo, Rs*invBo, Rv*invBg | d(Mgg+Mog), d(Mgo+Moo), d invBg, d invBo
//                           | d(Rs*invBo), d(Rv*invBg)   (14 values; stored as 16 floats)
struct GcPack0 { float v[12]; };
struct GcPack1 { float v[16]; };
template <bool SAVE>
__device__ __forceinline__ GcPack0 gc_pack0(const SrmDev& P, float x0) {
  GcPack0 o;
  float v[2], d[2], d2[2];
  d2[0] = d2[1] = 0.f;
  srm_pvt_ref<2, true, SAVE>(P, 0, x0, v, d, d2);          // InvBg, InvBo
  o.v[0] = v[0]; o.v[1] = v[1]; o.v[4] = d[0]; o.v[5] = d[1]; o.v[8] = d2[0]; o.v[9] = d2[1];
  d2[0] = d2[1] = 0.f;
  srm_pvt_ref<2, true, SAVE>(P, 4, x0, v, d, d2);          // Rs, Rv
  o.v[2] = v[0]; o.v[3] = v[1]; o.v[6] = d[0]; o.v[7] = d[1]; o.v[10] = d2[0]; o.v[11] = d2[1];
  return o;
}
template <bool SAVE>
__device__ __forceinline__ GcPack1 gc_pack1(const SrmDev& P, float x1) {
  GcPack1 o;
  float v[6], d[6], d2[6];
#pragma unroll
  for (int q = 0; q < 6; ++q) d[q] = 0.f;
  srm_pvt_ref<6, SAVE, false>(P, 0, x1, v, d, d2);         // InvBg, InvBo, Invug, Invuo, Rs, Rv
  const float a = v[0], b = v[1], ug = v[2], uo = v[3], rs = v[4], rv = v[5];
  const float r = __fmul_rn(rs, b), vv = __fmul_rn(rv, a);                     // physics_loss.py:388-389
  o.v[0] = __fmul_rn(a, ug);                                                   // :386
  o.v[1] = __fmul_rn(b, uo);                                                   // :387
  o.v[2] = __fmul_rn(r, uo);                                                   // :390
  o.v[3] = __fmul_rn(vv, ug);                                                  // :391
  o.v[4] = a; o.v[5] = b; o.v[6] = r; o.v[7] = vv;
  const float da = d[0], db = d[1], dug = d[2], duo = d[3], drs = d[4], drv = d[5];
  const float dr = __fmaf_rn(drs, b, __fmul_rn(rs, db)), dvv = __fmaf_rn(drv, a, __fmul_rn(rv, da));
  const float dMgg = __fmaf_rn(da, ug, __fmul_rn(a, dug)), dMoo = __fmaf_rn(db, uo, __fmul_rn(b, duo));
  const float dMgo = __fmaf_rn(dr, uo, __fmul_rn(r, duo)), dMog = __fmaf_rn(dvv, ug, __fmul_rn(vv, dug));
  o.v[8] = __fadd_rn(dMgg, dMog); o.v[9] = __fadd_rn(dMgo, dMoo);
  o.v[10] = da; o.v[11] = db; o.v[12] = dr; o.v[13] = dvv; o.v[14] = 0.f; o.v[15] = 0.f;
  return o;
}

// exact tabulation (SrmConfig.pvt_lut, see kernels_ref.cu): entry e = the packs of the fp32 pressure with bit
// pattern lut_lo_bits + e; 48 + 64 bytes per representable pressure.
__global__ void __launch_bounds__(kThreads) k_lut_build_gc(const __grid_constant__ SrmDev P, float4* __restrict__ t0,
                                                           float4* __restrict__ t1, float4* __restrict__ f0, float4* __restrict__ f1,
                                                           float2* __restrict__ fv, float4* __restrict__ a1) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P.lut_n) return;
  const float x = __uint_as_float(P.lut_lo_bits + e);
  const GcPack0 a = gc_pack0<true>(P, x);
  const GcPack1 b = gc_pack1<true>(P, x);
#pragma unroll
  for (int i = 0; i < 3; ++i) t0[(size_t)e * 3 + i] = make_float4(a.v[4 * i], a.v[4 * i + 1], a.v[4 * i + 2], a.v[4 * i + 3]);
  if (t1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) t1[(size_t)e * 4 + i] = make_float4(b.v[4 * i], b.v[4 * i + 1], b.v[4 * i + 2], b.v[4 * i + 3]);
  }
  if (f0) {     // the fused forward's 32-byte views (gc_fused.cuh): values and first derivatives of level n, products of level n+1
#pragma unroll
    for (int i = 0; i < 2; ++i) f0[(size_t)e * 2 + i] = make_float4(a.v[4 * i], a.v[4 * i + 1], a.v[4 * i + 2], a.v[4 * i + 3]);
    f1[(size_t)e * 2] = make_float4(b.v[0], b.v[2], b.v[1], b.v[3]);            // component order gg, go, oo, og
    f1[(size_t)e * 2 + 1] = make_float4(b.v[4], b.v[5], b.v[6], b.v[7]);
    fv[e] = make_float2(b.v[0] + b.v[3], b.v[2] + b.v[1]);                      // the adjoint's neighbour-visible sums
    // the adjoint's own-cell view of level n+1, one 32-byte sector: values, d(Mgg+Mog), d(Mgo+Moo) and the two sums
    // of derivatives its expressions use (d invBg + d(Rv invBg), d(Rs invBo) + d invBo)
    a1[(size_t)e * 2] = make_float4(b.v[4], b.v[5], b.v[6], b.v[7]);
    a1[(size_t)e * 2 + 1] = make_float4(b.v[8], b.v[9], b.v[10] + b.v[13], b.v[12] + b.v[11]);
  }
}

template <bool SAVE, bool LUT>
__global__ void __launch_bounds__(kThreads) k_stage_gc(const __grid_constant__ SrmDev P, int64_t total,
                                                       const float* __restrict__ p0, const float* __restrict__ p1,
                                                       const float* __restrict__ sg1, float* __restrict__ F) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  auto out = [&](int f, float v) { F[(int64_t)f * total + g] = v; };
  float m0, m1;
  const float x0 = srm_clamp(P, p0[g], m0);
  const float x1 = srm_clamp(P, p1[g], m1);
  GcPack0 a;
  GcPack1 b;
  const uint32_t e0 = __float_as_uint(x0) - P.lut_lo_bits, e1 = __float_as_uint(x1) - P.lut_lo_bits;
  if (LUT && e0 < P.lut_n) {
#pragma unroll
    for (int i = 0; i < (SAVE ? 3 : 2); ++i) {
      const float4 t = __ldg(P.lut0 + (size_t)e0 * 3 + i);
      a.v[4 * i] = t.x; a.v[4 * i + 1] = t.y; a.v[4 * i + 2] = t.z; a.v[4 * i + 3] = t.w;
    }
  } else {
    a = gc_pack0<SAVE>(P, x0);
  }
  if (LUT && e1 < P.lut_n) {
#pragma unroll
    for (int i = 0; i < (SAVE ? 4 : 2); ++i) {
      const float4 t = __ldg(P.lut1 + (size_t)e1 * 4 + i);
      b.v[4 * i] = t.x; b.v[4 * i + 1] = t.y; b.v[4 * i + 2] = t.z; b.v[4 * i + 3] = t.w;
    }
  } else {
    b = gc_pack1<SAVE>(P, x1);
  }
  out(F_A0, a.v[0]); out(F_B0, a.v[1]); out(F_RS0, a.v[2]); out(F_RV0, a.v[3]);
  out(F_DA0, a.v[4]); out(F_DB0, a.v[5]); out(F_DRS0, a.v[6]); out(F_DRV0, a.v[7]);
  out(F_MGG, b.v[0]); out(F_MOO, b.v[1]); out(F_MGO, b.v[2]); out(F_MOG, b.v[3]);
  out(F_A1, b.v[4]); out(F_B1, b.v[5]); out(F_R1, b.v[6]); out(F_V1, b.v[7]);
  float ko, kg, dko, dkg;
  corey(P, sg1[g], ko, kg, dko, dkg);                                            // :457
  out(F_KRG, kg); out(F_KRO, ko);
  if (SAVE) {
    out(F_D2A0, a.v[8] * m0); out(F_D2B0, a.v[9] * m0); out(F_D2RS0, a.v[10] * m0); out(F_D2RV0, a.v[11] * m0);
    out(F_DMG, b.v[8] * m1); out(F_DMO, b.v[9] * m1);
    out(F_DA1, b.v[10] * m1); out(F_DB1, b.v[11] * m1); out(F_DR1, b.v[12] * m1); out(F_DV1, b.v[13] * m1);
    out(F_DKRG, dkg); out(F_DKRO, dko);
  }
}

// ---- wells: forward-mode numbers carrying d/dp and d/dSg of the connection cell ----------------------
struct D2 { float v, a, b; };
__device__ __forceinline__ D2 mk(float v, float a = 0.f, float b = 0.f) { D2 r; r.v = v; r.a = a; r.b = b; return r; }
__device__ __forceinline__ D2 operator+(D2 x, D2 y) { return mk(__fadd_rn(x.v, y.v), x.a + y.a, x.b + y.b); }
__device__ __forceinline__ D2 operator-(D2 x, D2 y) { return mk(__fsub_rn(x.v, y.v), x.a - y.a, x.b - y.b); }
__device__ __forceinline__ D2 operator*(D2 x, D2 y) { return mk(__fmul_rn(x.v, y.v), x.a * y.v + x.v * y.a, x.b * y.v + x.v * y.b); }
__device__ __forceinline__ D2 operator/(D2 x, D2 y) {
  const float q = __fdiv_rn(x.v, y.v);
  return mk(q, (x.a - q * y.a) / y.v, (x.b - q * y.b) / y.v);
}
__device__ __forceinline__ D2 dnn2(D2 x, D2 y) { return (y.v == 0.f) ? mk(0.f) : x / y; }       // tf.math.divide_no_nan
__device__ __forceinline__ D2 min2(D2 x, D2 y) { return (x.v <= y.v) ? x : y; }                  // ties -> first argument
__device__ __forceinline__ D2 max2(D2 x, D2 y) { return (x.v >= y.v) ? x : y; }
__device__ __forceinline__ D2 clip2(D2 t, D2 lo, D2 hi) {                                         // tf.clip_by_value
  const bool below = t.v < lo.v, above = t.v > hi.v;
  const D2& g = below ? lo : (above ? hi : t);
  return mk(fmaxf(fminf(t.v, hi.v), lo.v), g.a, g.b);
}

// Corey relative permeabilities with first AND second derivative in Sg (the Newton iterations of the blocking-factor
// integral are differentiated through: the tangent of d cost/d Sg needs the second derivative); TF's routing as in corey()
__device__ __forceinline__ float d2pow_pinned(float x, float n, int ni) {
  if (ni > 0) return ni > 1 ? (float)(ni * (ni - 1)) * (ni > 2 ? powi_rt(x, ni - 2) : 1.f) : 0.f;
  return n * (n - 1.f) * pow_general(x, n - 2.f);
}
struct Corey2 { float ko, kg, dko, dkg, d2ko, d2kg; };
__device__ __forceinline__ Corey2 corey2(const SrmDev& P, float sg) {
  Corey2 r;
  const float so = __fsub_rn(__fsub_rn(1.0f, sg), P.swmin);
  const float xo = __fdiv_rn(__fsub_rn(so, P.sorg), P.kr_den_o);
  const float xg = __fdiv_rn(__fsub_rn(sg, P.sgc), P.kr_den_g);
  const float io = 1.0f / P.kr_den_o, ig = 1.0f / P.kr_den_g;
  r.ko = __fmul_rn(P.kro_somax, pow_pinned(xo, P.nog, P.nog_i));
  r.kg = __fmul_rn(P.krg_sorg, pow_pinned(xg, P.ng, P.ng_i));
  r.dko = P.kro_somax * dpow_pinned(xo, P.nog, P.nog_i) * (-io);
  r.dkg = P.krg_sorg * dpow_pinned(xg, P.ng, P.ng_i) * ig;
  r.d2ko = P.kro_somax * d2pow_pinned(xo, P.nog, P.nog_i) * io * io;
  r.d2kg = P.krg_sorg * d2pow_pinned(xg, P.ng, P.ng_i) * ig * ig;
  if (so <= P.kr_so_zero) { r.ko = 0.f; r.dko = 0.f; r.d2ko = 0.f; }
  if (sg > P.kr_sg_full) { r.kg = P.krg_swmin; r.dkg = 0.f; r.d2kg = 0.f; }
  if (!(r.ko <= P.kro_somax)) { r.ko = P.kro_somax; r.dko = 0.f; r.d2ko = 0.f; }
  if (!(r.ko >= 0.f)) { r.ko = 0.f; r.dko = 0.f; r.d2ko = 0.f; }
  if (!(r.kg <= P.krg_swmin)) { r.kg = P.krg_swmin; r.dkg = 0.f; r.d2kg = 0.f; }
  if (!(r.kg >= 0.f)) { r.kg = 0.f; r.dkg = 0.f; r.d2kg = 0.f; }
  return r;
}

// the six PVT properties at a (dual) pressure: values with d/dp, d/dSg carried through the clamp
struct Pvt6 { D2 invBg, invBo, invug, invuo, Rs, Rv; };
__device__ __forceinline__ Pvt6 pvt6_at(const SrmDev& P, D2 pb) {
  float m;
  const float x = srm_clamp(P, pb.v, m);
  float v[6], d[6], d2[6];
  srm_pvt_ref<6, true, false>(P, 0, x, v, d, d2);
  Pvt6 o;
  auto mkp = [&](int i) { return mk(v[i], d[i] * m * pb.a, d[i] * m * pb.b); };
  o.invBg = mkp(0); o.invBo = mkp(1); o.invug = mkp(2); o.invuo = mkp(3); o.Rs = mkp(4); o.Rv = mkp(5);
  return o;
}
// mobilities of one node at a (dual) saturation                        well_rate_bhp_Subclassed.py:900-906,926-934
__device__ __forceinline__ void node_mob(const Pvt6& q, D2 krog, D2 krgo, D2& mg, D2& mo) {
  const D2 mgg = krgo * q.invBg * q.invug;
  const D2 mgo = krog * q.invBo * q.invuo * q.Rs;
  const D2 moo = krog * q.invBo * q.invuo;
  const D2 mog = krgo * q.invBg * q.invug * q.Rv;
  mg = mgg + mgo;
  mo = moo + mog;
}
__device__ __forceinline__ D2 with_tangent(float v, float dv, D2 s) { return mk(v, dv * s.a, dv * s.b); }

// compute_blocking_integral_and_factor, GC branch                       well_rate_bhp_Subclassed.py:857-950
// Trapezoid over tf.linspace(p, pwf, n+1); at each node the saturation solves cost(Sg) = mo(Sg) mg_n1 - mo_n1 mg(Sg) = 0
// with the node's PVT.  Every root-finder iteration carries d/dp and d/dSg of the connection cell, as tf.while_loop's
// gradient does.
__device__ void blocking_integral_gc(const SrmDev& P, D2 p, D2 pwf, float krog_n1, D2 mg_n1, D2 mo_n1, D2& Ig, D2& Io) {
  const int n = P.n_int;
  const D2 delta = (pwf - p) / mk((float)n);
  const float sg_max = 1.0f - P.swmin;
  const D2 tiny = mk(1e-12f), zero = mk(0.f), smax = mk(sg_max);
  D2 sum_g = zero, sum_o = zero, mg_prev = mg_n1, mo_prev = mo_n1, pa = p;
  const bool cond = krog_n1 < 1e-3f;                                            // :897
  for (int i = 0; i < n; ++i) {
    const D2 pb = (i + 1 < n) ? p + delta * mk((float)(i + 1)) : pwf;
    const Pvt6 q = pvt6_at(P, pb);
    auto cost = [&](D2 s, D2* dcost) {
      const Corey2 c = corey2(P, s.v);
      D2 mg, mo;
      node_mob(q, with_tangent(c.ko, c.dko, s), with_tangent(c.kg, c.dkg, s), mg, mo);
      if (dcost) {                      // d cost / d s (the inner tape), itself a dual number
        D2 dmg, dmo;
        node_mob(q, with_tangent(c.dko, c.d2ko, s), with_tangent(c.dkg, c.d2kg, s), dmg, dmo);
        *dcost = dmo * mg_n1 - mo_n1 * dmg;
      }
      return mo * mg_n1 - mo_n1 * mg;
    };
    D2 s1;
    if (P.root_solver == SRM_ROOT_NEWTON) {                                      // _solve_newton :236-269
      s1 = mk(0.1f);
      for (int it = 0; it < P.n_root_iter; ++it) {
        D2 df;
        const D2 f = cost(s1, &df);
        s1 = clip2(s1 - f / (df + tiny), zero, smax);
      }
    } else {                                                                     // _solve_chandrupatla :272-324
      D2 lo = zero, hi = smax;
      D2 f_lo = cost(lo, nullptr), f_hi = cost(hi, nullptr);
      if (f_lo.v * f_hi.v > 0.f) { hi = lo + mk(1e-3f); f_hi = cost(hi, nullptr); }
      // the reference's loop runs while ANY element of the field is wider than tol; a converged element keeps stepping
      // with the others.  Per connection here: step while this bracket is wider than tol.
      for (int it = 0; it < P.n_root_iter && (hi.v - lo.v) > 1e-6f; ++it) {
        const D2 d = (f_hi - f_lo) / ((hi - lo) + tiny);
        const D2 guess = hi - f_hi / d;
        const D2 f_g = cost(guess, nullptr);
        if (f_lo.v * f_g.v < 0.f) { hi = guess; f_hi = f_g; } else { lo = guess; f_lo = f_g; }
      }
      s1 = mk(0.5f) * (lo + hi);
    }
    if (cond) s1 = smax;                                                         // :912
    const Corey2 c1 = corey2(P, s1.v);
    D2 mg1, mo1;
    node_mob(q, with_tangent(c1.ko, c1.dko, s1), with_tangent(c1.kg, c1.dkg, s1), mg1, mo1);
    const D2 dp = pa - pb;
    sum_g = sum_g + mk(0.5f) * (mg_prev + mg1) * dp;                             // :937
    sum_o = sum_o + mk(0.5f) * (mo_prev + mo1) * dp;                             // :938
    mg_prev = mg1; mo_prev = mo1; pa = pb;
  }
  Ig = sum_g; Io = sum_o;
}

__global__ void __launch_bounds__(128) k_wells_gc(const __grid_constant__ SrmDev P, int32_t B, int32_t R,
                                                  const float* __restrict__ kx, const int32_t* __restrict__ sample_real,
                                                  const float* __restrict__ pfield, const float* __restrict__ sgfield,
                                                  const float* __restrict__ t_days, float* __restrict__ W7,
                                                  float* __restrict__ pwfw) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int nw = P.n_wells;
  const int64_t tot = (int64_t)B * nw;
  if (g >= tot) return;
  const int b = (int)(g / nw), w = (int)(g % nw);
  const WellDev wd = P.wells[w];
  const int r = srm_real_of(sample_real, b, B, R);
  const float k = kx[(int64_t)r * P.N + wd.cell];
  const float pv = pfield[(int64_t)b * P.N + wd.cell], sv = sgfield[(int64_t)b * P.N + wd.cell];
  const float t = t_days[b];
  const float open = (t >= wd.shut_start && t <= wd.shut_stop) ? 0.f : 1.f;      // welldata_processor.py:349-354
  // Peaceman                                                  well_rate_bhp_Subclassed.py:782-788
  const float ky = __fmul_rn(P.kx_ky, k);
  const float ryx = __fdiv_rn(ky, k), rxy = __fdiv_rn(k, ky);
  const float num = sqrtf(__fadd_rn(__fmul_rn(sqrtf(ryx), __fmul_rn(P.dx, P.dx)), __fmul_rn(sqrtf(rxy), __fmul_rn(P.dy, P.dy))));
  const float den = __fadd_rn(powf(ryx, 0.25f), powf(rxy, 0.25f));
  const float ro = __fdiv_rn(__fmul_rn(0.28f, num), den);
  float ck = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(6.283185307179586f, wd.hc), k), P.dz), P.C);
  ck = __fdiv_rn(ck, logf(__fdiv_rn(ro, wd.rw)));
  const D2 Ck = mk(__fmul_rn(open, ck));
  // relative permeabilities and PVT at the connection cell      :791-795
  float ko, kg, dko, dkg;
  corey(P, sv, ko, kg, dko, dkg);
  const D2 krog = mk(ko, 0.f, dko), krgo = mk(kg, 0.f, dkg);
  float m1;
  const float x1 = srm_clamp(P, pv, m1);
  float v[6], d[6], d2[6];
  srm_pvt_ref<6, true, false>(P, 0, x1, v, d, d2);
  const D2 invBg = mk(v[0], d[0] * m1), invBo = mk(v[1], d[1] * m1), invug = mk(v[2], d[2] * m1), invuo = mk(v[3], d[3] * m1),
           Rs = mk(v[4], d[4] * m1), Rv = mk(v[5], d[5] * m1);
  const D2 mgg = krgo * invBg * invug;                                           // :802-807
  const D2 mgo = krog * invBo * invuo * Rs;
  const D2 moo = krog * invBo * invuo;
  const D2 mog = krgo * invBg * invug * Rv;
  const D2 mg = mgg + mgo, mo = moo + mog;
  const D2 p = mk(pv, 1.f, 0.f), pmin = mk(wd.pwf_min), qt = mk(wd.q_target), zero = mk(0.f), one = mk(1.f), tiny = mk(1e-12f);
  // _compute_phase_rates (:963-1007); without the blocking factor Ig = Io = 1 (:955-959)
  auto phase_rates = [&](D2 pwf_, D2& qg_, D2& qo_) {
    D2 ig = one, io = one;
    if (P.use_blk) blocking_integral_gc(P, p, pwf_, ko, mg, mo, ig, io);
    const D2 dp = (p - pwf_) + tiny;                                             // :987
    const D2 blk_g = P.use_blk ? dnn2(ig, mg * dp) : ig;                         // :990-995
    const D2 blk_o = P.use_blk ? dnn2(io, mo * dp) : io;
    qg_ = max2(min2(qt, Ck * blk_g * mg * dp), zero);                            // :997,1001
    const D2 qo_target = qg_ * (one / (Rv + tiny));                              // :1004
    qo_ = max2(min2(qo_target, Ck * blk_o * mo * dp), zero);                     // :998,1005
  };
  D2 pwf, qg, qo;
  if (!P.bhp_iterative) {
    // ---- _non_iterative_method (:614-724)
    D2 ig_max = one, io_max = one;
    if (P.use_blk) blocking_integral_gc(P, p, pmin, ko, mg, mo, ig_max, io_max);
    const D2 dp_max = (p - pmin) + tiny;                                         // :650
    const D2 blk_max = P.use_blk ? dnn2(ig_max, mg * dp_max) : ig_max;           // :654-657
    const D2 qg_max = Ck * blk_max * mg * dp_max;                                // :662
    const D2 qg_opt = max2(min2(qt, qg_max), zero);                              // :666
    const D2 lam = clip2(dnn2(qg_opt, Ck * blk_max * mg), zero, blk_max);        // :699
    pwf = clip2(p - lam * dp_max, pmin, p);                                      // :721-723
  } else {
    // ---- _iterative_method (:515-612): see wells.cuh for the stopping rule
    const D2 eps = mk(14.7f);
    pwf = pmin + mk(0.5f) * (p - pmin);                                          // :537
    for (int it = 0; it < P.bhp_max_iters; ++it) {
      D2 qg_it, qo_it, qg_plus, qo_plus;
      phase_rates(pwf, qg_it, qo_it);                                            // :566-569
      if (!(fabsf(__fsub_rn(qg_it.v, qt.v)) > P.bhp_tol)) break;
      phase_rates(pwf + eps, qg_plus, qo_plus);                                  // :571-574
      const D2 dq = (qg_plus - qg_it) / eps;                                     // :576
      pwf = clip2(pwf - (qg_it - qt) / (dq + tiny), pmin, p);                    // :584-586
    }
  }
  phase_rates(pwf, qg, qo);
  // ---- _split_condensate_components (:1010-1034)
  const D2 dg = (mgg + mgo) + tiny, dn = (moo + mog) + tiny;
  const D2 qgg = qg * (mgg / dg), qgo = qg * (mgo / dg), qoo = qo * (moo / dn), qog = qo * (mog / dn);
  W7[0 * tot + g] = qgg.v; W7[1 * tot + g] = qgo.v; W7[2 * tot + g] = qoo.v; W7[3 * tot + g] = qog.v;
  W7[4 * tot + g] = (qgg.a + qgo.a) + (qoo.a + qog.a);       // d(sum of the four rates)/dp
  W7[5 * tot + g] = (qgg.b + qgo.b) + (qoo.b + qog.b);       // d(sum)/dSg
  pwfw[g] = pwf.v;
}

// ---- residual ----------------------------------------------------------------------------------
struct CellIdx { int i, j, k; int n[6]; };   // neighbour cells W,E,S,N,D,U with edge replication
__device__ __forceinline__ CellIdx cell_index(const SrmDev& P, int c) {
  CellIdx x;
  x.i = c % P.W;
  const int t = c / P.W;
  x.j = t % P.H;
  x.k = t / P.H;
  const int HW = P.H * P.W;
  x.n[0] = (x.i > 0) ? c - 1 : c;
  x.n[1] = (x.i < P.W - 1) ? c + 1 : c;
  x.n[2] = (x.j > 0) ? c - P.W : c;
  x.n[3] = (x.j < P.H - 1) ? c + P.W : c;
  x.n[4] = (x.k > 0) ? c - HW : c;
  x.n[5] = (x.k < P.D - 1) ? c + HW : c;
  return x;
}
__device__ __forceinline__ float harm_ref(float kc, float kn) {                  // physics_loss.py:283-284
  return __fdiv_rn(__fmul_rn(__fmul_rn(2.0f, kc), kn), __fadd_rn(kc, kn));
}
// C*k_f for the six faces (W,E,S,N,D,U)
__device__ __forceinline__ void face_perms(const SrmDev& P, const float* __restrict__ kr, int c, const CellIdx& ix, float (&ckf)[6]) {
  const float kc = kr[c];
  const float kyc = __fmul_rn(P.kx_ky, kc), kzc = __fmul_rn(P.kv_kh, kc);
  ckf[0] = __fmul_rn(P.C, harm_ref(kc, kr[ix.n[0]]));
  ckf[1] = __fmul_rn(P.C, harm_ref(kr[ix.n[1]], kc));
  ckf[2] = __fmul_rn(P.C, harm_ref(kyc, __fmul_rn(P.kx_ky, kr[ix.n[2]])));
  ckf[3] = __fmul_rn(P.C, harm_ref(__fmul_rn(P.kx_ky, kr[ix.n[3]]), kyc));
  ckf[4] = __fmul_rn(P.C, harm_ref(kzc, __fmul_rn(P.kv_kh, kr[ix.n[4]])));
  ckf[5] = __fmul_rn(P.C, harm_ref(__fmul_rn(P.kv_kh, kr[ix.n[5]]), kzc));
}

// neighbour cells of (i, j, k) with edge replication, without the integer divisions of cell_index
__device__ __forceinline__ CellIdx cell_index3(const SrmDev& P, int i, int j, int k, int c) {
  CellIdx x;
  x.i = i; x.j = j; x.k = k;
  const int HW = P.H * P.W;
  x.n[0] = (i > 0) ? c - 1 : c;
  x.n[1] = (i < P.W - 1) ? c + 1 : c;
  x.n[2] = (j > 0) ? c - P.W : c;
  x.n[3] = (j < P.H - 1) ? c + P.W : c;
  x.n[4] = (k > 0) ? c - HW : c;
  x.n[5] = (k < P.D - 1) ? c + HW : c;
  return x;
}
// C*k_f of the six faces (W,E,S,N,D,U) from the per-realisation table k_faces_ref builds with krg = 1
// (fl(x*1) = x: the same bits as face_perms; image faces hold the harmonic mean of the cell with itself)
__device__ __forceinline__ void face_perms_tab(const FaceLay& L, const float* __restrict__ fr, int W, int H, int i, int j, int k,
                                               float (&ckf)[6]) {
  const float* FE = fr;
  const float* FN = fr + L.nE;
  const float* FU = FN + L.nN;
  const int64_t e = ((int64_t)k * H + j) * L.WP + i, n = ((int64_t)k * (H + 1) + j) * W + i, u = ((int64_t)k * H + j) * W + i;
  ckf[0] = __ldg(FE + e); ckf[1] = __ldg(FE + e + 1);
  ckf[2] = __ldg(FN + n); ckf[3] = __ldg(FN + n + W);
  ckf[4] = __ldg(FU + u); ckf[5] = __ldg(FU + u + (int64_t)H * W);
}
// does the (i, j) column hold a connection in any layer?  (short well lists: one scan per thread and march;
// long ones: every cell searches)
__device__ __forceinline__ bool column_has_well_gc(const SrmDev& P, int col, int HW) {
  if (P.n_wells <= 0) return false;
  if (P.n_wells > 64) return true;
  bool any = false;
  for (int w = 0; w < P.n_wells; ++w) any |= (P.wells[w].cell % HW) == col;
  return any;
}

struct GcFwd {
  const float* faces;
  const float* kx; const int32_t* sample_real;
  const float* p0; const float* p1; const float* sg0; const float* sg1; const float* so0; const float* so1;
  const float* dt1; const float* dt2;
  const float* F; const float* W7;
  float* divqw; float* dom; float* dom_out;
  double* sse; double* s_mg; double* s_qg; double* s_mo; double* s_qo;
  int32_t B, R, tiles_x;
};

__global__ void __launch_bounds__(kThreads, 3) k_resid_fwd_gc(const __grid_constant__ SrmDev P, const __grid_constant__ GcFwd A) {
  // one thread per (j, i) column, marching over z: per-sample scalars (five IEEE divisions), the column's position and
  // its well flag are formed once per march, and the seven partial sums are reduced once per thread instead of once per cell
  __shared__ double red[7 * 32];
  const int b = blockIdx.y;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = P.H * P.W;
  const int r = srm_real_of(A.sample_real, b, A.B, A.R);
  const int64_t total = (int64_t)A.B * P.N;
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};   // dom^2, ibc^2, trn^2, sum mg cells, sum mo cells, sum qg, sum qo
  if (col < HW) {
    const int64_t base = (int64_t)b * P.N;
    auto F = [&](int f, int cell) { return A.F[(int64_t)f * total + base + cell]; };
    const int cj = col / P.W, ci = col - cj * P.W;
    const FaceLay FL = face_layout(P.D, P.H, P.W);
    const float* __restrict__ fr = A.faces + (int64_t)r * FL.per_real;
    const bool col_wells = column_has_well_gc(P, col, HW);
    const float d1 = A.dt1[b], d2 = A.dt2[b];
    const float idl[6] = {P.idx, P.idx, P.idy, P.idy, P.idz, P.idz};
    const float idt = __fdiv_rn(1.0f, __fmul_rn(P.Dc, d1));
    const float rho1 = __fadd_rn(1.0f, (d1 == 0.f) ? 0.f : __fdiv_rn(d2, d1));
    const float den = __fadd_rn(__fmul_rn(d1, d2), __fmul_rn(d2, d2));
    const float rte_d1 = __fdiv_rn(2.5e-8f, d1);                                  // :439-440
    const float d12 = __fadd_rn(d1, d2);
    const float mfac = __fmul_rn(__fmul_rn(P.dv, idt), P.phi);
    for (int k = 0; k < P.D; ++k) {
    const int c = k * HW + col;
    const CellIdx ix = cell_index3(P, ci, cj, k, c);
    const float p0 = A.p0[base + c], p1 = A.p1[base + c];
    const float sg0 = A.sg0[base + c], sg1 = A.sg1[base + c], so0 = A.so0[base + c], so1 = A.so1[base + c];
    float ckf[6];
    face_perms_tab(FL, fr, P.W, P.H, ci, cj, k, ckf);
    const float Mc[4] = {F(F_MGG, c), F(F_MGO, c), F(F_MOO, c), F(F_MOG, c)};      // gg, go, oo, og
    const float krg_c = F(F_KRG, c), kro_c = F(F_KRO, c);
    float pn[6], a[4][6];
#pragma unroll
    for (int f = 0; f < 6; ++f) {
      const int cn = ix.n[f];
      pn[f] = A.p1[base + cn];
      // potentials as written (physics_loss.py:538-541): "plus" faces nbr - cell, "minus" faces cell - nbr
      const float pot = (f & 1) ? __fsub_rn(pn[f], p1) : __fsub_rn(p1, pn[f]);
      const bool own = pot <= 0.f;                                                // :543-551
      const float krg_f = own ? krg_c : F(F_KRG, cn), kro_f = own ? kro_c : F(F_KRO, cn);
      const float Mn[4] = {F(F_MGG, cn), F(F_MGO, cn), F(F_MOO, cn), F(F_MOG, cn)};
#pragma unroll
      for (int X = 0; X < 4; ++X) {
        const float Mf = __fmul_rn(__fadd_rn(Mc[X], Mn[X]), 0.5f);                // :517-525
        const float kr = (X == 0 || X == 3) ? krg_f : kro_f;                      // gg, og: gas phase; go, oo: oil phase
        a[X][f] = __fmul_rn(__fmul_rn(__fmul_rn(ckf[f], __fmul_rn(kr, Mf)), idl[f]), idl[f]);   // :563-583
      }
    }
    // wells in this cell (scatter_nd sums duplicates)
    float q4[4] = {0.f, 0.f, 0.f, 0.f}, mask = 0.f;
    int wfirst = 0;
    const int64_t wt = (int64_t)A.B * P.n_wells;
    if (col_wells) {
      wfirst = well_lower_bound(P, c);
      for (int w = wfirst; w < P.n_wells && P.wells[w].cell == c; ++w) {
#pragma unroll
        for (int X = 0; X < 4; ++X) q4[X] = __fadd_rn(q4[X], A.W7[X * wt + (int64_t)b * P.n_wells + w]);
        mask += 1.f;
      }
    }
    // divergence of each component                               physics_loss.py:590-611
    float divq[4];
#pragma unroll
    for (int X = 0; X < 4; ++X) {
      float s = __fadd_rn(-__fmul_rn(a[X][0], pn[0]), -__fmul_rn(a[X][2], pn[2]));
      const float asum = __fadd_rn(__fadd_rn(__fadd_rn(a[X][0], a[X][2]), a[X][1]), a[X][3]);
      s = __fadd_rn(s, __fmul_rn(asum, p1));
      s = __fadd_rn(s, -__fmul_rn(a[X][1], pn[1]));
      s = __fadd_rn(s, -__fmul_rn(a[X][3], pn[3]));
      s = __fadd_rn(s, __fadd_rn(__fmul_rn(a[X][4], __fsub_rn(p1, pn[4])), __fmul_rn(a[X][5], __fsub_rn(p1, pn[5]))));   // 3-D extension
      s = __fadd_rn(s, (mask != 0.f) ? __fdiv_rn(q4[X], P.dv) : 0.0f);          // off-well: 0/dv = +0 without the division
      divq[X] = __fmul_rn(P.dv, s);
    }
    // accumulation                                               physics_loss.py:465-466,506-514,557-586
    const float A0 = F(F_A0, c), B0 = F(F_B0, c), Rs0 = F(F_RS0, c), Rv0 = F(F_RV0, c);
    const float dA0 = F(F_DA0, c), dB0 = F(F_DB0, c), dRs0 = F(F_DRS0, c), dRv0 = F(F_DRV0, c);
    const float a1 = F(F_A1, c), b1 = F(F_B1, c), r1 = F(F_R1, c), v1 = F(F_V1, c);
    const float R0 = __fmul_rn(Rs0, B0), V0 = __fmul_rn(Rv0, A0);                 // :343-344
    const float dpc = __fsub_rn(p1, p0);
    const float dSg = (dpc == 0.f) ? 0.f : div_z(__fsub_rn(sg1, sg0), dpc);       // :465
    const float dSo = (dpc == 0.f) ? 0.f : div_z(__fsub_rn(so1, so0), dpc);       // :466
    const float dR0 = __fadd_rn(__fmul_rn(Rs0, dB0), __fmul_rn(B0, dRs0));        // :511
    const float dV0 = __fadd_rn(__fmul_rn(Rv0, dA0), __fmul_rn(A0, dRv0));        // :513
    auto cpX = [&](float prop1, float dS, float s0, float dprop0, float prop0) {
      const float cpr = __fmul_rn(P.phicf, prop0);                                // :557-560
      const float t1 = __fmul_rn(__fmul_rn(P.phi, prop1), dS);
      const float t2 = __fmul_rn(s0, __fadd_rn(__fmul_rn(P.phi, dprop0), cpr));
      return __fmul_rn(__fmul_rn(idt, __fadd_rn(t1, t2)), dpc);                   // :572-573,585-586
    };
    const float cpgg = cpX(a1, dSg, sg0, dA0, A0), cpgo = cpX(r1, dSo, so0, dR0, R0);
    const float cpoo = cpX(b1, dSo, so0, dB0, B0), cpog = cpX(v1, dSg, sg0, dV0, V0);
    const float dom_gg = __fadd_rn(divq[0], __fmul_rn(P.dv, cpgg)), dom_go = __fadd_rn(divq[1], __fmul_rn(P.dv, cpgo));
    const float dom_oo = __fadd_rn(divq[2], __fmul_rn(P.dv, cpoo)), dom_og = __fadd_rn(divq[3], __fmul_rn(P.dv, cpog));
    const float dom = __fadd_rn(__fadd_rn(dom_gg, dom_go), __fadd_rn(dom_oo, dom_og));      // :638
    const float divq_tot = __fadd_rn(__fadd_rn(divq[0], divq[1]), __fadd_rn(divq[2], divq[3]));
    const float ibc = __fmul_rn(mask, divq_tot);                                  // :650
    // masses and truncation terms                                physics_loss.py:419-441
    auto trnX = [&](float m0, float m1) {
      const float m2 = __fadd_rn(__fmul_rn(__fsub_rn(m1, m0), rho1), m0);
      const float num = __fsub_rn(__fadd_rn(__fmul_rn(d2, m0), __fmul_rn(d1, m2)), __fmul_rn(d12, m1));
      return __fmul_rn(P.dvDc, __fadd_rn(rte_d1, div_z(num, den)));
    };
    const float mg0 = __fmul_rn(P.phi, __fadd_rn(__fmul_rn(A0, sg0), __fmul_rn(R0, so0)));
    const float mo0 = __fmul_rn(P.phi, __fadd_rn(__fmul_rn(B0, so0), __fmul_rn(V0, sg0)));
    const float mg1 = __fmul_rn(P.phi, __fadd_rn(__fmul_rn(a1, sg1), __fmul_rn(r1, so1)));
    const float mo1 = __fmul_rn(P.phi, __fadd_rn(__fmul_rn(b1, so1), __fmul_rn(v1, sg1)));
    const float trn = __fadd_rn(trnX(mg0, mg1), trnX(mo0, mo1));                  // :637
    // material balance summands                                  physics_loss.py:655-662
    const float mb_gg = __fmul_rn(mfac, __fsub_rn(__fmul_rn(sg1, a1), __fmul_rn(sg0, A0)));
    const float mb_go = __fmul_rn(mfac, __fsub_rn(__fmul_rn(so1, r1), __fmul_rn(so0, R0)));
    const float mb_oo = __fmul_rn(mfac, __fsub_rn(__fmul_rn(so1, b1), __fmul_rn(so0, B0)));
    const float mb_og = __fmul_rn(mfac, __fsub_rn(__fmul_rn(sg1, v1), __fmul_rn(sg0, V0)));
    A.dom[base + c] = dom;
    if (A.dom_out) A.dom_out[base + c] = dom;
    if (mask != 0.f)
      for (int w = wfirst; w < P.n_wells && P.wells[w].cell == c; ++w) A.divqw[(int64_t)b * P.n_wells + w] = divq_tot;
    acc[0] += (double)dom * (double)dom;
    if (mask != 0.f) acc[1] += (double)ibc * (double)ibc;
    acc[2] += (double)trn * (double)trn;
    acc[3] += (double)__fadd_rn(mb_gg, mb_go);
    acc[4] += (double)__fadd_rn(mb_oo, mb_og);
    if (mask != 0.f) { acc[5] += (double)__fadd_rn(q4[0], q4[1]); acc[6] += (double)__fadd_rn(q4[2], q4[3]); }
    }   // march
  }
  block_reduce<7>(acc, red);
  if (threadIdx.x == 0) {
    atomicAdd(&A.sse[SRM_TERM_DOM], acc[0]);
    if (acc[1] != 0.0) atomicAdd(&A.sse[SRM_TERM_IBC], acc[1]);
    atomicAdd(&A.sse[SRM_TERM_CMBC], acc[2]);
    atomicAdd(&A.s_mg[b], acc[3]);
    atomicAdd(&A.s_mo[b], acc[4]);
    if (acc[5] != 0.0) atomicAdd(&A.s_qg[b], acc[5]);
    if (acc[6] != 0.0) atomicAdd(&A.s_qo[b], acc[6]);
  }
}

// mbc_b = (-sum qg - sum mbc_g cells) + (-sum qo - sum mbc_o cells); terms and counts      physics_loss.py:661-665,800-832
__global__ void k_finalize_fwd_gc(const __grid_constant__ SrmDev P, int32_t B, double* __restrict__ sse,
                                  const double* __restrict__ s_mg, const double* __restrict__ s_qg,
                                  const double* __restrict__ s_mo, const double* __restrict__ s_qo,
                                  float* __restrict__ mbc, float* __restrict__ terms_out) {
  __shared__ double red[32];
  double v[1] = {0.0};
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float mg = __fsub_rn(-(float)s_qg[b], (float)s_mg[b]);
    const float mo = __fsub_rn(-(float)s_qo[b], (float)s_mo[b]);
    const float m = __fadd_rn(mg, mo);
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
      if (t == SRM_TERM_DOM || t == SRM_TERM_IBC || t == SRM_TERM_CMBC) cnt = n;
      if (t == SRM_TERM_MBC) cnt = n;      // the reference counts mbc with the ic field's shape (physics_loss.py:830)
      terms_out[SRM_N_TERMS + t] = (float)cnt;
    }
  }
}

// ---- adjoint -----------------------------------------------------------------------------------
struct GcAdj {
  const float* faces;
  const float* kx; const int32_t* sample_real;
  const float* p0; const float* p1; const float* sg0; const float* sg1; const float* so0; const float* so1;
  const float* dt1; const float* dt2; const float* dterms;
  const float* F; const float* W7; const float* divqw; const float* dom; const float* mbc;
  float* gp0; float* gp1; float* gsg0; float* gsg1; float* gso0; float* gso1;
  double* gdt1_acc; double* gdt2_acc;
  int32_t B, R, tiles_x;
};

// per face, seen from cell c with neighbour n: Lc / Ln = sum over the four components of kr_sel * <M>_f as the
// cell / the neighbour selects the relative permeability; dLc_p, dLn_p = their derivative w.r.t. p1 of c (through
// <M>_f); dL_s = derivative w.r.t. Sg1 of c of whichever view selects c.
struct FaceAdj { float Lc, Ln, dLc_p, dLn_p, dLc_s, dLn_s; };
__device__ __forceinline__ FaceAdj face_adj(bool own, float krg_c, float kro_c, float krg_n, float kro_n, float Mg_c, float Mo_c,
                                            float Mg_n, float Mo_n, float dMg_c, float dMo_c, float dkrg_c, float dkro_c) {
  const float hMg = 0.5f * (Mg_c + Mg_n), hMo = 0.5f * (Mo_c + Mo_n);
  // c's view selects c when own, n's view selects n when own (same potential, see kernels_gc.cu header of this block)
  const float kg_c = own ? krg_c : krg_n, ko_c = own ? kro_c : kro_n;
  const float kg_n = own ? krg_n : krg_c, ko_n = own ? kro_n : kro_c;
  FaceAdj r;
  r.Lc = kg_c * hMg + ko_c * hMo;
  r.Ln = kg_n * hMg + ko_n * hMo;
  r.dLc_p = 0.5f * (kg_c * dMg_c + ko_c * dMo_c);
  r.dLn_p = 0.5f * (kg_n * dMg_c + ko_n * dMo_c);
  const float ds = dkrg_c * hMg + dkro_c * hMo;
  r.dLc_s = own ? ds : 0.f;
  r.dLn_s = own ? 0.f : ds;
  return r;
}

__global__ void __launch_bounds__(kThreads, 3) k_resid_adj_gc(const __grid_constant__ SrmDev P, const __grid_constant__ GcAdj A) {
  // one thread per (j, i) column, marching over z (see k_resid_fwd_gc)
  __shared__ double red[2 * 32];
  const int b = blockIdx.y;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = P.H * P.W;
  const int r = srm_real_of(A.sample_real, b, A.B, A.R);
  const int64_t total = (int64_t)A.B * P.N;
  double acc2[2] = {0.0, 0.0};
  if (col < HW) {
    const int64_t base = (int64_t)b * P.N;
    auto F = [&](int f, int cell) { return A.F[(int64_t)f * total + base + cell]; };
    const int cj = col / P.W, ci = col - cj * P.W;
    const FaceLay FL = face_layout(P.D, P.H, P.W);
    const float* __restrict__ fr = A.faces + (int64_t)r * FL.per_real;
    const bool col_wells = column_has_well_gc(P, col, HW);
    const float w_dom = A.dterms[SRM_TERM_DOM], w_mbc = A.dterms[SRM_TERM_MBC], w_trn = A.dterms[SRM_TERM_CMBC];
    const float d1 = A.dt1[b], d2 = A.dt2[b];
    const float smb = 2.f * w_mbc * A.mbc[b];                // dL/d mbc_b
    const float idl[6] = {P.idx, P.idx, P.idy, P.idy, P.idz, P.idz};
    const float idt = 1.0f / (P.Dc * d1);
    const float id1 = 1.0f / d1;
    const float mfac = P.dv * idt * P.phi;
    const float smf = smb * mfac;
    // forward's per-sample scalars of the truncation term, in the forward's op order
    const float rho1 = __fadd_rn(1.0f, (d1 == 0.f) ? 0.f : __fdiv_rn(d2, d1));
    const float den = __fadd_rn(__fmul_rn(d1, d2), __fmul_rn(d2, d2));
    const float rte_d1 = __fdiv_rn(2.5e-8f, d1);
    const float d12 = __fadd_rn(d1, d2);
    const float iden2 = 1.0f / (den * den);
    const float dE1c = -2.f * 2.5e-8f / (d1 * d1);
    for (int k = 0; k < P.D; ++k) {
    const int c = k * HW + col;
    const CellIdx ix = cell_index3(P, ci, cj, k, c);
    const float p0 = A.p0[base + c], p1 = A.p1[base + c];
    const float sg0 = A.sg0[base + c], sg1 = A.sg1[base + c], so0 = A.so0[base + c], so1 = A.so1[base + c];
    const float sc = 2.f * w_dom * A.dom[base + c];          // dL/d dom_c
    float ckf[6];
    face_perms_tab(FL, fr, P.W, P.H, ci, cj, k, ckf);
    const float Mg_c = F(F_MGG, c) + F(F_MOG, c), Mo_c = F(F_MGO, c) + F(F_MOO, c);
    const float krg_c = F(F_KRG, c), kro_c = F(F_KRO, c);
    const float dMg_c = F(F_DMG, c), dMo_c = F(F_DMO, c), dkrg_c = F(F_DKRG, c), dkro_c = F(F_DKRO, c);
    float g1 = 0.f, gs1 = 0.f;
    // ---- divergence part: gather over the cell's own residual and its six neighbours' residuals
#pragma unroll
    for (int f = 0; f < 6; ++f) {
      const int cn = ix.n[f];
      if (cn == c) continue;                                 // image face: both views identical, no net contribution
      const float pn = A.p1[base + cn];
      const float sn = 2.f * w_dom * A.dom[base + cn];
      const float pot = (f & 1) ? (pn - p1) : (p1 - pn);
      const bool own = pot <= 0.f;
      const FaceAdj fa = face_adj(own, krg_c, kro_c, F(F_KRG, cn), F(F_KRO, cn), Mg_c, Mo_c, F(F_MGG, cn) + F(F_MOG, cn),
                                  F(F_MGO, cn) + F(F_MOO, cn), dMg_c, dMo_c, dkrg_c, dkro_c);
      const float Tf = ckf[f] * idl[f] * idl[f];
      const float dpf = p1 - pn;
      g1 += Tf * ((sc * fa.Lc - sn * fa.Ln) + dpf * (sc * fa.dLc_p - sn * fa.dLn_p));
      gs1 += Tf * dpf * (sc * fa.dLc_s - sn * fa.dLn_s);
    }
    g1 *= P.dv;
    gs1 *= P.dv;
    // ---- local terms
    float m0;
    (void)srm_clamp(P, p0, m0);
    const float A0 = F(F_A0, c), B0 = F(F_B0, c), Rs0 = F(F_RS0, c), Rv0 = F(F_RV0, c);
    const float dA0 = F(F_DA0, c), dB0 = F(F_DB0, c), dRs0 = F(F_DRS0, c), dRv0 = F(F_DRV0, c);
    const float d2A0 = F(F_D2A0, c), d2B0 = F(F_D2B0, c), d2Rs0 = F(F_D2RS0, c), d2Rv0 = F(F_D2RV0, c);
    const float a1 = F(F_A1, c), b1 = F(F_B1, c), r1 = F(F_R1, c), v1 = F(F_V1, c);
    const float da1 = F(F_DA1, c), db1 = F(F_DB1, c), dr1 = F(F_DR1, c), dv1 = F(F_DV1, c);
    const float R0 = Rs0 * B0, V0 = Rv0 * A0;
    const float dR0 = Rs0 * dB0 + B0 * dRs0, dV0 = Rv0 * dA0 + A0 * dRv0;
    // d/dp0 of the n0 quantities (first derivatives masked by the clamp, second derivatives staged masked)
    const float pA0 = dA0 * m0, pB0 = dB0 * m0, pR0 = dR0 * m0, pV0 = dV0 * m0;
    const float pdA0 = d2A0, pdB0 = d2B0;
    const float pdR0 = 2.f * dRs0 * dB0 * m0 + Rs0 * d2B0 + B0 * d2Rs0;
    const float pdV0 = 2.f * dRv0 * dA0 * m0 + Rv0 * d2A0 + A0 * d2Rv0;
    const float dpc = p1 - p0;
    const float nz = (dpc == 0.f) ? 0.f : 1.f;               // divide_no_nan: the chord-slope terms vanish with dpc
    const float dSgS = (sg1 - sg0) * nz, dSoS = (so1 - so0) * nz;
    // acc = dv*idt*( phi*(a1+v1)*dSg*dpc + phi*(r1+b1)*dSo*dpc + dpc*(sg0*Kg + so0*Ko) ),  dS*dpc = S1-S0
    const float Kg = P.phi * (dA0 + dV0) + P.phicf * (A0 + V0);
    const float Ko = P.phi * (dR0 + dB0) + P.phicf * (R0 + B0);
    const float pKg = P.phi * (pdA0 + pdV0) + P.phicf * (pA0 + pV0);
    const float pKo = P.phi * (pdR0 + pdB0) + P.phicf * (pR0 + pB0);
    const float sacc = sc * P.dv * idt;
    g1 += sacc * (P.phi * ((da1 + dv1) * dSgS + (dr1 + db1) * dSoS) + (sg0 * Kg + so0 * Ko));
    float g0 = sacc * (dpc * (sg0 * pKg + so0 * pKo) - (sg0 * Kg + so0 * Ko));
    gs1 += sacc * P.phi * (a1 + v1) * nz;
    float gs0 = sacc * (dpc * Kg - P.phi * (a1 + v1) * nz);
    float go1 = sacc * P.phi * (r1 + b1) * nz;
    float go0 = sacc * (dpc * Ko - P.phi * (r1 + b1) * nz);
    const float acc_tot = P.dv * idt * (P.phi * ((a1 + v1) * dSgS + (r1 + b1) * dSoS) + dpc * (sg0 * Kg + so0 * Ko));
    // material balance: mbc_b = -sum q - sum mcell
    const float mcell = mfac * ((sg1 * a1 - sg0 * A0) + (so1 * r1 - so0 * R0) + (so1 * b1 - so0 * B0) + (sg1 * v1 - sg0 * V0));
    g1 -= smf * (sg1 * (da1 + dv1) + so1 * (dr1 + db1));
    g0 += smf * (sg0 * (pA0 + pV0) + so0 * (pR0 + pB0));
    gs1 -= smf * (a1 + v1);
    gs0 += smf * (A0 + V0);
    go1 -= smf * (r1 + b1);
    go0 += smf * (R0 + B0);
    // wells in this cell: sum of the four rates enters dom (+) and mbc (-)
    if (col_wells) {
      const int64_t wt = (int64_t)A.B * P.n_wells;
      const int first = well_lower_bound(P, c);
      for (int w = first; w < P.n_wells && P.wells[w].cell == c; ++w) {
        const float dqp = A.W7[4 * wt + (int64_t)b * P.n_wells + w], dqs = A.W7[5 * wt + (int64_t)b * P.n_wells + w];
        g1 += (sc - smb) * dqp;
        gs1 += (sc - smb) * dqs;
      }
    }
    // truncation term (cmbc): trn = (dv/Dc) * (2*rte/d1 + (Ng + No)/den); the brackets N vanish identically for the
    // linear extrapolation, so only the explicit time steps carry a gradient (N itself is rounding noise, evaluated in
    // the forward's op order)
    {
      const float R0f = __fmul_rn(Rs0, B0), V0f = __fmul_rn(Rv0, A0);
      auto numX = [&](float mm0, float mm1) {
        const float m2 = __fadd_rn(__fmul_rn(__fsub_rn(mm1, mm0), rho1), mm0);
        return __fsub_rn(__fadd_rn(__fmul_rn(d2, mm0), __fmul_rn(d1, m2)), __fmul_rn(d12, mm1));
      };
      const float mg0 = __fmul_rn(P.phi, __fadd_rn(__fmul_rn(A0, sg0), __fmul_rn(R0f, so0)));
      const float mo0 = __fmul_rn(P.phi, __fadd_rn(__fmul_rn(B0, so0), __fmul_rn(V0f, sg0)));
      const float mg1 = __fmul_rn(P.phi, __fadd_rn(__fmul_rn(a1, sg1), __fmul_rn(r1, so1)));
      const float mo1 = __fmul_rn(P.phi, __fadd_rn(__fmul_rn(b1, so1), __fmul_rn(v1, sg1)));
      const float Ng = numX(mg0, mg1), No = numX(mo0, mo1);
      const float trn = __fadd_rn(__fmul_rn(P.dvDc, __fadd_rn(rte_d1, div_z(Ng, den))),
                                  __fmul_rn(P.dvDc, __fadd_rn(rte_d1, div_z(No, den))));
      const float st = 2.f * w_trn * trn;
      const float dE1 = dE1c - (Ng + No) * d2 * iden2;
      const float dE2 = -(Ng + No) * (d1 + 2.f * d2) * iden2;
      acc2[0] += (double)((smb * mcell - sc * acc_tot) * id1 + st * P.dvDc * dE1);
      acc2[1] += (double)(st * P.dvDc * dE2);
    }
    A.gp0[base + c] = g0;
    A.gp1[base + c] = g1;
    A.gsg0[base + c] = gs0;
    A.gsg1[base + c] = gs1;
    A.gso0[base + c] = go0;
    A.gso1[base + c] = go1;
    }   // march
  }
  block_reduce<2>(acc2, red);
  if (threadIdx.x == 0) {
    atomicAdd(&A.gdt1_acc[b], acc2[0]);
    atomicAdd(&A.gdt2_acc[b], acc2[1]);
  }
}

// inner-boundary term L_ibc = w_ibc * sum (mask*divq_tot)^2: scatter d divq_tot(c)/d(p1, Sg1) to the well cell and
// its six neighbours (atomics: neighbouring well cells may hit the same target).
__global__ void __launch_bounds__(128) k_ibc_adj_gc(const __grid_constant__ SrmDev P, const __grid_constant__ GcAdj A) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int nw = P.n_wells;
  if (g >= (int64_t)A.B * nw) return;
  const int b = (int)(g / nw), w = (int)(g % nw);
  const int c = P.wells[w].cell;
  if (w > 0 && P.wells[w - 1].cell == c) return;   // one thread per distinct cell
  const int64_t wt = (int64_t)A.B * nw;
  float mask = 0.f, dqp = 0.f, dqs = 0.f;
  for (int u = w; u < nw && P.wells[u].cell == c; ++u) {
    mask += 1.f;
    dqp += A.W7[4 * wt + (int64_t)b * nw + u];
    dqs += A.W7[5 * wt + (int64_t)b * nw + u];
  }
  const float s = 2.f * A.dterms[SRM_TERM_IBC] * mask * mask * A.divqw[g];
  if (s == 0.f) return;
  const int r = srm_real_of(A.sample_real, b, A.B, A.R);
  const int64_t total = (int64_t)A.B * P.N, base = (int64_t)b * P.N;
  auto F = [&](int f, int cell) { return A.F[(int64_t)f * total + base + cell]; };
  const CellIdx ix = cell_index(P, c);
  float ckf[6];
  face_perms(P, A.kx + (int64_t)r * P.N, c, ix, ckf);
  const float idl[6] = {P.idx, P.idx, P.idy, P.idy, P.idz, P.idz};
  const float p1 = A.p1[base + c];
  const float Mg_c = F(F_MGG, c) + F(F_MOG, c), Mo_c = F(F_MGO, c) + F(F_MOO, c);
  const float krg_c = F(F_KRG, c), kro_c = F(F_KRO, c);
  float self_p = 0.f, self_s = 0.f;
  for (int f = 0; f < 6; ++f) {
    const int cn = ix.n[f];
    if (cn == c) continue;
    const float pn = A.p1[base + cn];
    const float pot = (f & 1) ? (pn - p1) : (p1 - pn);
    const bool own = pot <= 0.f;
    const float krg_n = F(F_KRG, cn), kro_n = F(F_KRO, cn);
    const float Mg_n = F(F_MGG, cn) + F(F_MOG, cn), Mo_n = F(F_MGO, cn) + F(F_MOO, cn);
    const float hMg = 0.5f * (Mg_c + Mg_n), hMo = 0.5f * (Mo_c + Mo_n);
    const float kg = own ? krg_c : krg_n, ko = own ? kro_c : kro_n;
    const float L = kg * hMg + ko * hMo;
    const float Tf = ckf[f] * idl[f] * idl[f] * P.dv;
    const float dpf = p1 - pn;
    // divq_tot(c) = dv * sum_f Tf' * L * (p_c - p_n) + sum q
    self_p += Tf * (L + dpf * 0.5f * (kg * F(F_DMG, c) + ko * F(F_DMO, c)));
    atomicAdd(&A.gp1[base + cn], s * Tf * (-L + dpf * 0.5f * (kg * F(F_DMG, cn) + ko * F(F_DMO, cn))));
    if (own) self_s += Tf * dpf * (F(F_DKRG, c) * hMg + F(F_DKRO, c) * hMo);
    else atomicAdd(&A.gsg1[base + cn], s * Tf * dpf * (F(F_DKRG, cn) * hMg + F(F_DKRO, cn) * hMo));
  }
  atomicAdd(&A.gp1[base + c], s * (self_p + dqp));
  atomicAdd(&A.gsg1[base + c], s * (self_s + dqs));
}

#include "gc_fused.cuh"

}  // namespace

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
int srm_launch_relperm(const SrmHandle* h, int64_t n, const float* sg, float* krog, float* krgo, float* dkrog, float* dkrgo, cudaStream_t s) {
  if (n <= 0) return SRM_OK;
  k_relperm<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, s>>>(h->dev, n, sg, krog, krgo, dkrog, dkrgo);
  SRM_CUDA_CHECK(cudaGetLastError());
  return SRM_OK;
}

int srm_build_pvt_lut_gc(SrmHandle* h, float lo, float hi) {
  SrmDev& P = h->dev;
  uint32_t lo_bits, hi_bits;
  memcpy(&lo_bits, &lo, 4);
  memcpy(&hi_bits, &hi, 4);
  if (!(lo > 0.f) || !(hi >= lo)) { srm_set_error("srm_create: pvt_lut range [%g, %g] must be positive and ascending", lo, hi); return SRM_ERR_INVALID; }
  const uint64_t n = (uint64_t)hi_bits - lo_bits + 1;
  if (n > (1ull << 31)) { srm_set_error("srm_create: pvt_lut range too wide"); return SRM_ERR_INVALID; }
  // fused pair (gc_fused.cuh): needs the table over the whole clamp range; SRM_NO_GC2 keeps the staged pipeline
  h->gc_fused = (h->lut_full && !getenv("SRM_NO_GC2")) ? 1 : 0;
  // staged: 3 + 4 float4 per pressure; fused: 3 (level n) + 2 + 2 (forward views) + 2 (adjoint view of level n+1)
  const uint64_t per = h->gc_fused ? 9 : 7;
  cudaError_t e = cudaMalloc((void**)&h->d_lut, n * per * sizeof(float4) + (h->gc_fused ? n * sizeof(float2) : 0));
  if (e != cudaSuccess) { srm_set_error("srm_create: pvt_lut (GC) needs %.1f MB of device memory: %s", n * per * 16e-6, cudaGetErrorString(e)); return SRM_ERR_CUDA; }
  P.lut_lo_bits = lo_bits;
  P.lut_n = (uint32_t)n;
  P.lut0 = h->d_lut;
  float4* t1 = h->gc_fused ? nullptr : h->d_lut + 3 * n;           // staged 64-byte pack of level n+1
  float4* f0 = h->gc_fused ? h->d_lut + 3 * n : nullptr;
  float4* f1 = h->gc_fused ? h->d_lut + 5 * n : nullptr;
  float4* a1 = h->gc_fused ? h->d_lut + 7 * n : nullptr;
  float2* fv = h->gc_fused ? reinterpret_cast<float2*>(h->d_lut + 9 * n) : nullptr;
  P.lut1 = h->gc_fused ? a1 : t1;                                  // fused: the adjoint's 32-byte view (gc_fused.cuh)
  P.lutf0 = reinterpret_cast<const float2*>(f0); P.lutf1 = reinterpret_cast<const float2*>(f1);
  P.gcv = fv;
  k_lut_build_gc<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads>>>(P, h->d_lut, t1, f0, f1, fv, a1);
  SRM_CUDA_CHECK(cudaGetLastError());
  SRM_CUDA_CHECK(cudaDeviceSynchronize());
  return SRM_OK;
}

int srm_forward_gc_impl(SrmHandle* h, int32_t B, int32_t R, const float* kx, const int32_t* sample_real, const float* p0,
                        const float* p1, const float* sg0, const float* sg1, const float* so0, const float* so1,
                        const float* dt1, const float* dt2, const float* t1, float* terms_out, float* dom_out,
                        const SrmWs& ws, bool save, cudaStream_t s) {
  const SrmDev& P = h->dev;
  const int64_t total = (int64_t)B * P.N;
  SRM_CUDA_CHECK(cudaMemsetAsync(ws.sse, 0, (char*)ws.mbc - (char*)ws.sse, s));   // sse and the four per-sample sums
  const unsigned sblocks = (unsigned)((total + kThreads - 1) / kThreads);
  const bool lut = P.lut_n > 0;
  const bool fused = h->gc_fused != 0;
  if (fused) {}
  else if (save) { if (lut) k_stage_gc<true, true><<<sblocks, kThreads, 0, s>>>(P, total, p0, p1, sg1, ws.gc);
              else k_stage_gc<true, false><<<sblocks, kThreads, 0, s>>>(P, total, p0, p1, sg1, ws.gc); }
  else      { if (lut) k_stage_gc<false, true><<<sblocks, kThreads, 0, s>>>(P, total, p0, p1, sg1, ws.gc);
              else k_stage_gc<false, false><<<sblocks, kThreads, 0, s>>>(P, total, p0, p1, sg1, ws.gc); }
  SRM_CUDA_CHECK(cudaGetLastError());
  const int64_t nwt = (int64_t)B * P.n_wells;
  if (nwt > 0) {
    k_wells_gc<<<(unsigned)((nwt + 127) / 128), 128, 0, s>>>(P, B, R, kx, sample_real, p1, sg1, t1, ws.gc_wells, ws.pwfw);
    SRM_CUDA_CHECK(cudaGetLastError());
  }
  GcFwd A;
  memset(&A, 0, sizeof(A));
  A.kx = kx; A.sample_real = sample_real; A.p0 = p0; A.p1 = p1; A.sg0 = sg0; A.sg1 = sg1; A.so0 = so0; A.so1 = so1;
  A.dt1 = dt1; A.dt2 = dt2; A.F = ws.gc; A.W7 = ws.gc_wells; A.divqw = ws.divqw; A.dom = ws.dom; A.dom_out = dom_out;
  A.sse = ws.sse; A.s_mg = ws.mb_sum; A.s_qg = ws.q_sum; A.s_mo = ws.gdt1_acc; A.s_qo = ws.gdt2_acc;   // the adjoint re-zeroes its accumulators
  A.B = B; A.R = R;
  {
    // static face coefficients C*k_f per realisation: k_faces_ref (ref_fused.cuh) with krg = 1, fl(x*1) = x
    SrmDev P1 = P;
    P1.krg = 1.0f;
    const FaceLay FL = face_layout(P.D, P.H, P.W);
    k_faces_ref<<<dim3((unsigned)((FL.per_real + 255) / 256), (unsigned)R), 256, 0, s>>>(P1, kx, ws.faces);
    SRM_CUDA_CHECK(cudaGetLastError());
  }
  A.faces = ws.faces;
  if (fused) {
    A.tiles_x = (P.W + G2X - 1) / G2X;
    const dim3 grid((unsigned)(A.tiles_x * ((P.H + G2Y - 1) / G2Y)), (unsigned)B);
    const bool dflt = P.nog_i == 3 && P.ng_i == 6;      // the reference's Corey exponents (default_configurations.py:266)
    if (P.n_wells > 64) { if (dflt) k_fwd_gc2<true, 3, 6><<<grid, kThreads, 0, s>>>(P, A); else k_fwd_gc2<true, 0, 0><<<grid, kThreads, 0, s>>>(P, A); }
    else { if (dflt) k_fwd_gc2<false, 3, 6><<<grid, kThreads, 0, s>>>(P, A); else k_fwd_gc2<false, 0, 0><<<grid, kThreads, 0, s>>>(P, A); }
  } else {
    const dim3 grid((unsigned)((P.H * P.W + kThreads - 1) / kThreads), (unsigned)B);
    k_resid_fwd_gc<<<grid, kThreads, 0, s>>>(P, A);
  }
  SRM_CUDA_CHECK(cudaGetLastError());
  k_finalize_fwd_gc<<<1, 256, 0, s>>>(P, B, ws.sse, ws.mb_sum, ws.q_sum, ws.gdt1_acc, ws.gdt2_acc, ws.mbc, terms_out);
  SRM_CUDA_CHECK(cudaGetLastError());
  return SRM_OK;
}

int srm_backward_gc_impl(SrmHandle* h, int32_t B, int32_t R, const float* kx, const int32_t* sample_real, const float* p0,
                         const float* p1, const float* sg0, const float* sg1, const float* so0, const float* so1,
                         const float* dt1, const float* dt2, const float* t1, const float* dterms, float* gp0, float* gp1,
                         float* gsg0, float* gsg1, float* gso0, float* gso1, float* gdt1, float* gdt2, const SrmWs& ws,
                         cudaStream_t s) {
  const SrmDev& P = h->dev;
  (void)t1;
  SRM_CUDA_CHECK(cudaMemsetAsync(ws.gdt1_acc, 0, (char*)ws.mbc - (char*)ws.gdt1_acc, s));
  GcAdj A;
  memset(&A, 0, sizeof(A));
  A.kx = kx; A.sample_real = sample_real; A.p0 = p0; A.p1 = p1; A.sg0 = sg0; A.sg1 = sg1; A.so0 = so0; A.so1 = so1;
  A.dt1 = dt1; A.dt2 = dt2; A.dterms = dterms; A.F = ws.gc; A.W7 = ws.gc_wells; A.divqw = ws.divqw; A.dom = ws.dom; A.mbc = ws.mbc;
  A.gp0 = gp0; A.gp1 = gp1; A.gsg0 = gsg0; A.gsg1 = gsg1; A.gso0 = gso0; A.gso1 = gso1;
  A.gdt1_acc = ws.gdt1_acc; A.gdt2_acc = ws.gdt2_acc; A.B = B; A.R = R;
  A.faces = ws.faces;                                        // built by the forward (state in the workspace)
  const bool fused = h->gc_fused != 0;
  if (fused) {
    A.tiles_x = (P.W + G2X - 1) / G2X;
    const dim3 grid((unsigned)(A.tiles_x * ((P.H + G2Y - 1) / G2Y)), (unsigned)B);
    const bool dflt = P.nog_i == 3 && P.ng_i == 6;
    if (P.n_wells > 64) { if (dflt) k_adj_gc2<true, 3, 6><<<grid, kThreads, 0, s>>>(P, A); else k_adj_gc2<true, 0, 0><<<grid, kThreads, 0, s>>>(P, A); }
    else { if (dflt) k_adj_gc2<false, 3, 6><<<grid, kThreads, 0, s>>>(P, A); else k_adj_gc2<false, 0, 0><<<grid, kThreads, 0, s>>>(P, A); }
  } else {
    const dim3 grid((unsigned)((P.H * P.W + kThreads - 1) / kThreads), (unsigned)B);
    k_resid_adj_gc<<<grid, kThreads, 0, s>>>(P, A);
  }
  SRM_CUDA_CHECK(cudaGetLastError());
  const int64_t n = (int64_t)B * P.n_wells;
  if (n > 0) {
    if (fused) k_ibc_adj_gc2<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(P, A);
    else k_ibc_adj_gc<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(P, A);
    SRM_CUDA_CHECK(cudaGetLastError());
  }
  k_finalize_adj<<<(unsigned)((B + 255) / 256), 256, 0, s>>>(B, ws.gdt1_acc, ws.gdt2_acc, gdt1, gdt2);
  SRM_CUDA_CHECK(cudaGetLastError());
  return SRM_OK;
}
