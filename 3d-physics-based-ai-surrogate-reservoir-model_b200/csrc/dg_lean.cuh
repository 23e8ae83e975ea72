// Pieces shared by the lean dry-gas kernels (kernels_dg4.cu: every warp loads, gathers and computes;
// kernels_dg5.cu: warp-specialised producers / consumers): exact-table gathers with the index arithmetic folded
// into the base, mbarrier helpers, the guarded-once fast division.
#pragma once
#include "ref_fused.cuh"

namespace {

template <bool V> struct BoolC { static constexpr bool value = V; };

// ---- exact-table gathers ------------------------------------------------------------------------------
// The table covers [p_min, p_max] (lut_full), so the entry index is the clamped pressure's bit pattern minus
// bits(p_min); the subtraction is folded into the base address.  NaN clamps to p_min like srm_clamp.
// Entries are interleaved (k_lut_build): forward 16 bytes {invBg, cp | invBg, G} (level n | level n+1),
// adjoint 32 bytes {invBg, invBg', invBg'', cp | invBg, G, invBg', G'}.
struct Tab { uintptr_t base; float lo, hi; };
__device__ __forceinline__ Tab make_tab(const void* t, const SrmDev& P, int entry_bytes) {
  Tab r;
  r.base = reinterpret_cast<uintptr_t>(t) - (uintptr_t)P.lut_lo_bits * (uintptr_t)entry_bytes;
  r.lo = P.p_min; r.hi = P.p_max;
  return r;
}
template <int SHIFT, int BYTE, class V>
__device__ __forceinline__ V gat(const Tab& t, float p, uint64_t keep) {
  const float x = fminf(fmaxf(p, t.lo), t.hi);
#if defined(SRM_D4_ABL) && SRM_D4_ABL == 1     // timing ablation (tools/build_variants.sh): no gather at all
  V v; memset(&v, 0, sizeof(v)); v.x = x * 1e-4f; v.y = x * 2e-4f; return v;
#elif defined(SRM_D4_ABL) && SRM_D4_ABL == 2   // timing ablation: every gather hits a 4 KB window (L1 resident)
  return ld_hint(reinterpret_cast<const V*>(t.base + ((((uintptr_t)__float_as_uint(t.lo) << SHIFT)) + (((uintptr_t)__float_as_uint(x) & 127u) << SHIFT)) + BYTE), keep);
#else
  return ld_hint(reinterpret_cast<const V*>(t.base + ((uintptr_t)__float_as_uint(x) << SHIFT) + BYTE), keep);
#endif
}
#define GATF0(t, p) gat<4, 0, float2>(t, p, keep)    /* forward {invBg, cp} at level n */
#define GATF1(t, p) gat<4, 8, float2>(t, p, keep)    /* forward {invBg, G} at level n+1 */
#define GATA0(t, p) gat<5, 0, float4>(t, p, keep)    /* adjoint pack0 */
#define GATA1(t, p) gat<5, 16, float4>(t, p, keep)   /* adjoint pack1 */
#define GATA1G(t, p) gat<5, 16, float2>(t, p, keep)  /* {invBg, G} of the adjoint's table (halo cells) */

// ---- split CTA barrier (mbarrier): arrive after the plane is stored, wait just before the neighbours are read;
// the cell-local work of plane k+2 and the issue of the next loads and gathers sit between the two, so a warp
// that is late to store does not stall the others for as long.  One arrival per warp.
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_warp(uint64_t* bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0)
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, int parity) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(a), "r"(parity) : "memory");
  } while (!done);
}

// div_c's fast path without its guards (ref_fused.cuh); the caller tests the operands
__device__ __forceinline__ float div_fast(float a, const DivC& d) {
  const float q0 = __fmul_rn(a, d.y);
  const float q1 = __fmaf_rn(__fmaf_rn(-d.b, q0, a), d.y, q0);
  return __fmaf_rn(__fmaf_rn(-d.b, q1, a), d.y, q1);
}
// operands for which div_fast == div.rn; +-0 is allowed where the quotient is only ADDED to a non-zero
// number afterwards (the sign of a zero quotient is then immaterial)
__device__ __forceinline__ bool div_operand_bad(float a) {
  const float aa = fabsf(a);
  return !(aa <= 0x1p60f) || (aa < 0x1p-60f && aa != 0.f);
}


}  // namespace
