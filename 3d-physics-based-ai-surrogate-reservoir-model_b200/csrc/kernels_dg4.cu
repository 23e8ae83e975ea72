// SRM_NUMERICS_REFERENCE, exact PVT table over the whole clamp range, grids with W % CPT == 0: the lean
// forward / adjoint pair of the dry-gas physics loss (physics_loss.py:79-208, 742-870).
//
// Same arithmetic as kernels_ref2.cu (forward fields bit-identical to the pinned oracle); what changes is
// the instruction count per cell-timestep, which -- together with the table gathers -- bounds this path
// (profiles/):
//   * every value a cell shares with a neighbour is formed once: the x-face coefficient a3 of a cell IS a1 of
//     its E neighbour, the upper z-face flux term of plane k IS minus the lower term of plane k+1
//     (fl(a*(p-q)) = -fl(a*(q-p))), so the march carries ONE float per cell across planes instead of the
//     plane below (p, G, face coefficient);
//   * cp (physics_loss.py:149-150) comes out of the table with the spline values it is built from;
//   * wells are a block-uniform specialisation: tiles without a connection column run no well code at all;
//   * the two per-cell divisions by per-sample constants share one range test per thread and plane;
//   * the adjoint forms each face's pair (X, Y) = u*(G_c+G_n), u*(p_c-p_n), u = (s_c-s_n)*T_f/2 once and
//     hands it to both cells of the face (in registers along x and z).
//
//   * the table gathers of plane k+2 are issued right after the barrier of plane k and consumed at the END of
//     that iteration (the cell-local part of plane k+2: accumulation + truncation term, which needs no
//     neighbour) or in the next one (G), so no warp reaches the barrier waiting on a gather.
//
//   CTA tile  : TW (x) x TY (y) cells, CX x TY threads, CPT x-adjacent cells per thread, marching over z
//   shared    : double-buffered haloed planes of p1, G = invBg*invug (adjoint: and the seed 2 w dom), one
//               barrier per plane
#include <cstdlib>
#include <cstring>
#include "ref_fused.cuh"
#include "dg_lean.cuh"
#include "well_tile.cuh"

namespace {

#ifndef SRM_D4_CX
#define SRM_D4_CX 16
#endif
#ifndef SRM_D4_CPT
#define SRM_D4_CPT 2
#endif
#ifndef SRM_D4_TY
#define SRM_D4_TY 8
#endif
#ifndef SRM_D4_OCCF
#define SRM_D4_OCCF 4
#endif
#ifndef SRM_D4_OCCA
#define SRM_D4_OCCA 4
#endif
constexpr int CX = SRM_D4_CX, CPT = SRM_D4_CPT, TW = CX * CPT, TY = SRM_D4_TY, NT = CX * TY;
static_assert(CPT == 2 || CPT == 4, "cells per thread");
static_assert(CX <= 32 && (CX & (CX - 1)) == 0, "x threads: a power of two inside one warp");
static_assert(2 * TW + 2 * TY <= NT, "one halo cell per thread");
static_assert(NT % 32 == 0 && NT <= 1024, "whole warps");
constexpr int XO = 4;                  // column of the tile's first cell in a shared row (16-byte aligned)
constexpr int SW = TW + 2 * XO;        // [.. pad][XO-1: W halo][XO .. XO+TW-1 cells][XO+TW: E halo][pad ..]
constexpr int SH = TY + 2;
constexpr int PLANE = SH * SW;

// ---- CPT-wide moves ---------------------------------------------------------------------------------
#ifndef SRM_D4_NOALLOC
#define SRM_D4_NOALLOC 1
#endif
#if SRM_D4_NOALLOC
#define LD_STRM ld_stream
#else
#define LD_STRM ld_hint
#endif
__device__ __forceinline__ float2 ld_hint2(const float* p, uint64_t pol) { return LD_STRM(reinterpret_cast<const float2*>(p), pol); }
__device__ __forceinline__ void st_hint(float2* p, float2 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
// streamed fields: evict-first in L2 (ref_fused.cuh)
__device__ __forceinline__ void ldgs(const float* p, float (&v)[CPT], uint64_t pol) {
  if constexpr (CPT == 4) { const float4 t = LD_STRM(reinterpret_cast<const float4*>(p), pol); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  else { const float2 t = ld_hint2(p, pol); v[0] = t.x; v[1] = t.y; }
}
__device__ __forceinline__ void stgs(float* p, const float (&v)[CPT], uint64_t pol) {
  if constexpr (CPT == 4) st_hint(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]), pol);
  else st_hint(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]), pol);
}
// static face coefficients: shared by the T samples of a realisation, default caching
__device__ __forceinline__ void ldgc(const float* p, float* v) {
#if defined(SRM_D4_ABL) && SRM_D4_ABL == 3     // timing ablation: no face-coefficient loads at all (results are wrong)
  for (int c = 0; c < CPT; ++c) v[c] = 1e-3f + 1e-9f * (float)(reinterpret_cast<uintptr_t>(p) & 1023u);
  return;
#endif
  if constexpr (CPT == 4) { const float4 t = __ldg(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  else { const float2 t = __ldg(reinterpret_cast<const float2*>(p)); v[0] = t.x; v[1] = t.y; }
}
__device__ __forceinline__ void ldsv(const float* p, float (&v)[CPT]) {
  if constexpr (CPT == 4) { const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  else { const float2 t = *reinterpret_cast<const float2*>(p); v[0] = t.x; v[1] = t.y; }
}
__device__ __forceinline__ void stsv(float* p, const float (&v)[CPT]) {
  if constexpr (CPT == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  else *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
}

// what every thread knows about its place in the tile
struct Tile4 {
  int cx, ty, x0, y0;
  bool valid, edgeE, halo;
  int oc;                    // first own cell inside a plane (clamped to the grid)
  int h_off, h_slot;         // halo duty: cell offset inside a plane (clamped = edge replication), shared slot
};
__device__ __forceinline__ Tile4 make_tile4(const SrmDev& P, int tiles_x) {
  Tile4 t;
  const int tid = threadIdx.x, W = P.W, H = P.H;
  t.cx = tid & (CX - 1); t.ty = tid / CX;
  const int tyi = blockIdx.x / tiles_x, txi = blockIdx.x - tyi * tiles_x;
  t.x0 = txi * TW; t.y0 = tyi * TY;
  const int x = t.x0 + CPT * t.cx, y = t.y0 + t.ty;
  t.valid = x < W && y < H;
  const int xc = min(x, W - CPT), yc = min(y, H - 1);
  t.edgeE = xc + CPT >= W;
  t.oc = yc * W + xc;
  // halo duty: threads [0,TW) row y0-1, [TW,2TW) row y0+TY, then TY threads column x0-1, TY threads column x0+TW
  t.halo = tid < 2 * TW + 2 * TY;
  int gx = 0, gy = 0, hr = 0, hc = 0;
  if (tid < TW) { gy = t.y0 - 1; gx = t.x0 + tid; hr = 0; hc = XO + tid; }
  else if (tid < 2 * TW) { gy = t.y0 + TY; gx = t.x0 + tid - TW; hr = TY + 1; hc = XO + tid - TW; }
  else if (tid < 2 * TW + TY) { gx = t.x0 - 1; gy = t.y0 + tid - 2 * TW; hr = tid - 2 * TW + 1; hc = XO - 1; }
  else if (t.halo) { gx = t.x0 + TW; gy = t.y0 + tid - 2 * TW - TY; hr = tid - 2 * TW - TY + 1; hc = XO + TW; }
  t.h_off = min(max(gy, 0), H - 1) * W + min(max(gx, 0), W - 1);
  t.h_slot = hr * SW + hc;
  return t;
}

using WTile = WellTile<NT, TW, TY, 8, 192>;

// marks the threads that own a cell column with a well connection (any layer)
__device__ __forceinline__ bool thread_has_well4(const SrmDev& P, const Tile4& t, unsigned char (*s_flag)[TW]) {
  unsigned char* flat = &s_flag[0][0];
  for (int i = threadIdx.x; i < TY * TW; i += NT) flat[i] = 0;
  __syncthreads();
  const int HW = P.H * P.W;
  for (int w = threadIdx.x; w < P.n_wells; w += NT) {
    const int rem = P.wells[w].cell % HW;
    const int j = rem / P.W, i = rem - j * P.W;
    if (i >= t.x0 && i < t.x0 + TW && j >= t.y0 && j < t.y0 + TY) s_flag[j - t.y0][i - t.x0] = 1;
  }
  __syncthreads();
  bool any = false;
  if (t.valid) {
#pragma unroll
    for (int c = 0; c < CPT; ++c) any |= s_flag[t.ty][CPT * t.cx + c] != 0;
  }
  return any;
}

// ------------------------------------------------------------------------------------------
// forward                                                       physics_loss.py:143-193,787-807
// ------------------------------------------------------------------------------------------
// PK: the forward gathers the ADJOINT's 32-byte table entries (the same sectors as the forward's 16-byte ones) and
// stages the six values the adjoint needs per cell -- cp, A0', A0'', G, A1', G' -- in the workspace, so that k_adj4<true>
// runs without a single table gather (+24 B per cell-timestep each way through otherwise idle HBM bandwidth).
template <bool PK>
__global__ void __launch_bounds__(NT, SRM_D4_OCCF) k_fwd4(const __grid_constant__ SrmDev P, const __grid_constant__ R2Args A) {
  __shared__ __align__(16) float s_p[2 * PLANE];
  __shared__ __align__(16) float s_G[2 * PLANE];
  __shared__ double red[4 * 32];
  __shared__ __align__(8) WTile s_wt;
  __shared__ __align__(8) uint64_t s_bar;
  const Tile4 t = make_tile4(P, A.tiles_x);
  const int b = blockIdx.y;
  const int r = srm_real_of(A.sample_real, b, A.B, A.R);
  if (threadIdx.x == 0) mbar_init(&s_bar, NT / 32);
  __syncthreads();
  // the tile's connections: column lists staged in shared memory (well_tile.cuh); lists that do not fit keep the search
  bool tile_wells = false, wt_overflow = false, has_well = false;
  uint32_t wslots = 0;
  if (P.n_wells > 0) {
    tile_wells = s_wt.build(well_cols_of(P), P.W, P.D, t.x0, t.y0, A.qw + (int64_t)b * P.n_wells, wt_overflow);
    if (wt_overflow) {
      has_well = thread_has_well4(P, t, reinterpret_cast<unsigned char (*)[TW]>(s_wt.slot_of));
      tile_wells = __syncthreads_or(has_well ? 1 : 0) != 0;
    } else if (tile_wells && t.valid) {
      wslots = s_wt.template slots_of<CPT>(t.ty * TW + CPT * t.cx);
    }
  }
  const int W = P.W, H = P.H, D = P.D, HW = H * W;
  const FaceLay FL = face_layout(D, H, W);
  const float* __restrict__ p0f = A.p0 + (int64_t)b * P.N;
  const float* __restrict__ p1f = A.p1 + (int64_t)b * P.N;
  float* __restrict__ domf = A.dom + (int64_t)b * P.N;
  const float* __restrict__ FB = A.faces + (int64_t)r * FL.per_real;      // [FE | FN | FU]
  const int yy = t.oc / W, xx = t.oc - yy * W;
  const int strE = H * FL.WP, strN = (H + 1) * W;
  // per-sample scalars                                   physics_loss.py:126,156,171,193
  const float d1 = A.dt1[b], d2 = A.dt2[b];
  const float rho = (d1 == 0.f) ? 0.f : __fdiv_rn(d2, d1);
  const float one_rho = __fadd_rn(1.0f, rho);
  const DivC by_d1 = make_divc(d1);
  const DivC by_den = make_divc(__fadd_rn(__fmul_rn(d1, d2), __fmul_rn(d2, d2)));
  const float c2e7 = __fdiv_rn(2e-7f, d1);
  const float d12 = __fadd_rn(d1, d2);
  const float mbfac = __fdiv_rn(1.0f, __fmul_rn(P.Dc, d1));
  // the fast division is only used where its result equals div.rn's; anything else takes div_c (ref_fused.cuh).
  // c2e7 == 0 would expose the sign of a zero quotient in E = c2e7 + q.
  const bool div_slow = !(by_d1.ok && by_den.ok) || c2e7 == 0.f || !P.cp_safe;

  const uint64_t keep = l2_evict_last(), strm = l2_evict_first();
  const Tab TF = make_tab(PK ? (const void*)P.lut0 : (const void*)P.lutf0, P, PK ? 32 : 16);
  const int own_s = (t.ty + 1) * SW + XO + CPT * t.cx;       // own cells inside a shared plane
  float* __restrict__ pkb = PK ? A.pk + (int64_t)b * P.N : nullptr;
  const int64_t pks = A.pk_stride;
  // one pressure pair -> forward values {invBg, cp} / {invBg, G}; PK: plus {A0', A0''} / {A1', G'} for the adjoint pack
  auto gat2 = [&](float p1v, float p0v, float2& e1, float2& e0, float2& x1, float2& x0) {
    if constexpr (PK) {
      const float4 a = GATA1(TF, p1v), c = GATA0(TF, p0v);
      e1 = make_float2(a.x, a.y); x1 = make_float2(a.z, a.w);
      e0 = make_float2(c.x, c.w); x0 = make_float2(c.y, c.z);
    } else {
      e1 = GATF1(TF, p1v); e0 = GATF0(TF, p0v);
    }
  };

  float a_dom = 0.f, a_tde = 0.f, a_mbf = 0.f;
  double a_ibc = 0.0, a_mb = 0.0;

  // cell-local part of one plane: L = acc (+ tde), physics_loss.py:156,171,175; sums of tde^2 and of the
  // material-balance cells (:193) when `count`
  auto local = [&](const float (&p1)[CPT], const float (&p0)[CPT], const float2 (&e0)[CPT], const float2 (&e1)[CPT],
                   const float2 (&x0)[CPT], const float2 (&x1)[CPT], int off, float (&L)[CPT], bool count) {
    if (PK && count) {          // the adjoint's pack of this plane: cp, A0', A0'', G, A1', G'
      float v[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) v[c] = e0[c].y;
      stgs(pkb + off, v, strm);
#pragma unroll
      for (int c = 0; c < CPT; ++c) v[c] = x0[c].x;
      stgs(pkb + pks + off, v, strm);
#pragma unroll
      for (int c = 0; c < CPT; ++c) v[c] = x0[c].y;
      stgs(pkb + 2 * pks + off, v, strm);
#pragma unroll
      for (int c = 0; c < CPT; ++c) v[c] = e1[c].y;
      stgs(pkb + 3 * pks + off, v, strm);
#pragma unroll
      for (int c = 0; c < CPT; ++c) v[c] = x1[c].x;
      stgs(pkb + 4 * pks + off, v, strm);
#pragma unroll
      for (int c = 0; c < CPT; ++c) v[c] = x1[c].y;
      stgs(pkb + 5 * pks + off, v, strm);
    }
    float dpv[CPT], numr[CPT], q1[CPT], q2[CPT];
    bool bad = div_slow;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      dpv[c] = __fsub_rn(p1[c], p0[c]);
      // p2 by linear extrapolation, truncation bracket                          physics_loss.py:126,171
      const float p2 = __fadd_rn(__fmul_rn(dpv[c], one_rho), p0[c]);
      numr[c] = __fsub_rn(__fadd_rn(__fmul_rn(d2, p0[c]), __fmul_rn(d1, p2)), __fmul_rn(d12, p1[c]));
      q1[c] = div_fast(e0[c].y, by_d1);
      q2[c] = div_fast(numr[c], by_den);
      bad |= div_operand_bad(numr[c]);
    }
    if (bad) {
#pragma unroll
      for (int c = 0; c < CPT; ++c) { q1[c] = div_c(e0[c].y, by_d1); q2[c] = div_c(numr[c], by_den); }
    }
    float tsum = 0.f, msum = 0.f;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const float cp = e0[c].y;
      const float a5t = __fmul_rn(P.invDc, q1[c]);
      const float E = __fadd_rn(c2e7, q2[c]);
      const float tde = __fmul_rn(__fmul_rn(P.dvDc, cp), E);
      const float acc = __fmul_rn(__fmul_rn(P.dv, a5t), dpv[c]);                  // physics_loss.py:175
      L[c] = P.tde_in_dom ? __fadd_rn(acc, tde) : acc;
      const float mb = __fmul_rn(__fmul_rn(P.dvSgi_phi, __fsub_rn(e1[c].x, e0[c].x)), mbfac);   // physics_loss.py:193
      tsum = fmaf(tde, tde, tsum);
      msum += mb;
    }
    if (count) { a_tde += tsum; a_mbf += msum; }
  };

  auto march = [&](auto WT) {
    constexpr int WELLS = decltype(WT)::value;      // 0: no connection in the tile, 1: staged lists, 2: search (lists did not fit)
    int off = t.oc, offE = yy * FL.WP + xx, offN = (int)FL.nE + t.oc, offU = (int)(FL.nE + FL.nN) + t.oc + HW;
    float pc[CPT], Gc[CPT], Lc[CPT], pn[CPT], Gn[CPT], Ln[CPT], tz[CPT], pq[CPT], p0q[CPT];
    // halo ring, one plane further ahead than the own cells need it: whatever a halo thread publishes before it arrives
    // on the plane barrier is on the critical path of the WHOLE CTA, so its load is issued two iterations and its gather
    // one iteration before the store (p1 of plane k+1 / its gathered G / p1 of plane k+2), and the store itself moves to
    // the top of the iteration, right after the wait: nothing stays live across the stencil that was not live before.
    // Measured: forward 13.35 -> 12.7 ms on config 5 (K = 8); the same change in the adjoint cost 1.5 %, not taken.
    float hp1 = 0.f, hG1 = 0.f, hp2 = 0.f;
    {
      const int s1 = (D > 1) ? HW : 0, s2 = (D > 2) ? 2 * HW : s1;
      float p0a[CPT], p0b[CPT];
      ldgs(p1f + off, pc, strm);
      ldgs(p1f + off + s1, pn, strm);
      ldgs(p0f + off, p0a, strm);
      ldgs(p0f + off + s1, p0b, strm);
      float hp0 = 0.f;
      if (t.halo) { hp0 = LD_STRM(p1f + t.h_off, strm); hp1 = LD_STRM(p1f + (s1 + t.h_off), strm); hp2 = LD_STRM(p1f + (s2 + t.h_off), strm); }
      ldgs(p1f + off + s2, pq, strm);
      ldgs(p0f + off + s2, p0q, strm);
      float2 e1a[CPT], e1b[CPT], e0a[CPT], e0b[CPT], x1a[CPT], x1b[CPT], x0a[CPT], x0b[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        gat2(pc[c], p0a[c], e1a[c], e0a[c], x1a[c], x0a[c]);
        gat2(pn[c], p0b[c], e1b[c], e0b[c], x1b[c], x0b[c]);
      }
      if (t.halo) {
        s_p[t.h_slot] = hp0; s_G[t.h_slot] = PK ? GATA1G(TF, hp0).y : GATF1(TF, hp0).y;
        hG1 = PK ? GATA1G(TF, hp1).y : GATF1(TF, hp1).y;
      }
#pragma unroll
      for (int c = 0; c < CPT; ++c) { Gc[c] = e1a[c].y; Gn[c] = e1b[c].y; tz[c] = -0.0f; }   // image face below plane 0: a5*(p - p) = +0
      local(pc, p0a, e0a, e1a, x0a, x1a, off, Lc, t.valid);
      local(pn, p0b, e0b, e1b, x0b, x1b, off + s1, Ln, t.valid && D > 1);
      stsv(s_p + own_s, pc);
      stsv(s_G + own_s, Gc);
      mbar_arrive_warp(&s_bar);
    }
    int sb = 0;

    for (int k = 0; k < D; ++k) {
      const float* sp = s_p + sb;
      const float* sG = s_G + sb;
      const int rem = D - 1 - k;                         // planes above this one
      // static face coefficients of plane k first: they are not queued behind the gathers
      float fx[CPT + 1], fS[CPT], fN[CPT], fU[CPT];
      ldgc(FB + offE, fx);
#if defined(SRM_D4_ABL) && SRM_D4_ABL == 3
      fx[CPT] = 1e-3f;
#else
      fx[CPT] = __ldg(FB + offE + CPT);
#endif
      ldgc(FB + offN, fS);
      ldgc(FB + offN + W, fN);
      ldgc(FB + offU, fU);
      float pnn[CPT], p0nn[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) { pnn[c] = pq[c]; p0nn[c] = p0q[c]; }
      float hq3 = 0.f;                                   // halo p1 of plane k+3
      {
        const int u3 = min(3, rem) * HW;
        ldgs(p1f + off + u3, pq, strm);
        ldgs(p0f + off + u3, p0q, strm);
        if (t.halo && rem >= 3) hq3 = LD_STRM(p1f + (off - t.oc + 3 * HW + t.h_off), strm);
      }
      // gathers of plane k+2 (own) and k+1 (halo): in flight during the stencil of plane k
      float2 e1nn[CPT], e0nn[CPT], x1nn[CPT], x0nn[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) gat2(pnn[c], p0nn[c], e1nn[c], e0nn[c], x1nn[c], x0nn[c]);
      float hG2 = 0.f;                                   // halo gather of plane k+2: a whole iteration to land
      if (t.halo && rem >= 2) hG2 = PK ? GATA1G(TF, hp2).y : GATF1(TF, hp2).y;
      mbar_wait(&s_bar, k & 1);                          // plane k (own + halo) is in buffer sb
      // every warp has arrived for plane k, i.e. finished reading the other buffer: the halo of plane k+1 goes in now
      if (t.halo && rem >= 1) { s_p[(sb ^ PLANE) + t.h_slot] = hp1; s_G[(sb ^ PLANE) + t.h_slot] = hG1; }
      float pS[CPT], pN[CPT], gS[CPT], gN[CPT];
      ldsv(sp + own_s - SW, pS);
      ldsv(sp + own_s + SW, pN);
      ldsv(sG + own_s - SW, gS);
      ldsv(sG + own_s + SW, gN);
      float pWe = __shfl_up_sync(0xffffffffu, pc[CPT - 1], 1, CX), gWe = __shfl_up_sync(0xffffffffu, Gc[CPT - 1], 1, CX);
      float pEe = __shfl_down_sync(0xffffffffu, pc[0], 1, CX), gEe = __shfl_down_sync(0xffffffffu, Gc[0], 1, CX);
      if (t.cx == 0) { pWe = sp[own_s - 1]; gWe = sG[own_s - 1]; }
      if (t.cx == CX - 1) { pEe = sp[own_s + CPT]; gEe = sG[own_s + CPT]; }
      if (t.edgeE) { pEe = pc[CPT - 1]; gEe = Gc[CPT - 1]; }
      // x-face coefficients, one per face: C*k_f*krg*G_f*(1/dx)*(1/dx)          physics_loss.py:147-148,152-155
      float ax[CPT + 1];
#pragma unroll
      for (int i = 0; i <= CPT; ++i) {
        const float gl = (i == 0) ? gWe : Gc[i - (i > 0)], gr = (i == CPT) ? gEe : Gc[i - (i == CPT)];
        const float Gf = __fmul_rn(__fadd_rn(gr, gl), 0.5f);
        ax[i] = __fmul_rn(__fmul_rn(__fmul_rn(fx[i], Gf), P.idx), P.idx);
      }
      float domv[CPT], dsum = 0.f;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const float p1 = pc[c], G = Gc[c];
        const float pW = (c == 0) ? pWe : pc[c - (c > 0)], pE = (c == CPT - 1) ? pEe : pc[c + (c < CPT - 1)];
        const float a1 = ax[c], a3 = ax[c + 1];
        const float a2 = __fmul_rn(__fmul_rn(__fmul_rn(fS[c], __fmul_rn(__fadd_rn(G, gS[c]), 0.5f)), P.idy), P.idy);
        const float a4 = __fmul_rn(__fmul_rn(__fmul_rn(fN[c], __fmul_rn(__fadd_rn(gN[c], G), 0.5f)), P.idy), P.idy);
        // z faces (3-D extension): the upper term of this plane is minus the lower term of the next
        const float a6 = __fmul_rn(__fmul_rn(__fmul_rn(fU[c], __fmul_rn(__fadd_rn(Gn[c], G), 0.5f)), P.idz), P.idz);
        const float tu = __fmul_rn(a6, __fsub_rn(p1, pn[c]));
        const float zt = __fadd_rn(-tz[c], tu);
        tz[c] = tu;
        // wells in this cell (scatter_nd sums duplicates)                        well_rate_bhp_Subclassed.py:128-132
        float qdv = 0.f, mask = 0.f;
        int wfirst = 0, wlast = 0;
        const int cell = off + c;
        if (WELLS == 1) {
          const uint32_t sl = (wslots >> (8 * c)) & 255u;
          if (sl) {
            float q = 0.f;
            s_wt.take((int)sl - 1, k, wfirst, wlast);
            for (int e = wfirst; e < wlast; ++e) { q = __fadd_rn(q, s_wt.val[e]); mask += 1.f; }
            if (mask != 0.f) qdv = __fdiv_rn(q, P.dv);
          }
        }
        if (WELLS == 2 && has_well) {
          float q = 0.f;
          wfirst = well_lower_bound(P, cell);
          for (int w = wfirst; w < P.n_wells && P.wells[w].cell == cell; ++w) {
            q = __fadd_rn(q, A.qw[(int64_t)b * P.n_wells + w]);
            mask += 1.f;
          }
          if (mask != 0.f) qdv = __fdiv_rn(q, P.dv);
        }
        // flux divergence                                                        physics_loss.py:174
        float s = __fadd_rn(-__fmul_rn(a1, pW), -__fmul_rn(a2, pS[c]));
        const float asum = __fadd_rn(__fadd_rn(__fadd_rn(a1, a2), a3), a4);
        s = __fadd_rn(s, __fmul_rn(asum, p1));
        s = __fadd_rn(s, -__fmul_rn(a3, pE));
        s = __fadd_rn(s, -__fmul_rn(a4, pN[c]));
        s = __fadd_rn(s, zt);
        s = __fadd_rn(s, qdv);
        const float divq = __fmul_rn(P.dv, s);
        const float dom = __fadd_rn(divq, Lc[c]);                                 // physics_loss.py:176
        domv[c] = dom;
        dsum = fmaf(dom, dom, dsum);
        if (WELLS && mask != 0.f && t.valid) {
          if (WELLS == 1) { for (int e = wfirst; e < wlast; ++e) A.divqw[(int64_t)b * P.n_wells + s_wt.w[e]] = divq; }
          else { for (int w = wfirst; w < P.n_wells && P.wells[w].cell == cell; ++w) A.divqw[(int64_t)b * P.n_wells + w] = divq; }
          const float ibc = __fmul_rn(mask, divq);                                // physics_loss.py:189
          a_ibc += (double)ibc * (double)ibc;
        }
      }
      if (t.valid) {
        stgs(domf + off, domv, strm);
        a_dom += dsum;
      }
      // plane k+1 into the other buffer
      {
        float* spn = s_p + (sb ^ PLANE);
        float* sGn = s_G + (sb ^ PLANE);
        stsv(spn + own_s, pn);
        stsv(sGn + own_s, Gn);
      }
      mbar_arrive_warp(&s_bar);
      hp1 = hp2; hG1 = hG2; hp2 = hq3;
      // cell-local part of plane k+2
      float Lnn[CPT];
      local(pnn, p0nn, e0nn, e1nn, x0nn, x1nn, off + 2 * HW, Lnn, t.valid && rem >= 2);
#pragma unroll
      for (int c = 0; c < CPT; ++c) { pc[c] = pn[c]; Gc[c] = Gn[c]; Lc[c] = Ln[c]; pn[c] = pnn[c]; Gn[c] = e1nn[c].y; Ln[c] = Lnn[c]; }
      off += HW; offE += strE; offN += strN; offU += HW;
      sb ^= PLANE;
      if ((k & 7) == 7) { a_mb += (double)a_mbf; a_mbf = 0.f; }
    }
  };
  if (!tile_wells) march(IntC<0>()); else if (!wt_overflow) march(IntC<1>()); else march(IntC<2>());

  double acc4[4] = {(double)a_dom, a_ibc, (double)a_tde, a_mb + (double)a_mbf};
  __syncthreads();
  block_reduce<4>(acc4, red);
  if (threadIdx.x == 0) {
    atomicAdd(&A.sse[SRM_TERM_DOM], acc4[0]);
    if (acc4[1] != 0.0) atomicAdd(&A.sse[SRM_TERM_IBC], acc4[1]);
    atomicAdd(&A.sse[SRM_TERM_TDE], acc4[2]);
    atomicAdd(&A.mb_sum[b], acc4[3]);
  }
}

// ------------------------------------------------------------------------------------------
// adjoint: hand-derived, what tape.gradient delivers (physics_loss.py:849-859)
// ------------------------------------------------------------------------------------------
// PK: the six table values per cell come from the pack the forward staged (coalesced streams); no table gathers
template <bool PK>
__global__ void __launch_bounds__(NT, SRM_D4_OCCA) k_adj4(const __grid_constant__ SrmDev P, const __grid_constant__ R2Args A) {
  __shared__ __align__(16) float s_p[2 * PLANE];
  __shared__ __align__(16) float s_G[2 * PLANE];
  __shared__ __align__(16) float s_s[2 * PLANE];
  __shared__ double red[2 * 32];
  __shared__ __align__(8) WTile s_wt;
  __shared__ __align__(8) uint64_t s_bar;
  const Tile4 t = make_tile4(P, A.tiles_x);
  const int b = blockIdx.y;
  const int r = srm_real_of(A.sample_real, b, A.B, A.R);
  if (threadIdx.x == 0) mbar_init(&s_bar, NT / 32);
  __syncthreads();
  // the tile's connections: column lists staged in shared memory (well_tile.cuh); lists that do not fit keep the search
  bool tile_wells = false, wt_overflow = false, has_well = false;
  uint32_t wslots = 0;
  if (P.n_wells > 0) {
    tile_wells = s_wt.build(well_cols_of(P), P.W, P.D, t.x0, t.y0, A.dqdp + (int64_t)b * P.n_wells, wt_overflow);
    if (wt_overflow) {
      has_well = thread_has_well4(P, t, reinterpret_cast<unsigned char (*)[TW]>(s_wt.slot_of));
      tile_wells = __syncthreads_or(has_well ? 1 : 0) != 0;
    } else if (tile_wells && t.valid) {
      wslots = s_wt.template slots_of<CPT>(t.ty * TW + CPT * t.cx);
    }
  }
  const int W = P.W, H = P.H, D = P.D, HW = H * W;
  const FaceLay FL = face_layout(D, H, W);
  const float* __restrict__ p0f = A.p0 + (int64_t)b * P.N;
  const float* __restrict__ p1f = A.p1 + (int64_t)b * P.N;
  const float* __restrict__ domf = A.dom + (int64_t)b * P.N;
  float* __restrict__ gp0f = A.gp0 + (int64_t)b * P.N;
  float* __restrict__ gp1f = A.gp1 + (int64_t)b * P.N;
  const float* __restrict__ FB = A.faces + (int64_t)r * FL.per_real;
  const int yy = t.oc / W, xx = t.oc - yy * W;
  const int strE = H * FL.WP, strN = (H + 1) * W;
  const float w_tde2 = 2.f * A.dterms[SRM_TERM_TDE];
  const float d1 = A.dt1[b], d2 = A.dt2[b];
  const float two_wd = 2.f * A.dterms[SRM_TERM_DOM];
  const float smb = 2.f * A.dterms[SRM_TERM_MBC] * A.mbc[b];              // dL/d mbc_b
  // forward's per-sample scalars (op order as the forward: E is dominated by the rounding of the bracket)
  const float rho = (d1 == 0.f) ? 0.f : __fdiv_rn(d2, d1);
  const float one_rho = __fadd_rn(1.0f, rho);
  const float den = __fadd_rn(__fmul_rn(d1, d2), __fmul_rn(d2, d2));
  const DivC by_den = make_divc(den);
  const float c2e7 = __fdiv_rn(2e-7f, d1);
  const float d12 = __fadd_rn(d1, d2);
  const float id1 = 1.0f / d1, iden2 = 1.0f / (den * den);
  const float mbk = P.dvSgi_phi / (P.Dc * d1);          // d mb_cells / d(A1-A0)
  const float hx2 = 0.5f * P.idx * P.idx, hy2 = 0.5f * P.idy * P.idy, hz2 = 0.5f * P.idz * P.idz;
  const float dE1c = -2e-7f * id1 * id1;
  const float dE1n = d2 * iden2, dE2n = (d1 + 2.f * d2) * iden2;
  const float dvi = P.dv * P.invDc * id1;               // d acc / d(cp * dp)
  const float smbk = smb * mbk;
  const float seed_in_tde = P.tde_in_dom ? 1.f : 0.f;
  const float Sphi = P.Sgi * P.phi, Sphicf = P.Sgi * P.phicf;
  const bool div_slow = !by_den.ok || c2e7 == 0.f;

  const uint64_t keep = l2_evict_last(), strm = l2_evict_first();
  const Tab TA = make_tab(P.lut0, P, 32);
  const int own_s = (t.ty + 1) * SW + XO + CPT * t.cx;
  const float plo = P.p_min, phi_ = P.p_max;
  const float* __restrict__ pkb = PK ? A.pk + (int64_t)b * P.N : nullptr;
  const int64_t pks = A.pk_stride;
  // the adjoint's table values of CPT cells: gathered, or (PK) streamed from the forward's pack at plane offset o
  auto entries = [&](const float (&p1v)[CPT], const float (&p0v)[CPT], int o, float4 (&e1)[CPT], float4 (&e0)[CPT]) {
    if constexpr (PK) {
      float cp[CPT], a0p[CPT], a0pp[CPT], g[CPT], a1p[CPT], gp[CPT];
      ldgs(pkb + o, cp, strm); ldgs(pkb + pks + o, a0p, strm); ldgs(pkb + 2 * pks + o, a0pp, strm);
      ldgs(pkb + 3 * pks + o, g, strm); ldgs(pkb + 4 * pks + o, a1p, strm); ldgs(pkb + 5 * pks + o, gp, strm);
#pragma unroll
      for (int c = 0; c < CPT; ++c) { e1[c] = make_float4(0.f, g[c], a1p[c], gp[c]); e0[c] = make_float4(0.f, a0p[c], a0pp[c], cp[c]); }
    } else {
#pragma unroll
      for (int c = 0; c < CPT; ++c) { e1[c] = GATA1(TA, p1v[c]); e0[c] = GATA0(TA, p0v[c]); }
    }
  };
  auto halo_G = [&](float hp, int o) { return PK ? LD_STRM(pkb + 3 * pks + o, strm) : GATA1G(TA, hp).y; };

  float a_g1 = 0.f, a_g2 = 0.f;      // per-thread partial sums of dL/ddt1, dL/ddt2 (flushed to fp64 every 8 planes)
  double d_g1 = 0.0, d_g2 = 0.0;

  // cell-local part of plane j: dL/dp0 (complete: stored), the local part of dL/dp1, the masked G'
  auto local = [&](const float (&p1)[CPT], const float (&p0)[CPT], const float (&s)[CPT], const float4 (&e0)[CPT],
                   const float4 (&e1)[CPT], float (&Gp)[CPT], float (&L1)[CPT], int off, bool count) {
    float numr[CPT], q2[CPT];
    bool bad = div_slow;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const float dp = __fsub_rn(p1[c], p0[c]);
      const float p2 = __fadd_rn(__fmul_rn(dp, one_rho), p0[c]);
      numr[c] = __fsub_rn(__fadd_rn(__fmul_rn(d2, p0[c]), __fmul_rn(d1, p2)), __fmul_rn(d12, p1[c]));
      q2[c] = div_fast(numr[c], by_den);
      bad |= div_operand_bad(numr[c]);
    }
    if (bad) {
#pragma unroll
      for (int c = 0; c < CPT; ++c) q2[c] = div_c(numr[c], by_den);
    }
    float g0v[CPT], t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const bool in1 = p1[c] >= plo && p1[c] <= phi_;      // clamp's gradient mask (PVT_Layer_Subclassed.py:165-167)
      const bool in0 = p0[c] >= plo && p0[c] <= phi_;
      Gp[c] = in1 ? e1[c].w : 0.f;
      const float M = in1 ? smbk * e1[c].z : 0.f;
      const float cp = e0[c].w;
      const float A0pm = in0 ? e0[c].y : 0.f;
      const float cpp = in0 ? fmaf(Sphi, e0[c].z, Sphicf * e0[c].y) : 0.f;      // d cp / d p0
      const float dva5t = dvi * cp;                                              // dv * a5t
      const float dp = p1[c] - p0[c];
      const float E = __fadd_rn(c2e7, q2[c]);
      const float cE = __fmul_rn(P.dvDc, cp);
      const float tde = __fmul_rn(cE, E);
      const float st = fmaf(seed_in_tde, s[c], w_tde2 * tde);                    // dL/d tde
      L1[c] = fmaf(s[c], dva5t, -M);
      g0v[c] = s[c] * (dvi * dp * cpp - dva5t) + st * P.dvDc * cpp * E + smbk * A0pm;
      // d/d dt1, d/d dt2 (the d bracket/d* pieces vanish identically; the bracket itself is rounding noise);
      // the material-balance part of dL/ddt1 is a per-sample scalar (k_finalize_adj4)
      const float stc = st * cE;
      t1 += stc * (dE1c - numr[c] * dE1n) - s[c] * dva5t * dp * id1;
      t2 -= stc * numr[c] * dE2n;
    }
    if (count) {
      stgs(gp0f + off, g0v, strm);
      a_g1 += t1; a_g2 += t2;
    }
  };

  auto march = [&](auto WT) {
    constexpr int WELLS = decltype(WT)::value;
    int off = t.oc, offE = yy * FL.WP + xx, offN = (int)FL.nE + t.oc, offU = (int)(FL.nE + FL.nN) + t.oc + HW;
    // planes k and k+1: p1, G, masked G', seed, local part of dL/dp1; carried upper-face pair of the plane below
    float pc[CPT], Gc[CPT], Gpc[CPT], sc[CPT], Lc[CPT], pn[CPT], Gn[CPT], Gpn[CPT], sn[CPT], Ln[CPT], Xz[CPT], Yz[CPT];
    float pq[CPT], p0q[CPT], hq = 0.f;
    {
      const int s1 = (D > 1) ? HW : 0, s2 = (D > 2) ? 2 * HW : s1;
      float p0a[CPT], p0b[CPT];
      ldgs(p1f + off, pc, strm);
      ldgs(p1f + off + s1, pn, strm);
      ldgs(p0f + off, p0a, strm);
      ldgs(p0f + off + s1, p0b, strm);
      ldgs(domf + off, sc, strm);
      ldgs(domf + off + s1, sn, strm);
      float hp0 = 0.f, hs0 = 0.f;
      if (t.halo) { hp0 = LD_STRM(p1f + t.h_off, strm); hs0 = LD_STRM(domf + t.h_off, strm); hq = LD_STRM(p1f + (s1 + t.h_off), strm); }
      ldgs(p1f + off + s2, pq, strm);
      ldgs(p0f + off + s2, p0q, strm);
      float4 e1a[CPT], e1b[CPT], e0a[CPT], e0b[CPT];
      entries(pc, p0a, off, e1a, e0a);
      entries(pn, p0b, off + s1, e1b, e0b);
      if (t.halo) { s_p[t.h_slot] = hp0; s_G[t.h_slot] = halo_G(hp0, t.h_off); s_s[t.h_slot] = two_wd * hs0; }
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        Gc[c] = e1a[c].y; Gn[c] = e1b[c].y; sc[c] *= two_wd; sn[c] *= two_wd;
        Xz[c] = 0.f; Yz[c] = 0.f;                            // image face below plane 0
      }
      local(pc, p0a, sc, e0a, e1a, Gpc, Lc, off, t.valid);
      local(pn, p0b, sn, e0b, e1b, Gpn, Ln, off + s1, t.valid && D > 1);
      stsv(s_p + own_s, pc);
      stsv(s_G + own_s, Gc);
      stsv(s_s + own_s, sc);
      mbar_arrive_warp(&s_bar);
    }
    int sb = 0;

    for (int k = 0; k < D; ++k) {
      const float* sp = s_p + sb;
      const float* sG = s_G + sb;
      const float* ss = s_s + sb;
      const int rem = D - 1 - k;
      float fx[CPT + 1], fS[CPT], fN[CPT], fU[CPT];
      ldgc(FB + offE, fx);
#if defined(SRM_D4_ABL) && SRM_D4_ABL == 3
      fx[CPT] = 1e-3f;
#else
      fx[CPT] = __ldg(FB + offE + CPT);
#endif
      ldgc(FB + offN, fS);
      ldgc(FB + offN + W, fN);
      ldgc(FB + offU, fU);
      float pnn[CPT], p0nn[CPT], snn[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) { pnn[c] = pq[c]; p0nn[c] = p0q[c]; }
      const float hp = hq;
      float hs = 0.f;
      {
        const int u2 = min(2, rem) * HW, u3 = min(3, rem) * HW;
        ldgs(p1f + off + u3, pq, strm);
        ldgs(p0f + off + u3, p0q, strm);
        ldgs(domf + off + u2, snn, strm);
        if (t.halo && rem >= 2) hq = LD_STRM(p1f + (off - t.oc + 2 * HW + t.h_off), strm);
        if (t.halo && rem >= 1) hs = LD_STRM(domf + (off - t.oc + HW + t.h_off), strm);
      }
      float4 e1nn[CPT], e0nn[CPT];
      entries(pnn, p0nn, off + min(2, rem) * HW, e1nn, e0nn);
      float hG = 0.f;
      if (t.halo && rem >= 1) hG = halo_G(hp, off - t.oc + HW + t.h_off);
      mbar_wait(&s_bar, k & 1);
      float pS[CPT], pN[CPT], gS[CPT], gN[CPT], sS[CPT], sN[CPT];
      ldsv(sp + own_s - SW, pS);
      ldsv(sp + own_s + SW, pN);
      ldsv(sG + own_s - SW, gS);
      ldsv(sG + own_s + SW, gN);
      ldsv(ss + own_s - SW, sS);
      ldsv(ss + own_s + SW, sN);
      float pWe = __shfl_up_sync(0xffffffffu, pc[CPT - 1], 1, CX), gWe = __shfl_up_sync(0xffffffffu, Gc[CPT - 1], 1, CX);
      float sWe = __shfl_up_sync(0xffffffffu, sc[CPT - 1], 1, CX);
      float pEe = __shfl_down_sync(0xffffffffu, pc[0], 1, CX), gEe = __shfl_down_sync(0xffffffffu, Gc[0], 1, CX);
      float sEe = __shfl_down_sync(0xffffffffu, sc[0], 1, CX);
      if (t.cx == 0) { pWe = sp[own_s - 1]; gWe = sG[own_s - 1]; sWe = ss[own_s - 1]; }
      if (t.cx == CX - 1) { pEe = sp[own_s + CPT]; gEe = sG[own_s + CPT]; sEe = ss[own_s + CPT]; }
      if (t.edgeE) { pEe = pc[CPT - 1]; gEe = Gc[CPT - 1]; sEe = sc[CPT - 1]; }
      // stencil part of dL/dp1: dv * sum_f [ +-X_f + G'_c * Y_f ],  u_f = (s_c - s_n) T_f/2,
      // X_f = u_f (G_c + G_n), Y_f = u_f (p_c - p_n); the neighbour's view of the same face is (-X_f, +Y_f)
      float sx[CPT], sy[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) { sx[c] = -Xz[c]; sy[c] = Yz[c]; }
#pragma unroll
      for (int i = 0; i <= CPT; ++i) {
        // face i lies between cell i-1 (or the W neighbour) and cell i (or the E neighbour)
        const float Tf = fx[i] * hx2;
        if (i < CPT) {
          const float pl = (i == 0) ? pWe : pc[i - (i > 0)], gl = (i == 0) ? gWe : Gc[i - (i > 0)], sl = (i == 0) ? sWe : sc[i - (i > 0)];
          const float u = (sc[i] - sl) * Tf;
          const float X = u * (Gc[i] + gl), Y = u * (pc[i] - pl);
          sx[i] += X; sy[i] += Y;
          if (i > 0) { sx[i - (i > 0)] -= X; sy[i - (i > 0)] += Y; }
        } else {
          const float u = (sc[CPT - 1] - sEe) * Tf;
          sx[CPT - 1] = fmaf(u, Gc[CPT - 1] + gEe, sx[CPT - 1]);
          sy[CPT - 1] = fmaf(u, pc[CPT - 1] - pEe, sy[CPT - 1]);
        }
      }
      float g1v[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const float p1 = pc[c], G = Gc[c], s0 = sc[c];
        {
          const float u = (s0 - sS[c]) * (fS[c] * hy2);
          sx[c] = fmaf(u, G + gS[c], sx[c]); sy[c] = fmaf(u, p1 - pS[c], sy[c]);
        }
        {
          const float u = (s0 - sN[c]) * (fN[c] * hy2);
          sx[c] = fmaf(u, G + gN[c], sx[c]); sy[c] = fmaf(u, p1 - pN[c], sy[c]);
        }
        {
          const float u = (s0 - sn[c]) * (fU[c] * hz2);
          const float X = u * (G + Gn[c]), Y = u * (p1 - pn[c]);
          sx[c] += X; sy[c] += Y;
          Xz[c] = X; Yz[c] = Y;
        }
        float g1 = fmaf(P.dv, fmaf(Gpc[c], sy[c], sx[c]), Lc[c]);
        if (WELLS == 1) {
          const uint32_t sl = (wslots >> (8 * c)) & 255u;
          if (sl) {
            int first, last;
            float dq = 0.f;
            s_wt.take((int)sl - 1, k, first, last);
            for (int e = first; e < last; ++e) dq += s_wt.val[e];
            g1 = fmaf(s0 - smb, dq, g1);
          }
        }
        if (WELLS == 2 && has_well) {
          const int cell = off + c;
          float dq = 0.f;
          const int first = well_lower_bound(P, cell);
          for (int w = first; w < P.n_wells && P.wells[w].cell == cell; ++w) dq += A.dqdp[(int64_t)b * P.n_wells + w];
          g1 = fmaf(s0 - smb, dq, g1);
        }
        g1v[c] = g1;
      }
      if (t.valid) stgs(gp1f + off, g1v, strm);
      {
        float* spn = s_p + (sb ^ PLANE);
        float* sGn = s_G + (sb ^ PLANE);
        float* ssn = s_s + (sb ^ PLANE);
        stsv(spn + own_s, pn);
        stsv(sGn + own_s, Gn);
        stsv(ssn + own_s, sn);
        if (t.halo && rem >= 1) { spn[t.h_slot] = hp; sGn[t.h_slot] = hG; ssn[t.h_slot] = two_wd * hs; }
      }
      mbar_arrive_warp(&s_bar);
      // cell-local part of plane k+2
      float Gpnn[CPT], Lnn[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) snn[c] *= two_wd;
      local(pnn, p0nn, snn, e0nn, e1nn, Gpnn, Lnn, off + 2 * HW, t.valid && rem >= 2);
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        pc[c] = pn[c]; Gc[c] = Gn[c]; Gpc[c] = Gpn[c]; sc[c] = sn[c]; Lc[c] = Ln[c];
        pn[c] = pnn[c]; Gn[c] = e1nn[c].y; Gpn[c] = Gpnn[c]; sn[c] = snn[c]; Ln[c] = Lnn[c];
      }
      off += HW; offE += strE; offN += strN; offU += HW;
      sb ^= PLANE;
      if ((k & 7) == 7) { d_g1 += (double)a_g1; d_g2 += (double)a_g2; a_g1 = 0.f; a_g2 = 0.f; }
    }
  };
  if (!tile_wells) march(IntC<0>()); else if (!wt_overflow) march(IntC<1>()); else march(IntC<2>());

  double acc2[2] = {d_g1 + (double)a_g1, d_g2 + (double)a_g2};
  __syncthreads();
  block_reduce<2>(acc2, red);
  if (threadIdx.x == 0) {
    atomicAdd(&A.gdt1_acc[b], acc2[0]);
    atomicAdd(&A.gdt2_acc[b], acc2[1]);
  }
}

// dL/ddt: block partials plus the material-balance part, mbc_b = -sum q - sum mb, mb ~ 1/dt1:
// d mbc_b / d dt1 = (sum mb)/dt1                                         physics_loss.py:193
__global__ void k_finalize_adj4(int32_t B, const double* __restrict__ a1, const double* __restrict__ a2,
                                const double* __restrict__ mb_sum, const float* __restrict__ mbc,
                                const float* __restrict__ dterms, const float* __restrict__ dt1,
                                float* __restrict__ gdt1, float* __restrict__ gdt2) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    const double smb = 2.0 * (double)dterms[SRM_TERM_MBC] * (double)mbc[b];
    gdt1[b] = (float)(a1[b] + smb * mb_sum[b] / (double)dt1[b]);
    gdt2[b] = (float)a2[b];
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// launchers (called by kernels_ref2.cu)
// ------------------------------------------------------------------------------------------
bool srm_dg4_applicable(const SrmHandle* h) {
  const SrmDev& P = h->dev;
  return h->lut_full && P.lut_n > 0 && P.W % CPT == 0 && P.W >= CPT;
}

// The table gathers miss the L1 by construction, and the number of misses an SM keeps in flight grows with the L1's
// capacity (tools/gather_probe2.cu: 1.1 SM-cycles per gather with the whole 256 KB array as L1, 2.4 with a 227 KB
// shared-memory carve-out).  Left alone the driver configures ~100 KB of shared memory for these kernels although four
// resident CTAs need 35-46 KB, so the carve-out is requested explicitly (percent of the maximum; SRM_D4_CARVEOUT overrides).
static cudaError_t dg4_set_carveout(int device) {
  static bool done_dev[64] = {};              // function attributes are per device
  bool& done = done_dev[device & 63];
  if (done) return cudaSuccess;
  int pct = 20;
  if (const char* e = getenv("SRM_D4_CARVEOUT")) pct = atoi(e);
  if (pct >= 0) {
    cudaError_t e1 = cudaFuncSetAttribute(k_fwd4<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    if (e1 != cudaSuccess) return e1;
    e1 = cudaFuncSetAttribute(k_fwd4<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    if (e1 != cudaSuccess) return e1;
    e1 = cudaFuncSetAttribute(k_adj4<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    if (e1 != cudaSuccess) return e1;
    // the gather-free adjoint has no table misses to keep in flight: it keeps the driver's default carve-out

  }
  done = true;
  return cudaSuccess;
}

cudaError_t srm_dg4_launch_fwd(const SrmHandle* h, const void* args, int32_t B, cudaStream_t s) {
  const SrmDev& P = h->dev;
  { cudaError_t ce = dg4_set_carveout(h->device); if (ce != cudaSuccess) return ce; }
  R2Args A = *reinterpret_cast<const R2Args*>(args);
  A.tiles_x = (P.W + TW - 1) / TW;
  const dim3 grid((unsigned)(A.tiles_x * ((P.H + TY - 1) / TY)), (unsigned)B);
  if (A.pk) k_fwd4<true><<<grid, NT, 0, s>>>(P, A); else k_fwd4<false><<<grid, NT, 0, s>>>(P, A);
  cudaError_t e = cudaGetLastError();
  // the residual field for the caller (tests, diagnostics): a copy of the adjoint's seed, off the hot path
  if (e == cudaSuccess && A.dom_out) e = cudaMemcpyAsync(A.dom_out, A.dom, sizeof(float) * (size_t)B * (size_t)P.N, cudaMemcpyDeviceToDevice, s);
  return e;
}

cudaError_t srm_dg4_launch_adj(const SrmHandle* h, const void* args, int32_t B, cudaStream_t s) {
  const SrmDev& P = h->dev;
  { cudaError_t ce = dg4_set_carveout(h->device); if (ce != cudaSuccess) return ce; }
  R2Args A = *reinterpret_cast<const R2Args*>(args);
  A.tiles_x = (P.W + TW - 1) / TW;
  const dim3 grid((unsigned)(A.tiles_x * ((P.H + TY - 1) / TY)), (unsigned)B);
  if (A.pk) k_adj4<true><<<grid, NT, 0, s>>>(P, A); else k_adj4<false><<<grid, NT, 0, s>>>(P, A);
  return cudaGetLastError();
}

cudaError_t srm_dg4_finalize_adj(const void* args, int32_t B, float* gdt1, float* gdt2, cudaStream_t s) {
  const R2Args& A = *reinterpret_cast<const R2Args*>(args);
  k_finalize_adj4<<<(unsigned)((B + 255) / 256), 256, 0, s>>>(B, A.gdt1_acc, A.gdt2_acc, A.mb_sum, A.mbc, A.dterms, A.dt1, gdt1, gdt2);
  return cudaGetLastError();
}
