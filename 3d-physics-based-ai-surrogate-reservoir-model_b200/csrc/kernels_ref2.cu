// SRM_NUMERICS_REFERENCE, tabulated-PVT path (SrmConfig.pvt_lut): fused forward and adjoint.
//
// Same arithmetic, op for op, as kernels_ref.cu (forward fields stay bit-identical to the pinned
// oracle) -- what changes is the data movement:
//   * PVT is two 16-byte gathers per cell and pass from the exact table (L2 resident over the
//     operating window) instead of seven staged fields: forward reads p0,p1 and writes dom (12 B per
//     cell-timestep), adjoint reads p0,p1,dom and writes gp0,gp1 (20 B).
//   * the static face coefficients fl(fl(C*k_f)*krg) with k_f = 2 k1 k2/(k1+k2)
//     (physics_loss.py:56-60,152-155) are built once per call and realisation (k_faces_ref) instead of six
//     IEEE divisions per cell and sample; boundary slots hold the edge-replicated (self) value.
//   * one CTA = one 32x16 (x,y) tile of one sample, marching over z: z neighbours live in registers,
//     x/y neighbours (p1, G = invBg*invug, adjoint seed) go through a double-buffered shared-memory
//     plane with ONE barrier per plane; the plane ahead is prefetched (loads + gathers in flight while
//     the current plane is computed).
//
//   k_faces_ref     static face coefficients                       physics_loss.py:56-60,152-155
//   k_fwd_ref2      physics_error_gas residual + SSE partials      physics_loss.py:143-193,787-807
//   k_adj_ref2      hand-derived adjoint (tape.gradient, physics_loss.py:849-859)
//   k_ibc_adj_ref2  inner-boundary (well-cell) part of the adjoint
#include <cstdlib>
#include <cstring>
#include "ref_fused.cuh"

namespace {

constexpr int TX = 32, TY = 16, NT = TX * TY;
constexpr int SW = TX + 2, SH = TY + 2;
constexpr int NHALO = 2 * TX + 2 * TY;

// tile bookkeeping shared by both kernels
struct Tile {
  int tx, ty, x0, y0;
  bool valid, halo;
  int oc;      // own column offset inside a plane (coordinates clamped to the grid)
  int oh;      // halo cell offset inside a plane (clamped = edge replication)
  int hx, hy;  // halo slot in the shared plane
};
__device__ __forceinline__ Tile make_tile(const SrmDev& P, int tiles_x) {
  Tile t;
  const int tid = threadIdx.x;
  t.tx = tid & 31; t.ty = tid >> 5;
  const int tyi = blockIdx.x / tiles_x, txi = blockIdx.x - tyi * tiles_x;
  t.x0 = txi * TX; t.y0 = tyi * TY;
  const int x = t.x0 + t.tx, y = t.y0 + t.ty;
  t.valid = x < P.W && y < P.H;
  t.oc = min(y, P.H - 1) * P.W + min(x, P.W - 1);
  int gx = 0, gy = 0;
  t.hx = 0; t.hy = 0;
  if (tid < TX) { t.hy = 0; t.hx = tid + 1; gy = t.y0 - 1; gx = t.x0 + tid; }
  else if (tid < 2 * TX) { t.hy = TY + 1; t.hx = tid - TX + 1; gy = t.y0 + TY; gx = t.x0 + tid - TX; }
  else if (tid < 2 * TX + TY) { t.hx = 0; t.hy = tid - 2 * TX + 1; gx = t.x0 - 1; gy = t.y0 + tid - 2 * TX; }
  else if (tid < NHALO) { t.hx = TX + 1; t.hy = tid - 2 * TX - TY + 1; gx = t.x0 + TX; gy = t.y0 + tid - 2 * TX - TY; }
  t.halo = tid < NHALO;
  gx = min(max(gx, 0), P.W - 1);
  gy = min(max(gy, 0), P.H - 1);
  t.oh = gy * P.W + gx;
  return t;
}

// marks the threads whose (x,y) column holds a well connection (any layer)
__device__ __forceinline__ bool column_has_well(const SrmDev& P, const Tile& t, unsigned char* s_flag) {
  s_flag[threadIdx.x] = 0;
  __syncthreads();
  const int HW = P.H * P.W;
  for (int w = threadIdx.x; w < P.n_wells; w += NT) {
    const int rem = P.wells[w].cell % HW;
    const int j = rem / P.W, i = rem - j * P.W;
    if (i >= t.x0 && i < t.x0 + TX && j >= t.y0 && j < t.y0 + TY) s_flag[(j - t.y0) * TX + (i - t.x0)] = 1;
  }
  __syncthreads();
  return s_flag[threadIdx.x] != 0 && t.valid;
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <bool FULL>
__global__ void __launch_bounds__(NT, 2) k_fwd_ref2(const __grid_constant__ SrmDev P, const __grid_constant__ R2Args A) {
  __shared__ float s_p[2][SH][SW];
  __shared__ float s_G[2][SH][SW];
  __shared__ double red[4 * 32];
  __shared__ unsigned char s_flag[NT];
  const Tile t = make_tile(P, A.tiles_x);
  const int b = blockIdx.y;
  const int r = srm_real_of(A.sample_real, b, A.B, A.R);
  const bool has_well = (P.n_wells > 0) ? column_has_well(P, t, s_flag) : false;
  const int W = P.W, H = P.H, D = P.D, HW = H * W;
  const FaceLay FL = face_layout(D, H, W);
  const float* __restrict__ p0f = A.p0 + (int64_t)b * P.N;
  const float* __restrict__ p1f = A.p1 + (int64_t)b * P.N;
  float* __restrict__ domf = A.dom + (int64_t)b * P.N;
  float* __restrict__ domo = A.dom_out ? A.dom_out + (int64_t)b * P.N : nullptr;
  const float* __restrict__ FE = A.faces + (int64_t)r * FL.per_real;
  const float* __restrict__ FN = FE + FL.nE;
  const float* __restrict__ FU = FN + FL.nN;
  const int yy = t.oc / W, xx = t.oc - yy * W;          // clamped coordinates
  int offE = yy * FL.WP + xx;                         // FE offset, plane stride H*(W+1)
  int offN = yy * W + xx;                               // FN offset, plane stride (H+1)*W
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

  const uint64_t keep = l2_evict_last();
  // plane 0 (own + halo), then the march
  int off = t.oc;                                       // own cell of the current plane
  int offh = t.oh;                                      // halo cell of the current plane
  float pc = p1f[off];
  const float2 e1 = pack1_val<FULL>(P, pc, keep);
  float Gc = e1.y, A1c = e1.x;
  float pm = pc, Gm = Gc;
  float hp = 0.f, hG = 0.f;
  if (t.halo) { hp = p1f[offh]; hG = pack1_val<FULL>(P, hp, keep).y; }
  float fD = FU[off];                                   // face below plane 0 (image)
  float a_dom = 0.f, a_tde = 0.f;                       // per-thread partial sums (<= D terms each)
  double a_ibc = 0.0, a_mb = 0.0;

  auto plane = [&](auto BUF, const int k) {
    constexpr int buf = decltype(BUF)::value;
    s_p[buf][t.ty + 1][t.tx + 1] = pc;
    s_G[buf][t.ty + 1][t.tx + 1] = Gc;
    if (t.halo) { s_p[buf][t.hy][t.hx] = hp; s_G[buf][t.hy][t.hx] = hG; }
    // prefetch plane k+1; past the top the march re-reads the last plane (= the edge-replicated image)
    const int up = (k + 1 < D) ? HW : 0;
    const float pn = p1f[off + up];
    if (t.halo) hp = p1f[offh + up];
    const float p0 = p0f[off];
    const float fW = FE[offE], fE = FE[offE + 1];
    const float fS = FN[offN], fN = FN[offN + W];
    const float fU = FU[off + HW];
    const float2 e0 = pack0_val<FULL>(P, p0, keep);
    const float2 en = pack1_val<FULL>(P, pn, keep);
    if (t.halo) hG = pack1_val<FULL>(P, hp, keep).y;
    __syncthreads();
    const float pW = s_p[buf][t.ty + 1][t.tx], pE = s_p[buf][t.ty + 1][t.tx + 2];
    const float pS = s_p[buf][t.ty][t.tx + 1], pN = s_p[buf][t.ty + 2][t.tx + 1];
    const float gW = s_G[buf][t.ty + 1][t.tx], gE = s_G[buf][t.ty + 1][t.tx + 2];
    const float gS = s_G[buf][t.ty][t.tx + 1], gN = s_G[buf][t.ty + 2][t.tx + 1];
    const float p1 = pc, G = Gc, Gn = en.y;
    const float GW = __fmul_rn(__fadd_rn(G, gW), 0.5f), GE = __fmul_rn(__fadd_rn(gE, G), 0.5f);
    const float GS = __fmul_rn(__fadd_rn(G, gS), 0.5f), GN = __fmul_rn(__fadd_rn(gN, G), 0.5f);
    const float GD = __fmul_rn(__fadd_rn(G, Gm), 0.5f), GU = __fmul_rn(__fadd_rn(Gn, G), 0.5f);
    // C*k_f*krg*G_f*(1/dl)*(1/dl)                       physics_loss.py:152-155
    const float a1 = __fmul_rn(__fmul_rn(__fmul_rn(fW, GW), P.idx), P.idx);
    const float a2 = __fmul_rn(__fmul_rn(__fmul_rn(fS, GS), P.idy), P.idy);
    const float a3 = __fmul_rn(__fmul_rn(__fmul_rn(fE, GE), P.idx), P.idx);
    const float a4 = __fmul_rn(__fmul_rn(__fmul_rn(fN, GN), P.idy), P.idy);
    const float a5 = __fmul_rn(__fmul_rn(__fmul_rn(fD, GD), P.idz), P.idz);
    const float a6 = __fmul_rn(__fmul_rn(__fmul_rn(fU, GU), P.idz), P.idz);
    // accumulation coefficient                           physics_loss.py:149-150,156
    const float A0 = e0.x, cp = e0.y, A1 = A1c;                    // cp: physics_loss.py:149-150, tabulated with the spline
    const float a5t = __fmul_rn(P.invDc, div_c(cp, by_d1));
    // wells in this cell (scatter_nd sums duplicates)    well_rate_bhp_Subclassed.py:128-132
    float qdv = 0.f, mask = 0.f;
    int wfirst = 0;
    if (has_well) {
      float q = 0.f;
      wfirst = well_lower_bound(P, off);
      for (int w = wfirst; w < P.n_wells && P.wells[w].cell == off; ++w) {
        q = __fadd_rn(q, A.qw[(int64_t)b * P.n_wells + w]);
        mask += 1.f;
      }
      if (mask != 0.f) qdv = __fdiv_rn(q, P.dv);
    }
    // p2 by linear extrapolation, truncation term        physics_loss.py:126,171
    const float p2 = __fadd_rn(__fmul_rn(__fsub_rn(p1, p0), one_rho), p0);
    const float numr = __fsub_rn(__fadd_rn(__fmul_rn(d2, p0), __fmul_rn(d1, p2)), __fmul_rn(d12, p1));
    const float E = __fadd_rn(c2e7, div_c(numr, by_den));
    const float tde = __fmul_rn(__fmul_rn(P.dvDc, cp), E);
    // flux divergence                                    physics_loss.py:174
    float s = __fadd_rn(-__fmul_rn(a1, pW), -__fmul_rn(a2, pS));
    const float asum = __fadd_rn(__fadd_rn(__fadd_rn(a1, a2), a3), a4);
    s = __fadd_rn(s, __fmul_rn(asum, p1));
    s = __fadd_rn(s, -__fmul_rn(a3, pE));
    s = __fadd_rn(s, -__fmul_rn(a4, pN));
    const float zt = __fadd_rn(__fmul_rn(a5, __fsub_rn(p1, pm)), __fmul_rn(a6, __fsub_rn(p1, pn)));   // 3-D extension
    s = __fadd_rn(s, zt);
    s = __fadd_rn(s, qdv);
    const float divq = __fmul_rn(P.dv, s);
    const float acc = __fmul_rn(__fmul_rn(P.dv, a5t), __fsub_rn(p1, p0));       // physics_loss.py:175
    const float dom = P.tde_in_dom ? __fadd_rn(divq, __fadd_rn(acc, tde)) : __fadd_rn(divq, acc);
    const float mb = __fmul_rn(__fmul_rn(P.dvSgi_phi, __fsub_rn(A1, A0)), mbfac);   // physics_loss.py:193
    if (t.valid) {
      domf[off] = dom;
      if (domo) domo[off] = dom;
      if (mask != 0.f) {
        for (int w = wfirst; w < P.n_wells && P.wells[w].cell == off; ++w) A.divqw[(int64_t)b * P.n_wells + w] = divq;
        const float ibc = __fmul_rn(mask, divq);                                // physics_loss.py:189
        a_ibc += (double)ibc * (double)ibc;
      }
      a_dom = fmaf(dom, dom, a_dom);
      a_tde = fmaf(tde, tde, a_tde);
      a_mb += (double)mb;
    }
    pm = pc; pc = pn; Gm = Gc; Gc = Gn; A1c = en.x; fD = fU;
    off += HW; offh += HW; offE += strE; offN += strN;
  };
  int k = 0;
  for (; k + 1 < D; k += 2) { plane(IntC<0>(), k); plane(IntC<1>(), k + 1); }
  if (k < D) plane(IntC<0>(), k);

  double acc4[4] = {(double)a_dom, a_ibc, (double)a_tde, a_mb};
  __syncthreads();
  block_reduce<4>(acc4, red);
  if (threadIdx.x == 0) {
    atomicAdd(&A.sse[SRM_TERM_DOM], acc4[0]);
    if (acc4[1] != 0.0) atomicAdd(&A.sse[SRM_TERM_IBC], acc4[1]);
    atomicAdd(&A.sse[SRM_TERM_TDE], acc4[2]);
    atomicAdd(&A.mb_sum[b], acc4[3]);
  }
}

// sum of every sample's well rates (fp64)                 physics_loss.py:193
__global__ void __launch_bounds__(128) k_qsum_ref2(int32_t B, int32_t nw, const float* __restrict__ qw, double* __restrict__ q_sum) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B) return;
  double s = 0.0;
  for (int w = lane; w < nw; w += 32) s += (double)qw[(int64_t)warp * nw + w];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) q_sum[warp] = s;
}

// ------------------------------------------------------------------------------------------
// adjoint
// ------------------------------------------------------------------------------------------
template <bool FULL>
__global__ void __launch_bounds__(NT, 2) k_adj_ref2(const __grid_constant__ SrmDev P, const __grid_constant__ R2Args A) {
  __shared__ float s_p[2][SH][SW];
  __shared__ float s_G[2][SH][SW];
  __shared__ float s_s[2][SH][SW];
  __shared__ double red[2 * 32];
  __shared__ unsigned char s_flag[NT];
  const Tile t = make_tile(P, A.tiles_x);
  const int b = blockIdx.y;
  const int r = srm_real_of(A.sample_real, b, A.B, A.R);
  const bool has_well = (P.n_wells > 0) ? column_has_well(P, t, s_flag) : false;
  const int W = P.W, H = P.H, D = P.D, HW = H * W;
  const FaceLay FL = face_layout(D, H, W);
  const float* __restrict__ p0f = A.p0 + (int64_t)b * P.N;
  const float* __restrict__ p1f = A.p1 + (int64_t)b * P.N;
  const float* __restrict__ domf = A.dom + (int64_t)b * P.N;
  float* __restrict__ gp0f = A.gp0 + (int64_t)b * P.N;
  float* __restrict__ gp1f = A.gp1 + (int64_t)b * P.N;
  const float* __restrict__ FE = A.faces + (int64_t)r * FL.per_real;
  const float* __restrict__ FN = FE + FL.nE;
  const float* __restrict__ FU = FN + FL.nN;
  const int yy = t.oc / W, xx = t.oc - yy * W;
  int offE = yy * FL.WP + xx, offN = yy * W + xx;
  const int strE = H * FL.WP, strN = (H + 1) * W;
  const float w_tde2 = 2.f * A.dterms[SRM_TERM_TDE];
  const float d1 = A.dt1[b], d2 = A.dt2[b];
  const float two_wd = 2.f * A.dterms[SRM_TERM_DOM];
  const float smb = 2.f * A.dterms[SRM_TERM_MBC] * A.mbc[b];              // dL/d mbc_b
  // forward's per-sample scalars (op order as the forward: E is dominated by the rounding of N)
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

  const uint64_t keep = l2_evict_last();
  int off = t.oc, offh = t.oh;
  float m1c;
  float pc = p1f[off];
  float4 e1 = pack1_at<FULL>(P, pc, m1c, keep);
  float sc = two_wd * domf[off];
  float pm = pc, Gm = e1.y, sm = sc;
  float hp = 0.f, hG = 0.f, hs = 0.f;
  if (t.halo) { hp = p1f[offh]; hG = pack1_val<FULL, true>(P, hp, keep).y; hs = two_wd * domf[offh]; }
  float fD = FU[off];
  double a_g1 = 0.0, a_g2 = 0.0;

  auto plane = [&](auto BUF, const int k) {
    constexpr int buf = decltype(BUF)::value;
    s_p[buf][t.ty + 1][t.tx + 1] = pc;
    s_G[buf][t.ty + 1][t.tx + 1] = e1.y;
    s_s[buf][t.ty + 1][t.tx + 1] = sc;
    if (t.halo) { s_p[buf][t.hy][t.hx] = hp; s_G[buf][t.hy][t.hx] = hG; s_s[buf][t.hy][t.hx] = hs; }
    const int up = (k + 1 < D) ? HW : 0;
    const float pn = p1f[off + up];
    const float sn = two_wd * domf[off + up];
    if (t.halo) { hp = p1f[offh + up]; hs = two_wd * domf[offh + up]; }
    const float p0 = p0f[off];
    const float fW = FE[offE], fE = FE[offE + 1];
    const float fS = FN[offN], fN = FN[offN + W];
    const float fU = FU[off + HW];
    float m0, m1n;
    const float4 e0 = pack0_at<FULL>(P, p0, m0, keep);
    const float4 en = pack1_at<FULL>(P, pn, m1n, keep);
    if (t.halo) hG = pack1_val<FULL, true>(P, hp, keep).y;
    __syncthreads();
    const float p1 = pc, G = e1.y, Gp = e1.w * m1c, A1 = e1.x, A1p = e1.z * m1c;
    // stencil gather: dv * sum_f (s_c - s_n) * T_f/2 * [(G_c + G_n) + G'_c (p_c - p_n)]; image faces: s_n == s_c
    float g1 = 0.f;
    {
      const float pW = s_p[buf][t.ty + 1][t.tx], gW = s_G[buf][t.ty + 1][t.tx], sW = s_s[buf][t.ty + 1][t.tx];
      g1 = fmaf((sc - sW) * (fW * hx2), (G + gW) + Gp * (p1 - pW), g1);
      const float pE = s_p[buf][t.ty + 1][t.tx + 2], gE = s_G[buf][t.ty + 1][t.tx + 2], sE = s_s[buf][t.ty + 1][t.tx + 2];
      g1 = fmaf((sc - sE) * (fE * hx2), (G + gE) + Gp * (p1 - pE), g1);
      const float pS = s_p[buf][t.ty][t.tx + 1], gS = s_G[buf][t.ty][t.tx + 1], sS = s_s[buf][t.ty][t.tx + 1];
      g1 = fmaf((sc - sS) * (fS * hy2), (G + gS) + Gp * (p1 - pS), g1);
      const float pN = s_p[buf][t.ty + 2][t.tx + 1], gN = s_G[buf][t.ty + 2][t.tx + 1], sN = s_s[buf][t.ty + 2][t.tx + 1];
      g1 = fmaf((sc - sN) * (fN * hy2), (G + gN) + Gp * (p1 - pN), g1);
      g1 = fmaf((sc - sm) * (fD * hz2), (G + Gm) + Gp * (p1 - pm), g1);
      g1 = fmaf((sc - sn) * (fU * hz2), (G + en.y) + Gp * (p1 - pn), g1);
    }
    g1 *= P.dv;
    // local terms
    const float A0 = e0.x, A0p = e0.y, A0pm = e0.y * m0, A0pp = e0.z * m0;
    const float cp = P.Sgi * (P.phi * A0p + P.phicf * A0);
    const float cpp = P.Sgi * (P.phi * A0pp + P.phicf * A0pm);   // d cp / d p0
    const float dva5t = dvi * cp;                                 // dv * a5t
    const float dp = p1 - p0;
    const float p2 = __fadd_rn(__fmul_rn(__fsub_rn(p1, p0), one_rho), p0);
    const float numr = __fsub_rn(__fadd_rn(__fmul_rn(d2, p0), __fmul_rn(d1, p2)), __fmul_rn(d12, p1));
    const float E = __fadd_rn(c2e7, div_c(numr, by_den));
    const float cE = P.dvDc * cp;
    const float tde = __fmul_rn(__fmul_rn(P.dvDc, cp), E);
    const float st = (P.tde_in_dom ? sc : 0.f) + w_tde2 * tde;   // dL/d tde
    float dq = 0.f;
    if (has_well) {
      const int first = well_lower_bound(P, off);
      for (int w = first; w < P.n_wells && P.wells[w].cell == off; ++w) dq += A.dqdp[(int64_t)b * P.n_wells + w];
    }
    g1 += sc * (dq + dva5t) - smb * dq - smbk * A1p;
    const float g0 = sc * (dvi * dp * cpp - dva5t) + st * P.dvDc * cpp * E + smbk * A0pm;
    if (t.valid) {
      gp0f[off] = g0;
      gp1f[off] = g1;
      // d/d dt1, d/d dt2 (the dN/d* pieces vanish identically; N itself is rounding noise)
      const float dE1 = dE1c - numr * dE1n;
      const float dE2 = -numr * dE2n;
      a_g1 += (double)((smbk * (A1 - A0) - sc * dva5t * dp) * id1 + st * cE * dE1);
      a_g2 += (double)(st * cE * dE2);
    }
    pm = pc; pc = pn; Gm = e1.y; e1 = en; m1c = m1n; sm = sc; sc = sn; fD = fU;
    off += HW; offh += HW; offE += strE; offN += strN;
  };
  int k = 0;
  for (; k + 1 < D; k += 2) { plane(IntC<0>(), k); plane(IntC<1>(), k + 1); }
  if (k < D) plane(IntC<0>(), k);

  double acc2[2] = {a_g1, a_g2};
  __syncthreads();
  block_reduce<2>(acc2, red);
  if (threadIdx.x == 0) {
    atomicAdd(&A.gdt1_acc[b], acc2[0]);
    atomicAdd(&A.gdt2_acc[b], acc2[1]);
  }
}

// inner-boundary term: L_ibc = w_ibc * sum (mask*divq)^2 ; scatter d divq_c / d p1 to the cell and
// its six neighbours (atomics: adjacent well cells may hit the same target).
template <bool FULL>
__global__ void __launch_bounds__(128) k_ibc_adj_ref2(const __grid_constant__ SrmDev P, const __grid_constant__ R2Args A) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int nw = P.n_wells;
  if (g >= (int64_t)A.B * nw) return;
  const int b = (int)(g / nw), w = (int)(g % nw);
  const int c = P.wells[w].cell;
  if (w > 0 && P.wells[w - 1].cell == c) return;   // one thread per distinct cell
  float mask = 0.f, dq = 0.f;
  for (int u = w; u < nw && P.wells[u].cell == c; ++u) { mask += 1.f; dq += A.dqdp[(int64_t)b * nw + u]; }
  const float s = 2.f * A.dterms[SRM_TERM_IBC] * mask * mask * A.divqw[g];
  if (s == 0.f) return;
  const int r = srm_real_of(A.sample_real, b, A.B, A.R);
  const int W = P.W, H = P.H, D = P.D, HW = H * W;
  const FaceLay FL = face_layout(D, H, W);
  const float* FE = A.faces + (int64_t)r * FL.per_real;
  const float* FN = FE + FL.nE;
  const float* FU = FN + FL.nN;
  const float* p1f = A.p1 + (int64_t)b * P.N;
  float* gp1 = A.gp1 + (int64_t)b * P.N;
  const int i = c % W, j = (c / W) % H, k = c / HW;
  float m1, mn;
  const float p1 = p1f[c];
  const uint64_t keep = l2_evict_last();
  const float4 e1 = pack1_at<FULL>(P, p1, m1, keep);
  const float G = e1.y, Gp = e1.w * m1;
  float self = 0.f;
  auto face = [&](bool inside, int cn, float f, float h2) {
    if (!inside) return;
    const float pn = p1f[cn];
    const float4 en = pack1_at<FULL>(P, pn, mn, keep);
    const float Tf = f * h2 * 2.f;
    const float af = Tf * 0.5f * (G + en.y);
    self += af + 0.5f * Tf * Gp * (p1 - pn);
    atomicAdd(&gp1[cn], s * P.dv * (-af + 0.5f * Tf * (en.w * mn) * (p1 - pn)));
  };
  const float hx2 = 0.5f * P.idx * P.idx, hy2 = 0.5f * P.idy * P.idy, hz2 = 0.5f * P.idz * P.idz;
  face(i > 0, c - 1, FE[((int64_t)k * H + j) * FL.WP + i], hx2);
  face(i < W - 1, c + 1, FE[((int64_t)k * H + j) * FL.WP + i + 1], hx2);
  face(j > 0, c - W, FN[((int64_t)k * (H + 1) + j) * W + i], hy2);
  face(j < H - 1, c + W, FN[((int64_t)k * (H + 1) + j + 1) * W + i], hy2);
  face(k > 0, c - HW, FU[(int64_t)k * HW + j * W + i], hz2);
  face(k < D - 1, c + HW, FU[(int64_t)(k + 1) * HW + j * W + i], hz2);
  atomicAdd(&gp1[c], s * (P.dv * self + dq));
}

}  // namespace

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
#ifdef SRM_WITH_DG5     // the warp-specialised forward experiment (kernels_dg5.cu); not part of the default build
bool srm_dg5_applicable(const SrmHandle* h);
cudaError_t srm_dg5_launch_fwd(const SrmHandle* h, const void* args, int32_t B, cudaStream_t s);
#endif
bool srm_dg4_applicable(const SrmHandle* h);
cudaError_t srm_dg4_launch_fwd(const SrmHandle* h, const void* args, int32_t B, cudaStream_t s);
cudaError_t srm_dg4_launch_adj(const SrmHandle* h, const void* args, int32_t B, cudaStream_t s);
cudaError_t srm_dg4_finalize_adj(const void* args, int32_t B, float* gdt1, float* gdt2, cudaStream_t s);

// the lean kernels (kernels_dg4.cu) need the table over the whole clamp range and vector-aligned fields.  The choice is
// a function of the handle and of pointer alignment only (no environment reads on the call path); the forward records
// its family with the saved state and a backward of the other family recomputes the forward in its own.
static bool use_dg4(const SrmHandle* h, const void* a, const void* b, const void* c, const void* d, const void* e, const void* f) {
  auto ok = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  return !h->no_dg4 && srm_dg4_applicable(h) && ok(a) && ok(b) && ok(c) && ok(d) && ok(e) && ok(f);
}
int srm_ref2_backward_family(const SrmHandle* h, const float* p0, const float* p1, const void* dom_ws, const float* gp0, const float* gp1) {
  return use_dg4(h, p0, p1, dom_ws, nullptr, gp0, gp1) ? 1 : 0;
}

size_t srm_ref2_face_floats(const SrmDev& P) { return (size_t)face_layout(P.D, P.H, P.W).per_real; }

static R2Args make_args(const SrmDev& P, int32_t B, int32_t R, const int32_t* sample_real, const float* p0,
                        const float* p1, const float* dt1, const float* dt2, const SrmWs& ws) {
  R2Args A;
  memset(&A, 0, sizeof(A));
  A.p0 = p0; A.p1 = p1; A.dt1 = dt1; A.dt2 = dt2; A.sample_real = sample_real;
  A.faces = ws.faces;
  A.qw = ws.qw; A.divqw = ws.divqw; A.dom = ws.dom; A.sse = ws.sse; A.mb_sum = ws.mb_sum;
  A.mbc = ws.mbc; A.dqdp = ws.dqdp; A.gdt1_acc = ws.gdt1_acc; A.gdt2_acc = ws.gdt2_acc;
  A.B = B; A.R = R; A.tiles_x = (P.W + TX - 1) / TX;
  return A;
}

int srm_forward_ref2(SrmHandle* h, int32_t B, int32_t R, const float* kx, const int32_t* sample_real,
                     const float* p0, const float* p1, const float* dt1, const float* dt2, const float* t1,
                     float* terms_out, float* dom_out, const SrmWs& ws, cudaStream_t s, int force_family, bool save) {
  const SrmDev& P = h->dev;
  SRM_CUDA_CHECK(cudaMemsetAsync(ws.sse, 0, (char*)ws.mbc - (char*)ws.sse, s));   // sse, mb_sum, q_sum, gdt accs
  const FaceLay FL = face_layout(P.D, P.H, P.W);
  k_faces_ref<<<dim3((unsigned)((FL.per_real + 255) / 256), (unsigned)R), 256, 0, s>>>(P, kx, ws.faces);
  SRM_CUDA_CHECK(cudaGetLastError());
  int rc = srm_launch_wells_ref(h, B, kx, sample_real, R, p1, t1, ws.qw, ws.pwfw, ws.dqdp, s);
  if (rc) return rc;
  if (P.n_wells > 0) {
    k_qsum_ref2<<<(unsigned)((B * 32 + 127) / 128), 128, 0, s>>>(B, P.n_wells, ws.qw, ws.q_sum);
    SRM_CUDA_CHECK(cudaGetLastError());
  }
  R2Args A = make_args(P, B, R, sample_real, p0, p1, dt1, dt2, ws);
  A.dom_out = dom_out;
  const bool lean = force_family >= 0 ? (force_family == 1 && use_dg4(h, p0, p1, ws.dom, dom_out, nullptr, nullptr))
                                      : use_dg4(h, p0, p1, ws.dom, dom_out, nullptr, nullptr);
  h->st_family = lean ? 1 : 0;
  A.pk = (lean && save) ? ws.pk : nullptr;            // the lean forward stages the adjoint's table values
  A.pk_stride = (int64_t)B * P.N;
#ifdef SRM_WITH_DG5
  if (lean && srm_dg5_applicable(h)) {
    SRM_CUDA_CHECK(srm_dg5_launch_fwd(h, &A, B, s));
  } else
#endif
  if (lean) {
    SRM_CUDA_CHECK(srm_dg4_launch_fwd(h, &A, B, s));
  } else {
    const dim3 grid((unsigned)(A.tiles_x * ((P.H + TY - 1) / TY)), (unsigned)B);
    if (h->lut_full) k_fwd_ref2<true><<<grid, NT, 0, s>>>(P, A);
    else k_fwd_ref2<false><<<grid, NT, 0, s>>>(P, A);
    SRM_CUDA_CHECK(cudaGetLastError());
  }
  k_finalize_fwd<<<1, 256, 0, s>>>(P, B, ws.sse, ws.mb_sum, ws.q_sum, ws.mbc, terms_out);
  SRM_CUDA_CHECK(cudaGetLastError());
  return SRM_OK;
}

int srm_backward_ref2(SrmHandle* h, int32_t B, int32_t R, const float* kx, const int32_t* sample_real,
                      const float* p0, const float* p1, const float* dt1, const float* dt2, const float* t1,
                      const float* dterms, float* gp0, float* gp1, float* gdt1, float* gdt2,
                      const SrmWs& ws, cudaStream_t s) {
  const SrmDev& P = h->dev;
  (void)kx; (void)t1;
  SRM_CUDA_CHECK(cudaMemsetAsync(ws.gdt1_acc, 0, (char*)ws.mbc - (char*)ws.gdt1_acc, s));
  R2Args A = make_args(P, B, R, sample_real, p0, p1, dt1, dt2, ws);
  A.dterms = dterms; A.gp0 = gp0; A.gp1 = gp1;
  const bool dg4 = use_dg4(h, p0, p1, ws.dom, nullptr, gp0, gp1);
  A.pk = dg4 ? ws.pk : nullptr;                        // written by the forward that saved this state (same family)
  A.pk_stride = (int64_t)B * P.N;
  if (dg4) {
    SRM_CUDA_CHECK(srm_dg4_launch_adj(h, &A, B, s));
  } else {
    const dim3 grid((unsigned)(A.tiles_x * ((P.H + TY - 1) / TY)), (unsigned)B);
    if (h->lut_full) k_adj_ref2<true><<<grid, NT, 0, s>>>(P, A);
    else k_adj_ref2<false><<<grid, NT, 0, s>>>(P, A);
    SRM_CUDA_CHECK(cudaGetLastError());
  }
  const int64_t n = (int64_t)B * P.n_wells;
  if (n > 0) {
    k_ibc_adj_ref2<false><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(P, A);
    SRM_CUDA_CHECK(cudaGetLastError());
  }
  if (dg4) {
    SRM_CUDA_CHECK(srm_dg4_finalize_adj(&A, B, gdt1, gdt2, s));
  } else {
    k_finalize_adj<<<(unsigned)((B + 255) / 256), 256, 0, s>>>(B, ws.gdt1_acc, ws.gdt2_acc, gdt1, gdt2);
    SRM_CUDA_CHECK(cudaGetLastError());
  }
  return SRM_OK;
}
