// SRM_NUMERICS_REFERENCE, exact PVT table over the whole clamp range, W % 4 == 0: warp-specialised forward
// of the dry-gas physics loss (physics_loss.py:79-208, 742-870).  Same arithmetic as kernels_dg4.cu and
// kernels_ref2.cu (forward fields bit-identical to the pinned oracle).
//
// Why: the exact-table gathers cost one L1 tag cycle per lane (tools/gather_probe2.cu: 1.1 SM-cycles per gathered
// entry, whatever the entry size), about as long as the arithmetic of the cells they feed.  In kernels_dg4.cu
// every warp alternates between the two, so neither the L1 nor the issue slots stay busy (both ~50 %).  Here
//   * 4 PRODUCER warps stream p1/p0 (adjoint: and the seed) from HBM, gather the table entries and publish
//     finished planes -- p1, G with halo, the cell-local inputs without -- into a ring of shared-memory stages;
//     they keep two half-planes of gathers in flight per thread and never wait on arithmetic;
//   * 8 CONSUMER warps (16 x 16 threads, four x-adjacent cells each, marching over z) read stages, do the
//     stencil and the cell-local arithmetic and write the results; they never wait on a gather.
// full/empty mbarriers per stage, one arrival per warp.
//
// STATUS: correct (bit-exact, tests/test_gpu_parity.py) but slower than kernels_dg4.cu on B200 (profiles/r1_dg5_*):
// one 384-thread CTA per SM leaves 8 consumer warps to issue ~115 instructions per cell at IPC ~0.35 per scheduler,
// both sides spend ~15 % of their time in each other's mbarrier, and nothing covers a CTA's pipeline fill and drain.
// Not dispatched unless SRM_DG5=1.
#include <cstdlib>
#include <cstring>
#include "ref_fused.cuh"
#include "dg_lean.cuh"

namespace {

constexpr int CX5 = 16, CPT5 = 4, TW5 = CX5 * CPT5, TY5 = 16;
constexpr int NCONS = CX5 * TY5, NPROD = 128, NT5 = NCONS + NPROD;
constexpr int XO5 = 4, SW5 = TW5 + 2 * XO5, SH5 = TY5 + 2;
constexpr int PLH = SH5 * SW5;            // haloed plane (floats)
constexpr int PL = TY5 * TW5;             // plain plane
constexpr int NHALO5 = 2 * TW5 + 2 * TY5; // 160
constexpr int NS5 = 4;                    // ring depth: the consumers hold planes k and k+1, the producers fill k+2, k+3
static_assert(NPROD == 128 && TW5 == 64 && TY5 == 16, "producer mapping: 16 quads x 8 rows, two half-planes");
static_assert(NHALO5 / 2 <= NPROD, "halo cells of a half-plane: one per producer thread");

__device__ __forceinline__ void lds4(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void sts4(float* p, float a, float b, float c, float d) { *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d); }
__device__ __forceinline__ void ldg4c(const float* p, float* v) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void ldg4s(const float* p, float (&v)[4], uint64_t pol) {
  const float4 t = ld_hint(reinterpret_cast<const float4*>(p), pol);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void stg4s(float* p, const float (&v)[4], uint64_t pol) {
  st_hint(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]), pol);
}
__device__ __forceinline__ void mbar_arrive_warp5(uint64_t* bar) { mbar_arrive_warp(bar); }

// forward stage: haloed p1 and G, plain p0, cp, A1 - A0
struct StageF { float p[PLH]; float G[PLH]; float p0[PL]; float cp[PL]; float dA[PL]; };
struct SmemF {
  StageF st[NS5];
  double red[4 * 32];
  uint64_t full[NS5], empty[NS5];
  unsigned char flag[TY5][TW5];
};

// halo cell h in [0, 160): rows y0-1 and y0+TY, columns x0-1 and x0+TW; clamped coordinates = edge replication
__device__ __forceinline__ void halo_cell(const SrmDev& P, int x0, int y0, int h, int& goff, int& slot) {
  int gx, gy, hr, hc;
  if (h < TW5) { gy = y0 - 1; gx = x0 + h; hr = 0; hc = XO5 + h; }
  else if (h < 2 * TW5) { gy = y0 + TY5; gx = x0 + h - TW5; hr = TY5 + 1; hc = XO5 + h - TW5; }
  else if (h < 2 * TW5 + TY5) { gx = x0 - 1; gy = y0 + h - 2 * TW5; hr = h - 2 * TW5 + 1; hc = XO5 - 1; }
  else { gx = x0 + TW5; gy = y0 + h - 2 * TW5 - TY5; hr = h - 2 * TW5 - TY5 + 1; hc = XO5 + TW5; }
  goff = min(max(gy, 0), P.H - 1) * P.W + min(max(gx, 0), P.W - 1);
  slot = hr * SW5 + hc;
}

// ------------------------------------------------------------------------------------------
// forward                                                       physics_loss.py:143-193,787-807
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT5, 1) k_fwd5(const __grid_constant__ SrmDev P, const __grid_constant__ R2Args A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemF& S = *reinterpret_cast<SmemF*>(smem_raw);
  const int tid = threadIdx.x;
  const int W = P.W, H = P.H, D = P.D, HW = H * W;
  const int tyi = blockIdx.x / A.tiles_x, txi = blockIdx.x - tyi * A.tiles_x;
  const int x0 = txi * TW5, y0 = tyi * TY5;
  const int b = blockIdx.y;
  const float* __restrict__ p0f = A.p0 + (int64_t)b * P.N;
  const float* __restrict__ p1f = A.p1 + (int64_t)b * P.N;
  const uint64_t keep = l2_evict_last(), strm = l2_evict_first();

  if (tid == 0) {
    for (int s = 0; s < NS5; ++s) { mbar_init(&S.full[s], NPROD / 32); mbar_init(&S.empty[s], NCONS / 32); }
  }
  // connection columns of this tile (any layer)
  for (int i = tid; i < TY5 * TW5; i += NT5) (&S.flag[0][0])[i] = 0;
  __syncthreads();
  for (int w = tid; w < P.n_wells; w += NT5) {
    const int rem = P.wells[w].cell % HW;
    const int j = rem / W, i = rem - j * W;
    if (i >= x0 && i < x0 + TW5 && j >= y0 && j < y0 + TY5) S.flag[j - y0][i - x0] = 1;
  }
  __syncthreads();

  double acc4[4] = {0.0, 0.0, 0.0, 0.0};

  if (tid >= NCONS) {
    // ================================ producers ================================
    const int pt = tid - NCONS;
    const int q = pt & 15, r0 = pt >> 4;
    const Tab TF = make_tab(P.lutf0, P, 16);
    const int xc = min(x0 + 4 * q, W - 4);
    // unit u = 2 * plane + j: rows r0 + 8 j of the tile, and halo cells 80 j + pt (pt < 80)
    int own0, own1, hoff0, hoff1, hslot0, hslot1;
    own0 = min(y0 + r0, H - 1) * W + xc;
    own1 = min(y0 + r0 + 8, H - 1) * W + xc;
    halo_cell(P, x0, y0, min(pt, NHALO5 - 1), hoff0, hslot0);
    halo_cell(P, x0, y0, min(80 + pt, NHALO5 - 1), hoff1, hslot1);
    const bool hal = pt < NHALO5 / 2;
    const int nu = 2 * D;
    // a: loaded pressures of unit u+2;  b: pressures + gathers in flight of unit u+1;  c: ... of unit u (to store)
    float a1[4], a0[4], ah = 0.f;
    float b1[4], b0[4], bh = 0.f; float2 be1[4], be0[4]; float bg = 0.f;
    float c1[4], c0[4], ch = 0.f; float2 ce1[4], ce0[4]; float cg = 0.f;
    auto load = [&](int u, float (&x1)[4], float (&x0_)[4], float& xh) {
      const int pl = (u >> 1) * HW, j = u & 1;
      ldg4s(p1f + pl + (j ? own1 : own0), x1, strm);
      ldg4s(p0f + pl + (j ? own1 : own0), x0_, strm);
      if (hal) xh = ld_hint(p1f + pl + (j ? hoff1 : hoff0), strm);
    };
    auto gather = [&](const float (&x1)[4], const float (&x0_)[4], float xh, float2 (&e1)[4], float2 (&e0)[4], float& g) {
#pragma unroll
      for (int c = 0; c < 4; ++c) { e1[c] = GATF1(TF, x1[c]); e0[c] = GATF0(TF, x0_[c]); }
      if (hal) g = GATF1(TF, xh).y;
    };
    // prologue: unit 0 gathered (c), unit 1 gathered (b), unit 2 loaded (a)
    load(0, c1, c0, ch);
    load(1, b1, b0, bh);
    if (2 < nu) load(2, a1, a0, ah);
    gather(c1, c0, ch, ce1, ce0, cg);
    gather(b1, b0, bh, be1, be0, bg);
    for (int u = 0; u < nu; ++u) {
      const int k = u >> 1, j = u & 1, s = k % NS5;
      if (j == 0) mbar_wait(&S.empty[s], ((k / NS5) & 1) ^ 1);
      // store unit u (c)
      StageF& st = S.st[s];
      const int row = r0 + 8 * j;
      const int oh = (row + 1) * SW5 + XO5 + 4 * q, op = row * TW5 + 4 * q;
      sts4(st.p + oh, c1[0], c1[1], c1[2], c1[3]);
      sts4(st.G + oh, ce1[0].y, ce1[1].y, ce1[2].y, ce1[3].y);
      sts4(st.p0 + op, c0[0], c0[1], c0[2], c0[3]);
      sts4(st.cp + op, ce0[0].y, ce0[1].y, ce0[2].y, ce0[3].y);
      sts4(st.dA + op, __fsub_rn(ce1[0].x, ce0[0].x), __fsub_rn(ce1[1].x, ce0[1].x), __fsub_rn(ce1[2].x, ce0[2].x), __fsub_rn(ce1[3].x, ce0[3].x));
      if (hal) { st.p[j ? hslot1 : hslot0] = ch; st.G[j ? hslot1 : hslot0] = cg; }
      if (j == 1) mbar_arrive_warp5(&S.full[s]);
      // rotate: b -> c, gather a -> b, load u+3 -> a
#pragma unroll
      for (int c = 0; c < 4; ++c) { c1[c] = b1[c]; c0[c] = b0[c]; ce1[c] = be1[c]; ce0[c] = be0[c]; }
      ch = bh; cg = bg;
      if (u + 2 < nu) {
#pragma unroll
        for (int c = 0; c < 4; ++c) { b1[c] = a1[c]; b0[c] = a0[c]; }
        bh = ah;
        gather(b1, b0, bh, be1, be0, bg);
      }
      if (u + 3 < nu) load(u + 3, a1, a0, ah);
    }
  } else {
    // ================================ consumers ================================
    const int cx = tid & (CX5 - 1), ty = tid / CX5;
    const int x = x0 + 4 * cx, y = y0 + ty;
    const bool valid = x < W && y < H;
    const int xcl = min(x, W - 4), ycl = min(y, H - 1);
    const bool edgeE = xcl + 4 >= W;
    const int oc = ycl * W + xcl;
    const int r = srm_real_of(A.sample_real, b, A.B, A.R);
    const FaceLay FL = face_layout(D, H, W);
    float* __restrict__ domf = A.dom + (int64_t)b * P.N;
    const float* __restrict__ FB = A.faces + (int64_t)r * FL.per_real;      // [FE | FN | FU]
    const int strE = H * FL.WP, strN = (H + 1) * W;
    bool has_well = false;
    if (valid) {
#pragma unroll
      for (int c = 0; c < 4; ++c) has_well |= S.flag[ty][4 * cx + c] != 0;
    }
    // per-sample scalars                                   physics_loss.py:126,156,171,193
    const float d1 = A.dt1[b], d2 = A.dt2[b];
    const float rho = (d1 == 0.f) ? 0.f : __fdiv_rn(d2, d1);
    const float one_rho = __fadd_rn(1.0f, rho);
    const DivC by_d1 = make_divc(d1);
    const DivC by_den = make_divc(__fadd_rn(__fmul_rn(d1, d2), __fmul_rn(d2, d2)));
    const float c2e7 = __fdiv_rn(2e-7f, d1);
    const float d12 = __fadd_rn(d1, d2);
    const float mbfac = __fdiv_rn(1.0f, __fmul_rn(P.Dc, d1));
    const bool div_slow = !(by_d1.ok && by_den.ok) || c2e7 == 0.f || !P.cp_safe;
    const int own_h = (ty + 1) * SW5 + XO5 + 4 * cx, own_p = ty * TW5 + 4 * cx;

    float a_dom = 0.f, a_tde = 0.f, a_mbf = 0.f;
    double a_ibc = 0.0, a_mb = 0.0;
    int off = oc, offE = (oc / W) * FL.WP + (oc % W), offN = (int)FL.nE + oc, offU = (int)(FL.nE + FL.nN) + oc + HW;
    float pc[4], Gc[4], tz[4];
    float fx[5], fS[4], fN[4], fU[4];
    ldg4c(FB + offE, fx); fx[4] = __ldg(FB + offE + 4);
    ldg4c(FB + offN, fS); ldg4c(FB + offN + W, fN); ldg4c(FB + offU, fU);
    mbar_wait(&S.full[0], 0);
    lds4(S.st[0].p + own_h, pc);
    lds4(S.st[0].G + own_h, Gc);
#pragma unroll
    for (int c = 0; c < 4; ++c) tz[c] = -0.0f;              // image face below plane 0: a5*(p - p) = +0

    for (int k = 0; k < D; ++k) {
      const StageF& st = S.st[k % NS5];
      const bool more = k + 1 < D;
      // everything this plane needs from shared memory, then the stage goes back to the producers
      float pn[4], Gn[4], pS[4], pN[4], gS[4], gN[4], p0[4], cp[4], dA[4];
      lds4(st.p + own_h - SW5, pS);
      lds4(st.p + own_h + SW5, pN);
      lds4(st.G + own_h - SW5, gS);
      lds4(st.G + own_h + SW5, gN);
      lds4(st.p0 + own_p, p0);
      lds4(st.cp + own_p, cp);
      lds4(st.dA + own_p, dA);
      float pWe = __shfl_up_sync(0xffffffffu, pc[3], 1, CX5), gWe = __shfl_up_sync(0xffffffffu, Gc[3], 1, CX5);
      float pEe = __shfl_down_sync(0xffffffffu, pc[0], 1, CX5), gEe = __shfl_down_sync(0xffffffffu, Gc[0], 1, CX5);
      if (cx == 0) { pWe = st.p[own_h - 1]; gWe = st.G[own_h - 1]; }
      if (cx == CX5 - 1) { pEe = st.p[own_h + 4]; gEe = st.G[own_h + 4]; }
      if (edgeE) { pEe = pc[3]; gEe = Gc[3]; }
      if (more) {
        const StageF& sn = S.st[(k + 1) % NS5];
        mbar_wait(&S.full[(k + 1) % NS5], ((k + 1) / NS5) & 1);
        lds4(sn.p + own_h, pn);
        lds4(sn.G + own_h, Gn);
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) { pn[c] = pc[c]; Gn[c] = Gc[c]; }        // edge-replicated image above the top plane
      }
      mbar_arrive_warp5(&S.empty[k % NS5]);
      // static face coefficients of the next plane: in flight during this plane's arithmetic
      float gx[5], gSn[4], gNn[4], gUn[4];
      {
        const int e2 = more ? strE : 0, n2 = more ? strN : 0, u2 = more ? HW : 0;
        ldg4c(FB + offE + e2, gx); gx[4] = __ldg(FB + offE + e2 + 4);
        ldg4c(FB + offN + n2, gSn); ldg4c(FB + offN + n2 + W, gNn); ldg4c(FB + offU + u2, gUn);
      }
      // ---- cell-local part: L = acc (+ tde)                                  physics_loss.py:156,171,175,193
      float L[4];
      {
        float dpv[4], numr[4], q1[4], q2[4];
        bool bad = div_slow;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          dpv[c] = __fsub_rn(pc[c], p0[c]);
          const float p2 = __fadd_rn(__fmul_rn(dpv[c], one_rho), p0[c]);
          numr[c] = __fsub_rn(__fadd_rn(__fmul_rn(d2, p0[c]), __fmul_rn(d1, p2)), __fmul_rn(d12, pc[c]));
          q1[c] = div_fast(cp[c], by_d1);
          q2[c] = div_fast(numr[c], by_den);
          bad |= div_operand_bad(numr[c]);
        }
        if (bad) {
#pragma unroll
          for (int c = 0; c < 4; ++c) { q1[c] = div_c(cp[c], by_d1); q2[c] = div_c(numr[c], by_den); }
        }
        float tsum = 0.f, msum = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float a5t = __fmul_rn(P.invDc, q1[c]);
          const float E = __fadd_rn(c2e7, q2[c]);
          const float tde = __fmul_rn(__fmul_rn(P.dvDc, cp[c]), E);
          const float acc = __fmul_rn(__fmul_rn(P.dv, a5t), dpv[c]);
          L[c] = P.tde_in_dom ? __fadd_rn(acc, tde) : acc;
          const float mb = __fmul_rn(__fmul_rn(P.dvSgi_phi, dA[c]), mbfac);
          tsum = fmaf(tde, tde, tsum);
          msum += mb;
        }
        if (valid) { a_tde += tsum; a_mbf += msum; }
      }
      // ---- stencil                                                           physics_loss.py:147-155,174-176
      float ax[5];
#pragma unroll
      for (int i = 0; i <= 4; ++i) {
        const float gl = (i == 0) ? gWe : Gc[i - (i > 0)], gr = (i == 4) ? gEe : Gc[i - (i == 4)];
        const float Gf = __fmul_rn(__fadd_rn(gr, gl), 0.5f);
        ax[i] = __fmul_rn(__fmul_rn(__fmul_rn(fx[i], Gf), P.idx), P.idx);
      }
      float domv[4], dsum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float p1 = pc[c], G = Gc[c];
        const float pW = (c == 0) ? pWe : pc[c - (c > 0)], pE = (c == 3) ? pEe : pc[c + (c < 3)];
        const float a1 = ax[c], a3 = ax[c + 1];
        const float a2 = __fmul_rn(__fmul_rn(__fmul_rn(fS[c], __fmul_rn(__fadd_rn(G, gS[c]), 0.5f)), P.idy), P.idy);
        const float a4 = __fmul_rn(__fmul_rn(__fmul_rn(fN[c], __fmul_rn(__fadd_rn(gN[c], G), 0.5f)), P.idy), P.idy);
        const float a6 = __fmul_rn(__fmul_rn(__fmul_rn(fU[c], __fmul_rn(__fadd_rn(Gn[c], G), 0.5f)), P.idz), P.idz);
        const float tu = __fmul_rn(a6, __fsub_rn(p1, pn[c]));
        const float zt = __fadd_rn(-tz[c], tu);
        tz[c] = tu;
        float qdv = 0.f, mask = 0.f;
        int wfirst = 0;
        const int cell = off + c;
        if (has_well) {
          float qq = 0.f;
          wfirst = well_lower_bound(P, cell);
          for (int w = wfirst; w < P.n_wells && P.wells[w].cell == cell; ++w) {
            qq = __fadd_rn(qq, A.qw[(int64_t)b * P.n_wells + w]);
            mask += 1.f;
          }
          if (mask != 0.f) qdv = __fdiv_rn(qq, P.dv);
        }
        float s = __fadd_rn(-__fmul_rn(a1, pW), -__fmul_rn(a2, pS[c]));
        const float asum = __fadd_rn(__fadd_rn(__fadd_rn(a1, a2), a3), a4);
        s = __fadd_rn(s, __fmul_rn(asum, p1));
        s = __fadd_rn(s, -__fmul_rn(a3, pE));
        s = __fadd_rn(s, -__fmul_rn(a4, pN[c]));
        s = __fadd_rn(s, zt);
        s = __fadd_rn(s, qdv);
        const float divq = __fmul_rn(P.dv, s);
        const float dom = __fadd_rn(divq, L[c]);
        domv[c] = dom;
        dsum = fmaf(dom, dom, dsum);
        if (has_well && mask != 0.f) {
          for (int w = wfirst; w < P.n_wells && P.wells[w].cell == cell; ++w) A.divqw[(int64_t)b * P.n_wells + w] = divq;
          const float ibc = __fmul_rn(mask, divq);
          a_ibc += (double)ibc * (double)ibc;
        }
      }
      if (valid) {
        stg4s(domf + off, domv, strm);
        a_dom += dsum;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) { pc[c] = pn[c]; Gc[c] = Gn[c]; fS[c] = gSn[c]; fN[c] = gNn[c]; fU[c] = gUn[c]; }
#pragma unroll
      for (int i = 0; i < 5; ++i) fx[i] = gx[i];
      off += HW; offE += strE; offN += strN; offU += HW;
      if ((k & 7) == 7) { a_mb += (double)a_mbf; a_mbf = 0.f; }
    }
    acc4[0] = (double)a_dom; acc4[1] = a_ibc; acc4[2] = (double)a_tde; acc4[3] = a_mb + (double)a_mbf;
  }

  __syncthreads();
  block_reduce<4>(acc4, S.red);
  if (tid == 0) {
    atomicAdd(&A.sse[SRM_TERM_DOM], acc4[0]);
    if (acc4[1] != 0.0) atomicAdd(&A.sse[SRM_TERM_IBC], acc4[1]);
    atomicAdd(&A.sse[SRM_TERM_TDE], acc4[2]);
    atomicAdd(&A.mb_sum[b], acc4[3]);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// launchers (called by kernels_ref2.cu)
// ------------------------------------------------------------------------------------------
bool srm_dg5_applicable(const SrmHandle* h) {
  const SrmDev& P = h->dev;
  return h->lut_full && P.lut_n > 0 && P.W % 4 == 0 && P.W >= 4 && P.D >= 2;
}

cudaError_t srm_dg5_launch_fwd(const SrmHandle* h, const void* args, int32_t B, cudaStream_t s) {
  const SrmDev& P = h->dev;
  R2Args A = *reinterpret_cast<const R2Args*>(args);
  A.tiles_x = (P.W + TW5 - 1) / TW5;
  const dim3 grid((unsigned)(A.tiles_x * ((P.H + TY5 - 1) / TY5)), (unsigned)B);
  static bool attr_dev[64] = {};              // function attributes are per device
  bool& attr = attr_dev[h->device & 63];
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_fwd5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemF));
    if (e != cudaSuccess) return e;
    attr = true;
  }
  k_fwd5<<<grid, NT5, sizeof(SmemF), s>>>(P, A);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && A.dom_out) e = cudaMemcpyAsync(A.dom_out, A.dom, sizeof(float) * (size_t)B * (size_t)P.N, cudaMemcpyDeviceToDevice, s);
  return e;
}
