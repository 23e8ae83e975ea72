// Fused gas-condensate pair (included by kernels_gc.cu, same namespace): the staged pipeline of kernels_gc.cu
// (k_stage_gc writes 18 / 30 fields per cell, the residual kernels re-read them for the cell and its six neighbours)
// collapsed into one forward and one adjoint kernel that gather the PVT packs from the exact table themselves.
// Needs the table over the whole clamp range (lut_full), so a lookup has no range test.
//
//   tile      32 x 8 columns per CTA (256 threads, one column each), marching over z
//   shared    the neighbour-visible values of the current plane, tile + halo ring, double-buffered: written for
//             plane k+1 while plane k is consumed, one barrier per plane.  Halo cells outside the grid take the
//             clamped coordinates: the edge-replicating pad of the reference (image faces see the cell itself).
//   registers the z neighbours: the thread's own values of planes k-1, k, k+1 rotate through registers
//   gathers   forward: {Mgg,Mgo,Moo,Mog} at p1 for tile + halo, {invBg,invBo,Rs*invBo,Rv*invBg} at p1 and the
//             32-byte level-n pack at p0 for the tile; adjoint: {Mgg+Mog, Mgo+Moo} at p1 for tile + halo, the rest of
//             the 64-byte and the 48-byte packs of the staged tables for the tile.  The halo ring's gathers are
//             cp.async copies straight into shared memory (no registers, landing while the plane is computed);
//             raw p0 / p1 / Sg1 (/ dom) of the next plane are loaded one plane ahead of the gathers they address.
//
// The arithmetic of each cell is the staged kernels' (same expressions, same order): the forward is bit-identical
// to k_stage_gc + k_resid_fwd_gc, the adjoint to k_resid_adj_gc.
#pragma once

constexpr int G2X = 32, G2Y = 8, G2PX = G2X + 2, G2PL = (G2Y + 2) * G2PX;
static_assert(G2X * G2Y == kThreads, "one column per thread");

using G2Wells = WellTile<kThreads, G2X, G2Y, 8, 192>;      // the tile's connections (well_tile.cuh)

struct G2Geom {
  int tx0, ty0, lt;                 // tile origin, own cell inside the tile (row-major)
  int gi, gj, ci, cj, col, so;      // own column: global, clamped, flat, shared slot
  bool active, halo;
  int hcol, hs;                     // halo duty of this thread: clamped flat column, shared slot
};
__device__ __forceinline__ G2Geom g2_geom(const SrmDev& P, int tiles_x) {
  G2Geom g;
  const int tile = blockIdx.x;
  const int ty0 = (tile / tiles_x) * G2Y, tx0 = (tile % tiles_x) * G2X;
  const int t = threadIdx.x, tx = t & (G2X - 1), ty = t / G2X;
  g.tx0 = tx0; g.ty0 = ty0; g.lt = ty * G2X + tx;
  g.gi = tx0 + tx; g.gj = ty0 + ty;
  g.active = g.gi < P.W && g.gj < P.H;
  g.ci = min(g.gi, P.W - 1); g.cj = min(g.gj, P.H - 1);
  g.col = g.cj * P.W + g.ci;
  g.so = (ty + 1) * G2PX + tx + 1;
  g.halo = t < 2 * G2X + 2 * G2Y;
  int hx, hy;
  if (t < G2X) { hx = t; hy = -1; }
  else if (t < 2 * G2X) { hx = t - G2X; hy = G2Y; }
  else if (t < 2 * G2X + G2Y) { hx = -1; hy = t - 2 * G2X; }
  else { hx = G2X; hy = t - 2 * G2X - G2Y; }
  g.hcol = min(max(ty0 + hy, 0), P.H - 1) * P.W + min(max(tx0 + hx, 0), P.W - 1);
  g.hs = (hy + 1) * G2PX + hx + 1;
  return g;
}

// asynchronous gather global -> shared (LDGSTS): the halo ring of plane k+1 lands while plane k is computed
__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
// Experiment switch: ask for the own-cell packs of the NEXT plane one plane ahead with prefetch.global (1: L2, 2: L1).
// Measured slower on B200 (cfg4: forward 16.6 -> 17.2 ms, adjoint 20.8 -> 25.3 ms): the prefetches are extra L2
// requests on a path that is already bound by the gather request rate.  Off by default.
#ifndef G2_UNROLL_F
#define G2_UNROLL_F 1
#endif
#ifndef G2_UNROLL_A
#define G2_UNROLL_A 2      // measured on config 4: adjoint 4.11 -> 4.07 ms at K = 8 (the forward loses 3.6 % with two planes per trip: 1)
#endif
#define G2_STR2(x) #x
#define G2_STR(x) G2_STR2(x)
#define G2_LOOP_F _Pragma(G2_STR(unroll G2_UNROLL_F))
#define G2_LOOP_A _Pragma(G2_STR(unroll G2_UNROLL_A))
#ifndef G2_PF
#define G2_PF 0
#endif
__device__ __forceinline__ void g2_prefetch(const void* p) {
#if G2_PF == 1
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#elif G2_PF == 2
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ uint32_t g2_entry(const SrmDev& P, float p, float& m) {
  const float x = srm_clamp(P, p, m);
  return __float_as_uint(x) - P.lut_lo_bits;
}

// ---- forward -------------------------------------------------------------------------------------
struct VisF { float4 m; float p1, krg, kro; };     // m = {Mgg, Mgo, Moo, Mog}
template <int NOG = 0, int NG = 0>
__device__ __forceinline__ VisF g2_vis_fwd(const SrmDev& P, const DivC* den, const float4* __restrict__ t1, float p1, float sg1) {
  float m1;
  const uint32_t e = g2_entry(P, p1, m1);
  VisF v;
  v.m = __ldg(t1 + 2 * (size_t)e);                  // {Mgg, Mgo, Moo, Mog}: the forward view is stored in component order
  float dko, dkg;
  corey<NOG, NG>(P, sg1, v.kro, v.krg, dko, dkg, den);
  v.p1 = p1;
  return v;
}

// LISTS: the tile's connections come from the staged column lists (well_tile.cuh; lattices of many connections);
// otherwise the few-connection path: exact column flags, a search inside the cell's layer
template <bool LISTS, int NOG, int NG>
__global__ void __launch_bounds__(kThreads, 2) k_fwd_gc2(const __grid_constant__ SrmDev P, const __grid_constant__ GcFwd A) {
  __shared__ float4 s_m[2][G2PL];
  __shared__ float4 s_q[2][G2PL];                  // {p1, krg, kro, -}
  __shared__ double red[7 * 32];
  const int b = blockIdx.y;
  const int HW = P.H * P.W;
  const G2Geom g = g2_geom(P, A.tiles_x);
  const DivC* kr_den = nullptr;      // forward: the IEEE intrinsic (the hoisted-reciprocal form costs this kernel registers: measured slower)
  const int r = srm_real_of(A.sample_real, b, A.B, A.R);
  const float4* __restrict__ T0 = reinterpret_cast<const float4*>(P.lutf0);
  const float4* __restrict__ T1 = reinterpret_cast<const float4*>(P.lutf1);
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};   // dom^2, ibc^2, trn^2, sum mg cells, sum mo cells, sum qg, sum qo
  const int64_t base = (int64_t)b * P.N;
  const float* __restrict__ P1 = A.p1 + base;
  const float* __restrict__ SG1 = A.sg1 + base;
  const FaceLay FL = face_layout(P.D, P.H, P.W);
  const float* __restrict__ FE = A.faces + (int64_t)r * FL.per_real;
  const float* __restrict__ FN = FE + FL.nE;
  const float* __restrict__ FU = FN + FL.nN;
  int fe = g.cj * FL.WP + g.ci, fn = g.col, fu = g.col;      // face offsets of plane 0 (per-realisation arrays: < 2^31 floats)
  // the tile's connections: column lists staged in shared memory; lists that do not fit keep the per-plane search
  __shared__ __align__(8) unsigned char s_wt_raw[LISTS ? sizeof(G2Wells) : 8];
  G2Wells& s_wt = *reinterpret_cast<G2Wells*>(s_wt_raw);
  bool col_wells = false;
  int wslot = 0;
  if constexpr (LISTS) {
    bool over = false;
    const bool any = s_wt.build(well_cols_of(P), P.W, P.D, g.tx0, g.ty0, nullptr, over);
    if (over) col_wells = g.active && column_has_well_gc(P, g.col, HW);
    else if (any && g.active) wslot = s_wt.slot_of[g.lt];
  } else {
    col_wells = g.active && column_has_well_gc(P, g.col, HW);
  }
  const float d1 = A.dt1[b], d2 = A.dt2[b];
  const float idl[6] = {P.idx, P.idx, P.idy, P.idy, P.idz, P.idz};
  const float idt = __fdiv_rn(1.0f, __fmul_rn(P.Dc, d1));
  const float rho1 = __fadd_rn(1.0f, (d1 == 0.f) ? 0.f : __fdiv_rn(d2, d1));
  const float den = __fadd_rn(__fmul_rn(d1, d2), __fmul_rn(d2, d2));
  const DivC den_c = make_divc(den);          // correctly rounded division by the per-sample constant (ref_fused.cuh)
  const float rte_d1 = __fdiv_rn(2.5e-8f, d1);                                  // :439-440
  const float d12 = __fadd_rn(d1, d2);
  const float mfac = __fmul_rn(__fmul_rn(P.dv, idt), P.phi);
  const int64_t wt = (int64_t)A.B * P.n_wells;

  // plane 0 into buffer 0; raw p1 / sg1 of the next plane travel one plane ahead of their gather
  float sg1c = SG1[g.col];
  VisF vC = g2_vis_fwd<NOG, NG>(P, nullptr, T1, P1[g.col], sg1c);
  s_m[0][g.so] = vC.m;
  s_q[0][g.so] = make_float4(vC.p1, vC.krg, vC.kro, 0.f);
  if (g.halo) {
    const VisF h = g2_vis_fwd<NOG, NG>(P, nullptr, T1, P1[g.hcol], SG1[g.hcol]);
    s_m[0][g.hs] = h.m;
    s_q[0][g.hs] = make_float4(h.p1, h.krg, h.kro, 0.f);
  }
  float p1n = 0.f, sg1n = 0.f, hp1 = 0.f, hsg1 = 0.f;
  float p0c = A.p0[base + g.col], p0n = 0.f;
  if (P.D > 1) {
    p1n = P1[HW + g.col]; sg1n = SG1[HW + g.col]; p0n = A.p0[base + HW + g.col];
    if (g.halo) { hp1 = P1[HW + g.hcol]; hsg1 = SG1[HW + g.hcol]; }
  }
  VisF vP = vC;
  __syncthreads();

  G2_LOOP_F
  for (int k = 0; k < P.D; ++k) {
    const int c = k * HW + g.col;
    const int buf = k & 1;
    // ---- stage plane k+1 (gathers issued before this plane's arithmetic)
    VisF vN = vC;
    float sg1nn = sg1c;
    const float p0 = p0c;
    if (k + 1 < P.D) {
      vN = g2_vis_fwd<NOG, NG>(P, nullptr, T1, p1n, sg1n);
      sg1nn = sg1n;
      {   // level-n pack of plane k+1
        float mm;
        g2_prefetch(T0 + 2 * (size_t)g2_entry(P, p0n, mm));
        p0c = p0n;
      }
      if (k + 2 < P.D) { p1n = P1[(k + 2) * HW + g.col]; sg1n = SG1[(k + 2) * HW + g.col]; p0n = A.p0[base + (k + 2) * HW + g.col]; }
      if (g.halo) {     // raw values arrived a plane ago: the gather goes straight to shared memory, asynchronously
        float hm, hko, hkg, hdko, hdkg;
        cp_async16(&s_m[buf ^ 1][g.hs], T1 + 2 * (size_t)g2_entry(P, hp1, hm));
        corey<NOG, NG>(P, hsg1, hko, hkg, hdko, hdkg, kr_den);
        s_q[buf ^ 1][g.hs] = make_float4(hp1, hkg, hko, 0.f);
        if (k + 2 < P.D) { hp1 = P1[(k + 2) * HW + g.hcol]; hsg1 = SG1[(k + 2) * HW + g.hcol]; }
      }
    }
    // ---- own-cell loads of plane k
    const float p1 = vC.p1, sg1 = sg1c;
    const float sg0 = A.sg0[base + c], so0 = A.so0[base + c], so1 = A.so1[base + c];
    float m0, m1;
    const uint32_t e0 = g2_entry(P, p0, m0), e1 = g2_entry(P, p1, m1);
    const float4 n0a = __ldg(T0 + 2 * (size_t)e0), n0b = __ldg(T0 + 2 * (size_t)e0 + 1);
    const float4 n1b = __ldg(T1 + 2 * (size_t)e1 + 1);
    float ckf[6];      // C*k_f of the six faces (W,E,S,N,D,U), see face_perms_tab; offsets advance by one plane per step
    ckf[0] = __ldg(FE + fe); ckf[1] = __ldg(FE + fe + 1);
    ckf[2] = __ldg(FN + fn); ckf[3] = __ldg(FN + fn + P.W);
    ckf[4] = __ldg(FU + fu); ckf[5] = __ldg(FU + fu + HW);
    fe += P.H * FL.WP; fn += (P.H + 1) * P.W; fu += HW;
    const float Mc[4] = {vC.m.x, vC.m.y, vC.m.z, vC.m.w};                          // gg, go, oo, og
    const float krg_c = vC.krg, kro_c = vC.kro;
    // neighbours W,E,S,N from the shared plane, D,U from registers
    float4 nm[6];
    float pn[6], nkg[6], nko[6];
    {
      const int off[4] = {-1, 1, -G2PX, G2PX};
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        nm[f] = s_m[buf][g.so + off[f]];
        const float4 q = s_q[buf][g.so + off[f]];
        pn[f] = q.x; nkg[f] = q.y; nko[f] = q.z;
      }
      nm[4] = vP.m; pn[4] = vP.p1; nkg[4] = vP.krg; nko[4] = vP.kro;
      nm[5] = vN.m; pn[5] = vN.p1; nkg[5] = vN.krg; nko[5] = vN.kro;
    }
    float a[4][6];
#pragma unroll
    for (int f = 0; f < 6; ++f) {
      // potentials as written (physics_loss.py:538-541): "plus" faces nbr - cell, "minus" faces cell - nbr
      const float pot = (f & 1) ? __fsub_rn(pn[f], p1) : __fsub_rn(p1, pn[f]);
      const bool own = pot <= 0.f;                                                // :543-551
      const float krg_f = own ? krg_c : nkg[f], kro_f = own ? kro_c : nko[f];
      const float Mn[4] = {nm[f].x, nm[f].y, nm[f].z, nm[f].w};
#pragma unroll
      for (int X = 0; X < 4; ++X) {
        const float Mf = __fmul_rn(__fadd_rn(Mc[X], Mn[X]), 0.5f);                // :517-525
        const float kr = (X == 0 || X == 3) ? krg_f : kro_f;                      // gg, og: gas phase; go, oo: oil phase
        a[X][f] = __fmul_rn(__fmul_rn(__fmul_rn(ckf[f], __fmul_rn(kr, Mf)), idl[f]), idl[f]);   // :563-583
      }
    }
    // wells in this cell (scatter_nd sums duplicates)
    float q4[4] = {0.f, 0.f, 0.f, 0.f}, mask = 0.f;
    int wfirst = 0, wlast = 0;
    if (LISTS && wslot) {
      s_wt.take(wslot - 1, k, wfirst, wlast);
      for (int e = wfirst; e < wlast; ++e) {
        const int w = s_wt.w[e];
#pragma unroll
        for (int X = 0; X < 4; ++X) q4[X] = __fadd_rn(q4[X], A.W7[X * wt + (int64_t)b * P.n_wells + w]);
        mask += 1.f;
      }
    } else if (col_wells) {
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
    const float A0 = n0a.x, B0 = n0a.y, Rs0 = n0a.z, Rv0 = n0a.w;
    const float dA0 = n0b.x, dB0 = n0b.y, dRs0 = n0b.z, dRv0 = n0b.w;
    const float a1 = n1b.x, b1 = n1b.y, r1 = n1b.z, v1 = n1b.w;
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
    auto trnX = [&](float mm0, float mm1) {
      const float m2 = __fadd_rn(__fmul_rn(__fsub_rn(mm1, mm0), rho1), mm0);
      const float num = __fsub_rn(__fadd_rn(__fmul_rn(d2, mm0), __fmul_rn(d1, m2)), __fmul_rn(d12, mm1));
      return __fmul_rn(P.dvDc, __fadd_rn(rte_d1, div_c(num, den_c)));
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
    if (g.active) {
      A.dom[base + c] = dom;
      if (A.dom_out) A.dom_out[base + c] = dom;
      if (mask != 0.f) {
        if (LISTS && wslot) { for (int e = wfirst; e < wlast; ++e) A.divqw[(int64_t)b * P.n_wells + s_wt.w[e]] = divq_tot; }
        else { for (int w = wfirst; w < P.n_wells && P.wells[w].cell == c; ++w) A.divqw[(int64_t)b * P.n_wells + w] = divq_tot; }
      }
      acc[0] += (double)dom * (double)dom;
      if (mask != 0.f) acc[1] += (double)ibc * (double)ibc;
      acc[2] += (double)trn * (double)trn;
      acc[3] += (double)__fadd_rn(mb_gg, mb_go);
      acc[4] += (double)__fadd_rn(mb_oo, mb_og);
      if (mask != 0.f) { acc[5] += (double)__fadd_rn(q4[0], q4[1]); acc[6] += (double)__fadd_rn(q4[2], q4[3]); }
    }
    // ---- publish plane k+1, rotate
    if (k + 1 < P.D) {
      s_m[buf ^ 1][g.so] = vN.m;
      s_q[buf ^ 1][g.so] = make_float4(vN.p1, vN.krg, vN.kro, 0.f);
    }
    if (k + 2 < P.D) { float mm; g2_prefetch(T1 + 2 * (size_t)g2_entry(P, p1n, mm)); }   // next iteration's stage gather
    cp_async_wait_all();
    __syncthreads();
    vP = vC; vC = vN; sg1c = sg1nn;
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

// ---- adjoint -------------------------------------------------------------------------------------
struct VisA { float p1, sn, Mg, Mo, krg, kro, dkrg, dkro; };    // dkrg, dkro: own cell only (not published)
template <int NOG = 0, int NG = 0>
__device__ __forceinline__ VisA g2_vis_adj(const SrmDev& P, const DivC* den, float p1, float sg1, float dom, float w2) {
  float m1;
  const uint32_t e = g2_entry(P, p1, m1);
  const float2 t = __ldg(P.gcv + e);                // {Mgg + Mog, Mgo + Moo}
  VisA v;
  corey<NOG, NG>(P, sg1, v.kro, v.krg, v.dkro, v.dkrg, den);
  v.p1 = p1;
  v.sn = w2 * dom;
  v.Mg = t.x;
  v.Mo = t.y;
  return v;
}

template <bool LISTS, int NOG, int NG>
__global__ void __launch_bounds__(kThreads, 2) k_adj_gc2(const __grid_constant__ SrmDev P, const __grid_constant__ GcAdj A) {
  __shared__ float4 s_v[2][G2PL];                  // {p1, 2 w_dom dom, krg, kro}
  __shared__ float2 s_k[2][G2PL];                  // {Mg, Mo}
  __shared__ double red[2 * 32];
  const int b = blockIdx.y;
  const int HW = P.H * P.W;
  const G2Geom g = g2_geom(P, A.tiles_x);
  const DivC kr_den[2] = {make_divc(P.kr_den_o), make_divc(P.kr_den_g)};      // Corey saturations: division by two per-handle constants
  const int r = srm_real_of(A.sample_real, b, A.B, A.R);
  double acc2[2] = {0.0, 0.0};
  const int64_t base = (int64_t)b * P.N;
  const float* __restrict__ P1 = A.p1 + base;
  const float* __restrict__ SG1 = A.sg1 + base;
  const float* __restrict__ DOM = A.dom + base;
  const FaceLay FL = face_layout(P.D, P.H, P.W);
  const float* __restrict__ FE = A.faces + (int64_t)r * FL.per_real;
  const float* __restrict__ FN = FE + FL.nE;
  const float* __restrict__ FU = FN + FL.nN;
  int fe = g.cj * FL.WP + g.ci, fn = g.col, fu = g.col;      // face offsets of plane 0 (per-realisation arrays: < 2^31 floats)
  // the tile's connections: column lists staged in shared memory; lists that do not fit keep the per-plane search
  __shared__ __align__(8) unsigned char s_wt_raw[LISTS ? sizeof(G2Wells) : 8];
  G2Wells& s_wt = *reinterpret_cast<G2Wells*>(s_wt_raw);
  bool col_wells = false;
  int wslot = 0;
  if constexpr (LISTS) {
    bool over = false;
    const bool any = s_wt.build(well_cols_of(P), P.W, P.D, g.tx0, g.ty0, nullptr, over);
    if (over) col_wells = g.active && column_has_well_gc(P, g.col, HW);
    else if (any && g.active) wslot = s_wt.slot_of[g.lt];
  } else {
    col_wells = g.active && column_has_well_gc(P, g.col, HW);
  }
  const float w_dom = A.dterms[SRM_TERM_DOM], w_mbc = A.dterms[SRM_TERM_MBC], w_trn = A.dterms[SRM_TERM_CMBC];
  const float w2 = 2.f * w_dom;
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
  const DivC den_c = make_divc(den);
  const float rte_d1 = __fdiv_rn(2.5e-8f, d1);
  const float d12 = __fadd_rn(d1, d2);
  const float iden2 = 1.0f / (den * den);
  const float dE1c = -2.f * 2.5e-8f / (d1 * d1);

  float sg1c = SG1[g.col];
  VisA vC = g2_vis_adj<NOG, NG>(P, kr_den, P1[g.col], sg1c, DOM[g.col], w2);
  s_v[0][g.so] = make_float4(vC.p1, vC.sn, vC.krg, vC.kro);
  s_k[0][g.so] = make_float2(vC.Mg, vC.Mo);
  if (g.halo) {
    const VisA h = g2_vis_adj<NOG, NG>(P, kr_den, P1[g.hcol], SG1[g.hcol], DOM[g.hcol], w2);
    s_v[0][g.hs] = make_float4(h.p1, h.sn, h.krg, h.kro);
    s_k[0][g.hs] = make_float2(h.Mg, h.Mo);
  }
  float p1n = 0.f, sg1n = 0.f, domn = 0.f, hp1 = 0.f, hsg1 = 0.f, hdom = 0.f;
  float p0c = A.p0[base + g.col], p0n = 0.f;
  if (P.D > 1) {
    p1n = P1[HW + g.col]; sg1n = SG1[HW + g.col]; domn = DOM[HW + g.col]; p0n = A.p0[base + HW + g.col];
    if (g.halo) { hp1 = P1[HW + g.hcol]; hsg1 = SG1[HW + g.hcol]; hdom = DOM[HW + g.hcol]; }
  }
  VisA vP = vC;
  __syncthreads();

  G2_LOOP_A
  for (int k = 0; k < P.D; ++k) {
    const int c = k * HW + g.col;
    const int buf = k & 1;
    VisA vN = vC;
    float sg1nn = sg1c;
    const float p0 = p0c;
    if (k + 1 < P.D) {
      vN = g2_vis_adj<NOG, NG>(P, kr_den, p1n, sg1n, domn, w2);
      sg1nn = sg1n;
      {   // own-cell packs of plane k+1: 48 bytes at p0 (two sectors at most), 32 at p1
        float mm;
        const float4* a0 = P.lut0 + 3 * (size_t)g2_entry(P, p0n, mm);
        const float4* a1 = P.lut1 + 2 * (size_t)g2_entry(P, p1n, mm);
        g2_prefetch(a0); g2_prefetch(a0 + 2);
        g2_prefetch(a1);
        p0c = p0n;
      }
      if (k + 2 < P.D) { p1n = P1[(k + 2) * HW + g.col]; sg1n = SG1[(k + 2) * HW + g.col]; domn = DOM[(k + 2) * HW + g.col]; p0n = A.p0[base + (k + 2) * HW + g.col]; }
      if (g.halo) {
        float hm, hko, hkg, hdko, hdkg;
        cp_async8(&s_k[buf ^ 1][g.hs], P.gcv + g2_entry(P, hp1, hm));
        corey<NOG, NG>(P, hsg1, hko, hkg, hdko, hdkg, kr_den);
        s_v[buf ^ 1][g.hs] = make_float4(hp1, w2 * hdom, hkg, hko);
        if (k + 2 < P.D) { const int hc = (k + 2) * HW + g.hcol; hp1 = P1[hc]; hsg1 = SG1[hc]; hdom = DOM[hc]; }
      }
    }
    // ---- own-cell loads of plane k
    const float p1 = vC.p1, sg1 = sg1c;
    const float sg0 = A.sg0[base + c], so0 = A.so0[base + c], so1 = A.so1[base + c];
    float m0, m1;
    const uint32_t e0 = g2_entry(P, p0, m0), e1 = g2_entry(P, p1, m1);
    const float4 n0a = __ldg(P.lut0 + 3 * (size_t)e0), n0b = __ldg(P.lut0 + 3 * (size_t)e0 + 1), n0c = __ldg(P.lut0 + 3 * (size_t)e0 + 2);
    const float4 n1b = __ldg(P.lut1 + 2 * (size_t)e1), n1c = __ldg(P.lut1 + 2 * (size_t)e1 + 1);   // {a1,b1,r1,v1}, {dMg,dMo,da1+dv1,dr1+db1}
    const float sc = vC.sn;                                  // dL/d dom_c
    float ckf[6];      // C*k_f of the six faces (W,E,S,N,D,U), see face_perms_tab; offsets advance by one plane per step
    ckf[0] = __ldg(FE + fe); ckf[1] = __ldg(FE + fe + 1);
    ckf[2] = __ldg(FN + fn); ckf[3] = __ldg(FN + fn + P.W);
    ckf[4] = __ldg(FU + fu); ckf[5] = __ldg(FU + fu + HW);
    fe += P.H * FL.WP; fn += (P.H + 1) * P.W; fu += HW;
    const float Mg_c = vC.Mg, Mo_c = vC.Mo, krg_c = vC.krg, kro_c = vC.kro;
    const float dMg_c = n1c.x * m1, dMo_c = n1c.y * m1, dkrg_c = vC.dkrg, dkro_c = vC.dkro;
    float4 nv[6];
    float2 nk[6];
    {
      const int off[4] = {-1, 1, -G2PX, G2PX};
#pragma unroll
      for (int f = 0; f < 4; ++f) { nv[f] = s_v[buf][g.so + off[f]]; nk[f] = s_k[buf][g.so + off[f]]; }
      nv[4] = make_float4(vP.p1, vP.sn, vP.krg, vP.kro); nk[4] = make_float2(vP.Mg, vP.Mo);
      nv[5] = make_float4(vN.p1, vN.sn, vN.krg, vN.kro); nk[5] = make_float2(vN.Mg, vN.Mo);
    }
    // image faces: both views identical, no net contribution
    const bool img[6] = {g.ci == 0, g.ci == P.W - 1, g.cj == 0, g.cj == P.H - 1, k == 0, k == P.D - 1};
    float g1 = 0.f, gs1 = 0.f;
    // ---- divergence part: gather over the cell's own residual and its six neighbours' residuals
#pragma unroll
    for (int f = 0; f < 6; ++f) {
      if (img[f]) continue;
      const float pn = nv[f].x;
      const float sn = nv[f].y;
      const float pot = (f & 1) ? (pn - p1) : (p1 - pn);
      const bool own = pot <= 0.f;
      const FaceAdj fa = face_adj(own, krg_c, kro_c, nv[f].z, nv[f].w, Mg_c, Mo_c, nk[f].x, nk[f].y, dMg_c, dMo_c, dkrg_c, dkro_c);
      const float Tf = ckf[f] * idl[f] * idl[f];
      const float dpf = p1 - pn;
      g1 += Tf * ((sc * fa.Lc - sn * fa.Ln) + dpf * (sc * fa.dLc_p - sn * fa.dLn_p));
      gs1 += Tf * dpf * (sc * fa.dLc_s - sn * fa.dLn_s);
    }
    g1 *= P.dv;
    gs1 *= P.dv;
    // ---- local terms
    const float A0 = n0a.x, B0 = n0a.y, Rs0 = n0a.z, Rv0 = n0a.w;
    const float dA0 = n0b.x, dB0 = n0b.y, dRs0 = n0b.z, dRv0 = n0b.w;
    const float d2A0 = n0c.x * m0, d2B0 = n0c.y * m0, d2Rs0 = n0c.z * m0, d2Rv0 = n0c.w * m0;
    const float a1 = n1b.x, b1 = n1b.y, r1 = n1b.z, v1 = n1b.w;
    const float dav1 = n1c.z * m1, drb1 = n1c.w * m1;      // d(a1 + v1)/dp1, d(r1 + b1)/dp1
    const float R0 = Rs0 * B0, V0 = Rv0 * A0;
    const float dR0 = Rs0 * dB0 + B0 * dRs0, dV0 = Rv0 * dA0 + A0 * dRv0;
    // d/dp0 of the n0 quantities (first derivatives masked by the clamp, second derivatives masked above)
    const float pA0 = dA0 * m0, pB0 = dB0 * m0, pR0 = dR0 * m0, pV0 = dV0 * m0;
    const float pdA0 = d2A0, pdB0 = d2B0;
    const float pdR0 = 2.f * dRs0 * dB0 * m0 + Rs0 * d2B0 + B0 * d2Rs0;
    const float pdV0 = 2.f * dRv0 * dA0 * m0 + Rv0 * d2A0 + A0 * d2Rv0;
    const float dpc = p1 - p0;
    const float nz = (dpc == 0.f) ? 0.f : 1.f;               // divide_no_nan: the chord-slope terms vanish with dpc
    const float dSgS = (sg1 - sg0) * nz, dSoS = (so1 - so0) * nz;
    const float Kg = P.phi * (dA0 + dV0) + P.phicf * (A0 + V0);
    const float Ko = P.phi * (dR0 + dB0) + P.phicf * (R0 + B0);
    const float pKg = P.phi * (pdA0 + pdV0) + P.phicf * (pA0 + pV0);
    const float pKo = P.phi * (pdR0 + pdB0) + P.phicf * (pR0 + pB0);
    const float sacc = sc * P.dv * idt;
    g1 += sacc * (P.phi * (dav1 * dSgS + drb1 * dSoS) + (sg0 * Kg + so0 * Ko));
    float g0 = sacc * (dpc * (sg0 * pKg + so0 * pKo) - (sg0 * Kg + so0 * Ko));
    gs1 += sacc * P.phi * (a1 + v1) * nz;
    float gs0 = sacc * (dpc * Kg - P.phi * (a1 + v1) * nz);
    float go1 = sacc * P.phi * (r1 + b1) * nz;
    float go0 = sacc * (dpc * Ko - P.phi * (r1 + b1) * nz);
    const float acc_tot = P.dv * idt * (P.phi * ((a1 + v1) * dSgS + (r1 + b1) * dSoS) + dpc * (sg0 * Kg + so0 * Ko));
    // material balance: mbc_b = -sum q - sum mcell
    const float mcell = mfac * ((sg1 * a1 - sg0 * A0) + (so1 * r1 - so0 * R0) + (so1 * b1 - so0 * B0) + (sg1 * v1 - sg0 * V0));
    g1 -= smf * (sg1 * dav1 + so1 * drb1);
    g0 += smf * (sg0 * (pA0 + pV0) + so0 * (pR0 + pB0));
    gs1 -= smf * (a1 + v1);
    gs0 += smf * (A0 + V0);
    go1 -= smf * (r1 + b1);
    go0 += smf * (R0 + B0);
    // wells in this cell: sum of the four rates enters dom (+) and mbc (-)
    if (LISTS && wslot) {
      const int64_t wt = (int64_t)A.B * P.n_wells;
      int first, last;
      s_wt.take(wslot - 1, k, first, last);
      for (int e = first; e < last; ++e) {
        const int w = s_wt.w[e];
        const float dqp = A.W7[4 * wt + (int64_t)b * P.n_wells + w], dqs = A.W7[5 * wt + (int64_t)b * P.n_wells + w];
        g1 += (sc - smb) * dqp;
        gs1 += (sc - smb) * dqs;
      }
    } else if (col_wells) {
      const int64_t wt = (int64_t)A.B * P.n_wells;
      const int first = well_lower_bound(P, c);
      for (int w = first; w < P.n_wells && P.wells[w].cell == c; ++w) {
        const float dqp = A.W7[4 * wt + (int64_t)b * P.n_wells + w], dqs = A.W7[5 * wt + (int64_t)b * P.n_wells + w];
        g1 += (sc - smb) * dqp;
        gs1 += (sc - smb) * dqs;
      }
    }
    // truncation term (see k_resid_adj_gc)
    if (g.active) {
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
      const float trn = __fadd_rn(__fmul_rn(P.dvDc, __fadd_rn(rte_d1, div_c(Ng, den_c))),
                                  __fmul_rn(P.dvDc, __fadd_rn(rte_d1, div_c(No, den_c))));
      const float st = 2.f * w_trn * trn;
      const float dE1 = dE1c - (Ng + No) * d2 * iden2;
      const float dE2 = -(Ng + No) * (d1 + 2.f * d2) * iden2;
      acc2[0] += (double)((smb * mcell - sc * acc_tot) * id1 + st * P.dvDc * dE1);
      acc2[1] += (double)(st * P.dvDc * dE2);
      A.gp0[base + c] = g0;
      A.gp1[base + c] = g1;
      A.gsg0[base + c] = gs0;
      A.gsg1[base + c] = gs1;
      A.gso0[base + c] = go0;
      A.gso1[base + c] = go1;
    }
    if (k + 1 < P.D) {
      s_v[buf ^ 1][g.so] = make_float4(vN.p1, vN.sn, vN.krg, vN.kro);
      s_k[buf ^ 1][g.so] = make_float2(vN.Mg, vN.Mo);
    }
    if (k + 2 < P.D) { float mm; g2_prefetch(P.gcv + g2_entry(P, p1n, mm)); }             // next iteration's stage gather
    cp_async_wait_all();
    __syncthreads();
    vP = vC; vC = vN; sg1c = sg1nn;
  }
  block_reduce<2>(acc2, red);
  if (threadIdx.x == 0) {
    atomicAdd(&A.gdt1_acc[b], acc2[0]);
    atomicAdd(&A.gdt2_acc[b], acc2[1]);
  }
}

// inner-boundary term without the staged fields: the well cell's and its neighbours' values straight from the table
struct IbcCell { float Mg, Mo, krg, kro, dMg, dMo, dkrg, dkro; };
__device__ __forceinline__ IbcCell g2_ibc_cell(const SrmDev& P, float p1, float sg1) {
  float m1;
  const uint32_t e = g2_entry(P, p1, m1);
  const float2 t = __ldg(P.gcv + e);
  const float4 d = __ldg(P.lut1 + 2 * (size_t)e + 1);
  IbcCell v;
  corey(P, sg1, v.kro, v.krg, v.dkro, v.dkrg);
  v.Mg = t.x;
  v.Mo = t.y;
  v.dMg = d.x * m1;
  v.dMo = d.y * m1;
  return v;
}
__global__ void __launch_bounds__(128) k_ibc_adj_gc2(const __grid_constant__ SrmDev P, const __grid_constant__ GcAdj A) {
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
  const int64_t base = (int64_t)b * P.N;
  const CellIdx ix = cell_index(P, c);
  float ckf[6];
  face_perms(P, A.kx + (int64_t)r * P.N, c, ix, ckf);
  const float idl[6] = {P.idx, P.idx, P.idy, P.idy, P.idz, P.idz};
  const float p1 = A.p1[base + c];
  const IbcCell cc = g2_ibc_cell(P, p1, A.sg1[base + c]);
  float self_p = 0.f, self_s = 0.f;
  for (int f = 0; f < 6; ++f) {
    const int cn = ix.n[f];
    if (cn == c) continue;
    const float pn = A.p1[base + cn];
    const IbcCell nn = g2_ibc_cell(P, pn, A.sg1[base + cn]);
    const float pot = (f & 1) ? (pn - p1) : (p1 - pn);
    const bool own = pot <= 0.f;
    const float hMg = 0.5f * (cc.Mg + nn.Mg), hMo = 0.5f * (cc.Mo + nn.Mo);
    const float kg = own ? cc.krg : nn.krg, ko = own ? cc.kro : nn.kro;
    const float L = kg * hMg + ko * hMo;
    const float Tf = ckf[f] * idl[f] * idl[f] * P.dv;
    const float dpf = p1 - pn;
    self_p += Tf * (L + dpf * 0.5f * (kg * cc.dMg + ko * cc.dMo));
    atomicAdd(&A.gp1[base + cn], s * Tf * (-L + dpf * 0.5f * (kg * nn.dMg + ko * nn.dMo)));
    if (own) self_s += Tf * dpf * (cc.dkrg * hMg + cc.dkro * hMo);
    else atomicAdd(&A.gsg1[base + cn], s * Tf * dpf * (nn.dkrg * hMg + nn.dkro * hMo));
  }
  atomicAdd(&A.gp1[base + c], s * (self_p + dqp));
  atomicAdd(&A.gsg1[base + c], s * (self_s + dqs));
}
