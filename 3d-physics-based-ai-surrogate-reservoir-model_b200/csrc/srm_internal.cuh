// Internal declarations shared by the kernels of libsrm_physics.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/srm_physics.h"

#define SRM_MAXK 64   // knots
#define SRM_MAXP 7    // PVT properties (GC)
#define SRM_EPS 1e-10f  // polyhm_splines.py:6

// One connection, sorted by flat cell index (so a block can binary-search its range).
struct WellDev {
  int32_t cell;   // (k*H + j)*W + i     welldata_processor.py:26-40 ([k,j,i] index rows)
  int32_t orig;   // position in the caller's well list (output tables use that order)
  float q_target, pwf_min, rw, hc, shut_start, shut_stop;
};

// Immutable parameters, passed to kernels by value (lands in the constant bank: the knot loop
// reads c[i], w[i] with uniform addresses).
struct SrmDev {
  int32_t D, H, W, N;  // N = D*H*W cells per sample
  float dx, dy, dz;
  float idx, idy, idz;  // fl(1/dx) ...
  float dv;             // (dx*dy)*dz          physics_loss.py:36
  float C, Dc, invDc;   // invDc = fl(1/Dc)    physics_loss.py:156
  float dvDc;           // fl(dv/Dc)           physics_loss.py:171
  float phi, cf, phicf; // phicf = fl(phi*cf)  physics_loss.py:149
  float Sgi, krg;
  float dvSgi_phi;      // (dv*Sgi)*phi        physics_loss.py:193
  float kx_ky, kv_kh;
  float p_min, p_max;
  int32_t tde_in_dom;
  int32_t use_blk, n_int;
  int32_t root_solver, n_root_iter;   // GC blocking-factor integral: SRM_ROOT_*, iterations per trapezoid node
  int32_t bhp_iterative, bhp_max_iters;   // _iterative_method instead of _non_iterative_method (well_rate_bhp_Subclassed.py:813-822)
  float bhp_tol;
  int32_t n_wells;
  const WellDev* wells;  // device, sorted by cell
  // the same connections grouped by (j, i) column, layers ascending inside a column (well_tile.cuh): what a z-marching
  // tile needs to find its connections without searching the cell-sorted table once per plane
  int32_t n_cols;
  const int32_t* col_rem;   // [n_cols]      j*W + i of the column, ascending
  const int32_t* col_ptr;   // [n_cols + 1]  range of the column in col_ent
  const int2* col_ent;      // [n_wells]     {layer k, position in the cell-sorted table}, sorted by (column, layer, position)
  const int32_t* layer_ptr; // [D + 1]       range of layer k in the cell-sorted table (the searches that remain stay inside one layer)
  // spline
  int32_t n_knots, order, n_props;   // polynomial fit: n_knots = number of coefficients, w[q][i] = coefficient i
  int32_t pvt_method;                // SRM_PVT_*
  float c[SRM_MAXK];
  float c2[SRM_MAXK];            // fl(c*c)
  float w[SRM_MAXP][SRM_MAXK];
  float v[SRM_MAXP][2];
  // SCAL (gas-condensate path): thresholds and denominators formed in fp32 the way relative_permeability.py:58-68 does
  float swmin, sorg, sgc, kro_somax, krg_sorg, krg_swmin, nog, ng;
  float kr_den_o, kr_den_g;      // (1-Swmin)-Sorg ; ((1-Sgc)-Swmin)-Sorg
  float kr_so_zero;              // Swmin + max(Sorg, Socr): krog = 0 at or below
  float kr_sg_full;              // 1 - (Swmin+Sorg): krgo = krg_Swmin above
  int32_t nog_i, ng_i;           // integer exponents (0: not integer-valued -> powf)
  int32_t fluid;
  // exact PVT tabulation (SrmConfig.pvt_lut); lut_n == 0: off
  // dry gas: entries interleaved, element e of each view at ptr + 2e (kernels_ref.cu, k_lut_build)
  const float4* lut0;    // {invBg, d/dp, d2/dp2, cp} at the fp32 value with bits lut_lo_bits + e
  const float4* lut1;    // {invBg, invBg*invug, d invBg/dp, d(invBg*invug)/dp}
  const float2* lutf0;   // {invBg, cp}             -- the forward's 8-byte views: half the L2 footprint
  const float2* lutf1;   // {invBg, invBg*invug}
  const float2* gcv;     // gas condensate, fused pair: {Mgg + Mog, Mgo + Moo} at level n+1 (the adjoint's neighbour-visible sums)
  uint32_t lut_lo_bits, lut_n;
  int32_t cp_safe;       // every tabulated |cp| lies in [2^-60, 2^60]: division by dt1 needs no per-cell range test
};

// realisation of sample b: the caller's map clamped into [0, R) (an index outside it must not become an out-of-bounds
// read of kx / the face coefficients), or realisation-major equal-sized groups without a map
__host__ __device__ __forceinline__ int srm_real_of(const int32_t* __restrict__ sample_real, int b, int B, int R) {
  if (!sample_real) return (int)(((int64_t)b * R) / B);
  const int r = sample_real[b];
  return r < 0 ? 0 : (r >= R ? R - 1 : r);
}

// Closed-form (piecewise-linear) tables for SRM_NUMERICS_CLOSED_FORM, device global memory
// (copied to shared memory by the tiled kernels).
//   interval k in [0, n]: k = number of knots <= x;  value_q(x) = f0[q][k] + slope[q][k]*(x - x0[k])
#define SRM_CF_MAXBUCKET 4096
struct SrmClosedForm {
  int32_t n;            // knots
  int32_t use_bucket;   // bucket[] valid (every bucket holds at most one knot)
  int32_t nb;           // buckets
  float inv_w;          // 1 / bucket width
  float2 lohi[SRM_MAXK + 1];            // [lo, hi) of interval k; lo(0) = -inf, hi(n) = +inf
  float4 ent[SRM_MAXK + 1];             // DG packing: {x0, f0[0], slope[0], f0[1]}
  float sM[SRM_MAXK + 1];               //             slope[1]
  float x0[SRM_MAXK + 1];
  float f0[SRM_MAXP][SRM_MAXK + 1];
  float slope[SRM_MAXP][SRM_MAXK + 1];
  double f0d[SRM_MAXK + 1];             // f0[0] (invBg) in fp64: exact anchor differences for the material balance
  unsigned char bucket[SRM_CF_MAXBUCKET];   // bucket b -> number of knots <= b*w
};

struct SrmHandle {
  SrmConfig cfg;       // host copy (pointers nulled)
  SrmDev dev;          // device parameter block
  WellDev* d_wells;    // device
  int32_t* d_wcols;    // device: col_rem | col_ptr | col_ent of the parameter block (one allocation)
  int64_t l2_persist_bytes;   // persisting L2 set-aside requested for the table gathers' evict_last policy (0: none)
  SrmClosedForm* d_cf; // device (closed-form tables), may be null
  void* d_cf2;         // device (tables of the lean closed-form pair, kernels_cf2.cu), may be null
  int cf_faces_ok, cf_grouped;   // closed form: which per-call scratch the last forward left in the workspace
  float4* d_lut;       // device (exact PVT tabulation), may be null
  int lut_full;        // the table covers the whole clamp range [p_min, p_max]
  int gc_fused;        // gas condensate: the fused pair (gc_fused.cuh) runs, no staged fields in the workspace
  int device;
  int sm_count;
  // fingerprint of the forward state held in a workspace (SRM_FLAG_SAVE_FOR_BACKWARD): every input pointer of the
  // call that produced it (dry gas leaves the saturation slots null), so a backward on other inputs recomputes
  const void* st_ptr[13];    // ws, kx, sample_real, p0, p1, sg0, sg1, so0, so1, dt1, dt2, t1, spare
  int32_t st_B, st_R; int32_t st_valid;
  int32_t st_family;   // fused reference path: kernel family of the forward that saved the state (1 lean kernels_dg4.cu, 0 generic)
  int32_t no_dg4;      // test knob SRM_NO_DG4, read ONCE at srm_create: the handle runs the generic fused kernels only
  int32_t adj_packs;   // lean family: the forward stages the adjoint's six table values per cell in the workspace (24 B per
                       // cell-timestep through otherwise idle HBM bandwidth), so the adjoint runs without table gathers
};
static inline void srm_state_set(SrmHandle* h, int32_t B, int32_t R, const void* ws, const void* kx, const void* sr, const void* p0,
                                 const void* p1, const void* sg0, const void* sg1, const void* so0, const void* so1,
                                 const void* dt1, const void* dt2, const void* t1) {
  const void* v[13] = {ws, kx, sr, p0, p1, sg0, sg1, so0, so1, dt1, dt2, t1, nullptr};
  for (int i = 0; i < 13; ++i) h->st_ptr[i] = v[i];
  h->st_B = B; h->st_R = R; h->st_valid = 1;
}
static inline bool srm_state_is(const SrmHandle* h, int32_t B, int32_t R, const void* ws, const void* kx, const void* sr, const void* p0,
                                const void* p1, const void* sg0, const void* sg1, const void* so0, const void* so1,
                                const void* dt1, const void* dt2, const void* t1) {
  const void* v[13] = {ws, kx, sr, p0, p1, sg0, sg1, so0, so1, dt1, dt2, t1, nullptr};
  if (!h->st_valid || h->st_B != B || h->st_R != R) return false;
  for (int i = 0; i < 13; ++i) if (h->st_ptr[i] != v[i]) return false;
  return true;
}

// ---- workspace carving ------------------------------------------------------------------
#define SRM_MAXR 65536
struct SrmWs {
  double* sse;        // [8]
  double* mb_sum;     // [B]  sum over cells of mb_cells
  double* q_sum;      // [B]  sum of well rates
  double* gdt1_acc;   // [B]
  double* gdt2_acc;   // [B]
  float* mbc;         // [B]
  float* qw;          // [B*nw] sorted-well order
  float* pwfw;        // [B*nw]
  float* dqdp;        // [B*nw]
  float* divqw;       // [B*nw]
  // closed-form scheduling: samples grouped by realisation, cut into segments
  int32_t* grp_cnt;   // [SRM_MAXR+1]  counts -> exclusive starts
  int32_t* grp_fill;  // [SRM_MAXR]
  int32_t* grp_list;  // [B]   sample ids, grouped by realisation, ascending within a group
  int32_t* seg;       // [3*(B+1)] (r, offset into grp_list, count)
  int32_t* ctl;       // [8]   ctl[0] = number of segments, ctl[1..2] = work counters
  float* faces;       // fused reference path and gas condensate: static face coefficients [R][face floats]
  float* gc;          // gas-condensate path: SRM_GC_NFIELDS staged fields [B*N] each, 4+1+2 extra well tables
  float* gc_wells;    // [7][B*nw]: qgg,qgo,qoo,qog (sorted order), d(sum q)/dp, d(sum q)/dSg, spare
  float* dom;         // field [B*N]
  float* A0;          // reference-order fields [B*N]
  float* A0p;
  float* A1;
  float* G1;
  float* A0pp;
  float* G1p;
  float* A1p;
  float* pk;          // fused reference path, lean family: adjoint packs [6][B*N] = cp, A0', A0'', G, A1', G'
  size_t bytes;
};

static inline size_t srm_align(size_t x) { return (x + 255) & ~size_t(255); }

// workspace flavours: staged reference order (7 PVT fields), fused reference order (tabulated PVT), closed form
enum { SRM_WS_REF_STAGED = 0, SRM_WS_REF_FUSED = 1, SRM_WS_CF = 2, SRM_WS_GC = 3, SRM_WS_GC_FUSED = 4 };
#define SRM_GC_NFIELDS 32
size_t srm_ref2_face_floats(const SrmDev& P);
static inline int srm_ws_mode(const SrmHandle* h) {
  if (h->cfg.fluid_type == SRM_FLUID_GC) return h->gc_fused ? SRM_WS_GC_FUSED : SRM_WS_GC;
  if (h->cfg.numerics == SRM_NUMERICS_CLOSED_FORM) return SRM_WS_CF;
  return h->dev.lut_n > 0 ? SRM_WS_REF_FUSED : SRM_WS_REF_STAGED;
}

static inline SrmWs srm_carve(void* base, int64_t B, int64_t R, int64_t N, int64_t nw, int mode, size_t face_floats, bool packs = false) {
  const bool closed_form = mode == SRM_WS_CF;
  SrmWs w;
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += srm_align(bytes); return r; };
  w.sse = (double*)take(8 * sizeof(double));
  w.mb_sum = (double*)take(B * sizeof(double));
  w.q_sum = (double*)take(B * sizeof(double));
  w.gdt1_acc = (double*)take(B * sizeof(double));
  w.gdt2_acc = (double*)take(B * sizeof(double));
  w.mbc = (float*)take(B * sizeof(float));
  size_t wt = (size_t)(B * (nw > 0 ? nw : 1)) * sizeof(float);
  w.qw = (float*)take(wt);
  w.pwfw = (float*)take(wt);
  w.dqdp = (float*)take(wt);
  w.divqw = (float*)take(wt);
  w.grp_cnt = w.grp_fill = w.grp_list = w.seg = w.ctl = nullptr;
  if (closed_form) {
    w.grp_cnt = (int32_t*)take((SRM_MAXR + 1) * sizeof(int32_t));
    w.grp_fill = (int32_t*)take(SRM_MAXR * sizeof(int32_t));
    w.grp_list = (int32_t*)take(B * sizeof(int32_t));
    w.seg = (int32_t*)take(3 * (B + 1) * sizeof(int32_t));
    w.ctl = (int32_t*)take(8 * sizeof(int32_t));
  }
  size_t fb = (size_t)(B * N) * sizeof(float);
  w.faces = nullptr;
  if (mode == SRM_WS_REF_FUSED || mode == SRM_WS_GC || mode == SRM_WS_GC_FUSED) w.faces = (float*)take((size_t)R * face_floats * sizeof(float));
  if (mode == SRM_WS_CF) w.faces = (float*)take((size_t)R * 3 * (size_t)N * sizeof(float));      // kernels_cf2.cu: [3][R][N]
  w.dom = (float*)take(fb);
  w.A0 = w.A0p = w.A1 = w.G1 = w.A0pp = w.G1p = w.A1p = nullptr;
  w.gc = w.gc_wells = nullptr;
  if (mode == SRM_WS_GC || mode == SRM_WS_GC_FUSED) w.gc_wells = (float*)take(7 * wt);
  if (mode == SRM_WS_GC) w.gc = (float*)take((size_t)SRM_GC_NFIELDS * fb);
  if (mode == SRM_WS_REF_STAGED) {
    w.A0 = (float*)take(fb);
    w.A0p = (float*)take(fb);
    w.A1 = (float*)take(fb);
    w.G1 = (float*)take(fb);
    w.A0pp = (float*)take(fb);
    w.G1p = (float*)take(fb);
    w.A1p = (float*)take(fb);
  }
  w.pk = nullptr;
  if (mode == SRM_WS_REF_FUSED && packs) w.pk = (float*)take(6 * fb);
  w.bytes = off;
  return w;
}

// ---- error plumbing ---------------------------------------------------------------------
void srm_set_error(const char* fmt, ...);
#define SRM_CUDA_CHECK(expr)                                                          \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      srm_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return SRM_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)

// ---- launchers implemented in the .cu files ------------------------------------------------
int srm_launch_pvt_eval_ref(const SrmHandle* h, int64_t n, const float* p, float* val, float* dval, cudaStream_t s);
int srm_launch_wells_ref(const SrmHandle* h, int32_t B, const float* kx, const int32_t* sample_real, int32_t R,
                         const float* p, const float* t_days, float* qw_sorted, float* pwfw_sorted,
                         float* dqdp_sorted, cudaStream_t s);
int srm_build_pvt_lut(SrmHandle* h, float lo, float hi);
int srm_build_pvt_lut_gc(SrmHandle* h, float lo, float hi);
bool srm_dg4_applicable(const SrmHandle* h);       // kernels_dg4.cu
int srm_ref2_backward_family(const SrmHandle* h, const float* p0, const float* p1, const void* dom_ws, const float* gp0, const float* gp1);
int srm_forward_ref2(SrmHandle* h, int32_t B, int32_t R, const float* kx, const int32_t* sample_real,
                     const float* p0, const float* p1, const float* dt1, const float* dt2, const float* t1,
                     float* terms_out, float* dom_out, const SrmWs& ws, cudaStream_t s, int force_family = -1, bool save = true);
int srm_backward_ref2(SrmHandle* h, int32_t B, int32_t R, const float* kx, const int32_t* sample_real,
                      const float* p0, const float* p1, const float* dt1, const float* dt2, const float* t1,
                      const float* dterms, float* gp0, float* gp1, float* gdt1, float* gdt2,
                      const SrmWs& ws, cudaStream_t s);
int srm_launch_relperm(const SrmHandle* h, int64_t n, const float* sg, float* krog, float* krgo, float* dkrog, float* dkrgo, cudaStream_t s);
int srm_forward_gc_impl(SrmHandle* h, int32_t B, int32_t R, const float* kx, const int32_t* sample_real, const float* p0,
                        const float* p1, const float* sg0, const float* sg1, const float* so0, const float* so1,
                        const float* dt1, const float* dt2, const float* t1, float* terms_out, float* dom_out,
                        const SrmWs& ws, bool save, cudaStream_t s);
int srm_backward_gc_impl(SrmHandle* h, int32_t B, int32_t R, const float* kx, const int32_t* sample_real, const float* p0,
                         const float* p1, const float* sg0, const float* sg1, const float* so0, const float* so1,
                         const float* dt1, const float* dt2, const float* t1, const float* dterms, float* gp0, float* gp1,
                         float* gsg0, float* gsg1, float* gso0, float* gso1, float* gdt1, float* gdt2, const SrmWs& ws,
                         cudaStream_t s);
int srm_forward_ref(SrmHandle* h, int32_t B, int32_t R, const float* kx, const int32_t* sample_real,
                    const float* p0, const float* p1, const float* dt1, const float* dt2, const float* t1,
                    float* terms_out, float* dom_out, const SrmWs& ws, bool save, cudaStream_t s);
int srm_backward_ref(SrmHandle* h, int32_t B, int32_t R, const float* kx, const int32_t* sample_real,
                     const float* p0, const float* p1, const float* dt1, const float* dt2, const float* t1,
                     const float* dterms, float* gp0, float* gp1, float* gdt1, float* gdt2,
                     const SrmWs& ws, cudaStream_t s);
