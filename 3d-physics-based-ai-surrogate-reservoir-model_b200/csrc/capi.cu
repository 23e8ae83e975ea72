// C ABI of libsrm_physics.so (see include/srm_physics.h).  Host-side validation, handle
// construction and dispatch to the kernel launchers.  No CPU compute path exists here.
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "srm_internal.cuh"
#ifndef SRM_L2_PERSIST_DEFAULT_MB
#define SRM_L2_PERSIST_DEFAULT_MB 0
#endif

static thread_local char g_err[512] = "";

void srm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int srm_launch_denorm_log(int64_t n, const float* x, float kmin, float kmax, float lo, float hi, float* out, cudaStream_t s);
int srm_launch_scatter_wells(const SrmHandle* h, int32_t B, const float* sorted, float* dense, cudaStream_t s);
int srm_launch_unsort_wells(const SrmHandle* h, int32_t B, const float* sorted, float* out, cudaStream_t s);
int srm_launch_selftest_rounding(int64_t n, uint64_t seed, int64_t* bad_host, cudaStream_t s);
int srm_build_closed_form(SrmHandle* h, const SrmConfig* cfg);
int srm_build_cf2(SrmHandle* h, const SrmConfig* cfg);
int srm_launch_pvt_eval_cf(const SrmHandle* h, int64_t n, const float* p, float* val, float* dval, cudaStream_t s);
int srm_launch_wells_cf(const SrmHandle* h, int32_t B, const float* kx, const int32_t* sample_real, int32_t R,
                        const float* p, const float* t_days, float* qw, float* pwfw, float* dqdp, cudaStream_t s);
int srm_forward_cf(SrmHandle* h, int32_t B, int32_t R, const float* kx, const int32_t* sample_real,
                   const float* p0, const float* p1, const float* dt1, const float* dt2, const float* t1,
                   float* terms_out, float* dom_out, const SrmWs& ws, bool save, cudaStream_t s);
int srm_backward_cf(SrmHandle* h, int32_t B, int32_t R, const float* kx, const int32_t* sample_real,
                    const float* p0, const float* p1, const float* dt1, const float* dt2, const float* t1,
                    const float* dterms, float* gp0, float* gp1, float* gdt1, float* gdt2,
                    const SrmWs& ws, cudaStream_t s);

extern "C" {

int srm_version(void) { return SRM_ABI_VERSION; }
const char* srm_last_error(void) { return g_err; }

int srm_create(const SrmConfig* cfg, SrmHandle** out) {
  if (!cfg || !out) { srm_set_error("srm_create: null argument"); return SRM_ERR_INVALID; }
  *out = nullptr;
  if (cfg->abi_version != SRM_ABI_VERSION) {
    srm_set_error("srm_create: abi_version %d != %d", cfg->abi_version, SRM_ABI_VERSION);
    return SRM_ERR_INVALID;
  }
  if (cfg->D < 1 || cfg->H < 1 || cfg->W < 1 || (int64_t)cfg->D * cfg->H * cfg->W > (int64_t)INT32_MAX / 2) {
    srm_set_error("srm_create: bad grid %d x %d x %d", cfg->D, cfg->H, cfg->W);
    return SRM_ERR_INVALID;
  }
  if (cfg->fluid_type != SRM_FLUID_DG && cfg->fluid_type != SRM_FLUID_GC) { srm_set_error("srm_create: unknown fluid_type %d", cfg->fluid_type); return SRM_ERR_INVALID; }
  if (cfg->fluid_type == SRM_FLUID_GC) {
    if (cfg->n_props != 7) { srm_set_error("srm_create: SRM_FLUID_GC needs the 7 GC properties (InvBg, InvBo, Invug, Invuo, Rs, Rv, Vro), got %d", cfg->n_props); return SRM_ERR_INVALID; }
    if (cfg->numerics != SRM_NUMERICS_REFERENCE) { srm_set_error("srm_create: SRM_FLUID_GC is built for SRM_NUMERICS_REFERENCE only"); return SRM_ERR_INVALID; }
    if (cfg->use_blocking_factor && (cfg->n_root_iter < 1 || cfg->n_root_iter > 200 ||
                                     (cfg->root_solver != SRM_ROOT_NEWTON && cfg->root_solver != SRM_ROOT_BRACKET))) {
      srm_set_error("srm_create: the GC blocking-factor integral needs root_solver SRM_ROOT_NEWTON|SRM_ROOT_BRACKET and 1 <= n_root_iter <= 200");
      return SRM_ERR_INVALID;
    }
  }
  if (cfg->bhp_iterative && (cfg->bhp_max_iters < 0 || cfg->bhp_max_iters > 1000 || !(cfg->bhp_tol >= 0.f))) {
    srm_set_error("srm_create: iterative BHP control needs 0 <= bhp_max_iters <= 1000 and bhp_tol >= 0");
    return SRM_ERR_INVALID;
  }
  const bool poly = cfg->pvt_method == SRM_PVT_POLYNOMIAL;
  if (cfg->pvt_method != SRM_PVT_SPLINE && !poly) { srm_set_error("srm_create: unknown pvt_method %d", cfg->pvt_method); return SRM_ERR_INVALID; }
  if (cfg->n_knots < (poly ? 1 : 2) || cfg->n_knots > SRM_MAXK || cfg->n_props < 2 || cfg->n_props > SRM_MAXP || !cfg->spline_w ||
      (!poly && (!cfg->knots || !cfg->spline_v))) {
    srm_set_error("srm_create: bad PVT table (n_knots=%d, n_props=%d)", cfg->n_knots, cfg->n_props);
    return SRM_ERR_INVALID;
  }
  if (!poly) {
    if (cfg->spline_order != 1 && cfg->spline_order != 2) { srm_set_error("srm_create: spline_order must be 1 or 2"); return SRM_ERR_INVALID; }
    for (int i = 1; i < cfg->n_knots; ++i)
      if (!(cfg->knots[i] > cfg->knots[i - 1])) { srm_set_error("srm_create: knots must be strictly ascending"); return SRM_ERR_INVALID; }
  }
  if (poly && cfg->numerics != SRM_NUMERICS_REFERENCE) { srm_set_error("srm_create: the polynomial PVT fit is built for SRM_NUMERICS_REFERENCE"); return SRM_ERR_INVALID; }
  if (cfg->n_wells < 0 || (cfg->n_wells > 0 && !cfg->wells)) { srm_set_error("srm_create: bad wells"); return SRM_ERR_INVALID; }
  if (cfg->numerics != SRM_NUMERICS_REFERENCE && cfg->numerics != SRM_NUMERICS_CLOSED_FORM) {
    srm_set_error("srm_create: unknown numerics %d", cfg->numerics);
    return SRM_ERR_INVALID;
  }
  if (cfg->numerics == SRM_NUMERICS_CLOSED_FORM && cfg->spline_order != 1) {
    srm_set_error("srm_create: SRM_NUMERICS_CLOSED_FORM needs spline_order 1 (order 2 is not piecewise polynomial)");
    return SRM_ERR_INVALID;
  }
  if (cfg->use_blocking_factor && (cfg->n_intervals < 1 || cfg->n_intervals > 64)) {
    srm_set_error("srm_create: n_intervals out of range");
    return SRM_ERR_INVALID;
  }
  for (int w = 0; w < cfg->n_wells; ++w) {
    const SrmWell& x = cfg->wells[w];
    if (x.i < 0 || x.i >= cfg->W || x.j < 0 || x.j >= cfg->H || x.k < 0 || x.k >= cfg->D) {
      srm_set_error("srm_create: well %d at (i=%d,j=%d,k=%d) outside the grid", w, x.i, x.j, x.k);
      return SRM_ERR_INVALID;
    }
  }
  int ndev = 0;
  SRM_CUDA_CHECK(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) { srm_set_error("srm_create: device %d of %d", cfg->device, ndev); return SRM_ERR_CUDA; }
  SRM_CUDA_CHECK(cudaSetDevice(cfg->device));

  SrmHandle* h = new (std::nothrow) SrmHandle();
  if (!h) { srm_set_error("srm_create: out of host memory"); return SRM_ERR_INVALID; }
  std::memset(h, 0, sizeof(*h));
  h->cfg = *cfg;
  h->cfg.knots = h->cfg.spline_w = h->cfg.spline_v = nullptr;
  h->cfg.wells = nullptr;
  h->device = cfg->device;
  h->st_family = -1;
  h->adj_packs = 0;
  h->no_dg4 = getenv("SRM_NO_DG4") != nullptr ? 1 : 0;      // test knob, read once here (tests compare the kernel families)
  SRM_CUDA_CHECK(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, cfg->device));

  SrmDev& P = h->dev;
  P.D = cfg->D; P.H = cfg->H; P.W = cfg->W; P.N = cfg->D * cfg->H * cfg->W;
  P.dx = cfg->dx; P.dy = cfg->dy; P.dz = cfg->dz;
  P.idx = 1.0f / cfg->dx; P.idy = 1.0f / cfg->dy; P.idz = 1.0f / cfg->dz;
  P.dv = (cfg->dx * cfg->dy) * cfg->dz;
  P.C = cfg->C; P.Dc = cfg->Dc; P.invDc = 1.0f / cfg->Dc;
  P.dvDc = P.dv / cfg->Dc;
  P.phi = cfg->phi; P.cf = cfg->cf; P.phicf = cfg->phi * cfg->cf;
  P.Sgi = cfg->Sgi; P.krg = cfg->krg;
  P.dvSgi_phi = (P.dv * cfg->Sgi) * cfg->phi;
  P.kx_ky = cfg->kx_ky; P.kv_kh = cfg->kv_kh;
  P.p_min = cfg->p_min; P.p_max = cfg->p_max;
  P.tde_in_dom = cfg->tde_in_dom;
  P.use_blk = cfg->use_blocking_factor; P.n_int = cfg->n_intervals;
  P.root_solver = cfg->root_solver; P.n_root_iter = cfg->n_root_iter;
  P.bhp_iterative = cfg->bhp_iterative ? 1 : 0; P.bhp_max_iters = cfg->bhp_max_iters; P.bhp_tol = cfg->bhp_tol;
  P.n_knots = cfg->n_knots; P.order = cfg->spline_order; P.n_props = cfg->n_props;
  // SCAL: constants formed in fp32 like relative_permeability.py:58-68
  P.fluid = cfg->fluid_type;
  P.swmin = cfg->Swmin; P.sorg = cfg->Sorg; P.sgc = cfg->Sgc;
  P.kro_somax = cfg->kro_Somax; P.krg_sorg = cfg->krg_Sorg; P.krg_swmin = cfg->krg_Swmin;
  P.nog = cfg->nog; P.ng = cfg->ng;
  P.kr_den_o = (1.0f - cfg->Swmin) - cfg->Sorg;
  P.kr_den_g = ((1.0f - cfg->Sgc) - cfg->Swmin) - cfg->Sorg;
  P.kr_so_zero = cfg->Swmin + std::max(cfg->Sorg, cfg->Socr);
  P.kr_sg_full = 1.0f - (cfg->Swmin + cfg->Sorg);
  auto as_int = [](float e) { return (e >= 1.f && e <= 16.f && e == std::floor(e)) ? (int)e : 0; };
  P.nog_i = as_int(cfg->nog); P.ng_i = as_int(cfg->ng);
  P.pvt_method = cfg->pvt_method;
  for (int i = 0; i < cfg->n_knots; ++i) {
    if (!poly) { P.c[i] = cfg->knots[i]; P.c2[i] = cfg->knots[i] * cfg->knots[i]; }
    for (int q = 0; q < cfg->n_props; ++q) P.w[q][i] = cfg->spline_w[q * cfg->n_knots + i];
  }
  if (!poly)
    for (int q = 0; q < cfg->n_props; ++q) { P.v[q][0] = cfg->spline_v[2 * q]; P.v[q][1] = cfg->spline_v[2 * q + 1]; }

  // wells: [k,j,i] -> flat cell, sorted (stable) by cell; integer work, bit-exact
  P.n_wells = cfg->n_wells;
  if (cfg->n_wells > 0) {
    std::vector<WellDev> wd(cfg->n_wells);
    for (int w = 0; w < cfg->n_wells; ++w) {
      const SrmWell& x = cfg->wells[w];
      wd[w].cell = (x.k * cfg->H + x.j) * cfg->W + x.i;
      wd[w].orig = w;
      wd[w].q_target = x.q_target; wd[w].pwf_min = x.pwf_min; wd[w].rw = x.rw; wd[w].hc = x.hc;
      wd[w].shut_start = x.shut_start; wd[w].shut_stop = x.shut_stop;
    }
    std::stable_sort(wd.begin(), wd.end(), [](const WellDev& a, const WellDev& b) { return a.cell < b.cell; });
    cudaError_t e = cudaMalloc((void**)&h->d_wells, sizeof(WellDev) * cfg->n_wells);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_wells, wd.data(), sizeof(WellDev) * cfg->n_wells, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { srm_set_error("srm_create: wells upload: %s", cudaGetErrorString(e)); srm_destroy(h); return SRM_ERR_CUDA; }
    P.wells = h->d_wells;
    // column lists (well_tile.cuh): connections grouped by (j, i), layers ascending; integer work
    {
      const int HW = cfg->H * cfg->W, nw = cfg->n_wells;
      std::vector<int> order(nw);
      for (int w = 0; w < nw; ++w) order[w] = w;
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return wd[a].cell % HW < wd[b].cell % HW; });   // wd is cell-sorted: layers stay ascending
      std::vector<int32_t> rem, ptr;
      std::vector<int32_t> ent(2 * (size_t)nw);
      for (int t = 0; t < nw; ++t) {
        const int w = order[t], r = wd[w].cell % HW;
        if (rem.empty() || rem.back() != r) { rem.push_back(r); ptr.push_back(t); }
        ent[2 * t] = wd[w].cell / HW; ent[2 * t + 1] = w;
      }
      ptr.push_back(nw);
      const size_t nc = rem.size();
      std::vector<int32_t> blob;
      blob.insert(blob.end(), rem.begin(), rem.end());
      blob.insert(blob.end(), ptr.begin(), ptr.end());
      if (blob.size() & 1) blob.push_back(0);            // col_ent is read as int2
      const size_t ent_at = blob.size();
      blob.insert(blob.end(), ent.begin(), ent.end());
      const size_t lay_at = blob.size();
      {
        int w = 0;
        for (int k = 0; k <= cfg->D; ++k) {                // first connection with cell >= k*HW
          while (w < nw && wd[w].cell < (int64_t)k * HW) ++w;
          blob.push_back(w);
        }
      }
      e = cudaMalloc((void**)&h->d_wcols, sizeof(int32_t) * blob.size());
      if (e == cudaSuccess) e = cudaMemcpy(h->d_wcols, blob.data(), sizeof(int32_t) * blob.size(), cudaMemcpyHostToDevice);
      if (e != cudaSuccess) { srm_set_error("srm_create: well columns upload: %s", cudaGetErrorString(e)); srm_destroy(h); return SRM_ERR_CUDA; }
      P.n_cols = (int32_t)nc;
      P.col_rem = h->d_wcols; P.col_ptr = h->d_wcols + nc; P.col_ent = reinterpret_cast<const int2*>(h->d_wcols + ent_at);
      P.layer_ptr = h->d_wcols + lay_at;
    }
  }
  if (!poly && cfg->spline_order == 1) {
    int rc = srm_build_closed_form(h, cfg);
    if (!rc && cfg->fluid_type == SRM_FLUID_DG) rc = srm_build_cf2(h, cfg);
    if (rc) { srm_destroy(h); return rc; }
  }
  if (cfg->pvt_lut && cfg->numerics == SRM_NUMERICS_REFERENCE) {
    const bool whole = !(cfg->lut_p_lo < cfg->lut_p_hi);
    const float lo = whole ? cfg->p_min : std::max(cfg->lut_p_lo, cfg->p_min);
    const float hi = whole ? cfg->p_max : std::min(cfg->lut_p_hi, cfg->p_max);
    h->lut_full = (lo <= cfg->p_min && hi >= cfg->p_max) ? 1 : 0;
    int rc = cfg->fluid_type == SRM_FLUID_GC ? srm_build_pvt_lut_gc(h, lo, hi) : srm_build_pvt_lut(h, lo, hi);
    if (rc) { srm_destroy(h); return rc; }
    // The table gathers carry an L2::evict_last policy; that priority only has a region of the L2 to live in when the
    // device's persisting set-aside is non-zero (cudaLimitPersistingL2CacheSize, 0 by default).  SRM_L2_PERSIST_MB
    // (read once, here) overrides the size; 0 leaves the device limit alone.
    {
      int max_persist = 0;
      cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, h->device);
      long mb = SRM_L2_PERSIST_DEFAULT_MB;
      if (const char* e = getenv("SRM_L2_PERSIST_MB")) mb = atol(e);
      if (mb > 0 && max_persist > 0) {
        size_t want = (size_t)mb << 20;
        if (want > (size_t)max_persist) want = (size_t)max_persist;
        cudaError_t ce = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
        if (ce != cudaSuccess) (void)cudaGetLastError();      // not fatal: the hints fall back to ordinary replacement
        h->l2_persist_bytes = (ce == cudaSuccess) ? (int64_t)want : 0;
      }
    }
  }
  // Optional (SRM_ADJ_PACKS=1, read once here): the lean dry-gas forward stages the adjoint's six table values per cell
  // through the workspace and the adjoint runs without table gathers.  Measured on B200 (cfg5, K = 8): adjoint
  // 18.4 -> 14.7 ms, forward 15.6 -> 20.0 ms -- the 24 B per cell of extra stores cost the forward more than the
  // adjoint gains, so it is off by default.
  h->adj_packs = (cfg->fluid_type == SRM_FLUID_DG && srm_ws_mode(h) == SRM_WS_REF_FUSED && srm_dg4_applicable(h) && !h->no_dg4 &&
                  getenv("SRM_ADJ_PACKS") != nullptr) ? 1 : 0;
  *out = h;
  return SRM_OK;
}

void srm_destroy(SrmHandle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->d_wells) cudaFree(h->d_wells);
  if (h->d_wcols) cudaFree(h->d_wcols);
  if (h->d_cf) cudaFree(h->d_cf);
  if (h->d_cf2) cudaFree(h->d_cf2);
  if (h->d_lut) cudaFree(h->d_lut);
  delete h;
}

static SrmWs carve(const SrmHandle* h, void* base, int32_t B, int32_t R) {
  return srm_carve(base, B, R, h->dev.N, h->dev.n_wells, srm_ws_mode(h), srm_ref2_face_floats(h->dev), h->adj_packs != 0);
}

size_t srm_workspace_bytes(const SrmHandle* h, int32_t B, int32_t R, int32_t flags) {
  if (!h || B < 0 || R < 0) return 0;
  (void)flags;
  return carve(h, nullptr, B, R).bytes;
}

int srm_pvt_eval(const SrmHandle* h, int64_t n, const float* p, float* val, float* dval, void* stream) {
  if (!h || n < 0 || (n > 0 && !p)) { srm_set_error("srm_pvt_eval: bad argument"); return SRM_ERR_INVALID; }
  SRM_CUDA_CHECK(cudaSetDevice(h->device));
  if (h->cfg.numerics == SRM_NUMERICS_CLOSED_FORM) return srm_launch_pvt_eval_cf(h, n, p, val, dval, (cudaStream_t)stream);
  return srm_launch_pvt_eval_ref(h, n, p, val, dval, (cudaStream_t)stream);
}

int srm_denormalize_log(int32_t device, int64_t n, const float* x_norm, float kmin, float kmax, float lo, float hi, float* out, void* stream) {
  if (n < 0 || (n > 0 && (!x_norm || !out)) || !(kmin > 0.f) || !(kmax > kmin) || !(hi > lo)) {
    srm_set_error("srm_denormalize_log: bad argument");
    return SRM_ERR_INVALID;
  }
  SRM_CUDA_CHECK(cudaSetDevice(device));
  return srm_launch_denorm_log(n, x_norm, kmin, kmax, lo, hi, out, (cudaStream_t)stream);
}

int srm_selftest_rounding(int32_t device, int64_t n, uint64_t seed, int64_t* mismatches, void* stream) {
  if (n < 0 || !mismatches) { srm_set_error("srm_selftest_rounding: bad argument"); return SRM_ERR_INVALID; }
  SRM_CUDA_CHECK(cudaSetDevice(device));
  return srm_launch_selftest_rounding(n, seed, mismatches, (cudaStream_t)stream);
}

int srm_relperm(const SrmHandle* h, int64_t n, const float* sg, float* krog, float* krgo, float* dkrog, float* dkrgo, void* stream) {
  if (!h || n < 0 || (n > 0 && !sg)) { srm_set_error("srm_relperm: bad argument"); return SRM_ERR_INVALID; }
  if (!(h->dev.kr_den_o > 0.f) || !(h->dev.kr_den_g > 0.f)) { srm_set_error("srm_relperm: the handle carries no SCAL end points"); return SRM_ERR_INVALID; }
  SRM_CUDA_CHECK(cudaSetDevice(h->device));
  return srm_launch_relperm(h, n, sg, krog, krgo, dkrog, dkrgo, (cudaStream_t)stream);
}

static int check_batch(const SrmHandle* h, int32_t B, int32_t R, const char* who) {
  if (!h) { srm_set_error("%s: null handle", who); return SRM_ERR_INVALID; }
  if (B < 1 || R < 1 || B > 65535 || R > SRM_MAXR) { srm_set_error("%s: B=%d (1..65535), R=%d (1..%d)", who, B, R, SRM_MAXR); return SRM_ERR_INVALID; }
  return SRM_OK;
}

int srm_wells(const SrmHandle* h, int32_t B, int32_t R, const float* kx, const int32_t* sample_real,
              const float* p, const float* t_days, float* qw, float* pwfw, float* dqdp, float* q_dense,
              float* pwf_dense, void* stream) {
  int rc = check_batch(h, B, R, "srm_wells");
  if (rc) return rc;
  if (!kx || !p || !t_days) { srm_set_error("srm_wells: null input"); return SRM_ERR_INVALID; }
  SRM_CUDA_CHECK(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t N = h->dev.N;
  if (q_dense) SRM_CUDA_CHECK(cudaMemsetAsync(q_dense, 0, sizeof(float) * B * N, s));
  if (pwf_dense) SRM_CUDA_CHECK(cudaMemsetAsync(pwf_dense, 0, sizeof(float) * B * N, s));
  const int nw = h->dev.n_wells;
  if (nw == 0) return SRM_OK;
  float* tmp = nullptr;
  SRM_CUDA_CHECK(cudaMallocAsync((void**)&tmp, sizeof(float) * 3 * (size_t)B * nw, s));
  float *tq = tmp, *tp = tmp + (size_t)B * nw, *td = tmp + 2 * (size_t)B * nw;
  rc = (h->cfg.numerics == SRM_NUMERICS_CLOSED_FORM) ? srm_launch_wells_cf(h, B, kx, sample_real, R, p, t_days, tq, tp, td, s)
                                                      : srm_launch_wells_ref(h, B, kx, sample_real, R, p, t_days, tq, tp, td, s);
  if (!rc && qw) rc = srm_launch_unsort_wells(h, B, tq, qw, s);
  if (!rc && pwfw) rc = srm_launch_unsort_wells(h, B, tp, pwfw, s);
  if (!rc && dqdp) rc = srm_launch_unsort_wells(h, B, td, dqdp, s);
  if (!rc && q_dense) rc = srm_launch_scatter_wells(h, B, tq, q_dense, s);
  if (!rc && pwf_dense) rc = srm_launch_scatter_wells(h, B, tp, pwf_dense, s);
  cudaFreeAsync(tmp, s);
  return rc;
}

int srm_forward(SrmHandle* h, int32_t B, int32_t R, const float* kx, const int32_t* sample_real,
                const float* p0, const float* p1, const float* dt1, const float* dt2, const float* t1,
                float* terms_out, float* dom_out, float* qw_out, float* pwfw_out, void* workspace,
                size_t workspace_bytes, int32_t flags, void* stream) {
  int rc = check_batch(h, B, R, "srm_forward");
  if (rc) return rc;
  if (!kx || !p0 || !p1 || !dt1 || !dt2 || !t1 || !terms_out || !workspace) {
    srm_set_error("srm_forward: null argument");
    return SRM_ERR_INVALID;
  }
  if (h->cfg.fluid_type != SRM_FLUID_DG) { srm_set_error("srm_forward: the handle is not dry gas (use srm_forward_gc)"); return SRM_ERR_INVALID; }
  const SrmWs ws = carve(h, workspace, B, R);
  if (ws.bytes > workspace_bytes) {
    srm_set_error("srm_forward: workspace %zu < required %zu bytes", workspace_bytes, ws.bytes);
    return SRM_ERR_WORKSPACE;
  }
  SRM_CUDA_CHECK(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  const bool save = (flags & SRM_FLAG_SAVE_FOR_BACKWARD) != 0;
  h->st_valid = 0;
  const int mode = srm_ws_mode(h);
  if (mode == SRM_WS_CF)
    rc = srm_forward_cf(h, B, R, kx, sample_real, p0, p1, dt1, dt2, t1, terms_out, dom_out, ws, save, s);
  else if (mode == SRM_WS_REF_FUSED)
    rc = srm_forward_ref2(h, B, R, kx, sample_real, p0, p1, dt1, dt2, t1, terms_out, dom_out, ws, s, -1, save);
  else
    rc = srm_forward_ref(h, B, R, kx, sample_real, p0, p1, dt1, dt2, t1, terms_out, dom_out, ws, save, s);
  if (rc) return rc;
  if (qw_out && h->dev.n_wells) { rc = srm_launch_unsort_wells(h, B, ws.qw, qw_out, s); if (rc) return rc; }
  if (pwfw_out && h->dev.n_wells) { rc = srm_launch_unsort_wells(h, B, ws.pwfw, pwfw_out, s); if (rc) return rc; }
  if (save) srm_state_set(h, B, R, workspace, kx, sample_real, p0, p1, nullptr, nullptr, nullptr, nullptr, dt1, dt2, t1);
  return SRM_OK;
}

int srm_backward(SrmHandle* h, int32_t B, int32_t R, const float* kx, const int32_t* sample_real,
                 const float* p0, const float* p1, const float* dt1, const float* dt2, const float* t1,
                 const float* dterms, float* gp0, float* gp1, float* gdt1, float* gdt2, void* workspace,
                 size_t workspace_bytes, int32_t flags, void* stream) {
  int rc = check_batch(h, B, R, "srm_backward");
  if (rc) return rc;
  (void)flags;
  if (!kx || !p0 || !p1 || !dt1 || !dt2 || !t1 || !dterms || !gp0 || !gp1 || !gdt1 || !gdt2 || !workspace) {
    srm_set_error("srm_backward: null argument");
    return SRM_ERR_INVALID;
  }
  if (h->cfg.fluid_type != SRM_FLUID_DG) { srm_set_error("srm_backward: the handle is not dry gas (use srm_backward_gc)"); return SRM_ERR_INVALID; }
  const SrmWs ws = carve(h, workspace, B, R);
  if (ws.bytes > workspace_bytes) {
    srm_set_error("srm_backward: workspace %zu < required %zu bytes", workspace_bytes, ws.bytes);
    return SRM_ERR_WORKSPACE;
  }
  SRM_CUDA_CHECK(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  bool have = srm_state_is(h, B, R, workspace, kx, sample_real, p0, p1, nullptr, nullptr, nullptr, nullptr, dt1, dt2, t1);
  const int mode = srm_ws_mode(h);
  const bool cf = mode == SRM_WS_CF;
  // fused reference path: the adjoint's kernel family follows from the alignment of ITS pointers (gp0, gp1 included);
  // a state saved by the other family's forward is recomputed in this one's, so the pair never mixes families
  const int fam = mode == SRM_WS_REF_FUSED ? srm_ref2_backward_family(h, p0, p1, ws.dom, gp0, gp1) : -1;
  if (have && fam >= 0 && h->st_family != fam) have = false;
  if (!have) {
    h->st_valid = 0;
    // recompute the forward state (PVT stage with derivatives, wells, residual field) into the workspace
    float* terms_tmp = nullptr;
    SRM_CUDA_CHECK(cudaMallocAsync((void**)&terms_tmp, sizeof(float) * 2 * SRM_N_TERMS, s));
    rc = cf ? srm_forward_cf(h, B, R, kx, sample_real, p0, p1, dt1, dt2, t1, terms_tmp, nullptr, ws, true, s)
         : mode == SRM_WS_REF_FUSED ? srm_forward_ref2(h, B, R, kx, sample_real, p0, p1, dt1, dt2, t1, terms_tmp, nullptr, ws, s, fam, true)
            : srm_forward_ref(h, B, R, kx, sample_real, p0, p1, dt1, dt2, t1, terms_tmp, nullptr, ws, true, s);
    cudaFreeAsync(terms_tmp, s);
    if (rc) return rc;
    // the workspace now holds THIS call's forward state (it overwrote whatever forward was saved there before)
    srm_state_set(h, B, R, workspace, kx, sample_real, p0, p1, nullptr, nullptr, nullptr, nullptr, dt1, dt2, t1);
  }
  rc = cf ? srm_backward_cf(h, B, R, kx, sample_real, p0, p1, dt1, dt2, t1, dterms, gp0, gp1, gdt1, gdt2, ws, s)
       : mode == SRM_WS_REF_FUSED ? srm_backward_ref2(h, B, R, kx, sample_real, p0, p1, dt1, dt2, t1, dterms, gp0, gp1, gdt1, gdt2, ws, s)
          : srm_backward_ref(h, B, R, kx, sample_real, p0, p1, dt1, dt2, t1, dterms, gp0, gp1, gdt1, gdt2, ws, s);
  return rc;
}

int srm_forward_gc(SrmHandle* h, int32_t B, int32_t R, const float* kx, const int32_t* sample_real,
                   const float* p0, const float* p1, const float* sg0, const float* sg1, const float* so0,
                   const float* so1, const float* dt1, const float* dt2, const float* t1, float* terms_out,
                   float* dom_out, float* q4w_out, float* pwfw_out, void* workspace, size_t workspace_bytes,
                   int32_t flags, void* stream) {
  int rc = check_batch(h, B, R, "srm_forward_gc");
  if (rc) return rc;
  if (!kx || !p0 || !p1 || !sg0 || !sg1 || !so0 || !so1 || !dt1 || !dt2 || !t1 || !terms_out || !workspace) {
    srm_set_error("srm_forward_gc: null argument");
    return SRM_ERR_INVALID;
  }
  if (h->cfg.fluid_type != SRM_FLUID_GC) { srm_set_error("srm_forward_gc: the handle is not gas condensate"); return SRM_ERR_INVALID; }
  const SrmWs ws = carve(h, workspace, B, R);
  if (ws.bytes > workspace_bytes) {
    srm_set_error("srm_forward_gc: workspace %zu < required %zu bytes", workspace_bytes, ws.bytes);
    return SRM_ERR_WORKSPACE;
  }
  SRM_CUDA_CHECK(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  const bool save = (flags & SRM_FLAG_SAVE_FOR_BACKWARD) != 0;
  h->st_valid = 0;
  rc = srm_forward_gc_impl(h, B, R, kx, sample_real, p0, p1, sg0, sg1, so0, so1, dt1, dt2, t1, terms_out, dom_out, ws, save, s);
  if (rc) return rc;
  const int nw = h->dev.n_wells;
  if (nw) {
    const size_t wt = (size_t)B * nw;
    if (q4w_out)
      for (int X = 0; X < 4; ++X) { rc = srm_launch_unsort_wells(h, B, ws.gc_wells + X * wt, q4w_out + X * wt, s); if (rc) return rc; }
    if (pwfw_out) { rc = srm_launch_unsort_wells(h, B, ws.pwfw, pwfw_out, s); if (rc) return rc; }
  }
  if (save) srm_state_set(h, B, R, workspace, kx, sample_real, p0, p1, sg0, sg1, so0, so1, dt1, dt2, t1);
  return SRM_OK;
}

int srm_backward_gc(SrmHandle* h, int32_t B, int32_t R, const float* kx, const int32_t* sample_real,
                    const float* p0, const float* p1, const float* sg0, const float* sg1, const float* so0,
                    const float* so1, const float* dt1, const float* dt2, const float* t1, const float* dterms,
                    float* gp0, float* gp1, float* gsg0, float* gsg1, float* gso0, float* gso1, float* gdt1,
                    float* gdt2, void* workspace, size_t workspace_bytes, int32_t flags, void* stream) {
  int rc = check_batch(h, B, R, "srm_backward_gc");
  if (rc) return rc;
  (void)flags;
  if (!kx || !p0 || !p1 || !sg0 || !sg1 || !so0 || !so1 || !dt1 || !dt2 || !t1 || !dterms || !gp0 || !gp1 || !gsg0 || !gsg1 ||
      !gso0 || !gso1 || !gdt1 || !gdt2 || !workspace) {
    srm_set_error("srm_backward_gc: null argument");
    return SRM_ERR_INVALID;
  }
  if (h->cfg.fluid_type != SRM_FLUID_GC) { srm_set_error("srm_backward_gc: the handle is not gas condensate"); return SRM_ERR_INVALID; }
  const SrmWs ws = carve(h, workspace, B, R);
  if (ws.bytes > workspace_bytes) {
    srm_set_error("srm_backward_gc: workspace %zu < required %zu bytes", workspace_bytes, ws.bytes);
    return SRM_ERR_WORKSPACE;
  }
  SRM_CUDA_CHECK(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  const bool have = srm_state_is(h, B, R, workspace, kx, sample_real, p0, p1, sg0, sg1, so0, so1, dt1, dt2, t1);
  if (!have) {
    h->st_valid = 0;
    float* terms_tmp = nullptr;
    SRM_CUDA_CHECK(cudaMallocAsync((void**)&terms_tmp, sizeof(float) * 2 * SRM_N_TERMS, s));
    rc = srm_forward_gc_impl(h, B, R, kx, sample_real, p0, p1, sg0, sg1, so0, so1, dt1, dt2, t1, terms_tmp, nullptr, ws, true, s);
    cudaFreeAsync(terms_tmp, s);
    if (rc) return rc;
    srm_state_set(h, B, R, workspace, kx, sample_real, p0, p1, sg0, sg1, so0, so1, dt1, dt2, t1);
  }
  return srm_backward_gc_impl(h, B, R, kx, sample_real, p0, p1, sg0, sg1, so0, so1, dt1, dt2, t1, dterms, gp0, gp1, gsg0, gsg1,
                              gso0, gso1, gdt1, gdt2, ws, s);
}

}  // extern "C"
