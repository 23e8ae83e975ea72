// SRM_NUMERICS_CLOSED_FORM kernels: the HBM-bound path.
//
// Same formulas as the reference (physics_loss.py:79-208; polyhm_splines.py:138-146), evaluated the
// way exact arithmetic would: the order-1 polyharmonic interpolant
//     f(x) = sum_i w_i |x - c_i| + v0 x + v1
// is piecewise linear, so PVT is one table lookup + one FMA per property; the flux is assembled in
// difference form a_f (p_c - p_n) (no a*p cancellation); the truncation bracket N, which vanishes
// identically for the linear extrapolation of p2, is taken as 0.  Results are closer to the fp64
// evaluation of the reference's formulas than the reference's own fp32 evaluation is (tests).
//
// Data movement (DG, per cell-timestep): forward reads p0,p1 and writes dom (12 B), adjoint reads
// p0,p1,dom and writes gp0,gp1 (20 B); kx is read once per (realisation, tile) and amortised over the
// time samples of the realisation (the face transmissibilities live in shared memory while the CTA
// walks the samples).
//
// Kernel structure (forward and adjoint alike):
//   work item  = (segment of samples of one realisation) x (z-chunk of DZ planes) x (32 x TY tile)
//   scheduling = persistent CTAs + atomic work counter
//   per item   : build TE/TN/TU (static face coefficients) in smem
//   per sample : march over the DZ+2 planes of the chunk; each plane arrives in a ring of smem stages by
//                TMA (cp.async.bulk.tensor, 4-D tensor maps over (W,H,D,B), out-of-bounds zero fill =
//                the zero-flux image cells once TE/TN/TU are zero on the boundary); z neighbours stay
//                in registers, x/y neighbours come from the stage (p1) and a triple-buffered G plane.
#include <cuda.h>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "wells.cuh"

// ------------------------------------------------------------------------------------------
// host: tables
// ------------------------------------------------------------------------------------------
int srm_build_closed_form(SrmHandle* h, const SrmConfig* cfg) {
  const int n = cfg->n_knots, P = cfg->n_props;
  std::vector<SrmClosedForm> host(1);
  SrmClosedForm& T = host[0];
  std::memset(&T, 0, sizeof(T));
  T.n = n;
  // interval k = number of knots <= x: k=0 is x < c0 (anchor c0), k>=1 is [c_{k-1}, c_k) (anchor c_{k-1})
  for (int k = 0; k <= n; ++k) {
    const double xa = (k == 0) ? (double)cfg->knots[0] : (double)cfg->knots[k - 1];
    T.x0[k] = (float)xa;
    T.lohi[k].x = (k == 0) ? -INFINITY : cfg->knots[k - 1];
    T.lohi[k].y = (k == n) ? INFINITY : cfg->knots[k];
    for (int q = 0; q < P; ++q) {
      const float* w = cfg->spline_w + (size_t)q * n;
      const double v0 = cfg->spline_v[2 * q], v1 = cfg->spline_v[2 * q + 1];
      double f = v0 * xa + v1, sl = v0;
      for (int i = 0; i < n; ++i) {
        const double ci = cfg->knots[i];
        f += (double)w[i] * std::fabs(xa - ci);
        sl += (i < k) ? (double)w[i] : -(double)w[i];   // knots with index < k lie to the left
      }
      T.f0[q][k] = (float)f;
      T.slope[q][k] = (float)sl;
      if (q == 0) T.f0d[k] = f;
    }
    T.ent[k] = make_float4(T.x0[k], T.f0[0][k], T.slope[0][k], T.f0[1][k]);
    T.sM[k] = T.slope[1][k];
  }
  // bucket table over (0, p_max]: width = smallest knot spacing inside the clamp window
  double wmin = 1e300;
  for (int i = 1; i < n; ++i)
    if (cfg->knots[i] >= cfg->p_min && cfg->knots[i - 1] <= cfg->p_max) wmin = std::fmin(wmin, (double)cfg->knots[i] - cfg->knots[i - 1]);
  T.use_bucket = 0;
  if (wmin < 1e300 && wmin > 0 && cfg->p_max > 0) {
    const int nb = (int)std::floor(cfg->p_max / wmin) + 2;
    if (nb <= SRM_CF_MAXBUCKET) {
      T.use_bucket = 1;
      T.nb = nb;
      T.inv_w = (float)(1.0 / wmin);
      for (int b = 0; b < nb; ++b) {
        const double edge = b * wmin;
        int cnt = 0;
        while (cnt < n && (double)cfg->knots[cnt] <= edge) ++cnt;
        T.bucket[b] = (unsigned char)cnt;
      }
    }
  }
  cudaError_t e = cudaMalloc((void**)&h->d_cf, sizeof(SrmClosedForm));
  if (e == cudaSuccess) e = cudaMemcpy(h->d_cf, &T, sizeof(SrmClosedForm), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { srm_set_error("closed-form table upload: %s", cudaGetErrorString(e)); return SRM_ERR_CUDA; }
  return SRM_OK;
}

namespace {

// ------------------------------------------------------------------------------------------
// closed-form PVT
// ------------------------------------------------------------------------------------------
// number of knots <= x (x already clamped, finite).  on_knot: x coincides with a knot (|x - c| below the
// reference's sqrt(EPSILON) = 1e-5 guard, polyhm_splines.py:77-80): the reference's gradient mask
// [r >= EPSILON] then drops that knot's term, i.e. the derivative there is the MEAN of the two slopes.
__device__ __forceinline__ int cf_interval(const SrmClosedForm* __restrict__ T, float x, bool& on_knot) {
  int k;
  if (T->use_bucket) {
    int b = (int)(x * T->inv_w);
    b = min(max(b, 0), T->nb - 1);
    k = T->bucket[b];
  } else {
    k = T->n >> 1;
  }
  float2 lh = T->lohi[k];
  if (x >= lh.y) { do { ++k; lh = T->lohi[k]; } while (x >= lh.y); }
  else if (x < lh.x) { do { --k; lh = T->lohi[k]; } while (x < lh.x); }
  on_knot = (x - lh.x) < 1e-5f;     // lh.x = -inf for k == 0 -> false
  return k;
}

__device__ __forceinline__ float cf_clamp(const SrmDev& P, float p, float& pass) {
  const float x = fminf(fmaxf(p, P.p_min), P.p_max);
  pass = (p >= P.p_min && p <= P.p_max) ? 1.f : 0.f;   // tf.maximum/minimum pass ties
  return x;
}

// invBg and its slope at p0
// (kA, dA): interval and increment sA*(x - x0[k]) over the interval's anchor value -- the material balance
// sums A1 - A0 = (f0[k1] - f0[k0]) + (dA1 - dA0), whose first part is exact in fp64 and zero when k1 == k0
__device__ __forceinline__ void cf_pvt_p0(const SrmDev& P, const SrmClosedForm* __restrict__ T, float p, float& A, float& Ap, float& pass,
                                          int& kA, float& dA) {
  const float x = cf_clamp(P, p, pass);
  bool on;
  const int k = cf_interval(T, x, on);
  const float4 e = T->ent[k];
  kA = k;
  dA = e.z * (x - e.x);
  A = e.y + dA;
  Ap = e.z;
  if (on) Ap = 0.5f * (e.z + T->ent[k - 1].z);
}
// invBg, G = invBg*invug and their slopes at p1
__device__ __forceinline__ void cf_pvt_p1(const SrmDev& P, const SrmClosedForm* __restrict__ T, float p, float& A, float& G, float& Ap, float& Gp,
                                          int& kA, float& dA) {
  float pass;
  const float x = cf_clamp(P, p, pass);
  bool on;
  const int k = cf_interval(T, x, on);
  const float4 e = T->ent[k];
  float sA = e.z, sM = T->sM[k];
  const float dx = x - e.x;
  kA = k;
  dA = sA * dx;
  A = e.y + dA;
  const float M = fmaf(sM, dx, e.w);
  G = A * M;
  if (on) { sA = 0.5f * (sA + T->ent[k - 1].z); sM = 0.5f * (sM + T->sM[k - 1]); }
  Ap = sA * pass;
  Gp = (sA * M + A * sM) * pass;
}
__device__ __forceinline__ float cf_G(const SrmDev& P, const SrmClosedForm* __restrict__ T, float p) {
  float pass;
  const float x = cf_clamp(P, p, pass);
  bool on;
  const int k = cf_interval(T, x, on);
  const float4 e = T->ent[k];
  const float dx = x - e.x;
  return fmaf(e.z, dx, e.y) * fmaf(T->sM[k], dx, e.w);
}

__global__ void __launch_bounds__(256) k_pvt_eval_cf(const __grid_constant__ SrmDev P, const SrmClosedForm* __restrict__ T,
                                                     int64_t n, const float* __restrict__ p, float* __restrict__ val,
                                                     float* __restrict__ dval) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  float pass;
  const float x = cf_clamp(P, p[g], pass);
  bool on;
  const int k = cf_interval(T, x, on);
  const float dx = x - T->x0[k];
  for (int q = 0; q < P.n_props; ++q) {
    if (val) val[(int64_t)q * n + g] = fmaf(T->slope[q][k], dx, T->f0[q][k]);
    // derivative w.r.t. the clamped input
    if (dval) dval[(int64_t)q * n + g] = on ? 0.5f * (T->slope[q][k] + T->slope[q][k - 1]) : T->slope[q][k];
  }
}

// A(p1) - A(p0) from the (interval, increment) pairs
__device__ __forceinline__ float cf_dA(const SrmClosedForm* __restrict__ T, int k1, float d1, int k0, float d0) {
  float r = d1 - d0;
  if (k1 != k0) r += (float)(T->f0d[k1] - T->f0d[k0]);
  return r;
}

struct MobilityCf {
  const SrmClosedForm* T;
  __device__ __forceinline__ Dual operator()(const SrmDev& P, Dual p) const {
    float A, G, Ap, Gp, dA;
    int kA;
    cf_pvt_p1(P, T, p.v, A, G, Ap, Gp, kA, dA);
    return dmk(P.krg * G, P.krg * Gp * p.d);
  }
};

// ------------------------------------------------------------------------------------------
// grouping: samples -> per-realisation lists (stable) -> segments of at most `tchunk` samples
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_group_samples(int32_t B, int32_t R, const int32_t* __restrict__ sample_real,
                                                        int32_t tchunk, int32_t* __restrict__ cnt /*[R+1]*/,
                                                        int32_t* __restrict__ fill /*[R]*/, int32_t* __restrict__ list,
                                                        int32_t* __restrict__ seg, int32_t* __restrict__ ctl) {
  __shared__ int32_t s_key[1024];
  __shared__ int32_t s_scan[1024];
  const int tid = threadIdx.x;
  auto key_of = [&](int b) { return srm_real_of(sample_real, b, B, R); };
  for (int r = tid; r <= R; r += 1024) cnt[r] = 0;
  for (int r = tid; r < R; r += 1024) fill[r] = 0;
  __syncthreads();
  for (int b = tid; b < B; b += 1024) {
    const int r = key_of(b);
    if (r >= 0 && r < R) atomicAdd(&cnt[r], 1);
  }
  __syncthreads();
  // exclusive scan of cnt[0..R) -> starts; cnt[R] = total
  int carry = 0;
  for (int base = 0; base < R; base += 1024) {
    const int r = base + tid;
    const int v = (r < R) ? cnt[r] : 0;
    s_scan[tid] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int t = (tid >= o) ? s_scan[tid - o] : 0;
      __syncthreads();
      s_scan[tid] += t;
      __syncthreads();
    }
    if (r < R) cnt[r] = carry + s_scan[tid] - v;
    const int tot = s_scan[1023];
    __syncthreads();
    carry += tot;
  }
  if (tid == 0) cnt[R] = carry;
  __syncthreads();
  // stable fill, 1024 samples at a time
  for (int base = 0; base < B; base += 1024) {
    const int b = base + tid;
    const int r = (b < B) ? key_of(b) : -1;
    s_key[tid] = r;
    __syncthreads();
    if (r >= 0 && r < R) {
      int rank = 0;
      for (int u = 0; u < tid; ++u) rank += (s_key[u] == r);
      list[cnt[r] + fill[r] + rank] = b;
    }
    __syncthreads();
    if (r >= 0 && r < R) atomicAdd(&fill[r], 1);
    __syncthreads();
  }
  // segments (serial over realisations: R is small compared with the field work)
  if (tid == 0) {
    int ns = 0;
    for (int r = 0; r < R; ++r) {
      const int start = cnt[r], n = cnt[r + 1] - cnt[r];
      for (int o = 0; o < n; o += tchunk) {
        seg[3 * ns + 0] = r;
        seg[3 * ns + 1] = start + o;
        seg[3 * ns + 2] = min(tchunk, n - o);
        ++ns;
      }
    }
    ctl[0] = ns;
    ctl[1] = 0;
    ctl[2] = 0;
  }
}

// ------------------------------------------------------------------------------------------
// TMA / mbarrier primitives
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z, int b) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z), "r"(b)
      : "memory");
}

// ------------------------------------------------------------------------------------------
// tiled kernels
// ------------------------------------------------------------------------------------------
constexpr int TX = 32;          // tile width (one warp per tile row)
constexpr int XO = 4;           // column of the tile's first cell inside a haloed box
constexpr int BX = TX + 2 * XO; // p1/dom box width: columns x0-4 .. x0+35.  TMA needs the box row (160 B) AND the start
                                // coordinate (x0-4) to be multiples of 16 bytes -- an odd start traps as "illegal instruction"

struct CfArgs {
  // inputs
  const float* p0; const float* p1; const float* kx; const float* dt1; const float* dt2;
  const int32_t* list; const int32_t* seg; int32_t* ctl;
  const float* qw; const float* dqdp; float* divqw;
  // forward outputs
  float* dom; float* dom_out; double* sse; double* mb_sum;
  // adjoint
  const float* dterms; const float* mbc; float* gp0; float* gp1; double* gdt1_acc;
  const SrmClosedForm* T;
  int32_t B, R;
  int32_t tiles_x, tiles_y, zchunks;
  int32_t ctl_slot;      // which work counter
};

template <int TY, int DZ, int S>
struct Lay {
  static constexpr int BY = TY + 2;
  static constexpr int NT = TX * TY;
  static constexpr int BOX1 = BX * BY;                       // floats of a haloed plane box
  static constexpr int BOX1_B = ((BOX1 * 4 + 127) / 128) * 128;
  static constexpr int BOX0 = TX * TY;
  static constexpr int BOX0_B = BOX0 * 4;                    // multiple of 128 (TX*4 = 128)
  static constexpr int TE_N = DZ * TY * 33;
  static constexpr int TN_N = DZ * (TY + 1) * TX;
  static constexpr int TU_N = (DZ + 1) * TY * TX;
  // byte offsets inside the dynamic shared-memory block (NB1 = haloed boxes per stage: 1 fwd, 2 adjoint)
  template <int NB1> struct Off {
    static constexpr int STAGE = NB1 * BOX1_B + BOX0_B;      // [p1 box][dom box (adjoint)][p0 box]
    static constexpr int GS = S * STAGE;
    static constexpr int TE = GS + 3 * BOX1_B;
    static constexpr int TN = TE + ((TE_N * 4 + 15) / 16) * 16;
    static constexpr int TU = TN + ((TN_N * 4 + 15) / 16) * 16;
    static constexpr int TAB = TU + ((TU_N * 4 + 15) / 16) * 16;
    static constexpr int FULL = TAB + (((int)sizeof(SrmClosedForm) + 15) / 16) * 16;
    static constexpr int COL = FULL + ((S * 8 + 15) / 16) * 16;
    static constexpr int RED = COL + ((NT + 15) / 16) * 16;
    static constexpr int TOTAL = RED + 32 * 4;
  };
};

// static face coefficients of one item:  0.5 * C * krg / dl^2 * harmonic(k_a, k_b), 0 on the grid boundary
template <int TY, int DZ>
__device__ __forceinline__ void build_statics(const SrmDev& P, const float* __restrict__ kr, int x0, int y0, int k0,
                                              float* TE, float* TN, float* TU) {
  const float cx = 0.5f * P.C * P.krg * P.idx * P.idx, cy = 0.5f * P.C * P.krg * P.idy * P.idy,
              cz = 0.5f * P.C * P.krg * P.idz * P.idz;
  const int HW = P.H * P.W, nt = TX * TY;
  auto kat = [&](int x, int y, int z) { return kr[(int64_t)z * HW + y * P.W + x]; };
  auto hm = [](float a, float b) { return __fdividef(2.f * a * b, a + b); };
  // TE[kk][ty][xx]: face between columns (x0+xx-1) and (x0+xx)
  for (int e = threadIdx.x; e < DZ * TY * 33; e += nt) {
    const int xx = e % 33, ty = (e / 33) % TY, kk = e / (33 * TY);
    const int xa = x0 + xx - 1, xb = x0 + xx, y = y0 + ty, z = k0 + kk;
    float v = 0.f;
    if (xa >= 0 && xb < P.W && y < P.H && z < P.D) v = cx * hm(kat(xa, y, z), kat(xb, y, z));
    TE[e] = v;
  }
  // TN[kk][yy][tx]: face between rows (y0+yy-1) and (y0+yy)
  for (int e = threadIdx.x; e < DZ * (TY + 1) * TX; e += nt) {
    const int tx = e % TX, yy = (e / TX) % (TY + 1), kk = e / (TX * (TY + 1));
    const int ya = y0 + yy - 1, yb = y0 + yy, x = x0 + tx, z = k0 + kk;
    float v = 0.f;
    if (ya >= 0 && yb < P.H && x < P.W && z < P.D) v = cy * hm(P.kx_ky * kat(x, ya, z), P.kx_ky * kat(x, yb, z));
    TN[e] = v;
  }
  // TU[kk2][ty][tx]: face between planes (k0+kk2-1) and (k0+kk2)
  for (int e = threadIdx.x; e < (DZ + 1) * TY * TX; e += nt) {
    const int tx = e % TX, ty = (e / TX) % TY, kk = e / (TX * TY);
    const int za = k0 + kk - 1, zb = k0 + kk, x = x0 + tx, y = y0 + ty;
    float v = 0.f;
    if (za >= 0 && zb < P.D && x < P.W && y < P.H) v = cz * hm(P.kv_kh * kat(x, y, za), P.kv_kh * kat(x, y, zb));
    TU[e] = v;
  }
}

// halo cell (x,y) in box coordinates served by thread h (rows y=0 and y=TY+1, columns x=XO-1 and x=XO+TX)
template <int TY>
__device__ __forceinline__ bool halo_cell(int h, int& x, int& y) {
  if (h < TX) { x = h + XO; y = 0; return true; }
  if (h < 2 * TX) { x = h - TX + XO; y = TY + 1; return true; }
  if (h < 2 * TX + TY) { x = XO - 1; y = h - 2 * TX + 1; return true; }
  if (h < 2 * TX + 2 * TY) { x = TX + XO; y = h - 2 * TX - TY + 1; return true; }
  return false;
}

// cooperative (non-TMA) load of a box [bx x by] at (xs, ys, z, b); zero outside the grid
__device__ __forceinline__ void load_box_manual(const SrmDev& P, const float* __restrict__ f, int b, int z, int xs, int ys,
                                                int bx, int by, float* dst, int nt) {
  for (int e = threadIdx.x; e < bx * by; e += nt) {
    const int x = xs + e % bx, y = ys + e / bx;
    float v = 0.f;
    if (x >= 0 && x < P.W && y >= 0 && y < P.H && z >= 0 && z < P.D) v = __ldg(&f[(int64_t)b * P.N + ((int64_t)z * P.H + y) * P.W + x]);
    dst[e] = v;
  }
}

template <int NV>
__device__ __forceinline__ void warp_sum(float (&v)[NV]) {
#pragma unroll
  for (int q = 0; q < NV; ++q)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_down_sync(0xffffffffu, v[q], o);
}

// ---------------------------------------------------------------- forward
template <int TY, int DZ, int S, bool TMA>
__global__ void __launch_bounds__(TX * TY, 1) k_fwd_cf(const __grid_constant__ SrmDev P, const __grid_constant__ CfArgs A,
                                                       const __grid_constant__ CUtensorMap map_p1,
                                                       const __grid_constant__ CUtensorMap map_p0) {
  using L = Lay<TY, DZ, S>;
  extern __shared__ __align__(1024) unsigned char smem[];
  using O = typename L::template Off<1>;
  auto stage_p1 = [&](int s) { return (float*)(smem + s * O::STAGE); };
  auto stage_p0 = [&](int s) { return (float*)(smem + s * O::STAGE + L::BOX1_B); };
  float* Gs = (float*)(smem + O::GS);
  float* TE = (float*)(smem + O::TE);
  float* TN = (float*)(smem + O::TN);
  float* TU = (float*)(smem + O::TU);
  SrmClosedForm* Ts = (SrmClosedForm*)(smem + O::TAB);
  uint64_t* full = (uint64_t*)(smem + O::FULL);
  unsigned char* colflag = smem + O::COL;
  float* red = (float*)(smem + O::RED);
  __shared__ int s_item;

  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  // closed-form tables -> smem
  for (int e = tid; e < (int)(sizeof(SrmClosedForm) / 4); e += L::NT) ((uint32_t*)Ts)[e] = ((const uint32_t*)A.T)[e];
  if (TMA && tid == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nseg = A.ctl[0];
  const int tiles = A.tiles_x * A.tiles_y;
  const int total = nseg * A.zchunks * tiles;
  uint32_t ring_phase = 0;   // bit s = parity of stage s
  float acc_dom = 0.f, acc_tde = 0.f;   // per-thread partial sums over the CTA's whole life
  double acc_ibc = 0.0;

  for (;;) {
    if (tid == 0) s_item = atomicAdd(&A.ctl[A.ctl_slot], 1);
    __syncthreads();
    const int item = s_item;
    if (item >= total) break;
    const int tile = item % tiles, zc = (item / tiles) % A.zchunks, sg = item / (tiles * A.zchunks);
    const int x0 = (tile % A.tiles_x) * TX, y0 = (tile / A.tiles_x) * TY, k0 = zc * DZ;
    const int r = A.seg[3 * sg], loff = A.seg[3 * sg + 1], nT = A.seg[3 * sg + 2];
    build_statics<TY, DZ>(P, A.kx + (int64_t)r * P.N, x0, y0, k0, TE, TN, TU);
    // which columns of this tile carry a well connection inside the chunk
    colflag[tid] = 0;
    __syncthreads();
    for (int w = tid; w < P.n_wells; w += L::NT) {
      const int c = P.wells[w].cell, wi = c % P.W, wj = (c / P.W) % P.H, wk = c / (P.W * P.H);
      if (wi >= x0 && wi < x0 + TX && wj >= y0 && wj < y0 + TY && wk >= k0 && wk < k0 + DZ) colflag[(wj - y0) * TX + (wi - x0)] = 1;
    }
    __syncthreads();
    const bool has_well = colflag[tid] != 0;
    const int gx = x0 + tx, gy = y0 + ty;
    const bool in_xy = (gx < P.W) && (gy < P.H);
    const int J = nT * (DZ + 2);

    auto issue = [&](int j) {   // TMA for job j (thread 0)
      const int s = j % S, ti = j / (DZ + 2), kk = j % (DZ + 2) - 1;
      const int b = A.list[loff + ti], z = k0 + kk;
      const bool want0 = (kk >= 0 && kk < DZ);
      mbar_expect_tx(&full[s], L::BOX1 * 4 + (want0 ? L::BOX0 * 4 : 0));
      tma_load_4d(stage_p1(s), &map_p1, &full[s], x0 - XO, y0 - 1, z, b);
      if (want0) tma_load_4d(stage_p0(s), &map_p0, &full[s], x0, y0, z, b);
    };
    if (TMA) {
      if (tid == 0) for (int j = 0; j < min(S, J); ++j) issue(j);
    }
    // registers carried along z
    float p_prev = 0.f, G_prev = 0.f, p_cur = 0.f, G_cur = 0.f, dA1_cur = 0.f;
    int kA1_cur = 0;
    float cA = 0.f, cT = 0.f, mbk = 0.f, mb_part = 0.f;
    int b_cur = 0;
    bool flush_pending = false;
    int flush_b = 0;
    float flush_k = 0.f;

    for (int j = 0; j < J; ++j) {
      const int s = j % S, ti = j / (DZ + 2), kk = j % (DZ + 2) - 1;
      if (kk == -1) {
        b_cur = A.list[loff + ti];
        const float d1 = A.dt1[b_cur];
        cA = P.dv * P.invDc / d1;
        cT = P.dvDc * 2e-7f / d1;
        mbk = P.dvSgi_phi / (P.Dc * d1);
        mb_part = 0.f;
      }
      if (TMA) {
        mbar_wait(&full[s], (ring_phase >> s) & 1u);
        ring_phase ^= (1u << s);
      } else {
        const int z = k0 + kk;
        load_box_manual(P, A.p1, b_cur, z, x0 - XO, y0 - 1, BX, L::BY, stage_p1(s), L::NT);
        if (kk >= 0 && kk < DZ) load_box_manual(P, A.p0, b_cur, z, x0, y0, TX, TY, stage_p0(s), L::NT);
        __syncthreads();
      }
      // G of the arriving plane: own cell + halo ring
      float* Gb = Gs + (j % 3) * (L::BOX1_B / 4);
      const float p_next = stage_p1(s)[(ty + 1) * BX + tx + XO];
      float A1_next, G_next, dummy1, dummy2, dA1_next;
      int kA1_next;
      cf_pvt_p1(P, Ts, p_next, A1_next, G_next, dummy1, dummy2, kA1_next, dA1_next);
      Gb[(ty + 1) * BX + tx + XO] = G_next;
      {
        int hx, hy;
        if (halo_cell<TY>(tid, hx, hy)) Gb[hy * BX + hx] = cf_G(P, Ts, stage_p1(s)[hy * BX + hx]);
      }
      __syncthreads();
      if (TMA && tid == 0 && j >= 2 && j - 2 + S < J) issue(j - 2 + S);   // stage (j-2)%S is free now
      if (flush_pending && tid == 0) {   // per-sample material-balance partial of the previous sample
        float t = 0.f;
        for (int w = 0; w < TY; ++w) t += red[w];
        atomicAdd(&A.mb_sum[flush_b], (double)(flush_k * t));
      }
      flush_pending = false;
      // stencil for plane m = kk-1
      const int m = kk - 1;
      if (m >= 0 && m < DZ) {
        const int sp1 = (j - 1) % S;
        const float* b1 = stage_p1(sp1);
        const float* Gm = Gs + ((j - 1) % 3) * (L::BOX1_B / 4);
        const int c = (ty + 1) * BX + tx + XO;
        const float pc = p_cur, Gc = G_cur;
        float flux = TE[(m * TY + ty) * 33 + tx] * (Gc + Gm[c - 1]) * (pc - b1[c - 1]);
        flux = fmaf(TE[(m * TY + ty) * 33 + tx + 1] * (Gc + Gm[c + 1]), pc - b1[c + 1], flux);
        flux = fmaf(TN[(m * (TY + 1) + ty) * TX + tx] * (Gc + Gm[c - BX]), pc - b1[c - BX], flux);
        flux = fmaf(TN[(m * (TY + 1) + ty + 1) * TX + tx] * (Gc + Gm[c + BX]), pc - b1[c + BX], flux);
        flux = fmaf(TU[(m * TY + ty) * TX + tx] * (Gc + G_prev), pc - p_prev, flux);
        flux = fmaf(TU[((m + 1) * TY + ty) * TX + tx] * (Gc + G_next), pc - p_next, flux);
        const int gz = k0 + m;
        if (in_xy && gz < P.D) {
          const float p0c = stage_p0(sp1)[ty * TX + tx];
          float A0, A0p, pass0, dA0;
          int kA0;
          cf_pvt_p0(P, Ts, p0c, A0, A0p, pass0, kA0, dA0);
          const float cp = P.Sgi * fmaf(P.phi, A0p, P.phicf * A0);
          const float acc = cA * cp * (pc - p0c);
          const float tde = cT * cp;
          float divq = P.dv * flux;
          const int cell = (gz * P.H + gy) * P.W + gx;
          if (has_well) {
            float q = 0.f, mask = 0.f;
            const int first = well_lower_bound(P, cell);
            for (int w = first; w < P.n_wells && P.wells[w].cell == cell; ++w) { q += A.qw[(int64_t)b_cur * P.n_wells + w]; mask += 1.f; }
            divq += q;
            if (mask != 0.f) {
              for (int w = first; w < P.n_wells && P.wells[w].cell == cell; ++w) A.divqw[(int64_t)b_cur * P.n_wells + w] = divq;
              const float ibc = mask * divq;
              acc_ibc += (double)ibc * (double)ibc;
            }
          }
          const float dom = divq + acc + (P.tde_in_dom ? tde : 0.f);
          const int64_t g = (int64_t)b_cur * P.N + cell;
          __stcs(&A.dom[g], dom);
          if (A.dom_out) __stcs(&A.dom_out[g], dom);
          acc_dom = fmaf(dom, dom, acc_dom);
          acc_tde = fmaf(tde, tde, acc_tde);
          mb_part += cf_dA(Ts, kA1_cur, dA1_cur, kA0, dA0);
        }
      }
      // rotate the z window
      p_prev = p_cur; G_prev = G_cur;
      p_cur = p_next; G_cur = G_next; dA1_cur = dA1_next; kA1_cur = kA1_next;
      if (kk == DZ) {   // sample finished: publish its material-balance partial
        float v[1] = {mb_part};
        warp_sum<1>(v);
        if (tx == 0) red[ty] = v[0];
        flush_pending = true;
        flush_b = b_cur;
        flush_k = mbk;
      }
    }
    __syncthreads();
    if (flush_pending && tid == 0) {
      float t = 0.f;
      for (int w = 0; w < TY; ++w) t += red[w];
      atomicAdd(&A.mb_sum[flush_b], (double)(flush_k * t));
    }
    __syncthreads();
  }
  // CTA-wide reduction of the squared-error partials
  {
    double v[3] = {(double)acc_dom, (double)acc_tde, acc_ibc};
    __shared__ double dred[3 * 32];
    block_reduce<3>(v, dred);
    if (tid == 0) {
      atomicAdd(&A.sse[SRM_TERM_DOM], v[0]);
      atomicAdd(&A.sse[SRM_TERM_TDE], v[1]);
      atomicAdd(&A.sse[SRM_TERM_IBC], v[2]);
    }
  }
}

// sum of well rates per sample (the forward's mbc needs -sum q)
__global__ void k_qsum_cf(int32_t B, int32_t nw, const float* __restrict__ qw, double* __restrict__ q_sum) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double s = 0.0;
  for (int w = 0; w < nw; ++w) s += (double)qw[(int64_t)b * nw + w];
  q_sum[b] = s;
}

// ---------------------------------------------------------------- adjoint
// gather form; boundary faces carry zero coefficient.  Per cell (with d2A/dp2 = 0 and N = 0):
//   gp1 = dv * sum_f (s_c - s_n) * [T_f (G_c+G_n) + T_f G'_c (p_c - p_n)] + s_c (dq + cA cp) + smb (-dq - mbk A1')
//   gp0 = s_c * cA * (cpp (p1-p0) - cp) + st cT cpp + smb mbk A0' pass0,      cpp = Sgi phicf A0' pass0
//   gdt1 += -(s_c acc + st tde)/dt1 + smb mb/dt1
template <int TY, int DZ, int S, bool TMA>
__global__ void __launch_bounds__(TX * TY, 1) k_adj_cf(const __grid_constant__ SrmDev P, const __grid_constant__ CfArgs A,
                                                       const __grid_constant__ CUtensorMap map_p1,
                                                       const __grid_constant__ CUtensorMap map_p0,
                                                       const __grid_constant__ CUtensorMap map_dom) {
  using L = Lay<TY, DZ, S>;
  extern __shared__ __align__(1024) unsigned char smem[];
  using O = typename L::template Off<2>;
  auto stage_p1 = [&](int s) { return (float*)(smem + s * O::STAGE); };
  auto stage_dm = [&](int s) { return (float*)(smem + s * O::STAGE + L::BOX1_B); };
  auto stage_p0 = [&](int s) { return (float*)(smem + s * O::STAGE + 2 * L::BOX1_B); };
  float* Gs = (float*)(smem + O::GS);
  float* TE = (float*)(smem + O::TE);
  float* TN = (float*)(smem + O::TN);
  float* TU = (float*)(smem + O::TU);
  SrmClosedForm* Ts = (SrmClosedForm*)(smem + O::TAB);
  uint64_t* full = (uint64_t*)(smem + O::FULL);
  unsigned char* colflag = smem + O::COL;
  float* red = (float*)(smem + O::RED);
  __shared__ int s_item;

  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  for (int e = tid; e < (int)(sizeof(SrmClosedForm) / 4); e += L::NT) ((uint32_t*)Ts)[e] = ((const uint32_t*)A.T)[e];
  if (TMA && tid == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const float w_dom = A.dterms[SRM_TERM_DOM], w_mbc = A.dterms[SRM_TERM_MBC], w_tde = A.dterms[SRM_TERM_TDE];
  const float sd = 2.f * w_dom;
  const int nseg = A.ctl[0];
  const int tiles = A.tiles_x * A.tiles_y;
  const int total = nseg * A.zchunks * tiles;
  uint32_t ring_phase = 0;

  for (;;) {
    if (tid == 0) s_item = atomicAdd(&A.ctl[A.ctl_slot], 1);
    __syncthreads();
    const int item = s_item;
    if (item >= total) break;
    const int tile = item % tiles, zc = (item / tiles) % A.zchunks, sg = item / (tiles * A.zchunks);
    const int x0 = (tile % A.tiles_x) * TX, y0 = (tile / A.tiles_x) * TY, k0 = zc * DZ;
    const int r = A.seg[3 * sg], loff = A.seg[3 * sg + 1], nT = A.seg[3 * sg + 2];
    build_statics<TY, DZ>(P, A.kx + (int64_t)r * P.N, x0, y0, k0, TE, TN, TU);
    colflag[tid] = 0;
    __syncthreads();
    for (int w = tid; w < P.n_wells; w += L::NT) {
      const int c = P.wells[w].cell, wi = c % P.W, wj = (c / P.W) % P.H, wk = c / (P.W * P.H);
      if (wi >= x0 && wi < x0 + TX && wj >= y0 && wj < y0 + TY && wk >= k0 && wk < k0 + DZ) colflag[(wj - y0) * TX + (wi - x0)] = 1;
    }
    __syncthreads();
    const bool has_well = colflag[tid] != 0;
    const int gx = x0 + tx, gy = y0 + ty;
    const bool in_xy = (gx < P.W) && (gy < P.H);
    const int J = nT * (DZ + 2);

    auto issue = [&](int j) {
      const int s = j % S, ti = j / (DZ + 2), kk = j % (DZ + 2) - 1;
      const int b = A.list[loff + ti], z = k0 + kk;
      const bool want0 = (kk >= 0 && kk < DZ);
      mbar_expect_tx(&full[s], 2 * L::BOX1 * 4 + (want0 ? L::BOX0 * 4 : 0));
      tma_load_4d(stage_p1(s), &map_p1, &full[s], x0 - XO, y0 - 1, z, b);
      tma_load_4d(stage_dm(s), &map_dom, &full[s], x0 - XO, y0 - 1, z, b);
      if (want0) tma_load_4d(stage_p0(s), &map_p0, &full[s], x0, y0, z, b);
    };
    if (TMA) {
      if (tid == 0) for (int j = 0; j < min(S, J); ++j) issue(j);
    }
    float p_prev = 0.f, G_prev = 0.f, d_prev = 0.f;
    float p_cur = 0.f, G_cur = 0.f, d_cur = 0.f, dA1_cur = 0.f, A1p_cur = 0.f, Gp_cur = 0.f;
    int kA1_cur = 0;
    float cA = 0.f, cT = 0.f, mbk = 0.f, inv_d1 = 0.f, smb = 0.f, g1_part = 0.f;
    int b_cur = 0;
    bool flush_pending = false;
    int flush_b = 0;

    for (int j = 0; j < J; ++j) {
      const int s = j % S, ti = j / (DZ + 2), kk = j % (DZ + 2) - 1;
      if (kk == -1) {
        b_cur = A.list[loff + ti];
        const float d1 = A.dt1[b_cur];
        inv_d1 = 1.f / d1;
        cA = P.dv * P.invDc * inv_d1;
        cT = P.dvDc * 2e-7f * inv_d1;
        mbk = P.dvSgi_phi / (P.Dc * d1);
        smb = 2.f * w_mbc * A.mbc[b_cur];
        g1_part = 0.f;
      }
      if (TMA) {
        mbar_wait(&full[s], (ring_phase >> s) & 1u);
        ring_phase ^= (1u << s);
      } else {
        const int z = k0 + kk;
        load_box_manual(P, A.p1, b_cur, z, x0 - XO, y0 - 1, BX, L::BY, stage_p1(s), L::NT);
        load_box_manual(P, A.dom, b_cur, z, x0 - XO, y0 - 1, BX, L::BY, stage_dm(s), L::NT);
        if (kk >= 0 && kk < DZ) load_box_manual(P, A.p0, b_cur, z, x0, y0, TX, TY, stage_p0(s), L::NT);
        __syncthreads();
      }
      float* Gb = Gs + (j % 3) * (L::BOX1_B / 4);
      const int c = (ty + 1) * BX + tx + XO;
      const float p_next = stage_p1(s)[c];
      const float d_next = stage_dm(s)[c];
      float A1_next, G_next, A1p_next, Gp_next, dA1_next;
      int kA1_next;
      cf_pvt_p1(P, Ts, p_next, A1_next, G_next, A1p_next, Gp_next, kA1_next, dA1_next);
      Gb[c] = G_next;
      {
        int hx, hy;
        if (halo_cell<TY>(tid, hx, hy)) Gb[hy * BX + hx] = cf_G(P, Ts, stage_p1(s)[hy * BX + hx]);
      }
      __syncthreads();
      if (TMA && tid == 0 && j >= 2 && j - 2 + S < J) issue(j - 2 + S);
      if (flush_pending && tid == 0) {
        float t = 0.f;
        for (int w = 0; w < TY; ++w) t += red[w];
        atomicAdd(&A.gdt1_acc[flush_b], (double)t);
      }
      flush_pending = false;
      const int m = kk - 1;
      if (m >= 0 && m < DZ) {
        const int sp1 = (j - 1) % S;
        const float* b1 = stage_p1(sp1);
        const float* dm = stage_dm(sp1);
        const float* Gm = Gs + ((j - 1) % 3) * (L::BOX1_B / 4);
        const float pc = p_cur, Gc = G_cur, dc = d_cur, Gpc = Gp_cur;
        float g1 = 0.f;
        auto face = [&](float Tf, float Gn, float pn, float dn) {
          const float dp = pc - pn;
          g1 = fmaf((dc - dn) * Tf, fmaf(Gpc, dp, Gc + Gn), g1);
        };
        face(TE[(m * TY + ty) * 33 + tx], Gm[c - 1], b1[c - 1], dm[c - 1]);
        face(TE[(m * TY + ty) * 33 + tx + 1], Gm[c + 1], b1[c + 1], dm[c + 1]);
        face(TN[(m * (TY + 1) + ty) * TX + tx], Gm[c - BX], b1[c - BX], dm[c - BX]);
        face(TN[(m * (TY + 1) + ty + 1) * TX + tx], Gm[c + BX], b1[c + BX], dm[c + BX]);
        face(TU[(m * TY + ty) * TX + tx], G_prev, p_prev, d_prev);
        face(TU[((m + 1) * TY + ty) * TX + tx], G_next, p_next, d_next);
        const int gz = k0 + m;
        if (in_xy && gz < P.D) {
          g1 *= sd * P.dv;
          const float sc = sd * dc;
          const float p0c = stage_p0(sp1)[ty * TX + tx];
          float A0, A0p, pass0, dA0;
          int kA0;
          cf_pvt_p0(P, Ts, p0c, A0, A0p, pass0, kA0, dA0);
          const float cp = P.Sgi * fmaf(P.phi, A0p, P.phicf * A0);
          const float cpp = P.Sgi * P.phicf * A0p * pass0;
          const float dp10 = pc - p0c;
          const float acc = cA * cp * dp10;
          const float tde = cT * cp;
          const float st = (P.tde_in_dom ? sc : 0.f) + 2.f * w_tde * tde;
          const int cell = (gz * P.H + gy) * P.W + gx;
          float dq = 0.f;
          if (has_well) {
            const int first = well_lower_bound(P, cell);
            for (int w = first; w < P.n_wells && P.wells[w].cell == cell; ++w) dq += A.dqdp[(int64_t)b_cur * P.n_wells + w];
          }
          g1 += sc * (dq + cA * cp) + smb * (-dq - mbk * A1p_cur);
          const float g0 = sc * cA * (cpp * dp10 - cp) + st * cT * cpp + smb * mbk * A0p * pass0;
          const int64_t g = (int64_t)b_cur * P.N + cell;
          __stcs(&A.gp0[g], g0);
          __stcs(&A.gp1[g], g1);
          g1_part += (-(sc * acc + st * tde) + smb * mbk * cf_dA(Ts, kA1_cur, dA1_cur, kA0, dA0)) * inv_d1;
        }
      }
      p_prev = p_cur; G_prev = G_cur; d_prev = d_cur;
      p_cur = p_next; G_cur = G_next; d_cur = d_next; dA1_cur = dA1_next; kA1_cur = kA1_next; A1p_cur = A1p_next; Gp_cur = Gp_next;
      if (kk == DZ) {
        float v[1] = {g1_part};
        warp_sum<1>(v);
        if (tx == 0) red[ty] = v[0];
        flush_pending = true;
        flush_b = b_cur;
      }
    }
    __syncthreads();
    if (flush_pending && tid == 0) {
      float t = 0.f;
      for (int w = 0; w < TY; ++w) t += red[w];
      atomicAdd(&A.gdt1_acc[flush_b], (double)t);
    }
    __syncthreads();
  }
}

// inner-boundary (well-cell) part of the adjoint, closed form (cf. k_ibc_adj_ref)
__global__ void __launch_bounds__(128) k_ibc_adj_cf(const __grid_constant__ SrmDev P, const SrmClosedForm* __restrict__ T,
                                                    int32_t B, int32_t R, const float* __restrict__ kx,
                                                    const int32_t* __restrict__ sample_real, const float* __restrict__ p1f,
                                                    const float* __restrict__ dterms, const float* __restrict__ divqw,
                                                    const float* __restrict__ dqdp, float* __restrict__ gp1) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int nw = P.n_wells;
  if (g >= (int64_t)B * nw) return;
  const int b = (int)(g / nw), w = (int)(g % nw);
  const int c = P.wells[w].cell;
  if (w > 0 && P.wells[w - 1].cell == c) return;
  float mask = 0.f, dq = 0.f;
  for (int u = w; u < nw && P.wells[u].cell == c; ++u) { mask += 1.f; dq += dqdp[(int64_t)b * nw + u]; }
  const float s = 2.f * dterms[SRM_TERM_IBC] * mask * mask * divqw[g];
  if (s == 0.f) return;
  const int r = srm_real_of(sample_real, b, B, R);
  const int64_t base = (int64_t)b * P.N;
  const float* kr = kx + (int64_t)r * P.N;
  const int i = c % P.W, j = (c / P.W) % P.H, k = c / (P.W * P.H), HW = P.H * P.W;
  float Ac, Gc, Apc, Gpc, dAc;
  int kAc;
  const float pc = p1f[base + c];
  cf_pvt_p1(P, T, pc, Ac, Gc, Apc, Gpc, kAc, dAc);
  const float kc = kr[c];
  float self = 0.f;
  auto hm = [](float a, float bb) { return 2.f * a * bb / (a + bb); };
  auto face = [&](bool ok, int cn, float ratio, float idl) {
    if (!ok) return;
    const float Tf = 0.5f * P.C * P.krg * idl * idl * hm(ratio * kc, ratio * kr[cn]);
    const float pn = p1f[base + cn];
    float An, Gn, Apn, Gpn, dAn;
    int kAn;
    cf_pvt_p1(P, T, pn, An, Gn, Apn, Gpn, kAn, dAn);
    const float af = Tf * (Gc + Gn);
    self += af + Tf * Gpc * (pc - pn);
    atomicAdd(&gp1[base + cn], s * P.dv * (-af + Tf * Gpn * (pc - pn)));
  };
  face(i > 0, c - 1, 1.f, P.idx);
  face(i < P.W - 1, c + 1, 1.f, P.idx);
  face(j > 0, c - P.W, P.kx_ky, P.idy);
  face(j < P.H - 1, c + P.W, P.kx_ky, P.idy);
  face(k > 0, c - HW, P.kv_kh, P.idz);
  face(k < P.D - 1, c + HW, P.kv_kh, P.idz);
  atomicAdd(&gp1[base + c], s * (P.dv * self + dq));
}

// ------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 4-D map over a (B, D, H, W) fp32 field, box = bx x by x 1 x 1
bool make_map(CUtensorMap* m, const float* base, int W, int H, int D, int B, int bx, int by) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * D * 4};
  cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)by, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int TY, int DZ, int S>
size_t smem_bytes(bool adj) {
  using L = Lay<TY, DZ, S>;
  return adj ? (size_t)L::template Off<2>::TOTAL : (size_t)L::template Off<1>::TOTAL;
}

struct Plan { int ty, dz, stages; bool tma; };

Plan choose_plan(const SrmDev& P, const void* a, const void* b, const void* c) {
  Plan pl;
  pl.dz = (P.D >= 16) ? 16 : (P.D >= 8 ? 8 : (P.D >= 4 ? 4 : 1));
  pl.ty = 8;
  pl.stages = 8;
  auto al16 = [](const void* p) { return p == nullptr || (((uintptr_t)p) & 15) == 0; };
  pl.tma = (P.W % 4 == 0) && al16(a) && al16(b) && al16(c) && get_encode() != nullptr && !getenv("SRM_NO_TMA");
  return pl;
}

template <int TY, int DZ, int S, bool TMA>
int launch_fwd(SrmHandle* h, const CfArgs& A, const CUtensorMap& m1, const CUtensorMap& m0, cudaStream_t s) {
  const size_t sm = smem_bytes<TY, DZ, S>(false);
  auto kern = k_fwd_cf<TY, DZ, S, TMA>;
  SRM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  int occ = 1;
  SRM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TX * TY, sm));
  if (occ < 1) { srm_set_error("k_fwd_cf does not fit on an SM (%zu B smem)", sm); return SRM_ERR_CUDA; }
  kern<<<h->sm_count * occ, TX * TY, sm, s>>>(h->dev, A, m1, m0);
  SRM_CUDA_CHECK(cudaGetLastError());
  return SRM_OK;
}
template <int TY, int DZ, int S, bool TMA>
int launch_adj(SrmHandle* h, const CfArgs& A, const CUtensorMap& m1, const CUtensorMap& m0, const CUtensorMap& md, cudaStream_t s) {
  const size_t sm = smem_bytes<TY, DZ, S>(true);
  auto kern = k_adj_cf<TY, DZ, S, TMA>;
  SRM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  int occ = 1;
  SRM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TX * TY, sm));
  if (occ < 1) { srm_set_error("k_adj_cf does not fit on an SM (%zu B smem)", sm); return SRM_ERR_CUDA; }
  kern<<<h->sm_count * occ, TX * TY, sm, s>>>(h->dev, A, m1, m0, md);
  SRM_CUDA_CHECK(cudaGetLastError());
  return SRM_OK;
}

#define SRM_DISPATCH_DZ(FN, TMAFLAG, ...)                                                   \
  switch (pl.dz) {                                                                          \
    case 16: rc = FN<8, 16, 8, TMAFLAG>(__VA_ARGS__); break;                                \
    case 8: rc = FN<8, 8, 8, TMAFLAG>(__VA_ARGS__); break;                                  \
    case 4: rc = FN<8, 4, 8, TMAFLAG>(__VA_ARGS__); break;                                  \
    default: rc = FN<8, 1, 8, TMAFLAG>(__VA_ARGS__); break;                                 \
  }

int tchunk_for(const SrmHandle* h, const Plan& pl, int B) {
  // enough items to keep every SM busy for several rounds, but long enough segments to amortise kx
  const SrmDev& P = h->dev;
  const int tiles = ((P.W + TX - 1) / TX) * ((P.H + pl.ty - 1) / pl.ty) * ((P.D + pl.dz - 1) / pl.dz);
  int t = 32;
  while (t > 1 && (int64_t)tiles * ((B + t - 1) / t) < 6 * (int64_t)h->sm_count) t >>= 1;
  return t;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
int srm_launch_pvt_eval_cf(const SrmHandle* h, int64_t n, const float* p, float* val, float* dval, cudaStream_t s) {
  if (n <= 0) return SRM_OK;
  if (!h->d_cf) { srm_set_error("closed-form tables missing"); return SRM_ERR_INVALID; }
  k_pvt_eval_cf<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(h->dev, h->d_cf, n, p, val, dval);
  SRM_CUDA_CHECK(cudaGetLastError());
  return SRM_OK;
}

int srm_launch_wells_cf(const SrmHandle* h, int32_t B, const float* kx, const int32_t* sample_real, int32_t R,
                        const float* p, const float* t_days, float* qw, float* pwfw, float* dqdp, cudaStream_t s) {
  const int64_t n = (int64_t)B * h->dev.n_wells;
  if (n <= 0) return SRM_OK;
  MobilityCf mob;
  mob.T = h->d_cf;
  k_wells<MobilityCf><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(h->dev, mob, B, R, kx, sample_real, p, t_days, qw, pwfw, dqdp);
  SRM_CUDA_CHECK(cudaGetLastError());
  return SRM_OK;
}

// the lean pair (kernels_cf2.cu)
bool srm_cf2_applicable(const SrmHandle* h, const void* a, const void* b, const void* c, const void* d, const void* e);
int srm_cf2_faces(const SrmHandle* h, int32_t R, const float* kx, float* faces, cudaStream_t s);
int srm_cf2_forward(const SrmHandle* h, int32_t B, int32_t R, const int32_t* sample_real, const float* p0, const float* p1,
                    const float* dt1, const SrmWs& ws, cudaStream_t s);
int srm_cf2_backward(const SrmHandle* h, int32_t B, int32_t R, const int32_t* sample_real, const float* p0, const float* p1,
                     const float* dt1, const float* dterms, float* gp0, float* gp1, float* gdt1, float* gdt2, const SrmWs& ws,
                     cudaStream_t s);

int srm_forward_cf(SrmHandle* h, int32_t B, int32_t R, const float* kx, const int32_t* sample_real,
                   const float* p0, const float* p1, const float* dt1, const float* dt2, const float* t1,
                   float* terms_out, float* dom_out, const SrmWs& ws, bool save, cudaStream_t s) {
  (void)save;
  const SrmDev& P = h->dev;
  SRM_CUDA_CHECK(cudaMemsetAsync(ws.sse, 0, (char*)ws.mbc - (char*)ws.sse, s));
  const Plan pl = choose_plan(P, p0, p1, nullptr);
  h->cf_faces_ok = h->cf_grouped = 0;
  int rc = srm_launch_wells_cf(h, B, kx, sample_real, R, p1, t1, ws.qw, ws.pwfw, ws.dqdp, s);
  if (rc) return rc;
  if (P.n_wells > 0) {
    k_qsum_cf<<<(B + 127) / 128, 128, 0, s>>>(B, P.n_wells, ws.qw, ws.q_sum);
    SRM_CUDA_CHECK(cudaGetLastError());
  }
  if (srm_cf2_applicable(h, p0, p1, ws.dom, ws.faces, nullptr)) {
    rc = srm_cf2_faces(h, R, kx, ws.faces, s);
    if (!rc) rc = srm_cf2_forward(h, B, R, sample_real, p0, p1, dt1, ws, s);
    if (rc) return rc;
    h->cf_faces_ok = 1;
    if (dom_out) SRM_CUDA_CHECK(cudaMemcpyAsync(dom_out, ws.dom, sizeof(float) * (size_t)B * (size_t)P.N, cudaMemcpyDeviceToDevice, s));
    k_finalize_fwd<<<1, 256, 0, s>>>(P, B, ws.sse, ws.mb_sum, ws.q_sum, ws.mbc, terms_out);
    SRM_CUDA_CHECK(cudaGetLastError());
    return SRM_OK;
  }
  k_group_samples<<<1, 1024, 0, s>>>(B, R, sample_real, tchunk_for(h, pl, B), ws.grp_cnt, ws.grp_fill, ws.grp_list, ws.seg, ws.ctl);
  SRM_CUDA_CHECK(cudaGetLastError());
  h->cf_grouped = 1;
  CfArgs A;
  std::memset(&A, 0, sizeof(A));
  A.p0 = p0; A.p1 = p1; A.kx = kx; A.dt1 = dt1; A.dt2 = dt2;
  A.list = ws.grp_list; A.seg = ws.seg; A.ctl = ws.ctl;
  A.qw = ws.qw; A.dqdp = ws.dqdp; A.divqw = ws.divqw;
  A.dom = ws.dom; A.dom_out = dom_out; A.sse = ws.sse; A.mb_sum = ws.mb_sum;
  A.T = h->d_cf; A.B = B; A.R = R;
  A.tiles_x = (P.W + TX - 1) / TX; A.tiles_y = (P.H + pl.ty - 1) / pl.ty; A.zchunks = (P.D + pl.dz - 1) / pl.dz;
  A.ctl_slot = 1;
  CUtensorMap m1, m0;
  std::memset(&m1, 0, sizeof(m1));
  std::memset(&m0, 0, sizeof(m0));
  bool tma = pl.tma;
  if (tma) tma = make_map(&m1, p1, P.W, P.H, P.D, B, BX, pl.ty + 2) && make_map(&m0, p0, P.W, P.H, P.D, B, TX, pl.ty);
  if (tma) { SRM_DISPATCH_DZ(launch_fwd, true, h, A, m1, m0, s); }
  else { SRM_DISPATCH_DZ(launch_fwd, false, h, A, m1, m0, s); }
  if (rc) return rc;
  k_finalize_fwd<<<1, 256, 0, s>>>(P, B, ws.sse, ws.mb_sum, ws.q_sum, ws.mbc, terms_out);
  SRM_CUDA_CHECK(cudaGetLastError());
  return SRM_OK;
}

int srm_backward_cf(SrmHandle* h, int32_t B, int32_t R, const float* kx, const int32_t* sample_real,
                    const float* p0, const float* p1, const float* dt1, const float* dt2, const float* t1,
                    const float* dterms, float* gp0, float* gp1, float* gdt1, float* gdt2,
                    const SrmWs& ws, cudaStream_t s) {
  (void)t1;
  const SrmDev& P = h->dev;
  SRM_CUDA_CHECK(cudaMemsetAsync(ws.gdt1_acc, 0, (char*)ws.mbc - (char*)ws.gdt1_acc, s));
  const Plan pl = choose_plan(P, p0, p1, ws.dom);
  if (srm_cf2_applicable(h, p0, p1, ws.dom, gp0, gp1) && srm_cf2_applicable(h, ws.faces, nullptr, nullptr, nullptr, nullptr)) {
    // the forward's kernel family left dom, mbc, the well tables -- and, if it was the lean one, the face planes
    int rc2 = h->cf_faces_ok ? SRM_OK : srm_cf2_faces(h, R, kx, ws.faces, s);
    if (!rc2) { h->cf_faces_ok = 1; rc2 = srm_cf2_backward(h, B, R, sample_real, p0, p1, dt1, dterms, gp0, gp1, gdt1, gdt2, ws, s); }
    if (rc2) return rc2;
    const int64_t nwb = (int64_t)B * P.n_wells;
    if (nwb > 0) {
      k_ibc_adj_cf<<<(unsigned)((nwb + 127) / 128), 128, 0, s>>>(P, h->d_cf, B, R, kx, sample_real, p1, dterms, ws.divqw, ws.dqdp, gp1);
      SRM_CUDA_CHECK(cudaGetLastError());
    }
    return SRM_OK;
  }
  if (!h->cf_grouped) {
    k_group_samples<<<1, 1024, 0, s>>>(B, R, sample_real, tchunk_for(h, pl, B), ws.grp_cnt, ws.grp_fill, ws.grp_list, ws.seg, ws.ctl);
    SRM_CUDA_CHECK(cudaGetLastError());
    h->cf_grouped = 1;
  }
  SRM_CUDA_CHECK(cudaMemsetAsync(ws.ctl + 2, 0, sizeof(int32_t), s));
  CfArgs A;
  std::memset(&A, 0, sizeof(A));
  A.p0 = p0; A.p1 = p1; A.kx = kx; A.dt1 = dt1; A.dt2 = dt2;
  A.list = ws.grp_list; A.seg = ws.seg; A.ctl = ws.ctl;
  A.qw = ws.qw; A.dqdp = ws.dqdp; A.divqw = ws.divqw;
  A.dom = ws.dom; A.dterms = dterms; A.mbc = ws.mbc; A.gp0 = gp0; A.gp1 = gp1; A.gdt1_acc = ws.gdt1_acc;
  A.T = h->d_cf; A.B = B; A.R = R;
  A.tiles_x = (P.W + TX - 1) / TX; A.tiles_y = (P.H + pl.ty - 1) / pl.ty; A.zchunks = (P.D + pl.dz - 1) / pl.dz;
  A.ctl_slot = 2;
  CUtensorMap m1, m0, md;
  std::memset(&m1, 0, sizeof(m1));
  std::memset(&m0, 0, sizeof(m0));
  std::memset(&md, 0, sizeof(md));
  bool tma = pl.tma;
  if (tma) tma = make_map(&m1, p1, P.W, P.H, P.D, B, BX, pl.ty + 2) && make_map(&m0, p0, P.W, P.H, P.D, B, TX, pl.ty) &&
                 make_map(&md, ws.dom, P.W, P.H, P.D, B, BX, pl.ty + 2);
  int rc;
  if (tma) { SRM_DISPATCH_DZ(launch_adj, true, h, A, m1, m0, md, s); }
  else { SRM_DISPATCH_DZ(launch_adj, false, h, A, m1, m0, md, s); }
  if (rc) return rc;
  const int64_t n = (int64_t)B * P.n_wells;
  if (n > 0) {
    k_ibc_adj_cf<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(P, h->d_cf, B, R, kx, sample_real, p1, dterms, ws.divqw, ws.dqdp, gp1);
    SRM_CUDA_CHECK(cudaGetLastError());
  }
  k_finalize_adj<<<(unsigned)((B + 255) / 256), 256, 0, s>>>(B, ws.gdt1_acc, ws.gdt2_acc, gdt1, gdt2);
  SRM_CUDA_CHECK(cudaGetLastError());
  return SRM_OK;
}
