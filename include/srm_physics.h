/*
 * srm_physics.h -- C ABI of the B200-native physics-loss library (libsrm_physics.so).
 *
 * The reference (molokwuvictor/3d-physics-based-ai-surrogate-reservoir-model) is pure
 * Python/TensorFlow and has NO foreign-function interface; this header is the boundary a
 * TensorFlow custom op (tf_op/srm_physics_op.cc) or any other host (ctypes, here) binds instead of
 * the reference's op-by-op graph.  Each entry point names the reference code it replaces
 * (file:line relative to the reference repo root).
 *
 * Conventions
 *   - All tensor pointers are DEVICE pointers on the handle's device unless marked "host".
 *   - Fields are fp32, layout (B, D, H, W) with W (x, index i) contiguous; kx is (R, D, H, W).
 *     "sample" b = one (realisation, time point) pair (training.py:187-204 flattens K x T into B).
 *   - sample_real[b] in [0, R) maps a sample to its permeability realisation (device int32;
 *     NULL means b * R / B, i.e. realisation-major equal-sized groups).  An index outside [0, R) is clamped
 *     into the range by the kernels (never an out-of-bounds read); dt1, dt2, t1, sample_real hold B elements.
 *   - The caller owns every buffer.  The library owns only the handle's immutable device tables.
 *     No allocation happens inside forward/backward: scratch comes from the caller's workspace
 *     (query srm_workspace_bytes once).
 *   - Every call returns 0 on success or a negative SrmStatus; the message is in
 *     srm_last_error() (thread-local).  Nothing throws across the ABI.  Launches are asynchronous
 *     on `stream` (a cudaStream_t passed as void*); asynchronous CUDA errors surface at the
 *     caller's next synchronisation.
 *   - A handle is read-only after creation: concurrent calls on different streams with different
 *     workspaces are safe.
 *   - There is NO CPU fallback.  If no CUDA device is usable srm_create fails with SRM_ERR_CUDA.
 */
#ifndef SRM_PHYSICS_H_
#define SRM_PHYSICS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRM_ABI_VERSION 5

typedef enum SrmStatus {
  SRM_OK = 0,
  SRM_ERR_INVALID = -1,   /* bad argument / unsupported configuration */
  SRM_ERR_CUDA = -2,      /* CUDA runtime error (message has the cudaError string) */
  SRM_ERR_WORKSPACE = -3, /* workspace too small */
  SRM_ERR_STATE = -4      /* backward without usable forward state and recompute disabled */
} SrmStatus;

/* loss-term slots of terms_out / dterms (default_configurations.py:63-83 key set) */
enum { SRM_ROOT_NEWTON = 0, SRM_ROOT_BRACKET = 1 };
enum { SRM_TERM_DOM = 0, SRM_TERM_IBC = 1, SRM_TERM_MBC = 2, SRM_TERM_TDE = 3,
       SRM_TERM_OBC = 4, SRM_TERM_IC = 5, SRM_TERM_TD = 6, SRM_TERM_CMBC = 7, SRM_N_TERMS = 8 };

enum { SRM_FLUID_DG = 0, SRM_FLUID_GC = 1 };
enum { SRM_PVT_SPLINE = 0, SRM_PVT_POLYNOMIAL = 1 };

/* Numerics of the PVT evaluation and the flux assembly.
 *   SRM_NUMERICS_REFERENCE : the reference's fp32 op order reproduced op by op (all 37 RBF terms in
 *       the x^2-2xc+c^2 form of polyhm_splines.py:91-96, a*p products of physics_loss.py:174, no FMA
 *       contraction).  Forward fields are bit-faithful to the pinned-order oracle.  Compute bound.
 *   SRM_NUMERICS_CLOSED_FORM : the same order-1 interpolant in its exact closed form (piecewise
 *       linear between knots) and the flux in difference form.  Closer to exact arithmetic than the
 *       fp32 reference itself, but not bit-faithful to its rounding noise.  HBM bound. */
enum { SRM_NUMERICS_REFERENCE = 0, SRM_NUMERICS_CLOSED_FORM = 1 };

/* forward/backward flags */
enum { SRM_FLAG_SAVE_FOR_BACKWARD = 1 };

/* One well connection (default_configurations.py:132-140; welldata_processor.py:39-107).
 * (i,j,k) are 0-based cell indices along (W,H,D); the library scatters at [k,j,i]
 * (welldata_processor.py:26-40).  q_target carries the sign rule (producer +, injector -). */
typedef struct SrmWell {
  int32_t i, j, k;
  float q_target;
  float pwf_min;
  float rw;
  float hc;          /* completion ratio */
  float shut_start;  /* shut in while shut_start <= t <= shut_stop (welldata_processor.py:349-354) */
  float shut_stop;
} SrmWell;

typedef struct SrmConfig {
  int32_t abi_version;   /* = SRM_ABI_VERSION */
  int32_t device;        /* CUDA device ordinal */
  /* grid (default_configurations.py:92-104) */
  int32_t D, H, W;
  float dx, dy, dz;
  /* unit constants (default_configurations.py:449-451) */
  float C, Dc;
  /* rock / SCAL scalars, host-computed in fp32 the way the reference does
   * (physics_loss.py:64-65,129; relative_permeability.py:49-75) */
  float phi, cf, Sgi, krg;
  float kx_ky, kv_kh;
  int32_t fluid_type;    /* SRM_FLUID_* */
  /* PVT (PVT_Layer_Subclassed.py:23-216; polyhm_splines.py) -- host pointers, copied at create */
  int32_t pvt_method;    /* SRM_PVT_*.  SRM_PVT_POLYNOMIAL (PVTLayer.evaluate_polynomial, PVT_Layer_Subclassed.py:218-266):
                          * n_knots = number of coefficients, spline_w[q][i] = coefficient a_i of property q
                          * (value = sum a_i p^i), knots / spline_v / spline_order unused */
  int32_t spline_order;  /* 1 (example) or 2 (default config) */
  int32_t n_knots;       /* <= 64 */
  int32_t n_props;       /* DG: 2 [invBg, invug] */
  const float* knots;    /* host [n_knots], ascending */
  const float* spline_w; /* host [n_props][n_knots]  RBF weights */
  const float* spline_v; /* host [n_props][2]        linear term */
  float p_min, p_max;    /* clamp (PVT_Layer_Subclassed.py:165-167) */
  /* wells -- host pointer, copied at create */
  int32_t n_wells;
  const SrmWell* wells;
  int32_t use_blocking_factor; /* well_rate_bhp_Subclassed.py:36 */
  int32_t n_intervals;         /* well_rate_bhp_Subclassed.py:39 */
  /* behaviour */
  int32_t numerics;      /* SRM_NUMERICS_* */
  int32_t tde_in_dom;    /* 1: legacy DG folds the truncation term into dom (physics_loss.py:175) */
  /* Exact PVT tabulation (SRM_NUMERICS_REFERENCE only).  The reference re-evaluates the 37-term
   * spline (polyhm_splines.py:138-146) and its tape derivatives (PVT_Layer_Subclassed.py:196-201) for
   * every cell of every call, although they are a pure function of ONE clamped fp32 pressure.
   * pvt_lut = 1 evaluates that function once, at create, for EVERY fp32 value of
   * [lut_p_lo, lut_p_hi] (48 bytes per value dry gas, 112-152 gas condensate; lo >= hi means the whole
   * clamp range [p_min, p_max], 78.7 M values for [14.7, 10000]) with the same reference-order code, and
   * the kernels index the table by the pressure's bit pattern.  Results are bit-identical to pvt_lut = 0;
   * pressures outside the tabulated range are evaluated directly.  A table over the whole clamp range
   * selects the fused kernels (dry gas: kernels_dg4.cu; gas condensate: gc_fused.cuh, whose workspace
   * then holds no staged fields). */
  int32_t pvt_lut;
  float lut_p_lo, lut_p_hi;
  /* SCAL end points and Corey exponents (relative_permeability.py:19-45; default_configurations.py:262-266).
   * Used by the gas-condensate path (SRM_FLUID_GC) and by srm_relperm; the dry-gas path only needs the
   * host-computed scalar krg above. */
  float Swmin, Sorg, Sgc, Socr, kro_Somax, krg_Sorg, krg_Swmin, nog, ng;
  /* root finder of the gas-condensate blocking-factor integral (well_rate_bhp_Subclassed.py:236-324, 909-911):
   * SRM_ROOT_NEWTON = _solve_newton (start 0.1, clip to [0, 1 - Swmin]), SRM_ROOT_BRACKET = _solve_chandrupatla
   * (regula-falsi bracket update, tol 1e-6); n_root_iter iterations per trapezoid node (reference default 20) */
  int32_t root_solver, n_root_iter;
  /* bottom-hole-pressure control (well_rate_bhp_Subclassed.py:41-44, 813-822): bhp_iterative = 0 is the reference's
   * default, _non_iterative_method (:614-724); 1 selects _iterative_method (:515-612) -- Newton-Raphson on the
   * bottom-hole pressure from min_bhp + (p - min_bhp)/2, one-sided difference quotient with eps = 14.7 psi, clip into
   * [min_bhp, p] after every step, at most bhp_max_iters steps (reference default 10) while |qg - q_target| > bhp_tol
   * (default 1e-6).  The reference stops the WHOLE batch together; every connection that has met its target is a fixed
   * point of the step (value and gradient), so the kernels iterate per connection.  d rate / d p (and d / d Sg) is carried
   * through every iteration, as tf.while_loop's gradient does. */
  int32_t bhp_iterative, bhp_max_iters;
  float bhp_tol;
} SrmConfig;

typedef struct SrmHandle SrmHandle;

/* library / error plumbing (no reference counterpart) */
int srm_version(void);
const char* srm_last_error(void);

/* Build the immutable device tables.  Replaces the per-call constant work of the reference:
 * the 39x39 spline solve redone inside every call (polyhm_splines.py:180 -- here solved once by
 * the host and passed in), scatter_nd of well scalars (well_rate_bhp_Subclassed.py:128-132),
 * Peaceman/relperm constants (physics_loss.py:61-77). */
int srm_create(const SrmConfig* cfg, SrmHandle** out);
void srm_destroy(SrmHandle* h);

/* Scratch needed by srm_forward/srm_backward for B samples of R realisations (bytes). */
size_t srm_workspace_bytes(const SrmHandle* h, int32_t B, int32_t R, int32_t flags);

/* PVTLayer.call (PVT_Layer_Subclassed.py:146-216) + PolyharmonicSplineInterpolationLayer.call
 * (polyhm_splines.py:152-196): clamp, value and d/dp of every property at n pressures.
 * val/dval are [n_props][n] (either may be NULL). */
int srm_pvt_eval(const SrmHandle* h, int64_t n, const float* p, float* val, float* dval, void* stream);

/* DataSummary.nonormalize, log branch used for permeability
 * (data_processing/data_processing_utils.py:1098-1106): kx = exp(ln(kmax/kmin)*((x-lo)/(hi-lo)) + ln kmin). */
int srm_denormalize_log(int32_t device, int64_t n, const float* x_norm, float kmin, float kmax, float lo, float hi,
                        float* out, void* stream);

/* Self-test of the library's own correctly rounded sqrt / division sequences (they replace the
 * slower IEEE intrinsics inside the reference-order spline): n pseudo-random operands from the
 * spline's operand ranges; mismatches[0..2] (host) = #sqrt, #division, #chained-division results
 * whose bits differ from sqrt.rn / div.rn.  Synchronises the stream.  No reference counterpart. */
int srm_selftest_rounding(int32_t device, int64_t n, uint64_t seed, int64_t* mismatches, void* stream);

/* WellRatesPressure.compute_rates_and_bhp (well_rate_bhp_Subclassed.py:727-837) incl.
 * _non_iterative_method (:614-724) or _iterative_method (:515-612, SrmConfig.bhp_iterative), _compute_phase_rates (:963-1007),
 * compute_blocking_integral_and_factor (:840-960), and the integer bookkeeping of
 * WellDataProcessor.scatter_y / conn_shutins_idx (welldata_processor.py:170-224,228-389),
 * evaluated sparsely at the connection cells.
 * p: (B,D,H,W) pressure; t_days: (B,) time used for the shut-in test.
 * Outputs (any may be NULL): qw,pwfw,dqdp [B][n_wells] per-connection tables;
 * q_dense,pwf_dense (B,D,H,W) zero off-well (scatter_nd semantics: duplicates sum). */
int srm_wells(const SrmHandle* h, int32_t B, int32_t R, const float* kx, const int32_t* sample_real,
              const float* p, const float* t_days, float* qw, float* pwfw, float* dqdp,
              float* q_dense, float* pwf_dense, void* stream);

/* physics_error_gas (physics_loss.py:79-208) + the SSE/count reduction of pinn_batch_sse_grad
 * (physics_loss.py:787-846), given the networks' outputs.
 *   p0,p1   (B,D,H,W) pressure at t_n and t_n+dt1         (physics_loss.py:88-95,111-115)
 *   dt1,dt2 (B,)      per-sample mean of the dt field      (physics_loss.py:102,122)
 *   t1      (B,)      time (days) at level n+1, for shut-ins
 * Outputs:
 *   terms_out [2][SRM_N_TERMS] fp32: row 0 = sum of squares per term, row 1 = element counts
 *   dom_out   (B,D,H,W) residual field, nullable
 *   qw_out, pwfw_out [B][n_wells], nullable
 */
int srm_forward(SrmHandle* h, int32_t B, int32_t R, const float* kx, const int32_t* sample_real,
                const float* p0, const float* p1, const float* dt1, const float* dt2, const float* t1,
                float* terms_out, float* dom_out, float* qw_out, float* pwfw_out,
                void* workspace, size_t workspace_bytes, int32_t flags, void* stream);

/* The gradient TF's tape delivers to the networks' outputs (physics_loss.py:849-859) for the loss
 * L = sum_k dterms[k] * SSE_k:  gp0,gp1 (B,D,H,W), gdt1,gdt2 (B,).  dterms is a DEVICE fp32[8].
 * If the workspace still holds the state of an srm_forward call with SRM_FLAG_SAVE_FOR_BACKWARD on
 * the same inputs it is reused; otherwise the forward state is recomputed first. */
int srm_backward(SrmHandle* h, int32_t B, int32_t R, const float* kx, const int32_t* sample_real,
                 const float* p0, const float* p1, const float* dt1, const float* dt2, const float* t1,
                 const float* dterms, float* gp0, float* gp1, float* gdt1, float* gdt2,
                 void* workspace, size_t workspace_bytes, int32_t flags, void* stream);

/* RelativePermeability.compute_krog_krgo (relative_permeability.py:49-75): Corey gas-oil relative
 * permeabilities with the end-point rules, and their d/dSg as TF's tape routes it (zero where a
 * where/min/max branch holds the value).  Any output may be NULL. */
int srm_relperm(const SrmHandle* h, int64_t n, const float* sg, float* krog, float* krgo, float* dkrog,
                float* dkrgo, void* stream);

/* physics_error_gas_oil (physics_loss.py:319-693) + the SSE reduction of pinn_batch_sse_grad, for a handle
 * created with fluid_type = SRM_FLUID_GC (n_props = 7: InvBg, InvBo, Invug, Invuo, Rs, Rv, Vro;
 * PVT_Layer_Subclassed.py:71-72).  Inputs are the networks' outputs at both time levels: pressure, gas
 * and oil saturation (physics_loss.py:330-332,372-374).  Well rates come from the GC branch of
 * WellRatesPressure (non-iterative or iterative control, _split_condensate_components;
 * well_rate_bhp_Subclassed.py:515-724,963-1034), with use_blocking_factor the blocking-factor integral and its root
 * finders (:857-950, 236-324; SrmConfig.root_solver, n_root_iter).
 *   terms_out [2][SRM_N_TERMS]: slots DOM, IBC, MBC and CMBC (= the truncation term, physics_loss.py:680)
 *   q4w_out   [4][B][n_wells]: qgg, qgo, qoo, qog per connection (nullable);  pwfw_out [B][n_wells] */
int srm_forward_gc(SrmHandle* h, int32_t B, int32_t R, const float* kx, const int32_t* sample_real,
                   const float* p0, const float* p1, const float* sg0, const float* sg1, const float* so0,
                   const float* so1, const float* dt1, const float* dt2, const float* t1, float* terms_out,
                   float* dom_out, float* q4w_out, float* pwfw_out, void* workspace, size_t workspace_bytes,
                   int32_t flags, void* stream);

/* Gradient of L = sum_k dterms[k] * SSE_k w.r.t. the eight differentiable inputs of srm_forward_gc
 * (what tape.gradient delivers to the networks, physics_loss.py:849-859).  Upstream selects of the
 * relative permeabilities carry no gradient (tf.cast, physics_loss.py:543-551). */
int srm_backward_gc(SrmHandle* h, int32_t B, int32_t R, const float* kx, const int32_t* sample_real,
                    const float* p0, const float* p1, const float* sg0, const float* sg1, const float* so0,
                    const float* so1, const float* dt1, const float* dt2, const float* t1, const float* dterms,
                    float* gp0, float* gp1, float* gsg0, float* gsg1, float* gso0, float* gso1, float* gdt1,
                    float* gdt2, void* workspace, size_t workspace_bytes, int32_t flags, void* stream);

/* ---- the element-wise glue either side of the physics kernels (both time levels in one pass) --------------------
 * HardLayer.call (Hard_Layer_Subclassed.py:219-242, reached through CompleteTrainableModule.call,
 * complete_trainable_module.py:142-176) and the per-sample mean of the time-step field (physics_loss.py:102,122):
 *   p_l[b,c] = init_value - alpha_t(tn_l[b]) ^ expo[c] * y_l[b,c],   alpha_t(t) = (t - t_lo) / (t_hi - t_lo)
 *   dt_l[b]  = mean_c dtf_l[b,c]
 * y_l (B,D,H,W): network output at level l = n, n+1;  expo (D,H,W): the layer's kernel_exponent (NULL: 1);
 * tn_l (B,): the layer's time input (normalised time with the default identity nonormalize_func; norm_limits =
 * [t_lo, t_hi]);  dtf_l (B,D,H,W): time-step field, NULL (with dt_l NULL) skips the mean.
 * workspace: srm_glue_workspace_bytes(B) bytes (fp64 partial sums), only read when a mean is requested. */
size_t srm_glue_workspace_bytes(int32_t B);
int srm_glue_forward(const SrmHandle* h, int32_t B, float init_value, float t_lo, float t_hi, const float* expo,
                     const float* tn0, const float* tn1, const float* y0, const float* y1, const float* dtf1,
                     const float* dtf2, float* p0, float* p1, float* dt1, float* dt2, void* workspace,
                     size_t workspace_bytes, void* stream);
/* Cotangents of srm_glue_forward (what tape.gradient hands to the networks and to kernel_exponent):
 *   gy_l = -alpha * gp_l;   gexpo[c] = sum_b sum_l gp_l * (-y_l * alpha * ln alpha_t)  (0 where alpha_t <= 0, as
 *   tf.pow's gradient);   gdtf_l[b,c] = gdt_l[b] / N;
 *   gtn_l[b] = sum_c gp_l * (-y_l * expo[c] * alpha_t^(expo[c]-1)) / (t_hi - t_lo): the layer's time input is
 *   differentiable -- at level n+1 it carries the time-step model's output (physics_loss.py:105-111,
 *   Hard_Layer_Subclassed.py:214-228).   gexpo, gdtf_l (with gdt_l), gtn_l (B,) may be NULL. */
int srm_glue_backward(const SrmHandle* h, int32_t B, float init_value, float t_lo, float t_hi, const float* expo,
                      const float* tn0, const float* tn1, const float* y0, const float* y1, const float* gp0,
                      const float* gp1, const float* gdt1, const float* gdt2, float* gy0, float* gy1, float* gexpo,
                      float* gdtf1, float* gdtf2, float* gtn0, float* gtn1, void* stream);

/* BatchGenerator.__getitem__ on a device-resident data set (training.py:110-143: tf.gather(x_all, batch_inds,
 * axis=0) after converting the WHOLE data set to a tensor every step): dst[r] = src[idx[r]] for r < n_idx, rows of
 * row_bytes bytes; idx is a DEVICE int32[n_idx] (<= 65535 rows per call); an index outside [0, n_rows) yields a zero
 * row, as tf.gather does on a GPU.  Byte work: bit-exact. */
int srm_gather_rows(int32_t device, const void* src, const int32_t* idx, int64_t n_idx, int64_t n_rows,
                    int64_t row_bytes, void* dst, void* stream);

/* Feature-tensor glue of the step.  x is (B, cells, C) fp32 with the channels innermost -- the
 * reference's (B,D,H,W,5) features [z,y,x,t,k].  One pass over x writes
 *   x1_out = x with channel t_channel += dn[b]       the time-shifted features of level n+1; the reference builds
 *            them with zeros_like + strided assign + add (physics_loss.py:105-110), dn = normalize_diff(dt1)
 *   kx_out[b][cell] = exp(ln(kmax/kmin)*((x[..,k_channel]-lo)/(hi-lo)) + ln kmin)   the de-normalised permeability
 *            (DataSummary.nonormalize, log branch; same arithmetic as srm_denormalize_log, NaN/Inf -> 0)
 * Either output may be NULL.  Rows with cells*C a multiple of 4 in 16-byte aligned tensors take the vector path. */
int srm_features_forward(int32_t device, const float* x, const float* dn, int32_t B, int64_t cells, int32_t C,
                         int32_t t_channel, int32_t k_channel, float kmin, float kmax, float lo, float hi,
                         float* x1_out, float* kx_out, void* stream);
/* Cotangent of dn: gdn[b] = sum over cells of gx1[b][cell][t_channel] (fp64 accumulation); the cotangent of x is gx1
 * itself. */
int srm_features_backward(int32_t device, const float* gx1, int32_t B, int64_t cells, int32_t C, int32_t t_channel,
                          float* gdn, void* stream);

/* Feature construction on the device: weave_tensors (data_processing/data_processing_utils.py:
 * 90-223; call site srm_data_processing.py:363-403) over [permx (K, cells), time (T), x, y, z (cells)] with
 * flatten_first_axes and the channel flip, fused with DataSummary.normalize ('lnk-linear-scaling',
 * data_processing_utils.py:1031-1042).  All pointers DEVICE fp32.  stats is [5][2] = (min, max) per output channel in
 * the order [z, y, x, t, k]; channels 0..3 are normalised linearly, the permeability channel logarithmically; NaN/Inf -> 0.
 *   out[(k*T + t)][cell][0..4]      -- the (K*T, D, H, W, 5) tensor every training step reads, written once in HBM */
int srm_weave_features(int32_t device, int32_t K, int32_t T, int64_t cells, const float* permx, const float* time,
                       const float* xg, const float* yg, const float* zg, const float* stats, float lo, float hi,
                       float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SRM_PHYSICS_H_ */
