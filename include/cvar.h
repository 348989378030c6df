/*
 * cvar.h -- C ABI of the B200-native per-day Value-at-Risk solve.
 *
 * Drop-in boundary for ONE path of Nassim-cha/copula-MSM-and-copula-Garch-VaR:
 * the per-out-of-sample-day solve of  P(w . r <= q) = alpha  for q, where P is a
 * Riemann sum of copula density x forecast marginal densities over the part of
 * a fixed non-uniform n x n grid below the line w . x = q.
 *
 * The reference is pure Python; the functions below are what a ctypes binding
 * inside the reference's utils/calc_var_class.py would call (INTEGRATION.md
 * shows the stub).  Each entry point names the reference interface it replaces.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - every function returns an int status: 0 = ok, < 0 = invalid argument (no
 *     CUDA work issued), > 0 = a cudaError_t value.  cvar_strerror() decodes both.
 *   - `*_host` entry points take HOST pointers and perform the H2D copy, the
 *     kernels and the D2H copy on the plan's own stream, returning when the
 *     results are in the caller's buffers.
 *   - `*_device` entry points take DEVICE pointers owned by the caller (e.g.
 *     torch tensors on the plan's device), enqueue on the given cudaStream_t
 *     (passed as void*; NULL = the legacy default stream) and do not synchronise.
 *   - a plan owns scratch buffers (launch order, iteration counts, host-path workspace): calls on ONE plan must not
 *     overlap in time (use one plan per concurrent stream / thread); different plans are independent.
 *   - there is no CPU fallback: without a CUDA device plan creation fails.
 *   - all floating-point data is IEEE binary64.  Two-asset portfolios (dim = 2).
 */
#ifndef CVAR_H
#define CVAR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVAR_ABI_VERSION 1

/* copula families   (reference: utils/factory.py:11-31 copula_type) */
#define CVAR_COPULA_GAUSSIAN 0 /* copulas/gaussian/gaussian.py:47-117  */
#define CVAR_COPULA_STUDENT 1  /* copulas/student/student.py:49-174    */
#define CVAR_COPULA_PLACKETT 2 /* copulas/plackett/plackett.py:35-71   */

/* marginal families (reference: utils/factory.py estimation_type) */
#define CVAR_MARGINAL_SINGLE 0  /* one normal per asset per day: 'garch', 'mean_reverting'
                                   (integration_functions/garch_integration_function.py:5-52) */
#define CVAR_MARGINAL_MIXTURE 1 /* q-state normal mixture: 'msm'
                                   (integration_functions/msm_integration_function.py:5-47)   */

/* compat_flags: 1 = reproduce the reference, bit cleared = mathematically intended behaviour.
 * CVAR_COMPAT_REFERENCE (all set) is the default and the only mode parity is claimed for. */
#define CVAR_COMPAT_SIGMA_SWAP 1u   /* Q3: mixture pdf of grid column d uses the vol states of asset 1-d
                                       (utils/calc_integral/create_grids.py:121,143)                        */
#define CVAR_COMPAT_CASE_C_SIGN 2u  /* Q6: first bisection update of bracket C subtracts the strip
                                       (utils/calc_var_class.py:132,241-246)                                */
#define CVAR_COMPAT_NAN_TO_NUM 4u   /* Q5: single-normal integrand passes through nan_to_num (cells whose
                                       copula quantile is infinite contribute 0); mixture strips that touch
                                       such a cell are NaN (garch_integration_function.py:47)               */
#define CVAR_COMPAT_REFERENCE 7u

/* status codes (< 0) */
#define CVAR_OK 0
#define CVAR_ERR_NULL (-1)          /* required pointer is NULL                          */
#define CVAR_ERR_COPULA (-2)        /* unknown copula / marginal id                      */
#define CVAR_ERR_GRID (-3)          /* n < 2, n > CVAR_MAX_N, or axis not strictly ascending */
#define CVAR_ERR_PARAM (-4)         /* |rho| >= 1, nu <= 0, theta <= 0, w0 == 0, q out of range ... */
#define CVAR_ERR_SIZE (-5)          /* negative T, n_alpha out of range                  */
#define CVAR_ERR_NO_DEVICE (-6)     /* no usable CUDA device (there is no CPU fallback)  */
#define CVAR_ERR_ABI (-7)           /* struct_size / abi_version mismatch                */
#define CVAR_ERR_SMEM (-8)          /* grid too large for the shared memory of one SM    */
#define CVAR_ERR_DIM (-10)          /* portfolio dimension other than 2 (see cvar_check_dim)           */
#define CVAR_ERR_TABLE (-9)         /* Student-t: the per-plan quantile table is less accurate than 1e-11 for this nu
                                       (checked against the iterative routine at plan creation); the plan is refused */

#define CVAR_MAX_N 8192 /* hard cap; this version keeps a day's axis data in one SM's shared memory, which
                           limits n to ~4900 on B200 (cvar_plan_create returns CVAR_ERR_SMEM beyond) */
#define CVAR_MAX_Q 32
#define CVAR_MAX_ALPHA 8
#define CVAR_MAX_ITER 30

/* bracket ids written by the solve (reference: utils/calc_var_class.py:147-155) */
#define CVAR_CASE_A 0         /* [min_var,   second_lo] */
#define CVAR_CASE_B 1         /* [second_lo, first]     */
#define CVAR_CASE_C 2         /* [second_hi, max_var]   */
#define CVAR_CASE_D 3         /* [first,     second_hi] */
#define CVAR_CASE_UNDEFINED 4 /* mass == alpha exactly or NaN: the reference leaves np.empty garbage (Q8);
                                 this library returns NaN for that day                      */

/*
 * Run-constant description of the solve.  Replaces the scattered literals and
 * attributes of the reference (utils/calc_var_class.py:16-17,95,111-112,201-202,257;
 * copula parameters from utils/model_estimation/copula/\*_estimation.py
 * `copula_integrations_params`).  Fill with cvar_desc_default() first.
 */
typedef struct cvar_desc {
    uint32_t struct_size; /* sizeof(cvar_desc_t), checked */
    uint32_t abi_version; /* CVAR_ABI_VERSION             */
    int32_t copula;       /* CVAR_COPULA_*                */
    int32_t marginal;     /* CVAR_MARGINAL_*              */
    int32_t n;            /* grid points per axis (num_points) */
    int32_t q;            /* vol states per asset; 1 for CVAR_MARGINAL_SINGLE */
    uint32_t compat_flags;
    int32_t max_iter;     /* bisection iterations recorded per solve; 0 = derive from tol (22) */
    double rho;           /* Gaussian / Student correlation */
    double nu;            /* Student degrees of freedom     */
    double theta;         /* Plackett parameter             */
    double w0, w1;        /* portfolio weights              */
    double clip_lo;       /* lower end of the grid, -5      */
    double neg_inf;       /* stand-in for -infinity, -100   */
    double first_guess;   /* -3   */
    double second_lo;     /* -3.5 */
    double second_hi;     /* -2   */
    double min_var;       /* -7.5 */
    double max_var;       /*  0   */
    double tol;           /* 1e-6 */
} cvar_desc_t;

typedef struct cvar_plan cvar_plan_t; /* opaque: device copies of the axis, tables, workspace, stream */

/* Per-plan facts a caller may want to read back. */
typedef struct cvar_plan_info {
    int32_t device;            /* CUDA device ordinal                                  */
    int32_t sm_count;          /* multiprocessors                                      */
    int32_t max_iter;          /* iterations the solve kernel records                  */
    int32_t ctas_per_sm;       /* resident solve CTAs per SM (occupancy query)         */
    int32_t threads_per_cta;
    int32_t smem_bytes_per_cta;
    double tq_table_max_rel_err; /* Student only: measured max of |table - iterative| / max(|iterative|, 0.1)
                                    over 4 off-node probes per table interval; 0 otherwise */
    double last_kernel_ms;     /* device time of the last *_host solve (CUDA events), ms */
    int32_t kernel_variant;    /* 0 Gaussian, 1 Student-t (generic log2/exp2 cell), 2 Plackett, 3/4/5/6 Student-t with the
                                  table-assisted power cell of degree 5/6/7/8 (picked from nu at plan creation) */
    int32_t cluster4_capacity; /* 4-CTA clusters of the solve kernel that can be resident at once: batches of at most
                                  this many days are split four ways, up to about twice as many two ways (n >= 1024) */
    int32_t pow_octaves;       /* Student-t power cell: octaves of the quadratic form covered by the one-lookup table */
    int64_t chunk_days;        /* days per launch of the solve kernel (size of the launch-order scratch, cvar_plan_reserve) */
} cvar_plan_info_t;

/*
 * Environment knobs read at cvar_plan_create (tuning and tests; defaults are measured choices):
 *   CVAR_CTA_THREADS=32..512   fixed CTA size of the solve / strip kernels (disables the launch-time choices below)
 *   CVAR_CLUSTER=1|2|4         fixed number of CTAs (thread-block cluster) per day; default: 4 / 2 while the whole batch
 *                              stays resident (n >= 1024), else 1
 *   CVAR_STUDENT_GENERIC=1     Student-t plans use the generic log2/exp2 cell instead of the table-assisted power cell
 *   CVAR_NO_SEGMENT_GUESS=1    row boundaries by plain bisection even on a piecewise-uniform axis
 *   CVAR_POW_OCTAVES=0..8      octaves of the Student-t cell's one-lookup power table (default: as many as fit without
 *                              costing the SM a resident warp, none below four)
 *   CVAR_CHUNK_DAYS=N          days per launch of the solve kernel a new plan reserves scratch for (cvar_plan_reserve)
 *   CVAR_TQ_BUDGET=x           accuracy budget of the Student-t quantile table (default 1e-11; tests force a refusal)
 *   CVAR_ORDER_MIN_WAVES=k     chunks of at most k waves of resident CTAs are started in their natural order (default 2)
 * Read by the Python layers: CVAR_BACKEND=b200|reference (factory default).
 * The Python loader additionally honours CVAR_B200_LIB=<path to an alternative libcvar_b200.so>.
 */

/*
 * Portfolio dimension.  DECISION: this backend solves TWO-asset portfolios and refuses anything else with CVAR_ERR_DIM;
 * it does not reproduce the reference's recursive grid for dim >= 3 "as written".  Reason: run unmodified on a three-asset
 * Gaussian + GARCH example (n = 24), the reference's create_nested_grid (utils/calc_integral/create_grids.py:125-171)
 * returns "probabilities" F(0.5) = 3.45 and F(20) = 4.32 -- its recursion multiplies every inner level by the full step
 * product -- so there is no meaningful parity target, and every BASELINE configuration has two assets.  cvar_desc_t
 * therefore carries exactly two weights; hosts that hold a `dim` (the reference's ValueAtRiskCalcualtion.dim,
 * utils/calc_var_class.py:31) check it here and surface the refusal (the Python mirror and `dropin` raise
 * NotImplementedError with this status' text).
 */
int cvar_check_dim(int32_t dim);

/* Fill *desc with the reference's defaults (everything except copula/marginal/n/q/params). */
void cvar_desc_default(cvar_desc_t* desc);

const char* cvar_strerror(int status);
int cvar_abi_version(void);

/*
 * Create a plan on CUDA device `device` (-1 = current device).
 *   x_host[n], dx_host[n]      : the axis and its right-endpoint steps (host pointers), exactly the arrays
 *                                of the reference's grids_generations_params (compute_normal_densities,
 *                                garch_estimation.py:148-188 / msm_estimation.py:283-330)
 *   sigma_states_host[2][q]    : merged vol states (integrations_params_static); NULL for single-normal
 */
int cvar_plan_create(const cvar_desc_t* desc, const double* x_host, const double* dx_host,
                     const double* sigma_states_host, int device, cvar_plan_t** plan_out);
int cvar_plan_destroy(cvar_plan_t* plan);
int cvar_plan_get_info(const cvar_plan_t* plan, cvar_plan_info_t* info_out);
/*
 * Size the plan's launch-order scratch (about 16 bytes per day) for batches of `days` days.  A batch larger than the
 * reserved chunk still runs, cut into chunks that are ordered and launched one after the other; the default chunk is
 * 65536 days (CVAR_CHUNK_DAYS).  The `*_host` solve reserves for its own batch; the `*_device` entry points never
 * allocate, so callers with larger device-resident batches reserve once after plan creation.  Only ever grows;
 * synchronises the plan's stream when it reallocates.  (No counterpart in the reference.)
 */
int cvar_plan_reserve(cvar_plan_t* plan, int64_t days);

/*
 * Strip masses: out[t] = S(bounds[t][0], bounds[t][1]) for day t.
 * Replaces ValueAtRiskCalcualtion.compute_integral (utils/calc_var_class.py:179-212), i.e.
 * calc_grids_and_integrals_results (utils/calc_integral/calc_integral.py:8-119).
 *   day_params : [T][2] sigma (single-normal) or [T][2][q] state probabilities (mixture)
 *   bounds     : [T][2]
 *   out        : [T]
 *   out_cells  : [T] number of grid cells in each strip, may be NULL
 */
int cvar_strip_mass_host(cvar_plan_t* plan, const double* day_params, int64_t T, const double* bounds,
                         double* out, uint64_t* out_cells);
int cvar_strip_mass_device(cvar_plan_t* plan, const double* day_params, int64_t T, const double* bounds,
                           double* out, uint64_t* out_cells, void* stream);

/*
 * The solve.  Replaces ValueAtRiskCalcualtion.calc_var + bisection_algorithm + adjust_integral
 * (utils/calc_var_class.py:95-177, 250-309, 214-248), one call per alpha in the reference; here all
 * alphas of a day are solved by the same CTA and share the per-day axis work.
 *
 * cvar_solve_device records, per (alpha, day), the whole decision sequence instead of a single number,
 * because the reference iterates every day until the slowest day of the batch has converged (Q7): the
 * iteration count K is a property of the full batch (possibly spread over several GPUs) and is applied
 * afterwards by cvar_finalize_*.
 *   traj  : [n_alpha][T][2] uint32
 *           word 0: bits 0..max_iter-1 = decision of iteration k (1: mass < alpha, lower end moves up),
 *                   bits 28..30 = bracket id CVAR_CASE_*, bit 31 = a running mass cancelled to rounding noise without
 *                   being exactly 0 (feeds CVAR_STATUS_ZERO_EXIT_AMBIGUOUS below)
 *           word 1: bits 0..max_iter-1 = running mass after iteration k was exactly 0 (early-exit test,
 *                   calc_var_class.py:293-295)
 *   mass  : [n_alpha][T] running mass after the last recorded iteration, may be NULL
 *   cells : [n_alpha][T] grid cells evaluated for that solve (algorithmic work counter), may be NULL
 * alphas is a HOST pointer in both variants (n_alpha <= CVAR_MAX_ALPHA).
 */
int cvar_solve_device(cvar_plan_t* plan, const double* day_params, int64_t T, const double* alphas,
                      int32_t n_alpha, uint32_t* traj, double* mass, uint64_t* cells, void* stream);

/*
 * Apply the global iteration count and produce VaR levels (calc_var_class.py:278,293-295,306,171).
 *   traj      : [n_alpha][T][2] as written by cvar_solve_device (T may be the concatenation of the
 *               blocks of several GPUs)
 *   forced_iterations : NULL, or n_alpha host ints; entry >= 0 overrides the batch-derived K for that alpha
 *   var_out   : [n_alpha][T]  solved quantile + ptf_mean (NaN for CVAR_CASE_UNDEFINED days)
 *   case_out  : [n_alpha][T] bracket ids, may be NULL
 *   iterations_out : device (for _device) / host (for _host) array of n_alpha ints receiving K, may be NULL
 */
int cvar_finalize_device(cvar_plan_t* plan, const uint32_t* traj, int64_t T, int32_t n_alpha,
                         const int32_t* forced_iterations, double ptf_mean, double* var_out,
                         int32_t* case_out, int32_t* iterations_out, void* stream);
/*
 * The same over trajectory words that arrive in blocks of `block_days` days, [n_blocks][n_alpha][block_days][2] with day d
 * in block d / block_days: exactly what an all-gather of the ranks' [n_alpha][block_days][2] arrays leaves behind when
 * every rank solves a contiguous block of ceil(T / ranks) days (the last block may be ragged: only its first
 * T - (n_blocks - 1) * block_days days are read).  No reshuffling kernel between the collective and the finalize.
 * block_days >= T is the plain layout of cvar_finalize_device.
 */
int cvar_finalize_blocked_device(cvar_plan_t* plan, const uint32_t* traj, int64_t T, int64_t block_days, int32_t n_alpha,
                                 const int32_t* forced_iterations, double ptf_mean, double* var_out,
                                 int32_t* case_out, int32_t* iterations_out, void* stream);

/*
 * Work counter for the roofline: the number of grid cells the plan's solve launches really evaluated since the last
 * reset, summed over days and alphas.  Differs from the `cells` output of cvar_solve_* (the cells of the reference's own
 * strip scheme, per solve) when several alphas are solved together: strips they have in common are evaluated once.
 * Synchronises the device.  (No counterpart in the reference.)
 */
int cvar_evaluated_cells_host(cvar_plan_t* plan, uint64_t* total_out, int reset);

/*
 * Status words of the plan's LAST finalize (cvar_finalize_device, or the finalize inside cvar_solve_host), one per alpha:
 *   CVAR_STATUS_ZERO_EXIT_TAKEN      the iteration count K was cut because the running mass of EVERY day of the batch was
 *                                    exactly 0 after iteration K -- the reference's early exit (calc_var_class.py:293-295)
 *   CVAR_STATUS_ZERO_EXIT_AMBIGUOUS  every day's running mass was 0 at some iteration either exactly or up to rounding (a
 *                                    difference of two equal sums that cancelled to <= 1e-12 of its operands).  Whether the
 *                                    reference takes its early exit on such a batch depends on the order in which its sums
 *                                    happened to be formed; the VaR levels returned here follow the
 *                                    no-exit branch, i.e. the bisection runs on to the quantile.  Only tiny batches whose
 *                                    alpha lies below all the mass the grid holds under the first midpoint can get here.
 * `_device`: status_out is a device array, the copy is enqueued on `stream` (after the finalize it reports on);
 * `_host`: status_out is a host array, the call synchronises the plan's stream.
 */
#define CVAR_STATUS_ZERO_EXIT_TAKEN 1
#define CVAR_STATUS_ZERO_EXIT_AMBIGUOUS 2
int cvar_finalize_status_device(cvar_plan_t* plan, int32_t* status_out, int32_t n_alpha, void* stream);
int cvar_finalize_status_host(cvar_plan_t* plan, int32_t* status_out, int32_t n_alpha);

/*
 * One-call host entry point: H2D of day_params, solve, finalize, D2H of the VaR levels.
 *   var_out [n_alpha][T], case_out [n_alpha][T] (may be NULL), cells_out [n_alpha][T] (may be NULL),
 *   iterations_out [n_alpha] (may be NULL).
 */
int cvar_solve_host(cvar_plan_t* plan, const double* day_params, int64_t T, const double* alphas,
                    int32_t n_alpha, const int32_t* forced_iterations, double ptf_mean, double* var_out,
                    int32_t* case_out, uint64_t* cells_out, int32_t* iterations_out);

/*
 * Device special functions exposed for testing (elementwise, host pointers, run on the plan's device).
 *   which: 0 = Student-t quantile via the plan's table (nu = desc.nu), 1 = iterative t quantile,
 *          2 = exp2 kernel primitive, 3 = log2 kernel primitive, 4 = Phi via erf (Q14), 5 = normal quantile
 */
int cvar_test_special_host(cvar_plan_t* plan, int32_t which, const double* in, int64_t count, double* out);

/*
 * Elementwise copula density c(u[i][0], u[i][1]) -- the calculators' `copula_density` hook
 * (copulas/gaussian/gaussian.py:47-61, copulas/student/student.py:49-79, copulas/plackett/plackett.py:35-71).
 * Not on the solve path (the solve never materialises cell lists); provided so that the plugin API is
 * complete.  NaN where a Gaussian / Student quantile is infinite (u == 0 or 1), like the reference.
 *   u : [count][2] host, out : [count] host, device : CUDA ordinal (-1 = current)
 */
int cvar_copula_density_host(int32_t copula, double rho, double nu, double theta, const double* u, int64_t count,
                             double* out, int device);

/*
 * Measure the FP64-pipe peak of `device` with a dependency-free DFMA micro-benchmark running for at
 * least `min_ms` milliseconds (the roofline denominator of this path: MEASURED_PEAKS.json only carries
 * HBM and bf16 tensor peaks).  *tflops_out = 2 * DFMA issued / time.
 */
int cvar_fp64_peak_host(int device, double min_ms, double* tflops_out, double* ms_out);

/* =====================================================================================================
 * Forecast producers: what fills `day_params` from rolling windows of centred returns.
 * Rolling window w of an asset is returns[w * window_stride .. w * window_stride + N) -- window_stride = 1 for the
 * reference's overlapping windows (data_loader/load_data.py:131-137), N for independent windows.
 * ===================================================================================================== */

/*
 * Binomial MSM(k): filtered state distribution at the end of every window, merged to the q distinct vol levels.
 * Replaces MSMEstimation.forecasts_array + sum_forecast_by_state (utils/model_estimation/model/msm_estimation.py:143-248),
 * i.e. calc_forecasts -> ProbEstimation.calc_state_prob (markov_switching_multifractal/calc_marginals.py:33-38,
 * calc_prob.py:8-69, 110-122).
 *   stay_prob      [n_assets][k]   p_c = 1 - gamma_c/2, gamma_c = 1-(1-gamma)^(b^c)   (calc_prob.py:90-101)   HOST
 *   vol_states     [n_assets][2^k] sigma_s in itertools.product order (calc_prob.py:103-108)
 *   level_of_state [n_assets][2^k] index of each state's merged vol level, 0..q-1
 *   returns        [n_assets][(T-1)*window_stride + N]
 *   probs_by_state [T][n_assets][q]   == the solve's day_params for CVAR_MARGINAL_MIXTURE with n_assets = 2
 *   state_probs    [n_assets][T][2^k] un-merged filtered probabilities, may be NULL
 *   status         set to 1 if a window's normalising constant was 0 (the reference then marks the run failed);
 *                  that window's probabilities are NaN
 * _device: vol_states, level_of_state, returns, outputs, status and workspace ([n_assets][L][2^k] doubles,
 *          L = (T-1)*window_stride + N) are DEVICE pointers; stay_prob is a host pointer; enqueues on `stream`.
 */
int cvar_msm_forecast_host(int32_t k, int32_t n_assets, const double* stay_prob, const double* vol_states,
                           const int32_t* level_of_state, int32_t q, const double* returns, int64_t T, int64_t N,
                           int64_t window_stride, double* probs_by_state, double* state_probs, int32_t* status_out,
                           double* kernel_ms_out, int device);
int cvar_msm_forecast_device(int32_t k, int32_t n_assets, const double* stay_prob, const double* vol_states,
                             const int32_t* level_of_state, int32_t q, const double* returns, int64_t T, int64_t N,
                             int64_t window_stride, double* probs_by_state, double* state_probs, double* workspace,
                             int32_t* status, void* stream);

/*
 * GARCH(p,q) one-step volatility forecast per window: replaces GarchEstimation.compute_forecast ->
 * calc_forecast (utils/model_estimation/model/garch_estimation.py:190-231, garch/forecast.py:5-18,
 * garch/estimation.py:40-65).  alpha / beta are [n_assets][8] (unused entries ignored), p, q <= 8.
 *   sigma_out [T][n_assets] == the solve's day_params for CVAR_MARGINAL_SINGLE with n_assets = 2
 * _device: returns and sigma_out are DEVICE pointers, the model parameters stay host pointers; enqueues on `stream`,
 *          so the forecast can feed cvar_solve_device on the same stream without a host round trip.
 */
int cvar_garch_forecast_host(int32_t n_assets, const double* omega, const int32_t* p, const int32_t* q, const double* alpha,
                             const double* beta, const double* returns, int64_t T, int64_t N, int64_t window_stride,
                             double* sigma_out, double* kernel_ms_out, int device);
int cvar_garch_forecast_device(int32_t n_assets, const double* omega, const int32_t* p, const int32_t* q, const double* alpha,
                               const double* beta, const double* returns, int64_t T, int64_t N, int64_t window_stride,
                               double* sigma_out, void* stream);

/*
 * Kalman mean-reverting log-volatility model: exp(last predicted state mean) of the scalar unscented filter per
 * window.  Replaces MeanRevertingEstimation.compute_forecast -> calc_forecast
 * (utils/model_estimation/model/mean_reverting_estimation.py:192-232, kalman_mean_reverting/forecast.py:5-12,
 * kalman_mean_reverting/estimate.py:231-281).  a, l, q are [n_assets]; ukf_alpha/beta/kappa default to 1.6, 2, 1.75
 * in the reference.  status 1 = the filter's normalising constant collapsed in some window (the reference returns
 * an error tuple there); that window's forecast is NaN.
 *   sigma_out [T][n_assets]
 * _device: returns, sigma_out and status (one int32, zeroed by the caller) are DEVICE pointers; enqueues on `stream`.
 */
int cvar_kalman_forecast_host(int32_t n_assets, const double* a, const double* l, const double* q, double ukf_alpha,
                              double ukf_beta, double ukf_kappa, const double* returns, int64_t T, int64_t N,
                              int64_t window_stride, double* sigma_out, int32_t* status_out, double* kernel_ms_out, int device);
int cvar_kalman_forecast_device(int32_t n_assets, const double* a, const double* l, const double* q, double ukf_alpha,
                                double ukf_beta, double ukf_kappa, const double* returns, int64_t T, int64_t N,
                                int64_t window_stride, double* sigma_out, int32_t* status, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CVAR_H */
