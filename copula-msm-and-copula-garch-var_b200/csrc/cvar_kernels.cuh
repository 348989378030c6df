// cvar_kernels.cuh -- the per-day VaR solve on sm_100a.
//
// One CTA per out-of-sample day (a 2- or 4-CTA thread-block cluster per day when the batch is smaller than the
// GPU, see solve_kernel).  The CTA
//   stage 0  evaluates every per-axis-point quantity once (marginal cdf/pdf, the MSM mixture over its
//            q states, the copula quantile) and leaves them in shared memory in the form the cell loop
//            wants (reference: integration_functions/*_integration_function.py,
//            copulas/*/*.py),
//   then, for every alpha, runs the reference's whole bracket + bisection state machine in-kernel
//   (utils/calc_var_class.py:95-177, 250-309): each step integrates only the strip between the previous
//   and the new boundary (create_grids.py:102-110), with FP64 warp-shuffle + block reductions, and no
//   host round trip per iteration.
//
// Cell membership is decided by the reference's exact expression  x[j] <= (q - x[i]*w1)/w0  evaluated
// with individually rounded operations (integration_algo.py:20), so the set of cells in every strip is
// identical to the reference's; only the cell *weights* are computed in a restructured (per-axis
// factored) form.
#pragma once
#include <cooperative_groups.h>

#include "cvar_math.cuh"

namespace cvar {

// CTA size is chosen per plan (cvar_api.cu): the smallest of 64 / 128 / 256 / 512 threads whose resident copies
// still give an SM 16 warps of row walkers; 512 threads when only one CTA fits an SM or the batch is small.
constexpr int CTA_THREADS_SMALL = 256;
constexpr int CTA_THREADS_LARGE = 512;
constexpr int MAX_CTA_WARPS = CTA_THREADS_LARGE / 32;
// Kernel variants (template parameter COPULA of the kernels below): the three copula families plus four
// Student-t variants whose cell uses the table-assisted power with a plan-fitted polynomial of fixed degree.
constexpr int KV_GAUSSIAN = 0, KV_STUDENT = 1, KV_PLACKETT = 2, KV_STUDENT_POW5 = 3, KV_STUDENT_POW6 = 4, KV_STUDENT_POW7 = 5,
              KV_STUDENT_POW8 = 6, KV_COUNT = 7;
__host__ __device__ constexpr bool kv_is_student(int kv) { return kv == KV_STUDENT || kv >= KV_STUDENT_POW5; }
__host__ __device__ constexpr int kv_pow_degree(int kv) {
    return kv == KV_STUDENT_POW5 ? 5 : kv == KV_STUDENT_POW6 ? 6 : kv == KV_STUDENT_POW7 ? 7 : kv == KV_STUDENT_POW8 ? 8 : 0;
}

// independent cells per thread per loop trip (FP64 latency hiding); the cheap cells need more of them
#ifndef CVAR_CIF_GAUSSIAN
#define CVAR_CIF_GAUSSIAN 8
#endif
#ifndef CVAR_CIF_STUDENT
#define CVAR_CIF_STUDENT 4
#endif
#ifndef CVAR_CIF_PLACKETT
#define CVAR_CIF_PLACKETT 8
#endif
template <int COPULA>
struct CellsInFlight {
    static constexpr int value = COPULA == 0 ? CVAR_CIF_GAUSSIAN : (kv_is_student(COPULA) ? CVAR_CIF_STUDENT : CVAR_CIF_PLACKETT);
};

typedef unsigned short u16;

// Alpha fusion: the alphas of a day walk the same first strips (F(first), the second probe, the first
// bisection steps while their decisions agree).  A strip is a pure function of its (lo, hi) pair, so the first
// MEMO_PER_ALPHA strips of every alpha are remembered and later alphas reuse the mass instead of re-integrating.
constexpr int MEMO_SIZE = 32;
constexpr int MEMO_PER_ALPHA = 7;

constexpr int MAX_AXIS_SEGMENTS = 8;

struct KernelParams {
    int copula, marginal, n, q;
    unsigned compat;
    int max_iter;
    int cmin;  // #{x <= clip_lo}: first inner index that can ever belong to a strip
    double rho, nu, theta, w0, w1;
    double rw0_exact;  // 1 / w0 when w0 is a power of two (division == multiplication, exactly), else 0
    // uniform segments of the axis (the reference's axis has five); nseg = 0: not piecewise uniform, plain bisection
    int nseg;
    int seg_first[MAX_AXIS_SEGMENTS + 1];     // first index of segment s; seg_first[nseg] = n
    double seg_x0[MAX_AXIS_SEGMENTS];         // x[seg_first[s]]; +inf beyond nseg
    double seg_inv_h[MAX_AXIS_SEGMENTS];      // 1 / spacing of segment s
    double neg_inf, first, second_lo, second_hi, min_var, max_var;
    // copula constants prepared on the host
    double g_in_scale;   // gaussian: sqrt(kappa*log2e)          student: 1/sqrt(nu(1-rho^2))
    double g_out_scale;  // gaussian: sgn(rho) sqrt(log2e/(2(1-rho^2)))   student: rho/sqrt(nu(1-rho^2))
    double g_const;      // gaussian: (1-rho^2)^(-1/2)            student: Gamma-ratio / sqrt(1-rho^2)
    double tq_tail_lc;   // student: log of the leading tail coefficient
    double negc;         // student: -(nu+2)/2, the exponent of the quadratic form
    double inv_nu;       // student: 1/nu
    unsigned seed_mask, seed_half;  // student pow variants: keep sign/exponent/top POW_BITS mantissa bits | interval midpoint bit
    double y_max;        // student pow variants: clamp of |T_nu^-1(u)| that keeps the quadratic form below 2^63
    double qc[CVAR_LOG2_1P_POLY_DEG + 1];  // student: negc * coefficients of log2(1+f)/f
    const double* logtab;  // student: negc * (-log2 r_i), LOGTAB_SIZE entries (global; staged to shared memory)
    const double* exptab;  // gaussian / student: 2^(i/EXPTAB_SIZE), EXPTAB_SIZE entries (global; staged to shared memory)
    double powc[POW_MAX_DEG + 1];  // student pow variants: polynomial for (1+f)^(-(nu+2)/2), |f| <= POW_FMAX
    const double* powtab;  // student pow variants: r_i^c (POW_MTAB) then 2^(-c e) (POW_ETAB)
    const double* powfast; // student pow variants: one-lookup table over the first pow_octaves octaves of t (cvar_math.cuh)
    int pow_octaves;       // 0: no one-lookup table
    double pow_fast_limit; // rows whose quadratic form stays below this use the one-lookup table
    const double* x;
    const double* dx;
    const double* sigma_states;  // [2][q] or nullptr
    const double* tq_table;      // student only
    const double* state_cdf;     // mixture: Phi(x_i / sigma_{a,s}), [2][q][n]
    const double* state_pdf;     // mixture: N(x_i; 0, sigma_{a,s}), [2][q][n]
    double thin_width;           // a bisection bracket narrower than this (in q) holds at most one grid point per row
    unsigned long long* evaluated_cells;   // plan-wide counter of the cells the launches really evaluated (strips shared
                                           // between the alphas of a day are evaluated once): the roofline's numerator
};

struct AlphaSet {
    int n_alpha;
    double a[8];
};

// dynamic shared memory carve-up
struct Smem {
    double* xs;     // [n] axis (membership searches)
    double2* in;    // [n] inner-axis pair (.x, .y), interleaved so one 16-byte load feeds a cell
    double* out0;   // [n] outer-axis array 0 (gaussian: scaled quantile, student: raw quantile, plackett: u)
    double* out1;   // [n] outer-axis array 1 (row weight)
    u16* c[3];      // [n] each: inner index bounds per outer row
    double* red;    // [2][MAX_CTA_WARPS]
    unsigned* redc; // [2][MAX_CTA_WARPS]
    int* live;      // [4]: dead-prefix / dead-suffix counts per axis
    double* cl_mass;    // [2] this CTA's strip partial, read by the other CTAs of a cluster
    unsigned* cl_cells; // [2]
    double* ltab;   // [LOGTAB_SIZE] student only
    double* etab;   // [EXPTAB_SIZE] gaussian / student
    double* memo;   // [MEMO_SIZE][4]: (lo, hi, mass, cells) of strips already integrated for an earlier alpha
    double* ptab;   // [POW_MTAB + POW_ETAB] student pow variants (aliases the ltab/etab region)
    unsigned ptab_s;  // shared-window address of ptab
    double* pfast;    // one-lookup power table (pow_octaves * POW_MTAB entries), 16-byte aligned
    unsigned pfast_s; // its shared-window address minus the index bias (see pow_neg_c_fast)
};

// doubles of per-variant lookup tables staged in shared memory (after the fixed part)
__host__ __device__ constexpr int table_doubles(int kv, int pow_octaves) {
    return kv == KV_GAUSSIAN ? EXPTAB_SIZE
         : kv == KV_STUDENT  ? LOGTAB_SIZE + EXPTAB_SIZE
         : kv == KV_PLACKETT ? 0
                             : POW_MTAB + POW_ETAB + pow_octaves * POW_MTAB * POW_FAST_ENTRY_DOUBLES;
}

__host__ __device__ inline size_t smem_bytes_for(int n, int kv, int pow_octaves) {
    size_t npad = (size_t)((n + 3) & ~3);
    return npad * 8 * 5 + npad * 2 * 3 + 2 * MAX_CTA_WARPS * 8 + 2 * MAX_CTA_WARPS * 4 + 16 + 64 + MEMO_SIZE * 4 * 8 + 8 +
           (size_t)table_doubles(kv, pow_octaves) * 8;
}

__device__ __forceinline__ Smem carve(unsigned char* base, int n, int kv, int pow_octaves) {
    size_t npad = (size_t)((n + 3) & ~3);
    Smem S;
    double* d = reinterpret_cast<double*>(base);
    S.xs = d;
    S.in = reinterpret_cast<double2*>(d + npad);
    S.out0 = d + 3 * npad;
    S.out1 = d + 4 * npad;
    S.red = d + 5 * npad;
    u16* h = reinterpret_cast<u16*>(S.red + 2 * MAX_CTA_WARPS);
    S.c[0] = h;
    S.c[1] = h + npad;
    S.c[2] = h + 2 * npad;
    S.redc = reinterpret_cast<unsigned*>(h + 3 * npad);
    S.live = reinterpret_cast<int*>(S.redc + 2 * MAX_CTA_WARPS);
    S.cl_mass = reinterpret_cast<double*>(S.live + 4);    // live[4] | cl_mass[2] | cl_cells[2] | pad: 64 bytes in all
    S.cl_cells = reinterpret_cast<unsigned*>(S.live + 8);
    S.memo = reinterpret_cast<double*>(S.live + 4 + 12);  // 64 bytes after `live`: stays 8-byte aligned
    double* tables = S.memo + MEMO_SIZE * 4;              // variant-specific tables, see table_doubles()
    tables += (reinterpret_cast<size_t>(tables) >> 3) & 1;   // 16-byte aligned (the u16 arrays leave it 8-byte aligned)
    S.ltab = tables;                                      // KV_STUDENT: ltab then etab
    S.etab = kv == KV_STUDENT ? tables + LOGTAB_SIZE : tables;
    S.ptab = tables;
    S.ptab_s = (unsigned)__cvta_generic_to_shared(tables);
    S.pfast = tables + POW_MTAB + POW_ETAB;
    S.pfast_s = S.ptab_s + (unsigned)((POW_MTAB + POW_ETAB) * 8) - (unsigned)((1023 << POW_BITS) * 8 * POW_FAST_ENTRY_DOUBLES);
    (void)pow_octaves;
    return S;
}

// ---------------------------------------------------------------------------------------------
// stage 0: per-axis quantities
// ---------------------------------------------------------------------------------------------
// One axis point of one asset: the two doubles the cell loop wants for it (reference:
// integration_functions/*_integration_function.py, copulas/*/*.py).
//   asset 1 (inner axis, the columns):  first = scaled copula quantile (Plackett: u), second = column weight
//   asset 0 (outer axis, the rows):     first = quantile / u,                         second = row weight
// `dead` is set to 1 (lower end) or 2 (upper end) when u saturated to exactly 0 or 1, where the reference's copula
// density is NaN (Q5/Q10); such points carry zero weight and are counted into the day's live window.
template <int COPULA>
__device__ __forceinline__ void axis_point(const KernelParams& P, const double* __restrict__ dayp, int d, int i,
                                           double& first, double& second, int& dead) {
    const int n = P.n, q = P.q;
    const double xi = P.x[i], dxi = P.dx[i];
    double u, a;
    dead = 0;
    if (P.marginal == 0) {
        // garch_integration_function.py:27-38  (u = Phi(x/sigma), pdf = phi(x/sigma)/sigma)
        const double sg = dayp[d];
        const double z = __ddiv_rn(xi, sg);
        u = phi_via_erf(z);
        a = (CVAR_INV_SQRT_2PI * exp(-0.5 * z * z) / sg) * dxi;
    } else {
        // msm_integration_function.py:32-36 (cdf mixture) and create_grids.py:121,143 (pdf mixture with the vol
        // states of the OTHER asset when the Q3 compat bit is set).  The vol states are run constants, so
        // Phi(x_i / sigma_{a,s}) and N(x_i; 0, sigma_{a,s}) do not depend on the day: they are tabulated once per plan
        // (state_table_kernel) and a day only forms the two probability-weighted sums per axis point -- q FMAs
        // instead of q erf + q exp evaluations.
        const bool swap = (P.compat & 1u) != 0;
        double su = 0.0, sa = 0.0;
        const double* pr = dayp + d * q;
        const double* cdf_t = P.state_cdf + (size_t)d * q * n + i;
        const double* pdf_t = P.state_pdf + (size_t)(swap ? (1 - d) : d) * q * n + i;
        for (int s = 0; s < q; ++s) {
            const double p = pr[s];
            su += p * cdf_t[(size_t)s * n];
            sa += p * pdf_t[(size_t)s * n];
        }
        u = su;
        a = dxi * sa;
    }
    if (COPULA == 2) {
        first = u;
        second = a;
        return;
    }
    double y = COPULA == 0 ? normcdfinv(u) : t_quantile_table(P.tq_table, P.nu, P.tq_tail_lc, u);
    if (!isfinite(y)) {  // u == 0 or u == 1
        dead = y < 0.0 ? 1 : 2;
        y = 0.0;
        a = 0.0;
    }
    // quantiles beyond ~1e9 (u below ~1e-19 even at nu = 2; such u only arise from mixture weights far below 1e-10)
    // are clamped so that the cell's power table never needs an exponent above 2^63; every factor of the cell is
    // formed from the clamped value, so the cell stays a density value of the same tail (it changes by a power of the
    // clamp ratio on a region whose total mass is below 1e-19)
    if (kv_pow_degree(COPULA) > 0) y = copysign(fmin(fabs(y), P.y_max), y);
    if (COPULA == 0) {
        if (d == 1) {
            first = P.g_in_scale * y;
            second = fmax(log2(a), -1100.0);
        } else {
            first = P.g_out_scale * y;
            second = a * P.g_const * exp(0.5 * y * y);
        }
    } else {
        const double hp = 0.5 * (P.nu + 1.0);
        const double c = 1.0 + y * y / P.nu;
        if (d == 1) {
            first = P.g_in_scale * y;
            if (kv_pow_degree(COPULA) > 0)   // the power variants multiply by the column weight instead of adding its log
                second = a * exp2(hp * log2(c));
            else
                second = fmax(fmax(log2(a), -1100.0) + hp * log2(c), -1100.0);
        } else {
            first = y;   // the row derives m0 = g_out_scale * y0 and c0 = 1 + y0^2 / nu when it is loaded
            second = a * P.g_const * exp2(hp * log2(c));
        }
    }
}

// Stage 0 of a solve: the day's axis arrays and the plan's lookup tables go to shared memory.  (Measured: running this
// stage as a separate, fully parallel kernel per batch that hands the arrays over through HBM is NOT faster -- inside
// the solve kernel its latency hides behind the cell loops of the SM's other resident CTA, while the hand-over adds a
// launch and 32 n bytes of traffic per day: c3 +1 %, c4 +2 %, c1 +15 %.)
// Returns true when every cell of the day has a quadratic form below pow_fast_limit (Student-t power variants).
template <int COPULA>
__device__ bool stage0(const KernelParams& P, const double* __restrict__ dayp, const Smem& S) {
    const int n = P.n;
    if (threadIdx.x < 4) S.live[threadIdx.x] = 0;
    if (COPULA == KV_STUDENT)
        for (int k = threadIdx.x; k < LOGTAB_SIZE; k += blockDim.x) S.ltab[k] = P.logtab[k];
    if (COPULA == KV_GAUSSIAN || COPULA == KV_STUDENT)
        for (int k = threadIdx.x; k < EXPTAB_SIZE; k += blockDim.x) S.etab[k] = P.exptab[k];
    if (kv_pow_degree(COPULA) > 0)
        for (int k = threadIdx.x; k < POW_MTAB + POW_ETAB; k += blockDim.x) S.ptab[k] = P.powtab[k];
    if (kv_pow_degree(COPULA) > 0 && POW_FAST_MODE > 0)
        for (int k = threadIdx.x; k < P.pow_octaves * POW_MTAB * POW_FAST_ENTRY_DOUBLES; k += blockDim.x) S.pfast[k] = P.powfast[k];
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        S.xs[i] = P.x[i];
        double f0, s0, f1, s1;
        int dead0, dead1;
        axis_point<COPULA>(P, dayp, 0, i, f0, s0, dead0);
        axis_point<COPULA>(P, dayp, 1, i, f1, s1, dead1);
        S.in[i] = make_double2(f1, s1);
        S.out0[i] = f0;
        S.out1[i] = s0;
        if (dead0) atomicAdd(&S.live[dead0 - 1], 1);
        if (dead1) atomicAdd(&S.live[2 + dead1 - 1], 1);
    }
    __syncthreads();
    if (kv_pow_degree(COPULA) > 0 && POW_FAST_MODE > 0) {
        if (P.pow_octaves == 0) return false;
        // the quadratic form is convex along a row, so its maximum over the live columns sits at one of their ends
        const int j_lo = S.live[2], j_hi = n - S.live[3];
        int slow = 0;
        if (j_hi > j_lo) {
            const double a_lo = S.in[j_lo].x, a_hi = S.in[j_hi - 1].x;
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const double y0 = S.out0[i];
                const double m0 = P.g_out_scale * y0, c0 = fma(y0 * y0, P.inv_nu, 1.0);
                const double d = fmax(fabs(a_lo - m0), fabs(a_hi - m0));
                slow |= !(fma(d, d, c0) < P.pow_fast_limit);
            }
        }
        return __syncthreads_or(slow) == 0;
    }
    return false;
}

// per-plan tables of the mixture marginals: one thread per (asset, state, axis point)
__global__ void state_table_kernel(int n, int q, const double* __restrict__ x, const double* __restrict__ sigma_states,
                                   double* __restrict__ cdf, double* __restrict__ pdf) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 2LL * q * n) return;
    const int i = (int)(idx % n);
    const int as = (int)(idx / n);  // a * q + s
    const double sg = sigma_states[as];
    const double z = __ddiv_rn(x[i], sg);
    cdf[idx] = phi_via_erf(z);                                        // utils/utils.py:17-22
    pdf[idx] = exp(-0.5 * z * z) / (2.5066282746310002 * sg);         // msm_estimation.py:328
}

// ---------------------------------------------------------------------------------------------
// membership: #{ j : x[j] <= g } by binary search inside [lo, hi]
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double inner_bound(double q, double xi, double w0, double w1, double rw0_exact) {
    // (q - x_outer * w[1]) / w[0], every operation individually rounded (integration_algo.py:20).  When w[0] is a
    // power of two (the reference's default 0.5) the division equals the multiplication by its exact reciprocal bit
    // for bit, which spares the ~20-instruction IEEE division per row and strip; rw0_exact is 0 otherwise.
    const double num = __dsub_rn(q, __dmul_rn(xi, w1));
    return rw0_exact != 0.0 ? __dmul_rn(num, rw0_exact) : __ddiv_rn(num, w0);
}

__device__ __forceinline__ int count_le(const double* __restrict__ xs, double g, int lo, int hi) {
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (xs[mid] <= g)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

// The same count when BOTH bracket ends are counts at bracketing levels (lo < hi, the bisection passes): the level
// sought is the midpoint of the two, so on a (locally) uniform axis the count is the midpoint of the bracket give or
// take one.  Two comparisons settle it in the common case; rows whose bracket straddles a change of spacing fall back
// to the bisection on what is left.  Exact like count_le: every decision is a comparison with an axis value.
__device__ __forceinline__ int count_between(const double* __restrict__ xs, double g, int lo, int hi) {
    int k = (lo + hi + 1) >> 1;   // lo < k <= hi
    const double below = xs[k - 1], above = xs[k < hi ? k : k - 1];   // two independent loads
    if (below <= g) {
        if (k < hi && above <= g) {
            ++k;
            if (k < hi && xs[k] <= g) k = count_le(xs, g, k + 1, hi);
        }
    } else {
        --k;
        if (k > lo && !(xs[k - 1] <= g)) k = count_le(xs, g, lo, k - 1);
    }
    return k;
}

// Row ownership.  Rows are dealt to warps in blocks of 32 (lane = row within the block, so adjacent lanes walk
// adjacent rows); the block-to-warp order alternates direction every round (0..W-1, W-1..0, ...) because
// the strip length falls off monotonically with the row index and a fixed order would always hand warp 0
// the longest rows.  The mapping is fixed for the whole solve: a thread only ever touches its own rows'
// boundary entries, which is why no barrier separates the boundary search from the cell walk.
//
// A day can also be split over a thread-block CLUSTER (small batches, see solve_kernel): the warps of all CTAs of the
// cluster then form one pool (`Part`), every CTA owns the rows of its warps, and the strip masses are combined
// through distributed shared memory (block_reduce).
struct Part {
    int rank, size;  // this CTA's rank in the cluster and the cluster size (0, 1 without a cluster)
    int rounds;      // rows per thread: ceil(n / (size * blockDim.x)), computed once (an integer division)
};
__device__ __forceinline__ Part make_part(int rank, int size, int n) {
    const int per = size * (int)blockDim.x;
    return Part{rank, size, (n + per - 1) / per};
}

__device__ __forceinline__ int owned_row(const Part& pt, int m) {
    const int w = pt.rank * (blockDim.x >> 5) + (threadIdx.x >> 5), l = threadIdx.x & 31;
    const int nw = pt.size * (blockDim.x >> 5);
    const int wb = (m & 1) ? (nw - 1 - w) : w;
    return ((m * nw + wb) << 5) + l;
}
__device__ __forceinline__ int owned_rounds(const Part& pt, int) { return pt.rounds; }

// c = max(#{x <= g_i(q)}, cmin) for outer row i; the search is confined to [lo, hi]
__device__ __forceinline__ int count_row(const KernelParams& P, const Smem& S, double q, int i, int lo, int hi) {
    const double g = inner_bound(q, S.xs[i], P.w0, P.w1, P.rw0_exact);
    // Stored bounds are counts at a smaller q, clipped at cmin.  With w0 > 0 the bound grows with q, so the count
    // sought is >= the stored one or both lie below cmin (and the result is clipped to cmin either way): [lo, hi]
    // always contains the answer.  Only a negative w0 reverses the order, and then the bracket is re-opened.
    if (P.w0 < 0.0 && lo > 0 && !(S.xs[lo - 1] <= g)) lo = 0;
    int k;
    if (P.nseg > 0 && hi - lo > 4) {
        // Piecewise-uniform axis: the count is known to +-1 from the segment's spacing, and the two comparisons that
        // settle it replace a chain of log2(hi - lo) dependent shared-memory loads.  The result is still decided by
        // comparisons with the axis values themselves, so it is exact whatever the quality of the guess.
        int sgm = 0;
#pragma unroll
        for (int t = 1; t < MAX_AXIS_SEGMENTS; ++t) sgm += (g >= P.seg_x0[t]) ? 1 : 0;
        const int first = P.seg_first[sgm], next = P.seg_first[sgm + 1];
        k = first + __double2int_rd((g - P.seg_x0[sgm]) * P.seg_inv_h[sgm]) + 1;
        k = min(max(k, first), next);
        k = min(max(k, lo), hi);
        while (k > lo && !(S.xs[k - 1] <= g)) --k;
        while (k < hi && S.xs[k] <= g) ++k;
    } else {
        k = count_le(S.xs, g, lo, hi);
    }
    return max(k, P.cmin);
}

// ctarget[i] = count_row(q) for every outer row this thread owns
__device__ __forceinline__ void count_rows(const KernelParams& P, const Smem& S, const Part& pt, double q, u16* ctarget,
                                           const u16* slo, const u16* shi) {
    const int n = P.n;
    for (int m = 0; m < owned_rounds(pt, n); ++m) {
        const int i = owned_row(pt, m);
        if (i < n) ctarget[i] = (u16)count_row(P, S, q, i, slo ? (int)slo[i] : 0, shi ? (int)shi[i] : n);
    }
}

// ---------------------------------------------------------------------------------------------
// cell weights
// ---------------------------------------------------------------------------------------------
template <int COPULA>
struct Row;  // per-outer-row constants held in registers

template <>
struct Row<0> {  // Gaussian:  W = rowfac * 2^( l1[j] - (y1'[j] - m0)^2 )
    double m0, fac;
    __device__ __forceinline__ void load(const KernelParams&, const Smem& S, int i) {
        m0 = S.out0[i];
        fac = S.out1[i];
    }
    __device__ __forceinline__ double cell(const KernelParams&, const Smem& S, double a, double b) const {
        const double d = a - m0;
        return exp2_tab(fma(-d, d, b), S.etab);
    }
    template <bool FAST>
    __device__ __forceinline__ double add_cell(const KernelParams&, const Smem& S, double a, double b, double acc) const {
        const double d = a - m0;
        return exp2_tab_add(fma(-d, d, b), S.etab, acc);
    }
    __device__ __forceinline__ double quad_form(double) const { return 0.0; }
};

template <>
struct Row<1> {  // Student-t: W = rowfac * 2^( l1[j] - (nu+2)/2 * log2( c0 + (y1'[j] - m0)^2 ) ), table-assisted log2
    double m0, c0, fac;
    __device__ __forceinline__ void load(const KernelParams& P, const Smem& S, int i) {
        const double y0 = S.out0[i];
        m0 = P.g_out_scale * y0;
        c0 = fma(y0 * y0, P.inv_nu, 1.0);
        fac = S.out1[i];
    }
    __device__ __forceinline__ double cell(const KernelParams& P, const Smem& S, double a, double b) const {
        const double d = a - m0;
        const double t = fma(d, d, c0);  // >= 1
        return exp2_tab(scaled_log2_plus(t, b, P.negc, P.qc, S.ltab), S.etab);
    }
    template <bool FAST>
    __device__ __forceinline__ double add_cell(const KernelParams& P, const Smem& S, double a, double b, double acc) const {
        const double d = a - m0;
        const double t = fma(d, d, c0);
        return exp2_tab_add(scaled_log2_plus(t, b, P.negc, P.qc, S.ltab), S.etab, acc);
    }
    __device__ __forceinline__ double quad_form(double) const { return 0.0; }
};

template <int DEG>
struct RowStudentPow {  // Student-t: W = rowfac * A1[j] * ( c0 + (y1'[j] - m0)^2 )^(-(nu+2)/2), table-assisted power
    double m0, c0, fac;
    __device__ __forceinline__ void load(const KernelParams& P, const Smem& S, int i) {
        const double y0 = S.out0[i];
        m0 = P.g_out_scale * y0;
#ifndef CVAR_NO_OPAQUE_M0
        asm volatile("" : "+d"(m0));   // keep the product: folded into the cells' a - m0 it costs a constant load per trip
#endif
        c0 = fma(y0 * y0, P.inv_nu, 1.0);
        fac = S.out1[i];
    }
    __device__ __forceinline__ double cell(const KernelParams& P, const Smem& S, double a, double b) const {
        const double d = a - m0;
        const double t = fma(d, d, c0);  // >= 1
        return pow_neg_c<DEG>(t, b, P.powc, S.ptab_s, P.seed_mask, P.seed_half);
    }
    template <bool FAST>
    __device__ __forceinline__ double add_cell(const KernelParams& P, const Smem& S, double a, double b, double acc) const {
        if (FAST) {
            const double d = a - m0;
            return acc + pow_neg_c_fast<DEG>(fma(d, d, c0), b, P.powc, S.pfast_s, P.seed_mask, P.seed_half);
        }
        return acc + cell(P, S, a, b);   // contracts to one FMA with the last product of pow_neg_c
    }
    __device__ __forceinline__ double quad_form(double a) const {
        const double d = a - m0;
        return fma(d, d, c0);
    }
};
template <> struct Row<KV_STUDENT_POW5> : RowStudentPow<5> {};
template <> struct Row<KV_STUDENT_POW6> : RowStudentPow<6> {};
template <> struct Row<KV_STUDENT_POW7> : RowStudentPow<7> {};
template <> struct Row<KV_STUDENT_POW8> : RowStudentPow<8> {};

template <>
struct Row<2> {  // Plackett (the reference's formula, plackett.py:66-69), u = row, v = column
    double n0, n1, qa, qb, qc, fac;
    __device__ __forceinline__ void load(const KernelParams& P, const Smem& S, int i) {
        const double u = S.out0[i];
        const double eta = P.theta - 1.0;
        const double p0 = 1.0 + eta * u;
        const double r0 = 1.0 + eta * (1.0 - u);
        n0 = P.theta * p0;
        n1 = P.theta * eta * (1.0 - 2.0 * u);
        // (p0 + eta v)(r0 - eta v) as one quadratic in v: two FMAs per cell instead of two FMAs and a multiply
        qa = -eta * eta;
        qb = eta * (r0 - p0);
        qc = p0 * r0;
        fac = S.out1[i];
    }
    __device__ __forceinline__ double cell(const KernelParams&, const Smem&, double v, double a1) const {
        const double num = fma(n1, v, n0);
        const double dd = fma(v, fma(v, qa, qb), qc);
        return (a1 * num) * rcp_cell(dd * dd);
    }
    template <bool FAST>
    __device__ __forceinline__ double add_cell(const KernelParams& P, const Smem& S, double v, double a1, double acc) const {
        return acc + cell(P, S, v, a1);
    }
    __device__ __forceinline__ double quad_form(double) const { return 0.0; }
};

struct StripResult {
    double mass;
    unsigned cells;
    bool poisoned;
    bool redo;   // strip_pass_thin met a row it cannot handle: the caller repeats the pass with strip_pass
};

struct Live {  // live window of rows / columns (cells outside have an infinite copula quantile)
    int i_lo, i_hi, j_lo, j_hi;
    bool full;      // the window is the whole grid (always, on mixture marginals whose tails do not saturate)
    bool day_fast;  // Student-t power variants: every cell of the day is in the range of the one-lookup table
};

// block-wide (cluster-wide when the day is split over a cluster) deterministic sum; every thread of every CTA
// receives the same value
__device__ __forceinline__ StripResult block_reduce(const Smem& S, const Part& pt, int& parity, double v, unsigned cells,
                                                    bool poison, bool redo = false) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    cells = __reduce_add_sync(0xffffffffu, cells);
    if (lane == 0) {
        S.red[parity * MAX_CTA_WARPS + warp] = v;
        S.redc[parity * MAX_CTA_WARPS + warp] = cells;
    }
    const int anyp = __syncthreads_or((poison ? 1 : 0) | (redo ? 2 : 0));
    StripResult r;
    r.mass = 0.0;
    r.cells = 0;
    const int nw = blockDim.x >> 5;
    for (int w = 0; w < nw; ++w) {   // fixed order: the sum does not depend on scheduling
        r.mass += S.red[parity * MAX_CTA_WARPS + w];
        r.cells += S.redc[parity * MAX_CTA_WARPS + w];
    }
    r.poisoned = (anyp & 1) != 0;
    r.redo = (anyp & 2) != 0;
    if (pt.size > 1) {
        // publish this CTA's partial, then read every CTA's slot in rank order through distributed shared memory.
        // The slots are double-buffered by `parity`: a slot is rewritten two strips later, after the cluster barrier
        // of the strip in between, which every CTA only passes once it has read the current values.
        namespace cg = cooperative_groups;
        cg::cluster_group cluster = cg::this_cluster();
        if (threadIdx.x == 0) {
            S.cl_mass[parity] = r.poisoned ? NAN : r.mass;   // a poisoned strip is NaN for the whole cluster
            S.cl_cells[parity] = r.cells;
        }
        cluster.sync();
        double tot = 0.0;
        unsigned cel = 0;
        for (int k = 0; k < pt.size; ++k) {
            tot += cluster.map_shared_rank(S.cl_mass, k)[parity];
            cel += cluster.map_shared_rank(S.cl_cells, k)[parity];
        }
        r.mass = tot;
        r.cells = cel;
        r.poisoned = tot != tot;
    }
    parity ^= 1;
    return r;
}

// One strip of the bisection.
//
// Every thread owns a fixed set of outer rows (owned_row) for the whole solve: it finds the row's new boundary
// index by an exact search (when q_new is given), then walks the row's cells [a, b) itself with
// CELLS_IN_FLIGHT independent cells per trip.  Adjacent lanes own adjacent rows, whose ranges are shifted
// by about one column, so the 16-byte shared-memory loads of a warp fall on consecutive addresses.
// There is no per-strip barrier besides the one inside the block reduction, and the boundary arrays are
// only ever touched by their owning thread.
//
// Blocks of long rows of the Plackett family take the column sweep below instead (sweep_rows).

// Sum of the cells [s, e) of one row.
template <int COPULA, bool FAST>
__device__ __forceinline__ double walk_row(const KernelParams& P, const Smem& S, const Row<COPULA>& row, int s, int e) {
    constexpr int CIF = CellsInFlight<COPULA>::value;
#ifdef CVAR_EXPERIMENT_NO_CELLS   // timing experiment: everything but the cell loops (results are meaningless)
    return 1e-9 * (double)(e - s);
#endif
    double acc[CIF];
#pragma unroll
    for (int c = 0; c < CIF; ++c) acc[c] = 0.0;
    int j = s;
    for (; j + CIF <= e; j += CIF) {
        double2 v[CIF];
#pragma unroll
        for (int c = 0; c < CIF; ++c) v[c] = S.in[j + c];
#pragma unroll
        for (int c = 0; c < CIF; ++c) acc[c] = row.template add_cell<FAST>(P, S, v[c].x, v[c].y, acc[c]);
    }
    for (; j < e; ++j) {
        const double2 v = S.in[j];
        acc[0] = row.template add_cell<FAST>(P, S, v.x, v.y, acc[0]);
    }
    double rowsum = acc[0];
#pragma unroll
    for (int c = 1; c < CIF; ++c) rowsum += acc[c];
    return rowsum;
}

// Column sweep for blocks of LONG rows.  When every one of the warp's 32 rows holds at least SWEEP_MIN_ROW cells, groups
// of G adjacent rows read the same column per trip (their 16-byte words steered into distinct banks, the groups skewed
// so that they start together): a warp-wide column load then costs two shared-memory wavefronts instead of four (micro-
// benchmark tools/micro/lds_patterns.cu).  Each lane walks what is left of its row at both ends by itself.
// Measured per copula (1000 days, n = 2048; 2 alphas at n = 4096): the Plackett cell -- 8 FP64 and one column load per
// cell, nothing else -- gains 3.1 % with pairs of rows (G = 2: 2.624 -> 2.542 ms; n = 4096: 9.30 -> 8.49 ms, -8.7 %) and 2.3 %
// with G = 8; the Gaussian cell gains 7.3 % at n = 4096 (11.37 -> 10.53 ms with G = 2, 10.60 with G = 8); the Student-t cell,
// whose loop is bound by issue slots (18 instructions per cell), loses 1-2 % with any G (c3 1.776 -> 1.79-1.81 ms, n = 4096
// 10.48 -> 10.57-10.72 ms), so it keeps the plain walk.  Override with -DCVAR_SWEEP_<FAMILY>=0|2|4|8|16|32.
#ifndef CVAR_SWEEP_GAUSSIAN
#define CVAR_SWEEP_GAUSSIAN 2
#endif
#ifndef CVAR_SWEEP_STUDENT
#define CVAR_SWEEP_STUDENT 0
#endif
#ifndef CVAR_SWEEP_PLACKETT
#define CVAR_SWEEP_PLACKETT 2
#endif
#ifndef CVAR_SWEEP_MIN_ROW
#define CVAR_SWEEP_MIN_ROW 128
#endif
template <int COPULA>
struct SweepGroup {
    static constexpr int value = COPULA == 0 ? CVAR_SWEEP_GAUSSIAN : (kv_is_student(COPULA) ? CVAR_SWEEP_STUDENT : CVAR_SWEEP_PLACKETT);
};

template <int COPULA, bool FAST>
__device__ __forceinline__ double sweep_rows(const KernelParams& P, const Smem& S, const Row<COPULA>& row, int s, int e) {
    constexpr int CIF = CellsInFlight<COPULA>::value;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int G = SweepGroup<COPULA>::value > 0 ? SweepGroup<COPULA>::value : 32, NG = 32 / G;
    double acc[CIF];
#pragma unroll
    for (int c = 0; c < CIF; ++c) acc[c] = 0.0;
    int off = 0;   // this lane reads column (trip - off)
    if (G < 32) {
        const int lane = threadIdx.x & 31, g = lane / G;
        off = __shfl_sync(FULL, s, 0) - __shfl_sync(FULL, s, lane & ~(G - 1));
        if (NG <= 8)
            off += (g - off) & (NG - 1);
        else   // 16 pairs: their words must fall on every 16-byte bank group exactly twice (offsets 2g + g/4 modulo 8)
            off += ((2 * g + (g >> 2)) - off) & 7;
    }
    const int t0 = __reduce_max_sync(FULL, s + off);
    const int t1 = __reduce_min_sync(FULL, e + off);
    int left_end = s, right_begin = s;
    if (t1 - t0 >= 2 * CIF) {
        const int trips = (t1 - t0) / CIF;
        const double2* col = S.in + (t0 - off);
        for (int k = 0; k < trips; ++k, col += CIF) {
            double2 v[CIF];
#pragma unroll
            for (int c = 0; c < CIF; ++c) v[c] = col[c];
#pragma unroll
            for (int c = 0; c < CIF; ++c) acc[c] = row.template add_cell<FAST>(P, S, v[c].x, v[c].y, acc[c]);
        }
        left_end = t0 - off;
        right_begin = t0 - off + trips * CIF;
    }
#pragma unroll 1
    for (int seg = 0; seg < 2; ++seg) {   // the ends are short: a plain loop
        const int je = seg ? e : left_end;
        for (int j = seg ? right_begin : s; j < je; ++j) {
            const double2 v = S.in[j];
            acc[0] = row.template add_cell<FAST>(P, S, v.x, v.y, acc[0]);
        }
    }
    double rowsum = acc[0];
#pragma unroll
    for (int c = 1; c < CIF; ++c) rowsum += acc[c];
    return rowsum;
}

// cnew[i] = count(q_new) for every owned row, searched inside [slo[i], shi[i]] (nullptr: open end), then the cells
// [ca[i], cb[i]) of the row are summed; ca == nullptr is the constant lower end cmin, and ca / cb may alias cnew, slo or
// shi -- the values then come from registers instead of a shared-memory round trip.
// BISECT = true is the pass of a bisection step with positive w0: both bracket ends exist, the boundary sought lies between
// them (midpoint search) and the strip is the lower (ca == slo) or the upper half of the bracket -- known at compile time,
// which spares every row the pointer comparisons and selects of the general form.
template <int COPULA, bool BISECT = false>
__device__ StripResult strip_pass(const KernelParams& P, const Smem& S, const Part& pt, const Live& L, int& parity,
                                  double q_new, u16* cnew, const u16* slo, const u16* shi, const u16* ca,
                                  const u16* cb, bool poison_mode) {
    const int n = P.n;
    double total = 0.0;
    unsigned cells = 0;
    bool poison = false;
    const bool between = BISECT || (slo && shi && P.w0 > 0.0);   // both bracket ends known and ordered: midpoint search
    const bool lower = ca == slo;
    for (int m = 0; m < owned_rounds(pt, n); ++m) {
        const int i = owned_row(pt, m);
        int s = 0, e = 0;
        if (i < n) {
            const double xi = S.xs[i];
            const int lo = (BISECT || slo) ? (int)slo[i] : 0, hi = (BISECT || shi) ? (int)shi[i] : n;
            int k = lo;
            if (BISECT) {
                if (lo != hi) {   // else: no grid point of this row between the bracket ends, nothing can move
                    k = max(count_between(S.xs, inner_bound(q_new, xi, P.w0, P.w1, P.rw0_exact), lo, hi), P.cmin);
                    s = lower ? lo : k;
                    e = lower ? k : hi;
                }
            } else if (!(slo && shi && lo == hi)) {
                if (between)
                    k = max(count_between(S.xs, inner_bound(q_new, xi, P.w0, P.w1, P.rw0_exact), lo, hi), P.cmin);
                else
                    k = count_row(P, S, q_new, i, lo, hi);
                s = !ca ? P.cmin : ca == cnew ? k : ca == slo ? lo : ca == shi ? hi : (int)ca[i];
                e = cb == cnew ? k : cb == slo ? lo : cb == shi ? hi : (int)cb[i];
            }
            cnew[i] = (u16)k;
#ifdef CVAR_DEBUG_ASSERT
            if (s < 0 || e > n || k > n) __trap();
#endif
            if (e > s) {
                cells += (unsigned)(e - s);
                if (!L.full) {
                    if (i < L.i_lo || i >= L.i_hi) {
                        poison = true;
                        e = s;
                    } else {
                        if (s < L.j_lo || e > L.j_hi) poison = true;
                        s = max(s, L.j_lo);
                        e = min(e, L.j_hi);
                    }
                }
            }
        }
        if (SweepGroup<COPULA>::value > 0 && __all_sync(0xffffffffu, e - s >= CVAR_SWEEP_MIN_ROW)) {   // every lane arrives here
            Row<COPULA> srow;
            srow.load(P, S, i);
            bool sfast = false;
            if (kv_pow_degree(COPULA) > 0 && POW_FAST_MODE > 0 && P.pow_octaves > 0) {
                sfast = L.day_fast;
                if (!sfast)
                    sfast = !__any_sync(0xffffffffu, !(fmax(srow.quad_form(S.in[s].x), srow.quad_form(S.in[e - 1].x)) < P.pow_fast_limit));
            }
            const double ssum = (kv_pow_degree(COPULA) > 0 && POW_FAST_MODE > 0 && sfast)
                                    ? sweep_rows<COPULA, true>(P, S, srow, s, e) : sweep_rows<COPULA, false>(P, S, srow, s, e);
            total = fma(srow.fac, ssum, total);
            continue;
        }
        if (e <= s) continue;
        Row<COPULA> row;
        row.load(P, S, i);
        double rowsum;
        bool fast = false;
        if (kv_pow_degree(COPULA) > 0 && POW_FAST_MODE > 0 && P.pow_octaves > 0) {
            // The quadratic form is convex along the row: check the two ends of this strip's range.  The lanes walking
            // rows together take the same form of the cell (both forms give the same bits; a warp split between the two
            // loops would run them one after the other).
            fast = L.day_fast;
            if (!fast)
                fast = !__any_sync(__activemask(), !(fmax(row.quad_form(S.in[s].x), row.quad_form(S.in[e - 1].x)) < P.pow_fast_limit));
        }
        if (kv_pow_degree(COPULA) > 0 && POW_FAST_MODE > 0 && fast)
            rowsum = walk_row<COPULA, true>(P, S, row, s, e);
        else
            rowsum = walk_row<COPULA, false>(P, S, row, s, e);
        total = fma(row.fac, rowsum, total);
    }
    StripResult r = block_reduce(S, pt, parity, total, cells, poison && poison_mode);
    if (r.poisoned) r.mass = NAN;
    return r;
}

// A pass of the late bisection steps, where the bracket is narrower than the finest axis spacing: every row's bracket
// [lo, hi] holds at most ONE grid point (hi <= lo + 1), so the new boundary is one comparison and the strip holds one
// cell of the row or none.  Straight-line code, a third of the instructions of the general pass for such rows (the
// bookkeeping of a pass is bound by instruction issue).  `lower`: the strip is the lower half
// [lo, mid) of the bracket, else the upper half [mid, hi).  A row with hi > lo + 1 (cannot happen for a bracket below
// thin_width; checked all the same) makes the whole CTA repeat the pass with strip_pass.
template <int COPULA>
__device__ StripResult strip_pass_thin(const KernelParams& P, const Smem& S, const Part& pt, const Live& L, int& parity,
                                       double q_new, u16* cnew, const u16* slo, const u16* shi, bool lower, bool poison_mode) {
    const int n = P.n;
    double total = 0.0;
    unsigned cells = 0;
    bool poison = false, redo = false;
    for (int m = 0; m < owned_rounds(pt, n); ++m) {
        const int i = owned_row(pt, m);
        if (i >= n) continue;
        const int lo = (int)slo[i], hi = (int)shi[i];
        int k = lo;
        if (lo != hi) {
            if (hi - lo > 1) redo = true;
            const double g = inner_bound(q_new, S.xs[i], P.w0, P.w1, P.rw0_exact);
            k = (S.xs[lo] <= g) ? hi : lo;   // counts are clipped at cmin already: lo >= cmin
            if (lower ? (k == hi) : (k == lo)) {   // the strip holds the cell (i, lo)
                cells += 1u;
                bool live = true;
                if (!L.full) {
                    live = i >= L.i_lo && i < L.i_hi && lo >= L.j_lo && lo < L.j_hi;
                    poison = poison || !live;
                }
                if (live) {
                    Row<COPULA> row;
                    row.load(P, S, i);
                    const double2 v = S.in[lo];
                    bool fast = false;
                    if (kv_pow_degree(COPULA) > 0 && POW_FAST_MODE > 0 && P.pow_octaves > 0)
                        fast = L.day_fast || row.quad_form(v.x) < P.pow_fast_limit;
                    const double w = (kv_pow_degree(COPULA) > 0 && POW_FAST_MODE > 0 && fast)
                                         ? row.template add_cell<true>(P, S, v.x, v.y, 0.0)
                                         : row.template add_cell<false>(P, S, v.x, v.y, 0.0);
                    total = fma(row.fac, w, total);
                }
            }
        }
        cnew[i] = (u16)k;
    }
    StripResult r = block_reduce(S, pt, parity, total, cells, poison && poison_mode, redo);
    if (r.poisoned) r.mass = NAN;
    return r;
}

// strip_pass with reuse across the alphas of a day.  `memo_n` (uniform across the CTA) counts the entries written
// so far (by thread 0, right after a strip's reduction); only the first `memo_visible` of them -- those of
// earlier alphas, published by the barrier at the top of the alpha loop -- are searched.
template <int COPULA>
__device__ __forceinline__ StripResult strip_memo(const KernelParams& P, const Smem& S, const Part& pt, const Live& L, int& parity,
                                                  bool use_memo, bool remember, int memo_visible, int& memo_n,
                                                  double a, double b,
                                                  double q_new, u16* cnew, const u16* slo, const u16* shi,
                                                  const u16* ca, const u16* cb, bool poison_mode,
                                                  unsigned long long& evaluated, bool thin = false, bool bisect = false) {
    if (use_memo) {
        for (int k = 0; k < memo_visible; ++k) {
            const double* e = S.memo + 4 * k;
            if (e[0] == a && e[1] == b) {
                count_rows(P, S, pt, q_new, cnew, slo, shi);  // the boundary indices are still needed downstream
                StripResult r;
                r.mass = e[2];
                r.cells = (unsigned)e[3];
                r.poisoned = e[2] != e[2];
                return r;
            }
        }
    }
    StripResult r;
    r.redo = true;
    if (thin) r = strip_pass_thin<COPULA>(P, S, pt, L, parity, q_new, cnew, slo, shi, ca == slo, poison_mode);
    if (r.redo) {
        if (bisect)
            r = strip_pass<COPULA, true>(P, S, pt, L, parity, q_new, cnew, slo, shi, ca, cb, poison_mode);
        else
            r = strip_pass<COPULA, false>(P, S, pt, L, parity, q_new, cnew, slo, shi, ca, cb, poison_mode);
    }
    evaluated += r.cells;
    if (use_memo && remember && memo_n < MEMO_SIZE) {
        if (threadIdx.x == 0) {
            double* e = S.memo + 4 * memo_n;
            e[0] = a; e[1] = b; e[2] = r.mass; e[3] = (double)r.cells;
        }
        ++memo_n;
    }
    return r;
}

// Bit 31 of a solve's first trajectory word: during the bisection a running mass cancelled to rounding noise (twelve or
// more digits lost) without being exactly 0.  In exact arithmetic such a mass IS 0; whether the reference's sums cancel
// exactly -- which decides its "every day's mass is 0" exit (calc_var_class.py:293-295) -- depends on the order in which
// it happened to add the cells.  finalize reports a batch in which every day is in that state (status word below).
constexpr unsigned TRAJ_FRAGILE_BIT = 0x80000000u;
constexpr int STATUS_ZERO_EXIT_TAKEN = 1;       // K was cut because every day's running mass was exactly 0 at iteration K
constexpr int STATUS_ZERO_EXIT_AMBIGUOUS = 2;   // every day's mass was 0 exactly or up to rounding at some iteration: the
                                                // reference's outcome for this batch depends on its summation order

// ---------------------------------------------------------------------------------------------
// the solve kernel
// ---------------------------------------------------------------------------------------------
// CLUSTER = true: launched with a cluster dimension of 2 or 4 -- the CTAs of a cluster share one day (each runs
// stage 0 for itself, owns a share of the outer rows and sees the same strip masses, hence takes the same decisions).
template <int COPULA, bool CLUSTER>
__global__ void __launch_bounds__(CTA_THREADS_LARGE, 1)
solve_kernel(KernelParams P, const double* __restrict__ day_params, long long day0, long long T, AlphaSet A,
             const int* __restrict__ order, unsigned* __restrict__ traj, double* __restrict__ mass_out,
             unsigned long long* __restrict__ cells_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem S = carve(smem_raw, P.n, COPULA, P.pow_octaves);
    Part pt = make_part(0, 1, P.n);
    if (CLUSTER) {
        cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
        pt = make_part((int)cluster.block_rank(), (int)cluster.num_blocks(), P.n);
    }
    // CTAs are dispatched in block-index order; `order` lists the days most expensive first (see order_key_kernel)
    // (the grid covers one chunk of the batch, whose first day is day0; `order` indexes within the chunk)
    const long long unit = blockIdx.x / pt.size;
    const long long day = day0 + (order ? order[unit] : unit);
    const int stride = (P.marginal == 0) ? 2 : 2 * P.q;
#ifdef CVAR_PROFILE_PHASES   // experiment: thread 0's clock per phase goes out through traj / mass / cells (results are overwritten)
    const long long prof_t0 = clock64();
    long long prof_thin = 0, prof_probe = 0;
#endif
#ifdef CVAR_PROFILE_TIMELINE   // experiment: where and when each CTA ran (results are overwritten): tools/timeline_profile.py
    unsigned long long tl_start;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tl_start));
#endif
    const bool day_fast = stage0<COPULA>(P, day_params + day * stride, S);
#ifdef CVAR_PROFILE_PHASES
    const long long prof_t1 = clock64();
#endif

    Live L;
    L.day_fast = day_fast;
    if (COPULA == 2) {
        L.i_lo = 0; L.i_hi = P.n; L.j_lo = 0; L.j_hi = P.n;
    } else {
        L.i_lo = S.live[0]; L.i_hi = P.n - S.live[1];
        L.j_lo = S.live[2]; L.j_hi = P.n - S.live[3];
    }
    L.full = L.i_lo == 0 && L.i_hi == P.n && L.j_lo == 0 && L.j_hi == P.n;
    // Q5: NaN cells are zeroed on the single-normal path, poison the strip on the mixture path
    const bool poison_mode = (COPULA != 2) && (P.marginal == 1) && ((P.compat & 4u) != 0);
    int parity = 0;

    unsigned long long evaluated = 0;   // cells this CTA's cluster really evaluated (all alphas)
    bool have_f3 = false;
    StripResult f3 = {0.0, 0u, false};
    const bool use_memo = A.n_alpha > 1;
    int memo_n = 0;

    for (int ia = 0; ia < A.n_alpha; ++ia) {
        const double alpha = A.a[ia];
        unsigned long long ncell = 0;
        if (ia > 0) __syncthreads();  // memo entries of the previous alphas are visible from here on
        const int memo_visible = memo_n;
        // --- probe 1: F(first)  (calc_var_class.py:114-119)
        if (!have_f3) {
            f3 = strip_pass<COPULA>(P, S, pt, L, parity, P.first, S.c[0], nullptr, nullptr, nullptr, S.c[0], poison_mode);
            have_f3 = true;
            evaluated += f3.cells;
        } else {
            count_rows(P, S, pt, P.first, S.c[0], nullptr, nullptr);
        }
        ncell += f3.cells;
        // --- probe 2  (:125-142)
        double lo2, hi2;
        if (f3.mass >= alpha) { lo2 = P.second_lo; hi2 = P.first; } else { lo2 = P.first; hi2 = P.second_hi; }
        double prev_upper = (lo2 == P.second_lo) ? P.second_lo : P.first;  // Q6
        StripResult s2;
        if (lo2 == P.first)   // strip (first, second_hi]: new upper boundary, searched above c[0]
            s2 = strip_memo<COPULA>(P, S, pt, L, parity, use_memo, true, memo_visible, memo_n, lo2, hi2, hi2, S.c[1], S.c[0], nullptr,
                                    S.c[0], S.c[1], poison_mode, evaluated);
        else                  // strip (second_lo, first]: new lower boundary, searched below c[0]
            s2 = strip_memo<COPULA>(P, S, pt, L, parity, use_memo, true, memo_visible, memo_n, lo2, hi2, lo2, S.c[1], nullptr, S.c[0],
                                    S.c[1], S.c[0], poison_mode, evaluated);
        ncell += s2.cells;
        double R = (lo2 == P.first) ? f3.mass + s2.mass : f3.mass - s2.mass;
        if ((P.compat & 2u) == 0 && lo2 == P.first) prev_upper = hi2;  // intended behaviour: R = F(hi2)
        // --- bracket  (:147-155)
        int kase;
        double lo, hi;
        u16 *cl, *ch, *cm;
        if (R > alpha && hi2 == P.second_hi) {
            kase = 3; lo = P.first; hi = P.second_hi; cl = S.c[0]; ch = S.c[1]; cm = S.c[2];
        } else if (R > alpha) {
            kase = 0; lo = P.min_var; hi = P.second_lo; ch = S.c[1]; cl = S.c[2]; cm = S.c[0];
            count_rows(P, S, pt, lo, cl, nullptr, ch);
        } else if (R < alpha && hi2 == P.first) {
            kase = 1; lo = P.second_lo; hi = P.first; cl = S.c[1]; ch = S.c[0]; cm = S.c[2];
        } else if (R < alpha && hi2 == P.second_hi) {
            kase = 2; lo = P.second_hi; hi = P.max_var; cl = S.c[1]; ch = S.c[2]; cm = S.c[0];
            count_rows(P, S, pt, hi, ch, cl, nullptr);
        } else {
            kase = 4; lo = hi = NAN; cl = ch = cm = S.c[0];
        }
        bool stack = !(hi == P.second_lo || hi == P.second_hi);  // :160
        unsigned dec = 0, zer = 0;
        bool fragile = false;   // some running mass was zero up to rounding, but not exactly (see TRAJ_FRAGILE_BIT)
        if (kase != 4) {
#ifdef CVAR_PROFILE_PHASES
            prof_probe = clock64() - prof_t1;
#endif
            for (int k = 0; k < P.max_iter; ++k) {
#ifdef CVAR_PROFILE_PHASES
                const long long prof_k0 = clock64();
#endif
                const double mid = (lo + hi) / 2;
                const double a = stack ? lo : mid, b = stack ? mid : hi;
                // bracket narrower than the finest axis spacing: at most one grid point per row between its ends
                const bool thin = !CLUSTER && P.thin_width > 0.0 && (hi - lo) < P.thin_width;
                const StripResult s = strip_memo<COPULA>(P, S, pt, L, parity, use_memo, k < MEMO_PER_ALPHA - 1, memo_visible, memo_n, a, b,
                                                         mid, cm, cl, ch, stack ? cl : cm, stack ? cm : ch, poison_mode, evaluated, thin,
                                                         P.w0 > 0.0);
                ncell += s.cells;
                const double r_prev = R;
                R = (a == prev_upper) ? R + s.mass : R - s.mass;  // adjust_integral (:241-246)
                if (R == 0.0) zer |= 1u << k;
                else if (fabs(R) <= 1e-12 * fmax(fabs(r_prev), fabs(s.mass))) fragile = true;
                stack = R < alpha;                                 // :298
                if (stack) { dec |= 1u << k; lo = mid; u16* t = cl; cl = cm; cm = t; }
                else       { hi = mid;       u16* t = ch; ch = cm; cm = t; }
                prev_upper = mid;
#ifdef CVAR_PROFILE_PHASES
                if (k >= 10) prof_thin += clock64() - prof_k0;
#endif
            }
        }
        if (threadIdx.x == 0 && pt.rank == 0) {
            const long long o = (long long)ia * T + day;
            traj[2 * o] = dec | ((unsigned)kase << 28) | (fragile ? TRAJ_FRAGILE_BIT : 0u);
            traj[2 * o + 1] = zer;
            if (mass_out) mass_out[o] = R;
            if (cells_out) cells_out[o] = ncell;
#ifdef CVAR_PROFILE_PHASES
            traj[2 * o] = (unsigned)((prof_t1 - prof_t0) >> 4);   // stage 0
            traj[2 * o + 1] = (unsigned)(prof_thin >> 4);         // bisection passes 10..
            if (mass_out) mass_out[o] = (double)prof_probe;       // probes + bracket set-up
            if (cells_out) cells_out[o] = (unsigned long long)(clock64() - prof_t0);
#endif
#ifdef CVAR_PROFILE_TIMELINE
            unsigned smid;
            unsigned long long tl_end;
            asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tl_end));
            traj[2 * o] = smid;
            traj[2 * o + 1] = blockIdx.x;
            if (mass_out) mass_out[o] = (double)tl_start;          // ns; exact below 2^53
            if (cells_out) cells_out[o] = tl_end;
#endif
        }
    }
    if (P.evaluated_cells && threadIdx.x == 0 && pt.rank == 0) atomicAdd(P.evaluated_cells, evaluated);
    // no CTA may retire while a peer can still read its partial sums
    if (CLUSTER) cooperative_groups::this_cluster().sync();
}

// ---------------------------------------------------------------------------------------------
// strip-mass kernel (parity seam for compute_integral)
// ---------------------------------------------------------------------------------------------
template <int COPULA>
__global__ void __launch_bounds__(CTA_THREADS_LARGE, 1)
strip_mass_kernel(KernelParams P, const double* __restrict__ day_params, const double* __restrict__ bounds,
                  double* __restrict__ out, unsigned long long* __restrict__ cells_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem S = carve(smem_raw, P.n, COPULA, P.pow_octaves);
    const long long day = blockIdx.x;
    const int stride = (P.marginal == 0) ? 2 : 2 * P.q;
    Live L;
    L.day_fast = stage0<COPULA>(P, day_params + day * stride, S);
    if (COPULA == 2) {
        L.i_lo = 0; L.i_hi = P.n; L.j_lo = 0; L.j_hi = P.n;
    } else {
        L.i_lo = S.live[0]; L.i_hi = P.n - S.live[1];
        L.j_lo = S.live[2]; L.j_hi = P.n - S.live[3];
    }
    L.full = L.i_lo == 0 && L.i_hi == P.n && L.j_lo == 0 && L.j_hi == P.n;
    const bool poison_mode = (COPULA != 2) && (P.marginal == 1) && ((P.compat & 4u) != 0);
    int parity = 0;
    const double lo = bounds[2 * day], hi = bounds[2 * day + 1];
    const Part pt = make_part(0, 1, P.n);
    count_rows(P, S, pt, lo, S.c[0], nullptr, nullptr);
    // an inverted pair yields an empty strip (cb <= ca), like the reference's empty np.where
    const StripResult s = strip_pass<COPULA>(P, S, pt, L, parity, hi, S.c[1], nullptr, nullptr, S.c[0], S.c[1], poison_mode);
    if (threadIdx.x == 0) {
        out[day] = s.mass;
        if (cells_out) cells_out[day] = s.cells;
    }
}

// ---------------------------------------------------------------------------------------------
// launch order: the cost of a solve falls with the day's portfolio volatility (calm days end in bracket C,
// whose strips cover ~4x more cells than bracket D's), so days are started in ascending order of a
// variance proxy and the last CTAs of the launch are the cheap ones.  Ordering never changes results.
// ---------------------------------------------------------------------------------------------
__global__ void order_key_kernel(KernelParams P, const double* __restrict__ day_params, long long T, double rho_eff,
                                 float* __restrict__ key, int* __restrict__ idx) {
    const long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= T) return;
    double v[2];
    if (P.marginal == 0) {
        v[0] = day_params[2 * d] * day_params[2 * d];
        v[1] = day_params[2 * d + 1] * day_params[2 * d + 1];
    } else {
        for (int a = 0; a < 2; ++a) {
            double acc = 0.0;
            for (int s = 0; s < P.q; ++s) {
                const double sg = P.sigma_states[a * P.q + s];
                acc += day_params[(2 * d + a) * P.q + s] * sg * sg;
            }
            v[a] = acc;
        }
    }
    const double var_p = P.w0 * P.w0 * v[0] + P.w1 * P.w1 * v[1] + 2.0 * rho_eff * P.w0 * P.w1 * sqrt(v[0] * v[1]);
    key[d] = (float)fmax(var_p, 0.0);
    idx[d] = (int)d;
}

// ---------------------------------------------------------------------------------------------
// finalize: global iteration count (Q7) and VaR reconstruction
// ---------------------------------------------------------------------------------------------
struct FinalizeParams {
    int max_iter;
    int need[4];  // iterations each bracket needs to reach the tolerance (A, B, C, D)
    double lo[4], hi[4];
    double ptf_mean;
    int forced[8];  // < 0: derive
};

// one block per alpha: K = min( max_d need[case_d], first k at which every day's mass was exactly 0 )
// The trajectory words may arrive in BLOCKS of `block` days -- [n_blocks][n_alpha][block][2], the layout an all-gather of
// the per-rank [n_alpha][block][2] arrays produces -- so that no kernel has to reshuffle them; block == T is the plain
// [n_alpha][T][2] layout.
__device__ __forceinline__ long long traj_index(long long block, int n_alpha, int ia, long long d) {
    return ((d / block) * n_alpha + ia) * block + d % block;
}

__global__ void finalize_reduce_kernel(FinalizeParams F, const unsigned* __restrict__ traj, long long T, long long block,
                                       int n_alpha, int* __restrict__ k_out, int* __restrict__ status_out) {
    __shared__ int s_need;
    __shared__ unsigned s_nonzero;
    __shared__ int s_firm, s_fragile;   // days whose masses were never near 0 / days with a rounding-noise mass
    const int ia = blockIdx.x;
    if (threadIdx.x == 0) {
        s_need = 0;
        s_nonzero = 0;
        s_firm = 0;
        s_fragile = 0;
    }
    __syncthreads();
    int need = 0;
    unsigned nonzero = 0;
    int firm = 0, fragile = 0;
    const unsigned mask = (F.max_iter >= 32) ? 0xffffffffu : ((1u << F.max_iter) - 1u);
    for (long long d = threadIdx.x; d < T; d += blockDim.x) {
        const long long o = traj_index(block, n_alpha, ia, d);
        const unsigned w0 = traj[2 * o], w1 = traj[2 * o + 1];
        const unsigned kase = (w0 >> 28) & 7u;
        if (kase < 4) {
            need = max(need, F.need[kase]);
            nonzero |= (~w1) & mask;
            if (w0 & TRAJ_FRAGILE_BIT) fragile = 1;
            else if ((w1 & mask) == 0u) firm = 1;
        } else {
            nonzero |= mask;  // NaN mass is never == 0
            firm = 1;
        }
    }
    atomicMax(&s_need, need);
    atomicOr(&s_nonzero, nonzero);
    if (firm) atomicOr(&s_firm, 1);
    if (fragile) atomicOr(&s_fragile, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int K = s_need, status = 0;
        const unsigned allzero = (~s_nonzero) & mask;
        if (T > 0 && allzero && __ffs(allzero) - 1 < K) {
            K = __ffs(allzero) - 1;
            status |= STATUS_ZERO_EXIT_TAKEN;
        }
        if (T > 0 && s_fragile && !s_firm) status |= STATUS_ZERO_EXIT_AMBIGUOUS;
        if (F.forced[ia] >= 0) K = F.forced[ia];
        k_out[ia] = min(K, F.max_iter);
        status_out[ia] = status;
    }
}

// Small batches: reduction and reconstruction in ONE launch, one block per alpha (a launch costs more than either kernel).
__global__ void __launch_bounds__(256) finalize_small_kernel(FinalizeParams F, const unsigned* __restrict__ traj, long long T,
                                                             long long block, int n_alpha, int* __restrict__ k_out,
                                                             int* __restrict__ status_out, double* __restrict__ var_out,
                                                             int* __restrict__ case_out) {
    __shared__ int s_need, s_firm, s_fragile, s_k;
    __shared__ unsigned s_nonzero;
    const int ia = blockIdx.x;
    if (threadIdx.x == 0) { s_need = 0; s_nonzero = 0; s_firm = 0; s_fragile = 0; }
    __syncthreads();
    int need = 0, firm = 0, fragile = 0;
    unsigned nonzero = 0;
    const unsigned mask = (F.max_iter >= 32) ? 0xffffffffu : ((1u << F.max_iter) - 1u);
    for (long long d = threadIdx.x; d < T; d += blockDim.x) {
        const long long o = traj_index(block, n_alpha, ia, d);
        const unsigned w0 = traj[2 * o], w1 = traj[2 * o + 1];
        const unsigned kase = (w0 >> 28) & 7u;
        if (kase < 4) {
            need = max(need, F.need[kase]);
            nonzero |= (~w1) & mask;
            if (w0 & TRAJ_FRAGILE_BIT) fragile = 1;
            else if ((w1 & mask) == 0u) firm = 1;
        } else {
            nonzero |= mask;
            firm = 1;
        }
    }
    atomicMax(&s_need, need);
    atomicOr(&s_nonzero, nonzero);
    if (firm) atomicOr(&s_firm, 1);
    if (fragile) atomicOr(&s_fragile, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int K = s_need, status = 0;
        const unsigned allzero = (~s_nonzero) & mask;
        if (T > 0 && allzero && __ffs(allzero) - 1 < K) {
            K = __ffs(allzero) - 1;
            status |= STATUS_ZERO_EXIT_TAKEN;
        }
        if (T > 0 && s_fragile && !s_firm) status |= STATUS_ZERO_EXIT_AMBIGUOUS;
        if (F.forced[ia] >= 0) K = F.forced[ia];
        s_k = min(K, F.max_iter);
        k_out[ia] = s_k;
        status_out[ia] = status;
    }
    __syncthreads();
    const int K = s_k;
    for (long long d = threadIdx.x; d < T; d += blockDim.x) {
        const unsigned w0 = traj[2 * traj_index(block, n_alpha, ia, d)];
        const unsigned kase = (w0 >> 28) & 7u;
        double v = NAN;
        if (kase < 4) {
            double lo = F.lo[kase], hi = F.hi[kase];
            for (int k = 0; k < K; ++k) {
                const double mid = (lo + hi) / 2;
                if ((w0 >> k) & 1u) lo = mid; else hi = mid;
            }
            v = (lo + hi) / 2 + F.ptf_mean;
        }
        var_out[(long long)ia * T + d] = v;
        if (case_out) case_out[(long long)ia * T + d] = (int)kase;
    }
}

__global__ void finalize_apply_kernel(FinalizeParams F, const unsigned* __restrict__ traj, long long T, long long block,
                                      int n_alpha, const int* __restrict__ k_in, double* __restrict__ var_out,
                                      int* __restrict__ case_out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * n_alpha) return;
    const int ia = (int)(idx / T);
    const unsigned w0 = traj[2 * traj_index(block, n_alpha, ia, idx - (long long)ia * T)];
    const unsigned kase = (w0 >> 28) & 7u;
    double v = NAN;
    if (kase < 4) {
        double lo = F.lo[kase], hi = F.hi[kase];
        const int K = k_in[ia];
        for (int k = 0; k < K; ++k) {
            const double mid = (lo + hi) / 2;
            if ((w0 >> k) & 1u) lo = mid; else hi = mid;
        }
        v = (lo + hi) / 2 + F.ptf_mean;
    }
    var_out[idx] = v;
    if (case_out) case_out[idx] = (int)kase;
}

// ---------------------------------------------------------------------------------------------
// special-function test kernel
// ---------------------------------------------------------------------------------------------
__global__ void special_kernel(int which, double nu, const double* __restrict__ table, double tail_lc,
                               const double* __restrict__ exptab, const double* __restrict__ in, long long count,
                               double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double v = in[i];
    double r;
    switch (which) {
        case 0: r = t_quantile_table(table, nu, tail_lc, v); break;
        case 1: { TDist D = tdist_make(nu); r = t_quantile_iterative(D, v); break; }
        case 2: r = exp2_fast(v); break;
        case 3: r = log2_fast(v); break;
        case 4: r = phi_via_erf(v); break;
        case 5: r = normcdfinv(v); break;
        case 6: r = rcp_fast(v); break;
        case 7: r = exptab ? exp2_tab(v, exptab) : NAN; break;
        default: r = NAN;
    }
    out[i] = r;
}

// ---------------------------------------------------------------------------------------------
// elementwise copula density c(u0, u1) (the calculators' `copula_density` hook; not on the solve path)
// ---------------------------------------------------------------------------------------------
__global__ void copula_density_kernel(int copula, double rho, double nu, double theta, double kc,
                                      const double* __restrict__ u, long long count, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double u0 = u[2 * i], u1 = u[2 * i + 1];
    double c;
    if (copula == 2) {
        const double eta = theta - 1.0;
        const double num = theta * (1.0 + eta * (u0 + u1 - 2.0 * u0 * u1));
        const double den = (1.0 + eta * (u0 + u1)) * (1.0 + eta * (1.0 - u0 - u1));
        c = num / (den * den);
    } else if (copula == 0) {
        const double y0 = normcdfinv(u0), y1 = normcdfinv(u1);
        const double om = 1.0 - rho * rho;
        c = (isfinite(y0) && isfinite(y1)) ? exp(-(rho * rho * (y0 * y0 + y1 * y1) - 2.0 * rho * y0 * y1) / (2.0 * om)) / sqrt(om)
                                           : NAN;
    } else {
        TDist D = tdist_make(nu);
        const double y0 = t_quantile_iterative(D, u0), y1 = t_quantile_iterative(D, u1);
        if (isfinite(y0) && isfinite(y1)) {
            const double om = 1.0 - rho * rho;
            const double qf = (y0 * y0 - 2.0 * rho * y0 * y1 + y1 * y1) / om;
            c = kc * exp(-0.5 * (nu + 2.0) * log1p(qf / nu) + 0.5 * (nu + 1.0) * (log1p(y0 * y0 / nu) + log1p(y1 * y1 / nu)));
        } else {
            c = NAN;
        }
    }
    out[i] = c;
}

// ---------------------------------------------------------------------------------------------
// FP64 roofline denominator: dependency-free DFMA streams (8 independent chains per thread)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(int iters, double seed, double* __restrict__ sink) {
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
    const double m = 0.999999, c = 1e-9 * threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 123.456) sink[0] = r;  // never true; keeps the chains alive
}

}  // namespace cvar
