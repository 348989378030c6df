// cvar_api.cu -- C ABI (include/cvar.h) over the sm_100a kernels in cvar_kernels.cuh.
//
// Host side only does: argument validation, run-constant preparation (copula constants, the
// Student-t quantile table, the axis on the device), launches, and the H2D/D2H copies of the
// `*_host` entry points.  There is no CPU implementation of the solve in this library.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/cvar.h"
#include "cvar_forecast.cuh"
#include "cvar_kernels.cuh"

using namespace cvar;

// ---------------------------------------------------------------------------------------------
struct cvar_plan {
    cvar_desc_t desc;
    KernelParams kp;
    FinalizeParams fp;
    int device;
    int sm_count;
    int ctas_per_sm;
    int cta_threads;
    bool cta_threads_forced;  // CVAR_CTA_THREADS was given: no launch-time adjustment
    int cluster_forced;       // CVAR_CLUSTER = 1, 2 or 4: fixed cluster size (0: chosen per launch)
    int max_clusters[5];      // [2], [4]: clusters of that size that can be co-resident (cudaOccupancyMaxActiveClusters)
    size_t smem_bytes;
    double tq_err;
    double last_kernel_ms;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    double* d_x;
    double* d_dx;
    double* d_sigma_states;
    double* d_tq_table;
    double* d_logtab;
    double* d_exptab;
    double* d_powtab;
    double* d_powfast;
    int kernel_variant;  // KV_* template instantiation used by this plan
    double* d_state_cdf;
    double* d_state_pdf;
    // growable workspace of the *_host entry points
    void* d_ws;
    size_t ws_bytes;
    int* d_k;  // [2][CVAR_MAX_ALPHA]: iteration counts, then status words of the last finalize
    int h_k[2 * CVAR_MAX_ALPHA];   // host copy of d_k after a *_host solve (one D2H for iterations and status)
    bool status_on_host;
    // launch-order scratch (keys, indices, sorted copies, cub temp), sized by cvar_plan_reserve for `chunk_days` days and
    // never grown inside a *_device call
    int64_t chunk_days;
    void* d_sched;
    size_t sched_bytes;
    size_t sort_tmp_bytes;
    double rho_eff;
    int order_min_waves;   // a chunk is ordered when it has more days than this many waves of resident CTAs
};

#define CU_TRY(expr)                          \
    do {                                      \
        cudaError_t _e = (expr);              \
        if (_e != cudaSuccess) return (int)_e; \
    } while (0)

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int needed_iterations(double width, double tol) {
    // while (upper - lower > tol) halve   (utils/calc_var_class.py:278); widths are dyadic => exact
    int k = 0;
    while (width > tol && k < 64) {
        width /= 2;
        ++k;
    }
    return k;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    return (int)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

int ensure_ws(cvar_plan* p, size_t bytes) {
    if (bytes <= p->ws_bytes) return 0;
    if (p->d_ws) CU_TRY(cudaFree(p->d_ws));
    p->d_ws = nullptr;
    p->ws_bytes = 0;
    size_t want = std::max(bytes, (size_t)1 << 20);
    CU_TRY(cudaMalloc(&p->d_ws, want));
    p->ws_bytes = want;
    return 0;
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// Near-minimax polynomial of (1+f)^(-c) on |f| <= fm: interpolation at the Chebyshev nodes of degree D, expanded to
// monomials in f (extended precision, coefficients then rounded to double).  Returns the largest relative error of
// the ROUNDED polynomial on a dense grid, which is what the plan compares with its budget.
double fit_pow_series(double c, double fm, int D, double* coef) {
    const long double PI = 3.141592653589793238462643383279502884L;
    long double g[POW_MAX_DEG + 1], a[POW_MAX_DEG + 1], mono[POW_MAX_DEG + 1] = {0};
    long double T[POW_MAX_DEG + 1][POW_MAX_DEG + 1] = {{0}};
    for (int k = 0; k <= D; ++k) g[k] = powl(1.0L + (long double)fm * cosl(PI * (k + 0.5L) / (D + 1)), -(long double)c);
    for (int j = 0; j <= D; ++j) {
        long double acc = 0.0L;
        for (int k = 0; k <= D; ++k) acc += g[k] * cosl(j * PI * (k + 0.5L) / (D + 1));
        a[j] = acc * (j == 0 ? 1.0L : 2.0L) / (D + 1);
    }
    T[0][0] = 1.0L;
    if (D >= 1) T[1][1] = 1.0L;
    for (int j = 2; j <= D; ++j)
        for (int i = 0; i <= j; ++i) T[j][i] = (i > 0 ? 2.0L * T[j - 1][i - 1] : 0.0L) - T[j - 2][i];
    for (int j = 0; j <= D; ++j)
        for (int i = 0; i <= j; ++i) mono[i] += a[j] * T[j][i];
    long double scale = 1.0L;
    for (int i = 0; i <= D; ++i) {
        coef[i] = (double)(mono[i] / scale);
        scale *= (long double)fm;
    }
    long double worst = 0.0L;
    for (int s = 0; s <= 512; ++s) {
        const long double f = (long double)fm * (-1.0L + s / 256.0L);
        long double p = coef[D];
        for (int i = D - 1; i >= 0; --i) p = p * f + (long double)coef[i];
        const long double ex = powl(1.0L + f, -(long double)c);
        worst = std::max(worst, fabsl(p - ex) / ex);
    }
    return (double)worst;
}

__global__ void tq_table_check_kernel(double nu, const double* __restrict__ table, double tail_lc,
                                      unsigned long long* __restrict__ max_err_bits) {
    // off-node probes in every interval: table vs the iterative routine
    const double probes[4] = {-0.83, -0.21, 0.37, 0.91};
    const int m = blockIdx.x, k = threadIdx.x;
    if (k >= 4) return;
    const double h = TQ_WMAX / TQ_INTERVALS;
    const double w = -TQ_WMAX + h * (m + 0.5 * (1.0 + probes[k]));
    const double p = 0.5 * erfc(-w * 0.7071067811865476);
    TDist D = tdist_make(nu);
    const double exact = t_quantile_mag_iterative(D, p);
    const double fast = t_quantile_mag_table(table, nu, tail_lc, p);
    const double err = fabs(fast - exact) / fmax(exact, 0.1);  // relative in the tails, absolute (x10) around the median
    atomicMax(max_err_bits, (unsigned long long)__double_as_longlong(err));  // err >= 0: bit order == value order
}

// Resident CTAs per SM of the solve kernel of variant `kv` at `threads` per CTA and `smem` bytes of dynamic shared memory.
template <int KV>
int occupancy_of(int threads, size_t smem, int* occ) {
    return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, solve_kernel<KV, false>, threads, smem);
}
int solve_occupancy(int kv, int threads, size_t smem, int* occ) {
    switch (kv) {
        case 0: return occupancy_of<0>(threads, smem, occ);
        case 1: return occupancy_of<1>(threads, smem, occ);
        case 2: return occupancy_of<2>(threads, smem, occ);
        case 3: return occupancy_of<3>(threads, smem, occ);
        case 4: return occupancy_of<4>(threads, smem, occ);
        case 5: return occupancy_of<5>(threads, smem, occ);
        case 6: return occupancy_of<6>(threads, smem, occ);
        default: return CVAR_ERR_COPULA;
    }
}

// The dynamic shared-memory limit is an attribute of a kernel instantiation on a device, not of a plan: every
// instantiation of the variant is opted in to the device maximum, so that plans of different grids can coexist.
template <int KV>
int opt_in_smem(size_t bytes) {
    int rc = set_smem(solve_kernel<KV, false>, bytes);
    if (!rc) rc = set_smem(solve_kernel<KV, true>, bytes);
    if (!rc) rc = set_smem(strip_mass_kernel<KV>, bytes);
    return rc;
}
int opt_in_smem_variant(int kv, size_t bytes) {
    switch (kv) {
        case 0: return opt_in_smem<0>(bytes);
        case 1: return opt_in_smem<1>(bytes);
        case 2: return opt_in_smem<2>(bytes);
        case 3: return opt_in_smem<3>(bytes);
        case 4: return opt_in_smem<4>(bytes);
        case 5: return opt_in_smem<5>(bytes);
        case 6: return opt_in_smem<6>(bytes);
        default: return CVAR_ERR_COPULA;
    }
}

// CTA size (measured on B200, tools/sweep_cta_threads.sh): registers cap an SM at 16 resident warps whatever the CTA
// size, and the more independent CTAs share those warps the better their barrier phases interleave.  So: the smallest
// of 64 / 128 / 256 / 512 threads whose resident CTAs (shared memory, registers) still add up to the most warps.
int best_cta_threads(int kv, size_t smem, int* threads_out, int* resident_out) {
    int best_threads = 0, best_resident = -1;
    for (int th = 64; th <= CTA_THREADS_LARGE; th *= 2) {
        int occ = 0;
        int rc = solve_occupancy(kv, th, smem, &occ);
        if (rc) return rc;
        const int resident = std::min(occ * th, 512);
        if (resident > best_resident) { best_resident = resident; best_threads = th; }
    }
    *threads_out = best_threads;
    *resident_out = best_resident;
    return 0;
}

size_t sched_layout(int64_t T, size_t b_tmp, size_t* b_key, size_t* b_idx) {
    *b_key = align256(sizeof(float) * T);
    *b_idx = align256(sizeof(int) * T);
    return 2 * *b_key + 2 * *b_idx + align256(b_tmp);
}

// (Re)allocate the launch-order scratch for chunks of `days` days.  Host-side set-up only: the *_device entry points
// never allocate, they cut a batch into chunks of at most chunk_days days (each with its own launch order).
int reserve_chunk(cvar_plan* p, int64_t days) {
    days = std::max<int64_t>(1, std::min<int64_t>(days, 0x7fffffffLL));
    if (days <= p->chunk_days) return 0;
    CU_TRY(cudaStreamSynchronize(p->stream));
    if (p->d_sched) CU_TRY(cudaFree(p->d_sched));
    p->d_sched = nullptr;
    p->chunk_days = 0;
    size_t b_tmp = 0, b_key, b_idx;
    CU_TRY(cub::DeviceRadixSort::SortPairs(nullptr, b_tmp, (const float*)nullptr, (float*)nullptr, (const int*)nullptr,
                                           (int*)nullptr, (int)days, 0, 32, (cudaStream_t)0));
    p->sort_tmp_bytes = b_tmp;
    p->sched_bytes = sched_layout(days, b_tmp, &b_key, &b_idx);
    CU_TRY(cudaMalloc(&p->d_sched, p->sched_bytes));
    p->chunk_days = days;
    return 0;
}

// Launch order of a chunk: ascending portfolio-variance proxy (most expensive solves first); nullptr when the chunk
// (nearly) fits the resident CTA slots anyway.  Works in the scratch reserved for chunk_days days.
int make_order(cvar_plan* p, const double* d_day, int64_t T, cudaStream_t st, const int** order_out) {
    *order_out = nullptr;
    const int64_t slots = (int64_t)p->sm_count * std::max(p->ctas_per_sm, 1);
    if (T <= p->order_min_waves * slots) return 0;
    size_t b_key, b_idx;
    sched_layout(p->chunk_days, p->sort_tmp_bytes, &b_key, &b_idx);
    size_t b_tmp = p->sort_tmp_bytes;
    char* base = (char*)p->d_sched;
    float* key_in = (float*)base;
    float* key_out = (float*)(base + b_key);
    int* idx_in = (int*)(base + 2 * b_key);
    int* idx_out = (int*)(base + 2 * b_key + b_idx);
    void* tmp = base + 2 * b_key + 2 * b_idx;
    order_key_kernel<<<(unsigned)((T + 255) / 256), 256, 0, st>>>(p->kp, d_day, (long long)T, p->rho_eff, key_in, idx_in);
    CU_TRY(cudaGetLastError());
    CU_TRY(cub::DeviceRadixSort::SortPairs(tmp, b_tmp, key_in, key_out, idx_in, idx_out, (int)T, 0, 32, st));
    *order_out = idx_out;
    return 0;
}

// Batches that leave SMs without a CTA (T <= #SMs) are latency-bound on one day's chain of strips: give each day
// 16 warps instead of 8 (measured at n = 2048, 125 days: 0.54 -> 0.45 ms; above #SMs days two 8-warp CTAs per SM win).
int launch_threads(const cvar_plan* p, int64_t T) {
    if (!p->cta_threads_forced && T <= p->sm_count && p->kp.n >= 1024) return CTA_THREADS_LARGE;
    return p->cta_threads;
}

// A batch that cannot even give every second (fourth) SM a day is split further: 2 or 4 CTAs of a thread-block
// cluster share one day (cvar_kernels.cuh, `Part`).  Measured at n = 2048: see DESIGN.md section 8.
int cluster_size(const cvar_plan* p, int64_t T) {
    if (p->cluster_forced > 0) return p->cluster_forced;
    if (p->cta_threads_forced || p->kp.n < 1024) return 1;
    if (T <= p->max_clusters[4]) return 4;   // only while every cluster of the batch is resident at once
    if (T <= p->max_clusters[2]) return 2;
    return 1;
}

template <typename K>
int launch_clustered(K kernel, int cluster, unsigned grid, unsigned block, size_t smem, cudaStream_t st, KernelParams kp,
                     const double* d_day, long long day0, long long T, AlphaSet A, const int* order, unsigned* traj,
                     double* mass, unsigned long long* cells) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, kernel, kp, d_day, day0, T, A, order, traj, mass, cells);
}

int day_stride(const cvar_plan* p) { return p->desc.marginal == CVAR_MARGINAL_SINGLE ? 2 : 2 * p->desc.q; }

int launch_solve_kernel(cvar_plan* p, int cluster, int threads, int64_t units, const double* d_day, int64_t day0, int64_t T,
                        const AlphaSet& A, const int* order, uint32_t* d_traj, double* d_mass, unsigned long long* d_cells,
                        cudaStream_t st) {
    int rc = 0;
    dim3 grid((unsigned)units), block(threads);
#define CVAR_LAUNCH_SOLVE(KV)                                                                                              \
    case KV:                                                                                                               \
        if (cluster > 1)                                                                                                   \
            rc = launch_clustered(solve_kernel<KV, true>, cluster, (unsigned)(units * cluster), block.x, p->smem_bytes, st, p->kp, \
                                  d_day, (long long)day0, (long long)T, A, order, d_traj, d_mass, d_cells);               \
        else                                                                                                               \
            solve_kernel<KV, false><<<grid, block, p->smem_bytes, st>>>(p->kp, d_day, (long long)day0, (long long)T, A, order, d_traj, d_mass, d_cells); \
        break;
    switch (p->kernel_variant) {
        CVAR_LAUNCH_SOLVE(0) CVAR_LAUNCH_SOLVE(1) CVAR_LAUNCH_SOLVE(2) CVAR_LAUNCH_SOLVE(3) CVAR_LAUNCH_SOLVE(4) CVAR_LAUNCH_SOLVE(5)
        CVAR_LAUNCH_SOLVE(6)
        default: return CVAR_ERR_COPULA;
    }
#undef CVAR_LAUNCH_SOLVE
    if (rc) return rc;
    return (int)cudaGetLastError();
}

// A batch runs as chunks of at most chunk_days days (the size of the launch-order scratch): order, solve kernel, on `st`.
// (Tried and left out: running the cheapest ~20 % of a chunk -- the last CTAs of the launch order -- as 2-CTA clusters in a
// second kernel on another stream, so that the tail of the launch is made of shorter units.  c3: 1.78 -> 1.82-1.91 ms for
// 100-400 tail days; the cluster barrier per strip and the second axis stage cost more than the idle SMs of the tail.
// Also tried: the remainder of the chunk modulo the resident CTA slots (112 of 1000 days) as CTAs of twice the threads, one
// per SM: 1.776 -> 1.804 ms -- a wide CTA only starts on an SM once BOTH of its slots have drained.
// The CTA timeline of c3 (tools/timeline_profile.py, profiles/r2_c3_timeline.txt): 1000 days on 296 slots, CTA durations
// 184 .. 710 us (mean 433) inside a 1647 us launch, 263 of 296 slots busy on average; all slots are busy until 85 % of the
// launch, the rest is the last day of each slot running out.  Replaying the launch with the measured durations gives
// 1644 us for this order and 1640 us for most-expensive-first on the TRUE durations: with 3.4 days per slot and a 4 x
// spread of durations the loss is the packing, not the proxy.  Tried on top: an arranged order that feeds the 112 slots
// which run a fourth day with cheap days only (1.686 -> 1.741 ms at 1000 days, 1.223 -> 1.266 at 700, 2.451 -> 2.393 at
// 1500; c4 unchanged) and ordering c2's 1.7 waves (1.225 -> 1.228 ms): neither kept.  Two batches in flight
// (ShardedSolver) are what fills the tail: DESIGN.md section 8.)
int launch_solve(cvar_plan* p, const double* d_day, int64_t T, const AlphaSet& A, uint32_t* d_traj, double* d_mass,
                 unsigned long long* d_cells, cudaStream_t st) {
    for (int64_t c0 = 0; c0 < T; c0 += p->chunk_days) {
        const int64_t Tc = std::min(p->chunk_days, T - c0);
        const int* order = nullptr;
        int rc = make_order(p, d_day + c0 * day_stride(p), Tc, st, &order);
        if (rc) return rc;
        rc = launch_solve_kernel(p, cluster_size(p, Tc), launch_threads(p, Tc), Tc, d_day, c0, T, A, order, d_traj, d_mass, d_cells, st);
        if (rc) return rc;
    }
    return 0;
}

int launch_strip(cvar_plan* p, const double* d_day, int64_t T, const double* d_bounds, double* d_out,
                 unsigned long long* d_cells, cudaStream_t st) {
    if (T == 0) return 0;
    dim3 grid((unsigned)T), block(launch_threads(p, T));
#define CVAR_LAUNCH_STRIP(KV) \
    case KV: strip_mass_kernel<KV><<<grid, block, p->smem_bytes, st>>>(p->kp, d_day, d_bounds, d_out, d_cells); break;
    switch (p->kernel_variant) {
        CVAR_LAUNCH_STRIP(0) CVAR_LAUNCH_STRIP(1) CVAR_LAUNCH_STRIP(2) CVAR_LAUNCH_STRIP(3) CVAR_LAUNCH_STRIP(4) CVAR_LAUNCH_STRIP(5)
        CVAR_LAUNCH_STRIP(6)
        default: return CVAR_ERR_COPULA;
    }
#undef CVAR_LAUNCH_STRIP
    return (int)cudaGetLastError();
}

int launch_finalize(cvar_plan* p, const uint32_t* d_traj, int64_t T, int64_t block, int32_t n_alpha, const int32_t* forced,
                    double ptf_mean, double* d_var, int32_t* d_case, cudaStream_t st) {
    FinalizeParams F = p->fp;
    F.ptf_mean = ptf_mean;
    for (int i = 0; i < CVAR_MAX_ALPHA; ++i) F.forced[i] = (forced && i < n_alpha) ? forced[i] : -1;
    const long long blk = std::max<int64_t>(block, 1);
    p->status_on_host = false;
    if (T <= 4096) {   // small batch: one launch instead of two
        finalize_small_kernel<<<n_alpha, 256, 0, st>>>(F, d_traj, (long long)T, blk, n_alpha, p->d_k, p->d_k + CVAR_MAX_ALPHA,
                                                       d_var, d_case);
        return (int)cudaGetLastError();
    }
    finalize_reduce_kernel<<<n_alpha, 1024, 0, st>>>(F, d_traj, (long long)T, blk, n_alpha, p->d_k, p->d_k + CVAR_MAX_ALPHA);
    CU_TRY(cudaGetLastError());
    const long long total = (long long)T * n_alpha;
    if (total > 0) {
        finalize_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(F, d_traj, (long long)T, blk, n_alpha, p->d_k,
                                                                              d_var, d_case);
        CU_TRY(cudaGetLastError());
    }
    return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
extern "C" {

int cvar_abi_version(void) { return CVAR_ABI_VERSION; }

int cvar_check_dim(int32_t dim) { return dim == 2 ? CVAR_OK : CVAR_ERR_DIM; }

void cvar_desc_default(cvar_desc_t* d) {
    if (!d) return;
    std::memset(d, 0, sizeof(*d));
    d->struct_size = (uint32_t)sizeof(cvar_desc_t);
    d->abi_version = CVAR_ABI_VERSION;
    d->copula = CVAR_COPULA_GAUSSIAN;
    d->marginal = CVAR_MARGINAL_SINGLE;
    d->n = 100;  // utils/calc_var_class.py:16
    d->q = 1;
    d->compat_flags = CVAR_COMPAT_REFERENCE;
    d->max_iter = 0;
    d->rho = d->nu = d->theta = NAN;
    d->w0 = d->w1 = 0.5;   // :17
    d->clip_lo = -5.0;     // :201
    d->neg_inf = -100.0;   // :114
    d->first_guess = -3.0; // :95
    d->second_lo = -3.5;
    d->second_hi = -2.0;
    d->min_var = -7.5;     // :111
    d->max_var = 0.0;      // :112
    d->tol = 1e-6;         // :257
}

const char* cvar_strerror(int status) {
    switch (status) {
        case CVAR_OK: return "ok";
        case CVAR_ERR_NULL: return "required pointer is NULL";
        case CVAR_ERR_COPULA: return "unknown copula or marginal family";
        case CVAR_ERR_GRID: return "bad grid: need 2 <= n <= CVAR_MAX_N and a strictly ascending axis";
        case CVAR_ERR_PARAM: return "parameter out of range (|rho| < 1, nu > 0, theta > 0, w0 != 0, 1 <= q <= CVAR_MAX_Q, finite sigmas)";
        case CVAR_ERR_SIZE: return "bad size (T < 0 or n_alpha outside 1..CVAR_MAX_ALPHA)";
        case CVAR_ERR_NO_DEVICE: return "no usable CUDA device (this library has no CPU fallback)";
        case CVAR_ERR_ABI: return "cvar_desc_t struct_size / abi_version mismatch";
        case CVAR_ERR_SMEM: return "grid too large for the shared memory of one SM";
        case CVAR_ERR_DIM: return "only two-asset portfolios are supported (the reference's dim >= 3 grid does not yield probabilities; see cvar.h)";
        case CVAR_ERR_TABLE: return "Student-t quantile table missed its accuracy budget for this nu (plan refused)";
        default: break;
    }
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "unknown status";
}

int cvar_plan_create(const cvar_desc_t* desc, const double* x, const double* dx, const double* sigma_states,
                     int device, cvar_plan_t** out) {
    if (!desc || !x || !dx || !out) return CVAR_ERR_NULL;
    *out = nullptr;
    if (desc->struct_size != sizeof(cvar_desc_t) || desc->abi_version != CVAR_ABI_VERSION) return CVAR_ERR_ABI;
    if (desc->copula < 0 || desc->copula > 2 || desc->marginal < 0 || desc->marginal > 1) return CVAR_ERR_COPULA;
    const int n = desc->n;
    if (n < 2 || n > CVAR_MAX_N) return CVAR_ERR_GRID;
    for (int i = 1; i < n; ++i)
        if (!(x[i] > x[i - 1])) return CVAR_ERR_GRID;
    for (int i = 0; i < n; ++i)
        if (!std::isfinite(x[i]) || !std::isfinite(dx[i])) return CVAR_ERR_GRID;
    const int q = desc->marginal == CVAR_MARGINAL_SINGLE ? 1 : desc->q;
    if (q < 1 || q > CVAR_MAX_Q) return CVAR_ERR_PARAM;
    if (desc->marginal == CVAR_MARGINAL_MIXTURE) {
        if (!sigma_states) return CVAR_ERR_NULL;
        for (int i = 0; i < 2 * q; ++i)
            if (!(sigma_states[i] > 0.0) || !std::isfinite(sigma_states[i])) return CVAR_ERR_PARAM;
    }
    if (!(desc->w0 != 0.0) || !std::isfinite(desc->w0) || !std::isfinite(desc->w1)) return CVAR_ERR_PARAM;
    if (desc->copula != CVAR_COPULA_PLACKETT && !(std::fabs(desc->rho) < 1.0)) return CVAR_ERR_PARAM;
    if (desc->copula == CVAR_COPULA_STUDENT && !(desc->nu > 0.0 && std::isfinite(desc->nu))) return CVAR_ERR_PARAM;
    if (desc->copula == CVAR_COPULA_PLACKETT && !(desc->theta > 0.0 && std::isfinite(desc->theta))) return CVAR_ERR_PARAM;
    if (!(desc->tol > 0.0) || desc->max_iter < 0 || desc->max_iter > CVAR_MAX_ITER) return CVAR_ERR_PARAM;
    if (!(desc->second_lo < desc->first_guess && desc->first_guess < desc->second_hi && desc->min_var < desc->second_lo &&
          desc->second_hi < desc->max_var && desc->neg_inf < desc->min_var))
        return CVAR_ERR_PARAM;

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return CVAR_ERR_NO_DEVICE;
    }
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) return CVAR_ERR_NO_DEVICE;
    }
    if (device >= ndev) return CVAR_ERR_NO_DEVICE;
    DeviceGuard guard(device);
    if (!guard.ok) return CVAR_ERR_NO_DEVICE;

    cvar_plan* p = new (std::nothrow) cvar_plan();
    if (!p) return (int)cudaErrorMemoryAllocation;
    std::memset(p, 0, sizeof(*p));
    p->desc = *desc;
    p->desc.q = q;
    p->device = device;

    cudaDeviceProp prop;
    int rc = (int)cudaGetDeviceProperties(&prop, device);
    if (rc) { delete p; return rc; }
    p->sm_count = prop.multiProcessorCount;
    // kernel variant: Student-t cells use the table-assisted power at the smallest polynomial degree that reproduces
    // (1+f)^(-(nu+2)/2) to 1e-15 relative on the table's interval (cell budget: 1e-13), else the generic log2/exp2 cell
    p->kernel_variant = desc->copula;
    double powc[POW_MAX_DEG + 1] = {0};
    if (desc->copula == CVAR_COPULA_STUDENT && !std::getenv("CVAR_STUDENT_GENERIC")) {
        const int degs[4] = {5, 6, 7, 8}, kvs[4] = {KV_STUDENT_POW5, KV_STUDENT_POW6, KV_STUDENT_POW7, KV_STUDENT_POW8};
        for (int i = 0; i < 4; ++i) {
            double trial[POW_MAX_DEG + 1] = {0};
            if (fit_pow_series(0.5 * (desc->nu + 2.0), POW_FMAX, degs[i], trial) < 1e-15) {
                std::memcpy(powc, trial, sizeof(powc));
                p->kernel_variant = kvs[i];
                break;
            }
        }
    }
    p->smem_bytes = smem_bytes_for(n, p->kernel_variant, 0);
    if (p->smem_bytes > (size_t)prop.sharedMemPerBlockOptin) { delete p; return CVAR_ERR_SMEM; }
    rc = opt_in_smem_variant(p->kernel_variant, prop.sharedMemPerBlockOptin);
    if (rc) { delete p; return rc; }
    // One-lookup power table (cvar_math.cuh): as many octaves of t as fit without costing the SM a resident warp.
    int pow_octaves = 0;
    if (kv_pow_degree(p->kernel_variant) > 0 && POW_FAST_MODE > 0) {
        int want = POW_FAST_MAX_OCTAVES;
        if (const char* env = std::getenv("CVAR_POW_OCTAVES")) want = std::max(0, std::min(POW_FAST_MAX_OCTAVES, std::atoi(env)));
        int th0 = 0, res0 = 0;
        rc = best_cta_threads(p->kernel_variant, p->smem_bytes, &th0, &res0);
        if (rc) { delete p; return rc; }
        // fewer than four octaves (t < 16) catch too few rows to pay for the per-row range check (measured on c2: n = 1024
        // has room for two octaves next to four resident CTAs and runs 1 % slower with them than without)
        for (int oct = want; oct >= (std::getenv("CVAR_POW_OCTAVES") ? 1 : 4); --oct) {
            const size_t bytes = smem_bytes_for(n, p->kernel_variant, oct);
            if (bytes > (size_t)prop.sharedMemPerBlockOptin) continue;
            int th = 0, res = 0;
            rc = best_cta_threads(p->kernel_variant, bytes, &th, &res);
            if (rc) { delete p; return rc; }
            if ((res >= res0 && th <= th0) || std::getenv("CVAR_POW_OCTAVES")) { pow_octaves = oct; break; }
        }
        p->smem_bytes = smem_bytes_for(n, p->kernel_variant, pow_octaves);
    }

    // ---- run constants -------------------------------------------------------------------
    KernelParams& kp = p->kp;
    std::memset(&kp, 0, sizeof(kp));
    kp.copula = desc->copula;
    kp.marginal = desc->marginal;
    kp.n = n;
    kp.q = q;
    kp.compat = desc->compat_flags;
    kp.rho = desc->rho; kp.nu = desc->nu; kp.theta = desc->theta;
    kp.w0 = desc->w0; kp.w1 = desc->w1;
    {
        int ex = 0;
        const double mant = std::frexp(std::fabs(desc->w0), &ex);
        // exact for every finite numerator unless the quotient leaves the normal range, which |q|, |x| <= 100 exclude
        kp.rw0_exact = (mant == 0.5 && ex > -500 && ex < 500) ? 1.0 / desc->w0 : 0.0;
    }
    kp.neg_inf = desc->neg_inf; kp.first = desc->first_guess;
    kp.second_lo = desc->second_lo; kp.second_hi = desc->second_hi;
    kp.min_var = desc->min_var; kp.max_var = desc->max_var;
    kp.cmin = (int)(std::upper_bound(x, x + n, desc->clip_lo) - x);
    {   // uniform segments of the axis, for the boundary guess (cvar_kernels.cuh, count_row)
        std::vector<int> first{0};
        for (int i = 2; i < n; ++i) {
            const double h0 = x[i - 1] - x[i - 2], h1 = x[i] - x[i - 1];
            if (std::fabs(h1 - h0) > 1e-6 * std::fabs(h0)) first.push_back(i - 1);   // spacing changes at x[i-1]
        }
        kp.nseg = 0;
        for (int s = 0; s < MAX_AXIS_SEGMENTS; ++s) kp.seg_x0[s] = INFINITY, kp.seg_inv_h[s] = 0.0, kp.seg_first[s] = n;
        kp.seg_first[MAX_AXIS_SEGMENTS] = n;
        if ((int)first.size() <= MAX_AXIS_SEGMENTS && std::getenv("CVAR_NO_SEGMENT_GUESS") == nullptr) {
            kp.nseg = (int)first.size();
            for (int s = 0; s < kp.nseg; ++s) {
                const int a = first[s], b = (s + 1 < kp.nseg) ? first[s + 1] : n - 1;   // segment spans x[a] .. x[b]
                kp.seg_first[s] = a;
                kp.seg_x0[s] = x[a];
                kp.seg_inv_h[s] = b > a ? (double)(b - a) / (x[b] - x[a]) : 0.0;
            }
        }
    }
    double dx_min = INFINITY;
    for (int i = 1; i < n; ++i) dx_min = std::min(dx_min, x[i] - x[i - 1]);
    // brackets whose image on the inner axis, (hi - lo) / w0, is shorter than the finest spacing hold at most one grid
    // point per row (strip_pass_thin); 1 % margin for the rounding of the two bounds
    kp.thin_width = (desc->w0 > 0.0 && std::getenv("CVAR_NO_THIN_PASS") == nullptr) ? 0.99 * dx_min * desc->w0 : 0.0;
    const double LOG2E = 1.4426950408889634;
    if (desc->copula == CVAR_COPULA_GAUSSIAN) {
        const double om = 1.0 - desc->rho * desc->rho;
        const double kappa = desc->rho * desc->rho / (2.0 * om);
        kp.g_in_scale = std::sqrt(kappa * LOG2E);
        kp.g_out_scale = (desc->rho < 0 ? -1.0 : 1.0) * std::sqrt(LOG2E / (2.0 * om));
        kp.g_const = 1.0 / std::sqrt(om);
    } else if (desc->copula == CVAR_COPULA_STUDENT) {
        const double nu = desc->nu, om = 1.0 - desc->rho * desc->rho;
        const double cs = 1.0 / std::sqrt(nu * om);
        kp.g_in_scale = cs;
        kp.g_out_scale = desc->rho * cs;
        kp.g_const = std::exp(std::lgamma(0.5 * (nu + 2.0)) + std::lgamma(0.5 * nu) - 2.0 * std::lgamma(0.5 * (nu + 1.0))) /
                     std::sqrt(om);
        const double lbeta = std::lgamma(0.5 * nu) + std::lgamma(0.5) - std::lgamma(0.5 * nu + 0.5);
        kp.tq_tail_lc = (-lbeta - 0.5 * std::log(nu)) + 0.5 * (nu - 1.0) * std::log(nu);
        // t = 1 + y0^2/nu + cs^2 (y1 - rho y0)^2 <= 1 + y_max^2 (1/nu + cs^2 (1+|rho|)^2) must stay below 2^63
        const double ar = 1.0 + std::fabs(desc->rho);
        kp.y_max = std::sqrt((std::ldexp(1.0, 63) - 1.0) / (1.0 / nu + cs * cs * ar * ar)) * (1.0 - 1e-9);
    }
    // correlation used by the launch-order proxy only (Plackett: Spearman's rho of the family)
    p->rho_eff = desc->rho;
    if (desc->copula == CVAR_COPULA_PLACKETT) {
        const double th = desc->theta;
        p->rho_eff = std::fabs(th - 1.0) < 1e-9 ? 0.0 : ((th + 1.0) / (th - 1.0) - 2.0 * th * std::log(th) / ((th - 1.0) * (th - 1.0)));
    }
    // iterations: every bracket's own requirement, and the largest of them is what the kernel records
    FinalizeParams& fp = p->fp;
    std::memset(&fp, 0, sizeof(fp));
    const double blo[4] = {desc->min_var, desc->second_lo, desc->second_hi, desc->first_guess};
    const double bhi[4] = {desc->second_lo, desc->first_guess, desc->max_var, desc->second_hi};
    int need_max = 0;
    for (int c = 0; c < 4; ++c) {
        fp.lo[c] = blo[c];
        fp.hi[c] = bhi[c];
        fp.need[c] = needed_iterations(bhi[c] - blo[c], desc->tol);
        need_max = std::max(need_max, fp.need[c]);
    }
    const int max_iter = desc->max_iter > 0 ? desc->max_iter : need_max;
    if (max_iter > 28) { delete p; return CVAR_ERR_PARAM; }  // decision bits share a word with the bracket id
    kp.max_iter = fp.max_iter = max_iter;
    p->desc.max_iter = max_iter;

#define PLAN_TRY(expr)                                   \
    do {                                                 \
        cudaError_t _e = (expr);                         \
        if (_e != cudaSuccess) { cvar_plan_destroy(p); return (int)_e; } \
    } while (0)

    PLAN_TRY(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    PLAN_TRY(cudaEventCreate(&p->ev0));
    PLAN_TRY(cudaEventCreate(&p->ev1));

    PLAN_TRY(cudaMalloc(&p->d_x, sizeof(double) * n));
    PLAN_TRY(cudaMalloc(&p->d_dx, sizeof(double) * n));
    // [2][CVAR_MAX_ALPHA] ints (iteration counts, status words), then the 8-byte evaluated-cells counter
    PLAN_TRY(cudaMalloc(&p->d_k, sizeof(int) * 2 * CVAR_MAX_ALPHA + sizeof(unsigned long long)));
    PLAN_TRY(cudaMemsetAsync(p->d_k, 0, sizeof(int) * 2 * CVAR_MAX_ALPHA + sizeof(unsigned long long), p->stream));
    p->kp.evaluated_cells = reinterpret_cast<unsigned long long*>(p->d_k + 2 * CVAR_MAX_ALPHA);
    PLAN_TRY(cudaMemcpyAsync(p->d_x, x, sizeof(double) * n, cudaMemcpyHostToDevice, p->stream));
    PLAN_TRY(cudaMemcpyAsync(p->d_dx, dx, sizeof(double) * n, cudaMemcpyHostToDevice, p->stream));
    if (desc->marginal == CVAR_MARGINAL_MIXTURE) {
        PLAN_TRY(cudaMalloc(&p->d_sigma_states, sizeof(double) * 2 * q));
        PLAN_TRY(cudaMemcpyAsync(p->d_sigma_states, sigma_states, sizeof(double) * 2 * q, cudaMemcpyHostToDevice, p->stream));
    }
    kp.x = p->d_x;
    kp.dx = p->d_dx;
    kp.sigma_states = p->d_sigma_states;
    if (desc->copula != CVAR_COPULA_PLACKETT) {
        PLAN_TRY(cudaMalloc(&p->d_exptab, sizeof(double) * EXPTAB_SIZE));
        exptab_build_kernel<<<1, EXPTAB_SIZE, 0, p->stream>>>(p->d_exptab);
        PLAN_TRY(cudaGetLastError());
        kp.exptab = p->d_exptab;
    }
    if (desc->marginal == CVAR_MARGINAL_MIXTURE) {
        const size_t cnt = (size_t)2 * q * n;
        PLAN_TRY(cudaMalloc(&p->d_state_cdf, sizeof(double) * cnt));
        PLAN_TRY(cudaMalloc(&p->d_state_pdf, sizeof(double) * cnt));
        state_table_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, p->stream>>>(n, q, p->d_x, p->d_sigma_states,
                                                                                p->d_state_cdf, p->d_state_pdf);
        PLAN_TRY(cudaGetLastError());
        kp.state_cdf = p->d_state_cdf;
        kp.state_pdf = p->d_state_pdf;
    }
    if (desc->copula == CVAR_COPULA_STUDENT) {
        PLAN_TRY(cudaMalloc(&p->d_tq_table, sizeof(double) * TQ_TABLE_DOUBLES));
        tq_table_build_kernel<<<TQ_INTERVALS, 32, 0, p->stream>>>(desc->nu, p->d_tq_table);
        PLAN_TRY(cudaGetLastError());
        unsigned long long* d_err = nullptr;
        PLAN_TRY(cudaMalloc(&d_err, sizeof(unsigned long long)));
        PLAN_TRY(cudaMemsetAsync(d_err, 0, sizeof(unsigned long long), p->stream));
        tq_table_check_kernel<<<TQ_INTERVALS, 32, 0, p->stream>>>(desc->nu, p->d_tq_table, kp.tq_tail_lc, d_err);
        cudaError_t ce = cudaGetLastError();
        unsigned long long bits = 0;
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(&bits, d_err, sizeof(bits), cudaMemcpyDeviceToHost, p->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(p->stream);
        cudaFree(d_err);
        PLAN_TRY(ce);
        std::memcpy(&p->tq_err, &bits, sizeof(double));
        // Accuracy budget of the table: 1e-11 relative keeps a cell weight inside its 1e-13 budget with two decades to
        // spare at the largest exponent (nu + 2) / 2 the table is used with.  Measured: <= 3e-14 for 2.01 <= nu <= 50.
        // A degree of freedom for which the table misses it (or whose check produced a NaN) is refused instead of being
        // solved with an inaccurate quantile.  (Falling back to the iterative routine inside the kernel was measured:
        // the call costs the solve kernel 6 registers and ~1 % on every Student-t plan.)
        double budget = 1e-11;
        if (const char* env = std::getenv("CVAR_TQ_BUDGET")) budget = std::atof(env);   // tests
        if (!(p->tq_err <= budget)) { cvar_plan_destroy(p); return CVAR_ERR_TABLE; }
        kp.tq_table = p->d_tq_table;
        // table-assisted log2 of the cell loop: constants scaled by -(nu+2)/2
        constexpr double q1p[] = CVAR_LOG2_1P_POLY;
        kp.negc = -0.5 * (desc->nu + 2.0);
        kp.inv_nu = 1.0 / desc->nu;
        kp.seed_mask = ~((1u << (20 - POW_BITS)) - 1u);
        kp.seed_half = 1u << (19 - POW_BITS);
        for (int k = 0; k <= CVAR_LOG2_1P_POLY_DEG; ++k) kp.qc[k] = kp.negc * q1p[k];
        PLAN_TRY(cudaMalloc(&p->d_logtab, sizeof(double) * LOGTAB_SIZE));
        logtab_build_kernel<<<1, LOGTAB_SIZE, 0, p->stream>>>(kp.negc, p->d_logtab);
        PLAN_TRY(cudaGetLastError());
        kp.logtab = p->d_logtab;
        for (int k = 0; k <= POW_MAX_DEG; ++k) kp.powc[k] = powc[k];
        PLAN_TRY(cudaMalloc(&p->d_powtab, sizeof(double) * (POW_MTAB + POW_ETAB)));
        powtab_build_kernel<<<1, POW_MTAB, 0, p->stream>>>(0.5 * (desc->nu + 2.0), p->d_powtab);
        PLAN_TRY(cudaGetLastError());
        kp.powtab = p->d_powtab;
        if (pow_octaves > 0) {
            const int entries = pow_octaves * POW_MTAB;
            PLAN_TRY(cudaMalloc(&p->d_powfast, sizeof(double) * entries * POW_FAST_ENTRY_DOUBLES));
            powfast_build_kernel<<<(entries + 255) / 256, 256, 0, p->stream>>>(0.5 * (desc->nu + 2.0), pow_octaves, p->d_powfast);
            PLAN_TRY(cudaGetLastError());
            kp.powfast = p->d_powfast;
            kp.pow_octaves = pow_octaves;
            kp.pow_fast_limit = std::ldexp(1.0, pow_octaves) * (1.0 - 1e-6);   // margin: the row check looks at the range ends only
        }
    }
    PLAN_TRY(cudaStreamSynchronize(p->stream));

    // opt in to the dynamic shared memory this grid needs (for the instantiation this plan uses) and size the CTAs.
    // CTA size (measured on B200, tools/sweep_cta_threads.sh): registers cap an SM at 16 resident warps whatever the
    // CTA size, and the more independent CTAs share those warps the better their barrier phases interleave.  So: the
    // smallest of 64 / 128 / 256 / 512 threads whose resident CTAs (shared memory, registers) still add up to 16 warps
    // -- 64 threads up to n = 512, 128 up to n ~ 1150, 256 up to n ~ 2300, 512 above (one CTA per SM).  Tiny grids
    // (n < 192, typically one wave of CTAs) prefer 128 threads for their axis stage.
    int occ = 0;
    p->cta_threads = n < 192 ? 128 : 0;   // 0: chosen below from the occupancy of this kernel instantiation
    if (const char* env = std::getenv("CVAR_CTA_THREADS")) {   // tuning knob: 32..512, multiple of 32
        const int v = std::atoi(env);
        if (v >= 32 && v <= CTA_THREADS_LARGE && v % 32 == 0) {
            p->cta_threads = v;
            p->cta_threads_forced = true;
        }
    }
    if (const char* env = std::getenv("CVAR_CLUSTER")) {        // tuning knob: 1, 2 or 4 CTAs per day
        const int v = std::atoi(env);
        if (v == 1 || v == 2 || v == 4) p->cluster_forced = v;
    }
    if (p->cta_threads == 0) {
        int resident = 0;
        PLAN_TRY((cudaError_t)best_cta_threads(p->kernel_variant, p->smem_bytes, &p->cta_threads, &resident));
    }
    PLAN_TRY((cudaError_t)solve_occupancy(p->kernel_variant, p->cta_threads, p->smem_bytes, &occ));
    p->ctas_per_sm = occ;
    // how many 2- and 4-CTA clusters of the large CTA fit the device at once (GPC boundaries make this less than
    // #SMs / size); a failure here only disables the cluster split
#define CVAR_CLUSTER_OCC(KV)                                                                              \
    case KV:                                                                                              \
        for (int cs = 2; cs <= 4; cs *= 2) {                                                              \
            cudaLaunchConfig_t cfg = {};                                                                  \
            cfg.gridDim = dim3((unsigned)(cs * p->sm_count));                                             \
            cfg.blockDim = dim3(CTA_THREADS_LARGE);                                                       \
            cfg.dynamicSmemBytes = p->smem_bytes;                                                         \
            cudaLaunchAttribute attr[1];                                                                  \
            attr[0].id = cudaLaunchAttributeClusterDimension;                                             \
            attr[0].val.clusterDim.x = (unsigned)cs;                                                      \
            attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;                                      \
            cfg.attrs = attr;                                                                             \
            cfg.numAttrs = 1;                                                                             \
            int num = 0;                                                                                  \
            if (cudaOccupancyMaxActiveClusters(&num, solve_kernel<KV, true>, &cfg) == cudaSuccess) p->max_clusters[cs] = num; \
            else cudaGetLastError();                                                                      \
        }                                                                                                 \
        break;
    switch (p->kernel_variant) {
        CVAR_CLUSTER_OCC(0) CVAR_CLUSTER_OCC(1) CVAR_CLUSTER_OCC(2) CVAR_CLUSTER_OCC(3) CVAR_CLUSTER_OCC(4) CVAR_CLUSTER_OCC(5)
        CVAR_CLUSTER_OCC(6)
        default: break;
    }
#undef CVAR_CLUSTER_OCC
    p->order_min_waves = 2;
    if (const char* env = std::getenv("CVAR_ORDER_MIN_WAVES")) p->order_min_waves = std::max(1, std::atoi(env));   // tuning knobs
    {   // launch-order scratch for the default chunk; cvar_plan_reserve (or a *_host call) sizes it for larger batches
        int64_t days = 65536;
        if (const char* env = std::getenv("CVAR_CHUNK_DAYS")) days = std::max(1LL, std::atoll(env));
        PLAN_TRY((cudaError_t)reserve_chunk(p, days));
    }
#undef PLAN_TRY
    *out = p;
    return CVAR_OK;
}

int cvar_plan_destroy(cvar_plan_t* p) {
    if (!p) return CVAR_OK;
    DeviceGuard guard(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    cudaFree(p->d_x);
    cudaFree(p->d_dx);
    cudaFree(p->d_sigma_states);
    cudaFree(p->d_tq_table);
    cudaFree(p->d_logtab);
    cudaFree(p->d_exptab);
    cudaFree(p->d_powtab);
    cudaFree(p->d_powfast);
    cudaFree(p->d_state_cdf);
    cudaFree(p->d_state_pdf);
    cudaFree(p->d_ws);
    cudaFree(p->d_k);
    cudaFree(p->d_sched);
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
    return CVAR_OK;
}

int cvar_plan_reserve(cvar_plan_t* p, int64_t days) {
    if (!p) return CVAR_ERR_NULL;
    if (days < 0) return CVAR_ERR_SIZE;
    DeviceGuard guard(p->device);
    return reserve_chunk(p, days);
}

int cvar_plan_get_info(const cvar_plan_t* p, cvar_plan_info_t* info) {
    if (!p || !info) return CVAR_ERR_NULL;
    info->device = p->device;
    info->sm_count = p->sm_count;
    info->max_iter = p->kp.max_iter;
    info->ctas_per_sm = p->ctas_per_sm;
    info->threads_per_cta = p->cta_threads;
    info->smem_bytes_per_cta = (int32_t)p->smem_bytes;
    info->tq_table_max_rel_err = p->tq_err;
    info->last_kernel_ms = p->last_kernel_ms;
    info->kernel_variant = p->kernel_variant;
    info->cluster4_capacity = p->max_clusters[4];
    info->pow_octaves = p->kp.pow_octaves;
    info->chunk_days = p->chunk_days;
    return CVAR_OK;
}

// ---- strip masses ---------------------------------------------------------------------------
int cvar_strip_mass_device(cvar_plan_t* p, const double* day_params, int64_t T, const double* bounds, double* out,
                           uint64_t* out_cells, void* stream) {
    if (!p || (T > 0 && (!day_params || !bounds || !out))) return CVAR_ERR_NULL;
    if (T < 0 || T > 0x7fffffffLL) return CVAR_ERR_SIZE;
    DeviceGuard guard(p->device);
    return launch_strip(p, day_params, T, bounds, out, (unsigned long long*)out_cells, (cudaStream_t)stream);
}

int cvar_strip_mass_host(cvar_plan_t* p, const double* day_params, int64_t T, const double* bounds, double* out,
                         uint64_t* out_cells) {
    if (!p || (T > 0 && (!day_params || !bounds || !out))) return CVAR_ERR_NULL;
    if (T < 0 || T > 0x7fffffffLL) return CVAR_ERR_SIZE;
    if (T == 0) return CVAR_OK;
    DeviceGuard guard(p->device);
    const size_t b_day = align256(sizeof(double) * T * day_stride(p));
    const size_t b_bnd = align256(sizeof(double) * T * 2);
    const size_t b_out = align256(sizeof(double) * T);
    const size_t b_cel = align256(sizeof(uint64_t) * T);
    int rc = ensure_ws(p, b_day + b_bnd + b_out + b_cel);
    if (rc) return rc;
    char* ws = (char*)p->d_ws;
    double* d_day = (double*)ws;
    double* d_bnd = (double*)(ws + b_day);
    double* d_out = (double*)(ws + b_day + b_bnd);
    unsigned long long* d_cel = (unsigned long long*)(ws + b_day + b_bnd + b_out);
    CU_TRY(cudaMemcpyAsync(d_day, day_params, sizeof(double) * T * day_stride(p), cudaMemcpyHostToDevice, p->stream));
    CU_TRY(cudaMemcpyAsync(d_bnd, bounds, sizeof(double) * T * 2, cudaMemcpyHostToDevice, p->stream));
    rc = launch_strip(p, d_day, T, d_bnd, d_out, out_cells ? d_cel : nullptr, p->stream);
    if (rc) return rc;
    CU_TRY(cudaMemcpyAsync(out, d_out, sizeof(double) * T, cudaMemcpyDeviceToHost, p->stream));
    if (out_cells) CU_TRY(cudaMemcpyAsync(out_cells, d_cel, sizeof(uint64_t) * T, cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    return CVAR_OK;
}

// ---- solve ----------------------------------------------------------------------------------
static int check_alphas(const double* alphas, int32_t n_alpha, AlphaSet* A) {
    if (!alphas) return CVAR_ERR_NULL;
    if (n_alpha < 1 || n_alpha > CVAR_MAX_ALPHA) return CVAR_ERR_SIZE;
    A->n_alpha = n_alpha;
    for (int i = 0; i < CVAR_MAX_ALPHA; ++i) A->a[i] = i < n_alpha ? alphas[i] : 0.0;
    for (int i = 0; i < n_alpha; ++i)
        if (!(alphas[i] > 0.0 && alphas[i] < 1.0)) return CVAR_ERR_PARAM;
    return 0;
}

int cvar_solve_device(cvar_plan_t* p, const double* day_params, int64_t T, const double* alphas, int32_t n_alpha,
                      uint32_t* traj, double* mass, uint64_t* cells, void* stream) {
    if (!p || (T > 0 && (!day_params || !traj))) return CVAR_ERR_NULL;
    if (T < 0 || T > 0x7fffffffLL) return CVAR_ERR_SIZE;
    AlphaSet A;
    int rc = check_alphas(alphas, n_alpha, &A);
    if (rc) return rc;
    DeviceGuard guard(p->device);
    return launch_solve(p, day_params, T, A, traj, mass, (unsigned long long*)cells, (cudaStream_t)stream);
}

int cvar_finalize_device(cvar_plan_t* p, const uint32_t* traj, int64_t T, int32_t n_alpha, const int32_t* forced,
                         double ptf_mean, double* var_out, int32_t* case_out, int32_t* iterations_out, void* stream) {
    return cvar_finalize_blocked_device(p, traj, T, T, n_alpha, forced, ptf_mean, var_out, case_out, iterations_out, stream);
}

int cvar_finalize_blocked_device(cvar_plan_t* p, const uint32_t* traj, int64_t T, int64_t block_days, int32_t n_alpha,
                                 const int32_t* forced, double ptf_mean, double* var_out, int32_t* case_out,
                                 int32_t* iterations_out, void* stream) {
    if (!p || (T > 0 && (!traj || !var_out))) return CVAR_ERR_NULL;
    if (T < 0 || n_alpha < 1 || n_alpha > CVAR_MAX_ALPHA || (T > 0 && block_days < 1)) return CVAR_ERR_SIZE;
    if (forced)
        for (int i = 0; i < n_alpha; ++i)
            if (forced[i] > p->kp.max_iter) return CVAR_ERR_PARAM;
    DeviceGuard guard(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_finalize(p, traj, T, block_days, n_alpha, forced, ptf_mean, var_out, case_out, st);
    if (rc) return rc;
    if (iterations_out)
        CU_TRY(cudaMemcpyAsync(iterations_out, p->d_k, sizeof(int) * n_alpha, cudaMemcpyDeviceToDevice, st));
    return CVAR_OK;
}

int cvar_evaluated_cells_host(cvar_plan_t* p, uint64_t* total_out, int reset) {
    if (!p || !total_out) return CVAR_ERR_NULL;
    DeviceGuard guard(p->device);
    CU_TRY(cudaDeviceSynchronize());   // launches on any stream
    unsigned long long v = 0;
    CU_TRY(cudaMemcpy(&v, p->kp.evaluated_cells, sizeof(v), cudaMemcpyDeviceToHost));
    if (reset) CU_TRY(cudaMemset(p->kp.evaluated_cells, 0, sizeof(v)));
    *total_out = v;
    return CVAR_OK;
}

int cvar_finalize_status_device(cvar_plan_t* p, int32_t* status_out, int32_t n_alpha, void* stream) {
    if (!p || !status_out) return CVAR_ERR_NULL;
    if (n_alpha < 1 || n_alpha > CVAR_MAX_ALPHA) return CVAR_ERR_SIZE;
    DeviceGuard guard(p->device);
    CU_TRY(cudaMemcpyAsync(status_out, p->d_k + CVAR_MAX_ALPHA, sizeof(int) * n_alpha, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return CVAR_OK;
}

int cvar_finalize_status_host(cvar_plan_t* p, int32_t* status_out, int32_t n_alpha) {
    if (!p || !status_out) return CVAR_ERR_NULL;
    if (n_alpha < 1 || n_alpha > CVAR_MAX_ALPHA) return CVAR_ERR_SIZE;
    if (p->status_on_host) {   // the last finalize was a *_host solve: its status words came back with the results
        std::memcpy(status_out, p->h_k + CVAR_MAX_ALPHA, sizeof(int) * n_alpha);
        return CVAR_OK;
    }
    DeviceGuard guard(p->device);
    CU_TRY(cudaMemcpyAsync(status_out, p->d_k + CVAR_MAX_ALPHA, sizeof(int) * n_alpha, cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    return CVAR_OK;
}

int cvar_solve_host(cvar_plan_t* p, const double* day_params, int64_t T, const double* alphas, int32_t n_alpha,
                    const int32_t* forced, double ptf_mean, double* var_out, int32_t* case_out, uint64_t* cells_out,
                    int32_t* iterations_out) {
    if (!p || (T > 0 && (!day_params || !var_out))) return CVAR_ERR_NULL;
    if (T < 0 || T > 0x7fffffffLL) return CVAR_ERR_SIZE;
    AlphaSet A;
    int rc = check_alphas(alphas, n_alpha, &A);
    if (rc) return rc;
    if (forced)
        for (int i = 0; i < n_alpha; ++i)
            if (forced[i] > p->kp.max_iter) return CVAR_ERR_PARAM;
    DeviceGuard guard(p->device);
    { int rr = reserve_chunk(p, T); if (rr) return rr; }
    const size_t na = (size_t)n_alpha;
    const size_t b_day = align256(sizeof(double) * T * day_stride(p));
    const size_t b_trj = align256(sizeof(uint32_t) * 2 * T * na);
    const size_t b_var = align256(sizeof(double) * T * na);
    const size_t b_cas = align256(sizeof(int32_t) * T * na);
    const size_t b_cel = align256(sizeof(uint64_t) * T * na);
    rc = ensure_ws(p, b_day + b_trj + b_var + b_cas + b_cel);
    if (rc) return rc;
    char* ws = (char*)p->d_ws;
    double* d_day = (double*)ws;
    uint32_t* d_trj = (uint32_t*)(ws + b_day);
    double* d_var = (double*)(ws + b_day + b_trj);
    int32_t* d_cas = (int32_t*)(ws + b_day + b_trj + b_var);
    unsigned long long* d_cel = (unsigned long long*)(ws + b_day + b_trj + b_var + b_cas);
    if (T > 0) CU_TRY(cudaMemcpyAsync(d_day, day_params, sizeof(double) * T * day_stride(p), cudaMemcpyHostToDevice, p->stream));
    CU_TRY(cudaEventRecord(p->ev0, p->stream));
    rc = launch_solve(p, d_day, T, A, d_trj, nullptr, cells_out ? d_cel : nullptr, p->stream);
    if (rc) return rc;
    rc = launch_finalize(p, d_trj, T, T, n_alpha, forced, ptf_mean, d_var, case_out ? d_cas : nullptr, p->stream);
    if (rc) return rc;
    CU_TRY(cudaEventRecord(p->ev1, p->stream));
    if (T > 0) {
        CU_TRY(cudaMemcpyAsync(var_out, d_var, sizeof(double) * T * na, cudaMemcpyDeviceToHost, p->stream));
        if (case_out) CU_TRY(cudaMemcpyAsync(case_out, d_cas, sizeof(int32_t) * T * na, cudaMemcpyDeviceToHost, p->stream));
        if (cells_out) CU_TRY(cudaMemcpyAsync(cells_out, d_cel, sizeof(uint64_t) * T * na, cudaMemcpyDeviceToHost, p->stream));
    }
    CU_TRY(cudaMemcpyAsync(p->h_k, p->d_k, sizeof(p->h_k), cudaMemcpyDeviceToHost, p->stream));   // iterations + status
    CU_TRY(cudaStreamSynchronize(p->stream));
    p->status_on_host = true;
    if (iterations_out) std::memcpy(iterations_out, p->h_k, sizeof(int) * n_alpha);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p->ev0, p->ev1) == cudaSuccess) p->last_kernel_ms = ms;
    return CVAR_OK;
}

// ---- special functions (tests) ---------------------------------------------------------------
int cvar_test_special_host(cvar_plan_t* p, int32_t which, const double* in, int64_t count, double* out) {
    if (!p || (count > 0 && (!in || !out))) return CVAR_ERR_NULL;
    if (count < 0) return CVAR_ERR_SIZE;
    if (count == 0) return CVAR_OK;
    if ((which == 0 || which == 1) && p->desc.copula != CVAR_COPULA_STUDENT) return CVAR_ERR_COPULA;
    DeviceGuard guard(p->device);
    const size_t b = align256(sizeof(double) * count);
    int rc = ensure_ws(p, 2 * b);
    if (rc) return rc;
    double* d_in = (double*)p->d_ws;
    double* d_out = (double*)((char*)p->d_ws + b);
    CU_TRY(cudaMemcpyAsync(d_in, in, sizeof(double) * count, cudaMemcpyHostToDevice, p->stream));
    special_kernel<<<(unsigned)((count + 127) / 128), 128, 0, p->stream>>>(which, p->desc.nu, p->d_tq_table, p->kp.tq_tail_lc,
                                                                          p->d_exptab, d_in, (long long)count, d_out);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(out, d_out, sizeof(double) * count, cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    return CVAR_OK;
}

// ---- elementwise copula density ----------------------------------------------------------------
int cvar_copula_density_host(int32_t copula, double rho, double nu, double theta, const double* u, int64_t count,
                             double* out, int device) {
    if (count > 0 && (!u || !out)) return CVAR_ERR_NULL;
    if (count < 0) return CVAR_ERR_SIZE;
    if (copula < 0 || copula > 2) return CVAR_ERR_COPULA;
    if (copula != CVAR_COPULA_PLACKETT && !(std::fabs(rho) < 1.0)) return CVAR_ERR_PARAM;
    if (copula == CVAR_COPULA_STUDENT && !(nu > 0.0)) return CVAR_ERR_PARAM;
    if (copula == CVAR_COPULA_PLACKETT && !(theta > 0.0)) return CVAR_ERR_PARAM;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return CVAR_ERR_NO_DEVICE;
    }
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return CVAR_ERR_NO_DEVICE;
    if (device >= ndev) return CVAR_ERR_NO_DEVICE;
    if (count == 0) return CVAR_OK;
    DeviceGuard guard(device);
    double kc = 0.0;
    if (copula == CVAR_COPULA_STUDENT)
        kc = std::exp(std::lgamma(0.5 * (nu + 2.0)) + std::lgamma(0.5 * nu) - 2.0 * std::lgamma(0.5 * (nu + 1.0))) /
             std::sqrt(1.0 - rho * rho);
    double *d_u = nullptr, *d_o = nullptr;
    CU_TRY(cudaMalloc(&d_u, sizeof(double) * 2 * count));
    cudaError_t e = cudaMalloc(&d_o, sizeof(double) * count);
    if (e != cudaSuccess) { cudaFree(d_u); return (int)e; }
    e = cudaMemcpy(d_u, u, sizeof(double) * 2 * count, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        copula_density_kernel<<<(unsigned)((count + 127) / 128), 128>>>(copula, rho, nu, theta, kc, d_u, (long long)count, d_o);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, d_o, sizeof(double) * count, cudaMemcpyDeviceToHost);
    cudaFree(d_u);
    cudaFree(d_o);
    return (int)e;
}

// ---- FP64 pipe peak (roofline denominator) -----------------------------------------------------
int cvar_fp64_peak_host(int device, double min_ms, double* tflops_out, double* ms_out) {
    if (!tflops_out) return CVAR_ERR_NULL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return CVAR_ERR_NO_DEVICE;
    }
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return CVAR_ERR_NO_DEVICE;
    if (device >= ndev) return CVAR_ERR_NO_DEVICE;
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    double* d_sink = nullptr;
    CU_TRY(cudaMalloc(&d_sink, sizeof(double)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    int iters = 2000;
    float ms = 0.f;
    cudaError_t err = cudaSuccess;
    for (int attempt = 0; attempt < 12; ++attempt) {
        fp64_peak_kernel<<<blocks, threads>>>(iters, 1.0, d_sink);  // warm-up at this size
        cudaEventRecord(e0);
        fp64_peak_kernel<<<blocks, threads>>>(iters, 1.0, d_sink);
        cudaEventRecord(e1);
        err = cudaEventSynchronize(e1);
        if (err != cudaSuccess) break;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms >= min_ms) break;
        iters *= 2;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_sink);
    if (err != cudaSuccess) return (int)err;
    const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;  // 64 DFMA per thread per iteration
    *tflops_out = flops / (ms * 1e-3) / 1e12;
    if (ms_out) *ms_out = ms;
    return CVAR_OK;
}

}  // extern "C"

// ---- forecast producers ---------------------------------------------------------------
namespace {

template <int K>
void launch_msm(const MsmAsset& A, const double* lik, int64_t T, int N, int64_t stride, const int* level, int q,
                double* out, int64_t out_stride, double* state_probs, int* status, cudaStream_t st) {
    const unsigned blocks = (unsigned)((T + MSM_WARPS_PER_CTA - 1) / MSM_WARPS_PER_CTA);
    msm_filter_kernel<K><<<blocks, 32 * MSM_WARPS_PER_CTA, 0, st>>>(A, lik, (long long)T, N, (long long)stride, level, q, out,
                                                                    (long long)out_stride, state_probs, status);
}

int msm_dispatch(int k, const MsmAsset& A, const double* lik, int64_t T, int N, int64_t stride, const int* level, int q,
                 double* out, int64_t out_stride, double* state_probs, int* status, cudaStream_t st) {
    switch (k) {
        case 1: launch_msm<1>(A, lik, T, N, stride, level, q, out, out_stride, state_probs, status, st); break;
        case 2: launch_msm<2>(A, lik, T, N, stride, level, q, out, out_stride, state_probs, status, st); break;
        case 3: launch_msm<3>(A, lik, T, N, stride, level, q, out, out_stride, state_probs, status, st); break;
        case 4: launch_msm<4>(A, lik, T, N, stride, level, q, out, out_stride, state_probs, status, st); break;
        case 5: launch_msm<5>(A, lik, T, N, stride, level, q, out, out_stride, state_probs, status, st); break;
        case 6: launch_msm<6>(A, lik, T, N, stride, level, q, out, out_stride, state_probs, status, st); break;
        case 7: launch_msm<7>(A, lik, T, N, stride, level, q, out, out_stride, state_probs, status, st); break;
        case 8: launch_msm<8>(A, lik, T, N, stride, level, q, out, out_stride, state_probs, status, st); break;
        case 9: launch_msm<9>(A, lik, T, N, stride, level, q, out, out_stride, state_probs, status, st); break;
        case 10: launch_msm<10>(A, lik, T, N, stride, level, q, out, out_stride, state_probs, status, st); break;
        default: return CVAR_ERR_PARAM;
    }
    return (int)cudaGetLastError();
}

int pick_device(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return -1;
    }
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return -1;
    return device < ndev ? device : -1;
}

}  // namespace

extern "C" {

int cvar_msm_forecast_device(int32_t k, int32_t n_assets, const double* stay_prob, const double* vol_states,
                             const int32_t* level_of_state, int32_t q, const double* returns, int64_t T, int64_t N,
                             int64_t window_stride, double* probs_by_state, double* state_probs, double* workspace,
                             int32_t* status, void* stream) {
    if (!stay_prob || !vol_states || !level_of_state || !returns || !probs_by_state || !workspace || !status) return CVAR_ERR_NULL;
    if (k < 1 || k > MSM_MAX_K || n_assets < 1 || q < 1 || q > (1 << k)) return CVAR_ERR_PARAM;
    if (T < 0 || N < 1 || window_stride < 1) return CVAR_ERR_SIZE;
    if (T == 0) return CVAR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int S = 1 << k;
    const int64_t L = (T - 1) * window_stride + N;  // returns per asset
    for (int a = 0; a < n_assets; ++a) {
        MsmAsset A;
        for (int c = 0; c < MSM_MAX_K; ++c) A.stay[c] = c < k ? stay_prob[a * k + c] : 1.0;
        double* lik = workspace + (size_t)a * L * S;
        const long long cnt = (long long)L * S;
        msm_likelihood_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(returns + a * L, (long long)L, vol_states + a * S, S, lik);
        CU_TRY(cudaGetLastError());
        int rc = msm_dispatch(k, A, lik, T, (int)N, window_stride, level_of_state + a * S, q, probs_by_state + a * q,
                              (int64_t)n_assets * q, state_probs ? state_probs + (size_t)a * T * S : nullptr, status, st);
        if (rc) return rc;
    }
    return CVAR_OK;
}

int cvar_msm_forecast_host(int32_t k, int32_t n_assets, const double* stay_prob, const double* vol_states,
                           const int32_t* level_of_state, int32_t q, const double* returns, int64_t T, int64_t N,
                           int64_t window_stride, double* probs_by_state, double* state_probs, int32_t* status_out,
                           double* kernel_ms_out, int device) {
    if (!stay_prob || !vol_states || !level_of_state || !returns || !probs_by_state) return CVAR_ERR_NULL;
    if (k < 1 || k > MSM_MAX_K || n_assets < 1 || q < 1 || q > (1 << k)) return CVAR_ERR_PARAM;
    if (T < 0 || N < 1 || window_stride < 1) return CVAR_ERR_SIZE;
    device = pick_device(device);
    if (device < 0) return CVAR_ERR_NO_DEVICE;
    if (T == 0) return CVAR_OK;
    DeviceGuard guard(device);
    const int S = 1 << k;
    const int64_t L = (T - 1) * window_stride + N;
    const size_t b_ret = align256(sizeof(double) * n_assets * L), b_vol = align256(sizeof(double) * n_assets * S);
    const size_t b_lvl = align256(sizeof(int) * n_assets * S), b_stay = 0;
    const size_t b_out = align256(sizeof(double) * T * n_assets * q);
    const size_t b_sp = state_probs ? align256(sizeof(double) * n_assets * T * S) : 0;
    const size_t b_ws = align256(sizeof(double) * n_assets * L * S), b_st = 256;
    (void)b_stay;
    char* base = nullptr;
    CU_TRY(cudaMalloc(&base, b_ret + b_vol + b_lvl + b_out + b_sp + b_ws + b_st));
    double* d_ret = (double*)base;
    double* d_vol = (double*)(base + b_ret);
    int* d_lvl = (int*)(base + b_ret + b_vol);
    double* d_out = (double*)(base + b_ret + b_vol + b_lvl);
    double* d_sp = state_probs ? (double*)(base + b_ret + b_vol + b_lvl + b_out) : nullptr;
    double* d_ws = (double*)(base + b_ret + b_vol + b_lvl + b_out + b_sp);
    int* d_st = (int*)(base + b_ret + b_vol + b_lvl + b_out + b_sp + b_ws);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaError_t e = cudaMemcpy(d_ret, returns, sizeof(double) * n_assets * L, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_vol, vol_states, sizeof(double) * n_assets * S, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_lvl, level_of_state, sizeof(int) * n_assets * S, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(d_st, 0, sizeof(int));
    int rc = (int)e;
    if (!rc) {
        cudaEventRecord(e0);
        rc = cvar_msm_forecast_device(k, n_assets, stay_prob, d_vol, d_lvl, q, d_ret, T, N, window_stride, d_out, d_sp, d_ws,
                                      d_st, nullptr);
        cudaEventRecord(e1);
    }
    if (!rc) rc = (int)cudaMemcpy(probs_by_state, d_out, sizeof(double) * T * n_assets * q, cudaMemcpyDeviceToHost);
    if (!rc && state_probs) rc = (int)cudaMemcpy(state_probs, d_sp, sizeof(double) * n_assets * T * S, cudaMemcpyDeviceToHost);
    int st_host = 0;
    if (!rc) rc = (int)cudaMemcpy(&st_host, d_st, sizeof(int), cudaMemcpyDeviceToHost);
    if (status_out) *status_out = st_host;
    float ms = 0.f;
    if (!rc && kernel_ms_out && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) *kernel_ms_out = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(base);
    return rc;
}

int cvar_kalman_forecast_device(int32_t n_assets, const double* a, const double* l, const double* q, double ukf_alpha,
                                double ukf_beta, double ukf_kappa, const double* returns, int64_t T, int64_t N,
                                int64_t window_stride, double* sigma_out, int32_t* status, void* stream) {
    if (!a || !l || !q || !returns || !sigma_out || !status) return CVAR_ERR_NULL;
    if (n_assets < 1 || T < 0 || N < 1 || window_stride < 1) return CVAR_ERR_SIZE;
    if (T == 0) return CVAR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t L = (T - 1) * window_stride + N;
    for (int k = 0; k < n_assets; ++k) {
        KalmanAsset K{a[k], l[k], q[k], ukf_alpha, ukf_beta, ukf_kappa};
        kalman_forecast_kernel<<<(unsigned)((T + 127) / 128), 128, 0, st>>>(K, returns + k * L, (long long)T, (int)N,
                                                                           (long long)window_stride, sigma_out + k,
                                                                           (long long)n_assets, status);
    }
    return (int)cudaGetLastError();
}

int cvar_kalman_forecast_host(int32_t n_assets, const double* a, const double* l, const double* q, double ukf_alpha,
                              double ukf_beta, double ukf_kappa, const double* returns, int64_t T, int64_t N,
                              int64_t window_stride, double* sigma_out, int32_t* status_out, double* kernel_ms_out, int device) {
    if (!a || !l || !q || !returns || !sigma_out) return CVAR_ERR_NULL;
    if (n_assets < 1 || T < 0 || N < 1 || window_stride < 1) return CVAR_ERR_SIZE;
    device = pick_device(device);
    if (device < 0) return CVAR_ERR_NO_DEVICE;
    if (T == 0) return CVAR_OK;
    DeviceGuard guard(device);
    const int64_t L = (T - 1) * window_stride + N;
    const size_t b_ret = align256(sizeof(double) * n_assets * L), b_out = align256(sizeof(double) * T * n_assets);
    char* base = nullptr;
    CU_TRY(cudaMalloc(&base, b_ret + b_out + 256));
    double* d_ret = (double*)base;
    double* d_out = (double*)(base + b_ret);
    int* d_st = (int*)(base + b_ret + b_out);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaError_t e = cudaMemcpy(d_ret, returns, sizeof(double) * n_assets * L, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(d_st, 0, sizeof(int));
    int rc = (int)e;
    if (!rc) {
        cudaEventRecord(e0);
        rc = cvar_kalman_forecast_device(n_assets, a, l, q, ukf_alpha, ukf_beta, ukf_kappa, d_ret, T, N, window_stride, d_out,
                                         d_st, nullptr);
        cudaEventRecord(e1);
    }
    if (!rc) rc = (int)cudaMemcpy(sigma_out, d_out, sizeof(double) * T * n_assets, cudaMemcpyDeviceToHost);
    int st = 0;
    if (!rc) rc = (int)cudaMemcpy(&st, d_st, sizeof(int), cudaMemcpyDeviceToHost);
    if (status_out) *status_out = st;
    float ms = 0.f;
    if (!rc && kernel_ms_out && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) *kernel_ms_out = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(base);
    return rc;
}

static int check_garch(int32_t n_assets, const double* omega, const int32_t* p, const int32_t* q, const double* alpha,
                       const double* beta, const double* returns, int64_t T, int64_t N, int64_t window_stride,
                       const double* sigma_out) {
    if (!omega || !p || !q || !alpha || !beta || !returns || !sigma_out) return CVAR_ERR_NULL;
    if (n_assets < 1 || T < 0 || N < 1 || window_stride < 1) return CVAR_ERR_SIZE;
    for (int a = 0; a < n_assets; ++a)
        if (p[a] < 1 || q[a] < 1 || p[a] > GARCH_MAX_ORDER || q[a] > GARCH_MAX_ORDER || p[a] > N || q[a] > N) return CVAR_ERR_PARAM;
    return CVAR_OK;
}

int cvar_garch_forecast_device(int32_t n_assets, const double* omega, const int32_t* p, const int32_t* q, const double* alpha,
                               const double* beta, const double* returns, int64_t T, int64_t N, int64_t window_stride,
                               double* sigma_out, void* stream) {
    int rc = check_garch(n_assets, omega, p, q, alpha, beta, returns, T, N, window_stride, sigma_out);
    if (rc || T == 0) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t L = (T - 1) * window_stride + N;
    for (int a = 0; a < n_assets; ++a) {
        GarchAsset G;
        std::memset(&G, 0, sizeof(G));
        G.omega = omega[a]; G.p = p[a]; G.q = q[a];
        for (int i = 0; i < p[a]; ++i) G.alpha[i] = alpha[a * GARCH_MAX_ORDER + i];
        for (int j = 0; j < q[a]; ++j) G.beta[j] = beta[a * GARCH_MAX_ORDER + j];
        garch_forecast_kernel<<<(unsigned)((T + 127) / 128), 128, 0, st>>>(G, returns + a * L, (long long)T, (int)N,
                                                                          (long long)window_stride, sigma_out + a,
                                                                          (long long)n_assets);
    }
    return (int)cudaGetLastError();
}

int cvar_garch_forecast_host(int32_t n_assets, const double* omega, const int32_t* p, const int32_t* q, const double* alpha,
                             const double* beta, const double* returns, int64_t T, int64_t N, int64_t window_stride,
                             double* sigma_out, double* kernel_ms_out, int device) {
    int rc = check_garch(n_assets, omega, p, q, alpha, beta, returns, T, N, window_stride, sigma_out);
    if (rc) return rc;
    device = pick_device(device);
    if (device < 0) return CVAR_ERR_NO_DEVICE;
    if (T == 0) return CVAR_OK;
    DeviceGuard guard(device);
    const int64_t L = (T - 1) * window_stride + N;
    const size_t b_ret = align256(sizeof(double) * n_assets * L);
    char* base = nullptr;
    CU_TRY(cudaMalloc(&base, b_ret + sizeof(double) * T * n_assets));
    double* d_ret = (double*)base;
    double* d_out = (double*)(base + b_ret);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    rc = (int)cudaMemcpy(d_ret, returns, sizeof(double) * n_assets * L, cudaMemcpyHostToDevice);
    if (!rc) {
        cudaEventRecord(e0);
        rc = cvar_garch_forecast_device(n_assets, omega, p, q, alpha, beta, d_ret, T, N, window_stride, d_out, nullptr);
        cudaEventRecord(e1);
    }
    if (!rc) rc = (int)cudaMemcpy(sigma_out, d_out, sizeof(double) * T * n_assets, cudaMemcpyDeviceToHost);
    float ms = 0.f;
    if (!rc && kernel_ms_out && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) *kernel_ms_out = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(base);
    return rc;
}

}  // extern "C"
