// cvar_math.cuh -- FP64 device primitives of the VaR solve (sm_100a).
//
// Two classes of functions live here:
//   * the per-CELL primitives (exp2_fast, log2_fast, rcp_fast): branch-free, FP64-pipe only,
//     no special-value handling -- callers guarantee finite in-range arguments;
//   * the per-AXIS-POINT functions used once per (day, axis point): Phi via erf exactly as the
//     reference forms it (utils/utils.py:17-22, quirk Q14), the normal quantile, and the
//     Student-t quantile (iterative reference routine + Chebyshev table in the normal score).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "cvar_coeffs.h"

namespace cvar {

// ---------------------------------------------------------------------------------------------
// per-cell primitives
// ---------------------------------------------------------------------------------------------

// Polynomial coefficients live in constant memory so that every Horner step is ONE instruction
// (DFMA with a constant-bank operand); as literals they cost two UMOV per coefficient per use.
__constant__ double kExp2Coef[CVAR_EXP2_POLY_DEG + 1] = CVAR_EXP2_POLY;
__constant__ double kAtanhCoef[CVAR_ATANH_POLY_DEG + 1] = CVAR_ATANH_POLY;

// 2^t for finite t <= ~1000.  Results below 2^-1021 flush to ~1e-308 (never garbage).  Max relative
// error ~4e-16 (degree-10 near-minimax on [-1/2, 1/2]).
__device__ __forceinline__ double exp2_fast(double t) {
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: adding it rounds t to the nearest integer
    const double kf = __dadd_rn(t, MAGIC);
    int k = __double2loint(kf);
    const double r = __dadd_rn(t, -__dadd_rn(kf, -MAGIC));  // r in [-1/2, 1/2], exact
    double p = kExp2Coef[CVAR_EXP2_POLY_DEG];
#pragma unroll
    for (int i = CVAR_EXP2_POLY_DEG - 1; i >= 0; --i) p = fma(p, r, kExp2Coef[i]);
    k = max(k, -1021);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// 2^t through a 256-entry table of 2^(i/256) (shared memory) and a degree-4 polynomial on |r| <= 2^-9:
// 8 FP64 instructions instead of 13.  Same argument contract as exp2_fast.
#ifndef CVAR_EXPTAB_BITS
#define CVAR_EXPTAB_BITS 8   // 10 (with the degree-3 polynomial) measured 4 % slower: the larger table costs more shared-memory wavefronts
#endif
constexpr int EXPTAB_BITS = CVAR_EXPTAB_BITS;
constexpr int EXPTAB_SIZE = 1 << EXPTAB_BITS;
#if CVAR_EXPTAB_BITS >= 10
#define CVAR_EXP2_TAB_POLY_DEG CVAR_EXP2_TINY_POLY_DEG
__constant__ double kExp2SmallCoef[CVAR_EXP2_TINY_POLY_DEG + 1] = CVAR_EXP2_TINY_POLY;
#else
#define CVAR_EXP2_TAB_POLY_DEG CVAR_EXP2_SMALL_POLY_DEG
__constant__ double kExp2SmallCoef[CVAR_EXP2_SMALL_POLY_DEG + 1] = CVAR_EXP2_SMALL_POLY;
#endif

__global__ void exptab_build_kernel(double* __restrict__ tab) {
    const int i = threadIdx.x;
    if (i < EXPTAB_SIZE) tab[i] = exp2((double)i / EXPTAB_SIZE);
}

__device__ __forceinline__ double exp2_tab(double t, const double* __restrict__ tab) {
    const double MAGIC = 6755399441055744.0 / EXPTAB_SIZE;  // rounds t to the nearest multiple of 2^-EXPTAB_BITS
    const double kf = __dadd_rn(t, MAGIC);
    const int k256 = __double2loint(kf);                     // round(t * 2^EXPTAB_BITS), two's complement
    const double r = __dadd_rn(t, -__dadd_rn(kf, -MAGIC));   // |r| <= 2^-(EXPTAB_BITS+1), exact
    double p = kExp2SmallCoef[CVAR_EXP2_TAB_POLY_DEG];
#pragma unroll
    for (int i = CVAR_EXP2_TAB_POLY_DEG - 1; i >= 0; --i) p = fma(p, r, kExp2SmallCoef[i]);
    const double v = tab[k256 & (EXPTAB_SIZE - 1)] * p;      // in [1, 2.01)
    const int k = max(k256 >> EXPTAB_BITS, -1021);
    return __hiloint2double(__double2hiint(v) + (k << 20), __double2loint(v));
}

// acc + 2^t: the table entry takes the exponent first (an integer add, off the dependency chain), so the polynomial is
// followed by ONE fused multiply-add instead of a multiply, the exponent add and an add.
__device__ __forceinline__ double exp2_tab_add(double t, const double* __restrict__ tab, double acc) {
    const double MAGIC = 6755399441055744.0 / EXPTAB_SIZE;
    const double kf = __dadd_rn(t, MAGIC);
    const int k256 = __double2loint(kf);
    const double r = __dadd_rn(t, -__dadd_rn(kf, -MAGIC));
    double p = kExp2SmallCoef[CVAR_EXP2_TAB_POLY_DEG];
#pragma unroll
    for (int i = CVAR_EXP2_TAB_POLY_DEG - 1; i >= 0; --i) p = fma(p, r, kExp2SmallCoef[i]);
    const double tv = tab[k256 & (EXPTAB_SIZE - 1)];                        // in [1, 2)
    const int k = max(k256 >> EXPTAB_BITS, -1021);
    const double ts = __hiloint2double(__double2hiint(tv) + (k << 20), __double2loint(tv));   // 2^k * table entry, normal
    return fma(ts, p, acc);
}

// 1/a for finite normal a: MUFU.RCP64H seed (2^-23) + two Newton steps on the FP64 pipe.
__device__ __forceinline__ double rcp_fast(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// 1/a with ONE Newton step on the 2^-23 seed: relative error e0^2 <= 1.4e-14, always of the same sign.  Enough for
// a cell weight (budget 1e-13 relative), and two FP64 instructions cheaper than rcp_fast.
__device__ __forceinline__ double rcp_cell(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    return fma(r, fma(-a, r, 1.0), r);
}

// log2(t) for positive normal t.  log2(t) = e + (2/ln2) * atanh(s) with m = t * 2^-e in
// [sqrt(1/2), sqrt(2)), s = (m-1)/(m+1).  Max relative error ~4e-16.
__device__ __forceinline__ double log2_fast(double t) {
    const int hi = __double2hiint(t);
    // e = floor(log2 t), +1 when the mantissa is above ~sqrt(2) (the polynomial range has 2 % slack)
    const int e = ((hi + (0x00100000 - 0x0006a09e)) >> 20) - 1023;
    const double scale = __hiloint2double((1023 - e) << 20, 0);  // 2^-e
    const double f = fma(t, scale, -1.0);                         // m - 1, exact
    const double den = fma(t, scale, 1.0);                        // m + 1
    double rc;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(den));
    rc = fma(rc, fma(-den, rc, 1.0), rc);                         // 2^-46
    double s = f * rc;
    s = fma(fma(-den, s, f), rc, s);                              // s = f/den to ~1 ulp
    const double z = s * s;
    double p = kAtanhCoef[6];
#pragma unroll
    for (int i = 5; i >= 0; --i) p = fma(p, z, kAtanhCoef[i]);
    const double sc = s * CVAR_TWO_OVER_LN2;
    return fma(sc, z * p, sc) + (double)e;
}

// ----- table-assisted  c * log2(t)  for the Student-t cell -----------------------------------------
// t = 2^e * m, m in [1, 2).  The top LOGTAB_BITS (7) mantissa bits select an interval; its midpoint reciprocal r_i comes
// from MUFU.RCP64H (deterministic, so the table built with the same instruction matches it exactly) and
// f = m * r_i - 1 is exact in one DFMA with |f| <= 2^-8.  Then
//     c * log2(t) = c * e + c * (-log2 r_i) + f * (c * Q(f)),   Q(f) = log2(1+f)/f  (degree 5)
// with c folded into the table (LOGTAB_SIZE doubles in shared memory) and into the polynomial (plan constants).
constexpr int LOGTAB_BITS = 7;
constexpr int LOGTAB_SIZE = 1 << LOGTAB_BITS;

__device__ __forceinline__ double seed_recip(int idx, int bits) {
    // reciprocal of the midpoint of mantissa interval idx (of 2^bits), exactly as the cell loop obtains it
    const double mid = __hiloint2double(0x3ff00000 | (idx << (20 - bits)) | (1 << (19 - bits)), 0);
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(mid));
    return r;
}
__device__ __forceinline__ double logtab_recip(int idx) { return seed_recip(idx, LOGTAB_BITS); }

__global__ void logtab_build_kernel(double c, double* __restrict__ tab) {
    const int i = threadIdx.x;
    if (i < LOGTAB_SIZE) tab[i] = -c * log2(logtab_recip(i));
}

// returns c*log2(t) + b for positive normal t;  qc[k] = c * (coefficients of Q), tab = c * (-log2 r_i)
__device__ __forceinline__ double scaled_log2_plus(double t, double b, double c, const double* __restrict__ qc,
                                                   const double* __restrict__ tab) {
    const int hi = __double2hiint(t);
    const int e = (hi >> 20) - 1023;
    const int idx = (hi >> (20 - LOGTAB_BITS)) & (LOGTAB_SIZE - 1);
    double r = logtab_recip(idx);
    r = __hiloint2double(__double2hiint(r) - (e << 20), 0);  // r_i * 2^-e (the seed's low word is zero)
    const double f = fma(t, r, -1.0);
    double base = fma(c, (double)e, b);
    base += tab[idx];
    double q = qc[CVAR_LOG2_1P_POLY_DEG];
#pragma unroll
    for (int k = CVAR_LOG2_1P_POLY_DEG - 1; k >= 0; --k) q = fma(q, f, qc[k]);
    return fma(f, q, base);
}

// ----- table-assisted  t^(-c)  for the Student-t cell (no log, no exp) ---------------------------------
// With t = 2^e * m and r_i the seed reciprocal (MUFU.RCP64H) of the midpoint of m's interval (the top POW_BITS
// mantissa bits), m = (1 + f) / r_i with |f| <= 2^-(POW_BITS+1), so
//     t^(-c) = 2^(-c e) * r_i^c * (1 + f)^(-c).
// 2^(-c e) and r_i^c come from two per-plan tables in shared memory (POW_ETAB + POW_MTAB doubles) and (1 + f)^(-c)
// is a near-minimax polynomial in f that the plan fits for its c (Chebyshev interpolation, cvar_api.cu) at the
// smallest degree with a relative error below 1e-15: degree 5 for nu <= 8, 6 up to nu = 30, 7 up to 50, 8 up to 100.
// Larger tables need lower degrees but their lookups broadcast less (more shared-memory wavefronts per warp).
// 3 + DEG FP64 instructions in total instead of ~17 for c*log2(t) followed by exp2.
#ifndef CVAR_POW_BITS
#define CVAR_POW_BITS 8    // measured on B200 (c3 / c2, ms): 7 bits 2.07 / 1.41, 8 bits 1.99 / 1.42, 9 bits 2.03 / 1.44, 10 bits 2.10 / 1.46
#endif
constexpr int POW_BITS = CVAR_POW_BITS;
constexpr int POW_MTAB = 1 << POW_BITS;
constexpr int POW_ETAB = 64;           // exponents 0..63 (t >= 1 always: t = C0 + d^2)
constexpr int POW_MAX_DEG = 8;
constexpr double POW_FMAX = 1.03 / (1 << (POW_BITS + 1));   // |f| bound incl. the seed's own error

__global__ void powtab_build_kernel(double c, double* __restrict__ tab) {
    const int i = threadIdx.x;
    if (i < POW_MTAB) tab[i] = pow(seed_recip(i, POW_BITS), c);
    if (i < POW_ETAB) tab[POW_MTAB + i] = exp2(-c * (double)i);
}

__device__ __forceinline__ double lds_f64(unsigned shared_addr) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(shared_addr));
    return v;
}

// Requires 1 <= t < 2^POW_ETAB (stage 0 clamps the Student-t quantiles so that this holds, see KernelParams::y_max):
// the exponent table is indexed without a range check.  Index arithmetic stays "in place": the masked mantissa
// and exponent fields of t's high word are shifted straight into byte offsets, one integer op each.  `tab_s` is the
// shared-window address of the table (cvta.to.shared, hoisted by the caller): with a generic pointer ptxas rebuilds
// the address from two loop-invariant halves for every lookup.
// Returns b * t^(-c); the caller's column weight b is folded into the table product so that only one FMA of the
// caller follows the polynomial on the dependency chain.
template <int DEG>
__device__ __forceinline__ double pow_neg_c(double t, double b, const double* __restrict__ kc, unsigned tab_s,
                                            unsigned seed_mask, unsigned seed_half) {
    const unsigned hi = (unsigned)__double2hiint(t);
    const unsigned mb = hi & (unsigned)((POW_MTAB - 1) << (20 - POW_BITS));   // interval index, still in place
    const unsigned eb = hi & 0x7ff00000u;                                     // biased exponent, still at bit 20
#ifdef CVAR_DEBUG_ASSERT
    if (!(t >= 1.0) || (eb >> 20) < 1023u || (eb >> 20) >= 1023u + POW_ETAB) __trap();   // table range (see y_max)
#endif
    // MUFU.RCP64H reads the high word only and is exponent-transparent (rcp(2^e m) == 2^-e rcp(m) bit for bit;
    // checked for every interval and exponent by tools/mufu_check.cu), so the seed of the interval midpoint WITH
    // t's exponent is r_i * 2^-e directly
    // seed_mask / seed_half arrive as kernel parameters (not literals) so that (hi & mask) | half is ONE LOP3
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(__hiloint2double((int)((hi & seed_mask) | seed_half), 0)));
    const double f = fma(t, r, -1.0);
    double p = kc[DEG];
#pragma unroll
    for (int k = DEG - 1; k >= 0; --k) p = fma(p, f, kc[k]);
    const double pm = lds_f64(tab_s + (mb >> (20 - POW_BITS - 3)));
    const double pe = lds_f64(tab_s + (unsigned)((POW_MTAB - 1023) * 8) + (eb >> 17));
    return ((pm * pe) * b) * p;
}

// ----- the same power over the first octaves of t, one lookup ------------------------------------------
// Almost every cell has a small quadratic form (t < 2^6 on mixture marginals, where the tails are fat).  For
// t < 2^octaves one table indexed by t's exponent AND interval bits together -- (high word >> 12) - (1023 << 8),
// 256 entries per octave -- replaces the two lookups and their product:
//   CVAR_POW_FAST = 1: 8-byte entries  U = r_i^c * 2^(-c e)  (the rounded product of the two table values, so the cell
//                      is bit-identical to the two-table form); the seed still comes from MUFU.RCP64H;
//   CVAR_POW_FAST = 2: 16-byte entries {r_i 2^-e, U}: the seed comes with the same load, which also removes the mask,
//                      the MUFU and the move that zeroes the seed's low word (3 issue slots per cell).
// Row blocks that reach t >= 2^octaves anywhere on their ranges take the two-table form (warp-uniform choice).
#ifndef CVAR_POW_FAST
#define CVAR_POW_FAST 1
#endif
constexpr int POW_FAST_MODE = CVAR_POW_FAST;
constexpr int POW_FAST_MAX_OCTAVES = 8;
constexpr int POW_FAST_ENTRY_DOUBLES = POW_FAST_MODE == 2 ? 2 : 1;

__global__ void powfast_build_kernel(double c, int octaves, double* __restrict__ tab) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= octaves * POW_MTAB) return;
    const int e = idx / POW_MTAB, i = idx % POW_MTAB;
    // the seed exactly as the cell loop obtains it: MUFU.RCP64H of the interval midpoint carrying t's exponent
    const double mid = __hiloint2double(((1023 + e) << 20) | (i << (20 - POW_BITS)) | (1 << (19 - POW_BITS)), 0);
    double rs;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rs) : "d"(mid));
    const double u = pow(seed_recip(i, POW_BITS), c) * exp2(-c * (double)e);   // == pm * pe of the two-table form
    if (POW_FAST_MODE == 2) {
        tab[2 * idx] = rs;
        tab[2 * idx + 1] = u;
    } else {
        tab[idx] = u;
    }
}

// b * t^(-c) for 1 <= t < 2^octaves.  `fast_s` is the shared-window address of the table minus the bias of the index
// (entry size * (1023 << POW_BITS)), so that the shifted high word of t is the byte offset itself.
template <int DEG>
__device__ __forceinline__ double pow_neg_c_fast(double t, double b, const double* __restrict__ kc, unsigned fast_s,
                                                 unsigned seed_mask, unsigned seed_half) {
    const unsigned hi = (unsigned)__double2hiint(t);
    double r, u;
    if (POW_FAST_MODE == 2) {
        const unsigned off = (hi >> (20 - POW_BITS - 4)) & ~15u;
        asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r), "=d"(u) : "r"(fast_s + off));
    } else {
        // the seed word (interval bits kept, midpoint bit set, lower bits clear) shifted down is 8 * index + 4: the byte
        // offset of the entry needs no mask of its own
        const unsigned sh = (hi & seed_mask) | seed_half;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(__hiloint2double((int)sh, 0)));
        u = lds_f64(fast_s - 4u + (sh >> (20 - POW_BITS - 3)));
    }
    const double f = fma(t, r, -1.0);
    double p = kc[DEG];
#pragma unroll
    for (int k = DEG - 1; k >= 0; --k) p = fma(p, f, kc[k]);
    return (u * b) * p;
}

// ---------------------------------------------------------------------------------------------
// per-axis-point functions
// ---------------------------------------------------------------------------------------------

// Phi(z) formed exactly like the reference: 0.5 * (1 + erf(z / sqrt(2)))  (absolute accuracy, Q14).
// The left tail of u is quantised in steps of 2^-54 by the rounding of erf towards -1, and saturates to
// exactly 0 / 1 beyond |z| ~ 8.3.  To land on the same quantum as a correctly rounded erf, |erf| >= ~0.84
// is formed as 1 - erfc(|v|) (one rounding of an accurately known small number) instead of CUDA's 2-ulp erf.
__device__ __forceinline__ double erf_like_reference(double v) {
    const double a = fabs(v);
    if (a < 1.0) return erf(v);
    const double e = 1.0 - erfc(a);
    return v < 0.0 ? -e : e;
}

__device__ __forceinline__ double phi_via_erf(double z) {
    return 0.5 * (1.0 + erf_like_reference(__ddiv_rn(z, 1.4142135623730951)));
}

// ----- Student-t distribution: iterative reference routine -------------------------------------

// Continued fraction of the regularised incomplete beta function (modified Lentz).
__device__ __noinline__ double ibeta_cf(double a, double b, double x) {
    const double TINY = 1e-300;
    double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0;
    double d = 1.0 - qab * x / qap;
    if (fabs(d) < TINY) d = TINY;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 2000; ++m) {
        double m2 = 2.0 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < TINY) d = TINY;
        c = 1.0 + aa / c;
        if (fabs(c) < TINY) c = TINY;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < TINY) d = TINY;
        c = 1.0 + aa / c;
        if (fabs(c) < TINY) c = TINY;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < 2e-16) break;
    }
    return h;
}

// log of I_x(a, b) evaluated through the continued fraction at x (caller picks the convergent side).
__device__ __noinline__ double log_ibeta_direct(double a, double b, double x, double lbeta) {
    return a * log(x) + b * log1p(-x) - lbeta - log(a) + log(ibeta_cf(a, b, x));
}

struct TDist {
    double nu;
    double lbeta;    // log B(nu/2, 1/2)
    double log_pdf0; // log of the density at 0
};

__device__ __forceinline__ TDist tdist_make(double nu) {
    TDist d;
    d.nu = nu;
    d.lbeta = lgamma(0.5 * nu) + 0.5723649429247001 /* lgamma(1/2) */ - lgamma(0.5 * nu + 0.5);
    d.log_pdf0 = -d.lbeta - 0.5 * log(nu);
    return d;
}

// log F(-tau), tau > 0 (lower tail), relative accuracy everywhere.
__device__ __noinline__ double t_log_lower_tail(const TDist& D, double tau) {
    double nu = D.nu, a = 0.5 * nu, b = 0.5;
    double t2 = tau * tau;
    double x = nu / (nu + t2);  // I_x(a, b) = 2 F(-tau)
    if (x < (a + 1.0) / (a + b + 2.0)) {
        return log_ibeta_direct(a, b, x, D.lbeta) - 0.6931471805599453;
    }
    double xc = t2 / (nu + t2);  // 1 - x without cancellation
    double ic = exp(log_ibeta_direct(b, a, xc, D.lbeta));
    return log(0.5 - 0.5 * ic);
}

// F(tau) - 1/2 for tau >= 0, absolute accuracy near 0.
__device__ __noinline__ double t_central_mass(const TDist& D, double tau) {
    double nu = D.nu, a = 0.5 * nu, b = 0.5;
    double t2 = tau * tau;
    if (t2 == 0.0) return 0.0;
    double xc = t2 / (nu + t2);
    if (xc < (b + 1.0) / (a + b + 2.0)) return 0.5 * exp(log_ibeta_direct(b, a, xc, D.lbeta));
    double x = nu / (nu + t2);
    return 0.5 - 0.5 * exp(log_ibeta_direct(a, b, x, D.lbeta));
}

__device__ __forceinline__ double t_log_pdf(const TDist& D, double tau) {
    return D.log_pdf0 - 0.5 * (D.nu + 1.0) * log1p(tau * tau / D.nu);
}

// |T_nu^{-1}(p)| for p in (0, 1/2]: Newton on log F in log|t| for the tail, Newton on F - 1/2 in the
// centre.  Used to build the table and as the test reference; never on the per-solve path.
__device__ __noinline__ double t_quantile_mag_iterative(const TDist& D, double p) {
    if (!(p > 0.0)) return INFINITY;
    if (p >= 0.5) return 0.0;
    double nu = D.nu;
    if (p < 0.2) {
        // start right of the root: leading-order tail  F(-tau) <= c nu^{(nu-1)/2} tau^{-nu}
        double lp = log(p);
        double lc = D.log_pdf0 + 0.5 * (nu - 1.0) * log(nu);
        double s = (lc - lp) / nu;  // log tau
        double wn = -normcdfinv(p);
        if (s < log(wn)) s = log(wn);
        for (int it = 0; it < 200; ++it) {
            double tau = exp(s);
            double lF = t_log_lower_tail(D, tau);
            double slope = -exp(t_log_pdf(D, tau) - lF) * tau;  // d log F / d log tau
            double ds = (lp - lF) / slope;
            if (ds > 2.0) ds = 2.0;
            if (ds < -2.0) ds = -2.0;
            s += ds;
            if (fabs(ds) < 4e-16 * fmax(1.0, fabs(s))) break;
        }
        return exp(s);
    }
    double delta = 0.5 - p;  // exact for p in [1/4, 1/2]
    double tau = -normcdfinv(p);
    for (int it = 0; it < 200; ++it) {
        double G = t_central_mass(D, tau) - delta;
        double dt = -G / exp(t_log_pdf(D, tau));
        double tn = tau + dt;
        if (tn <= 0.0) tn = 0.5 * tau;
        double change = fabs(tn - tau);
        tau = tn;
        if (change < 4e-16 * tau) break;
    }
    return tau;
}

// ----- Student-t quantile table ------------------------------------------------------------------
// |T_nu^{-1}(Phi(w))| = |w| * rho(w) on w in [-W_MAX, 0]; rho is tabulated per interval as a Chebyshev
// series of degree TQ_DEG (nu is a run constant, so the table is built once per plan).
constexpr int TQ_INTERVALS = 256;
constexpr int TQ_DEG = 9;
constexpr double TQ_WMAX = 9.5;
constexpr int TQ_TABLE_DOUBLES = TQ_INTERVALS * (TQ_DEG + 1);

__device__ __forceinline__ double tq_ratio_exact(const TDist& D, double w /* < 0 */) {
    double p = 0.5 * erfc(-w * 0.7071067811865476);
    return t_quantile_mag_iterative(D, p) / (-w);
}

// one thread per (interval, node) computes a sample, then one thread per interval does the DCT
__global__ void tq_table_build_kernel(double nu, double* __restrict__ table) {
    __shared__ double f[TQ_DEG + 1];
    const int N = TQ_DEG + 1;
    int m = blockIdx.x;
    int k = threadIdx.x;
    const double h = TQ_WMAX / TQ_INTERVALS;
    if (k < N) {
        TDist D = tdist_make(nu);
        double sk = cospi((k + 0.5) / N);
        double w = -TQ_WMAX + h * (m + 0.5 * (1.0 + sk));
        f[k] = tq_ratio_exact(D, w);
    }
    __syncthreads();
    if (k < N) {
        double acc = 0.0;
        for (int i = 0; i < N; ++i) acc += f[i] * cospi(k * (i + 0.5) / N);
        acc *= 2.0 / N;
        if (k == 0) acc *= 0.5;
        table[m * N + k] = acc;
    }
}

// |T_nu^{-1}(p)| for p in (0, 1/2] from the table (leading-order tail beyond it).
__device__ __forceinline__ double t_quantile_mag_table(const double* __restrict__ table, double nu, double tail_lc,
                                                       double p) {
    double w = normcdfinv(p);  // <= 0
    if (w < -TQ_WMAX) return exp((tail_lc - log(p)) / nu);
    const double inv_h = TQ_INTERVALS / TQ_WMAX;
    double pos = (w + TQ_WMAX) * inv_h;
    int m = min((int)pos, TQ_INTERVALS - 1);
    double s = 2.0 * (pos - m) - 1.0;  // [-1, 1]
    const double* c = table + m * (TQ_DEG + 1);
    double b1 = 0.0, b2 = 0.0, s2 = 2.0 * s;
#pragma unroll
    for (int k = TQ_DEG; k >= 1; --k) {
        double b0 = fma(s2, b1, c[k] - b2);
        b2 = b1;
        b1 = b0;
    }
    double ratio = fma(s, b1, c[0] - b2);
    return -w * ratio;
}

// signed Student-t quantile of u in [0, 1]: -inf at 0, +inf at 1, NaN outside.
__device__ __forceinline__ double t_quantile_table(const double* __restrict__ table, double nu, double tail_lc,
                                                   double u) {
    if (!(u >= 0.0 && u <= 1.0)) return NAN;
    bool low = u < 0.5;
    double p = low ? u : 1.0 - u;
    if (p == 0.0) return low ? -INFINITY : INFINITY;
    double mag = t_quantile_mag_table(table, nu, tail_lc, p);
    return low ? -mag : mag;
}

__device__ __forceinline__ double t_quantile_iterative(const TDist& D, double u) {
    if (!(u >= 0.0 && u <= 1.0)) return NAN;
    bool low = u < 0.5;
    double p = low ? u : 1.0 - u;
    if (p == 0.0) return low ? -INFINITY : INFINITY;
    double mag = t_quantile_mag_iterative(D, p);
    return low ? -mag : mag;
}

}  // namespace cvar
