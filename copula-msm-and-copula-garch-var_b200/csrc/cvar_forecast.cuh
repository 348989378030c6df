// cvar_forecast.cuh -- forecast producers feeding the VaR solve.
//
// MSM: the per-day input of the mixture solve is the filtered state distribution of a binomial Markov-switching
// multifractal at the end of each rolling window (reference: markov_switching_multifractal/calc_prob.py:8-69
// `calc_state_prob_numba`, calc_marginals.py:33-38 `calc_forecasts`, driven date by date from
// utils/model_estimation/model/msm_estimation.py:143-203).  The reference runs a dense 2^k x 2^k mat-vec per return
// (N * 4^k flops per window); here
//   * one WARP owns one (window, asset) and keeps the 2^k-vector in registers (state s = v*32 + lane),
//   * the transition matrix is applied as the Kronecker product of its k 2x2 factors -- k butterfly stages,
//     in-register for the high bits, __shfl_xor for the low five (2 k 2^k flops instead of 4^k),
//   * the state likelihoods N(r_t; 0, sigma_s) are tabulated once per return (rolling windows overlap in all but
//     one return) and read coalesced,
//   * normalisation is a warp shuffle reduction; there is no block-level synchronisation at all,
//   * the merge of the 2^k states into the k+1 distinct vol levels (msm_estimation.py:205-248) is fused at the end,
//     writing the solve's day_params[T][2][q] layout directly.
// GARCH: one thread per (window, asset) runs the conditional-variance recursion (garch/estimation.py:40-65) and the
// one-step forecast (garch/forecast.py:5-18).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace cvar {

constexpr int MSM_MAX_K = 10;
constexpr int MSM_WARPS_PER_CTA = 4;

struct MsmAsset {
    double stay[MSM_MAX_K];  // p_c = 1 - gamma_c / 2 for component c (c = 0 is the most significant state bit)
};

// lik[t][s] = N(r_t; 0, sigma_s)   (calc_prob.py:116-118)
__global__ void msm_likelihood_kernel(const double* __restrict__ returns, long long L, const double* __restrict__ vol_states,
                                      int S, double* __restrict__ lik) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L * S) return;
    const int s = (int)(idx % S);
    const double sg = vol_states[s];
    const double z = returns[idx / S] / sg;
    lik[idx] = (1.0 / (sg * 2.5066282746310002)) * exp(-0.5 * (z * z));
}

template <int K>
__global__ void __launch_bounds__(32 * MSM_WARPS_PER_CTA)
msm_filter_kernel(MsmAsset A, const double* __restrict__ lik, long long T, int N, long long window_stride,
                  const int* __restrict__ level_of_state, int q, double* __restrict__ probs_by_state, long long out_stride,
                  double* __restrict__ state_probs, int* __restrict__ status) {
    constexpr int S = 1 << K;
    constexpr int VPL = S >= 32 ? S / 32 : 1;  // values per lane
    constexpr int LANE_BITS = K < 5 ? K : 5;
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * MSM_WARPS_PER_CTA + (threadIdx.x >> 5);
    if (w >= T) return;
    const bool live = (S >= 32) || (lane < S);
    double pi[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) pi[v] = live ? 1.0 / S : 0.0;  // uniform prior (calc_prob.py:12-13)
    const double* row = lik + (w * window_stride) * S;
    bool degenerate = false;
    for (int i = 0; i < N; ++i, row += S) {
        // predict: pi <- (A_0 (x) A_1 (x) ... (x) A_{K-1}) pi, one 2x2 factor per state bit
#pragma unroll
        for (int bit = 0; bit < K; ++bit) {
            const double p = A.stay[K - 1 - bit], qq = 1.0 - p;
            if (bit < LANE_BITS) {
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    const double other = __shfl_xor_sync(0xffffffffu, pi[v], 1 << bit);
                    pi[v] = fma(p, pi[v], qq * other);
                }
            } else {
                const int m = 1 << (bit - 5);
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    if ((v & m) == 0) {
                        const double a = pi[v], b = pi[v | m];
                        pi[v] = fma(p, a, qq * b);
                        pi[v | m] = fma(p, b, qq * a);
                    }
                }
            }
        }
        // update with the likelihood of return i and renormalise (calc_prob.py:57-68)
        double sum = 0.0;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int s = v * 32 + lane;
            pi[v] = live ? pi[v] * row[s < S ? s : 0] : 0.0;
            sum += pi[v];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (sum == 0.0) {  // the reference flags the whole run as failed (returns -1 everywhere)
            degenerate = true;
            break;
        }
        const double inv = 1.0 / sum;
#pragma unroll
        for (int v = 0; v < VPL; ++v) pi[v] *= inv;
    }
    if (degenerate) {
        if (lane == 0) atomicExch(status, 1);
#pragma unroll
        for (int v = 0; v < VPL; ++v) pi[v] = NAN;
    }
    if (state_probs) {
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int s = v * 32 + lane;
            if (s < S) state_probs[w * S + s] = pi[v];
        }
    }
    // merge states of equal volatility level (msm_estimation.py:225-236)
    for (int l = 0; l < q; ++l) {
        double acc = 0.0;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int s = v * 32 + lane;
            if (s < S && level_of_state[s] == l) acc += pi[v];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) probs_by_state[w * out_stride + l] = acc;
    }
}

// GARCH(p,q) one-step volatility forecast per rolling window (garch/estimation.py:40-65, garch/forecast.py:5-18)
constexpr int GARCH_MAX_ORDER = 8;
struct GarchAsset {
    double omega;
    int p, q;
    double alpha[GARCH_MAX_ORDER];
    double beta[GARCH_MAX_ORDER];
};

__global__ void garch_forecast_kernel(GarchAsset G, const double* __restrict__ returns, long long T, int N,
                                      long long window_stride, double* __restrict__ sigma_out, long long out_stride) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= T) return;
    const double* r = returns + w * window_stride;
    double sa = 0.0, sb = 0.0;
    for (int i = 0; i < G.p; ++i) sa += G.alpha[i];
    for (int j = 0; j < G.q; ++j) sb += G.beta[j];
    double hist[GARCH_MAX_ORDER];  // hist[j] = sigma2[t - 1 - j]
    for (int j = 0; j < GARCH_MAX_ORDER; ++j) hist[j] = 0.0;
    hist[0] = G.omega / (1.0 - sa - sb);
    for (int t = 1; t < N; ++t) {
        double s2 = G.omega;
        const int pm = G.p < t ? G.p : t, qm = G.q < t ? G.q : t;
        for (int i = 0; i < pm; ++i) s2 += G.alpha[i] * (r[t - i - 1] * r[t - i - 1]);
        for (int j = 0; j < qm; ++j) s2 += G.beta[j] * hist[j];
        s2 = fmax(s2, 1e-7);
        for (int j = GARCH_MAX_ORDER - 1; j > 0; --j) hist[j] = hist[j - 1];
        hist[0] = s2;
    }
    // forecast = omega + sum(alpha * returns[-p:]**2) + sum(beta * sigma2[-q:])  -- note the pairing: alpha[0]
    // meets the OLDEST of the last p returns (quirk Q16), summed in NumPy's left-to-right order
    double fa = 0.0, fb = 0.0;
    for (int i = 0; i < G.p; ++i) fa += G.alpha[i] * (r[N - G.p + i] * r[N - G.p + i]);
    for (int j = 0; j < G.q; ++j) fb += G.beta[j] * hist[G.q - 1 - j];
    sigma_out[w * out_stride] = sqrt(G.omega + fa + fb);
}

// Kalman mean-reverting log-vol model: scalar unscented filter, forecast = exp(last PREDICTED state mean)
// (kalman_mean_reverting/estimate.py:231-281 `calculate_loglikelihood`, forecast.py:5-12).  One thread per
// (window, asset).  The reference's peculiarities are kept: sigma-point weights built with L = 2 also for the
// one-dimensional update (three points, weights that do not sum to one), measurement h = phi(eta)|eta|,
// eta = r / exp(x), covariance regularisation 1e-8 when the variance is not positive.
struct KalmanAsset {
    double a, l, q;          // mean reversion speed, long-run mean, vol of the log-vol process
    double alpha, beta, kappa;  // UKF tuning (defaults 1.6, 2, 1.75)
};

__global__ void kalman_forecast_kernel(KalmanAsset K, const double* __restrict__ returns, long long T, int N,
                                       long long window_stride, double* __restrict__ sigma_out, long long out_stride,
                                       int* __restrict__ status) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= T) return;
    const double* r = returns + w * window_stride;
    const double L = 2.0;
    const double lambda = K.alpha * K.alpha * (L + K.kappa) - L;
    const double phi = sqrt(L + lambda);
    const double w_rest = 1.0 / (2.0 * (L + lambda));
    const double wm0 = lambda / (L + lambda);
    const double wc0 = wm0 + (1.0 - K.alpha * K.alpha + K.beta);
    double mean = K.l, var = K.q;  // init_log_vol = l, init_var = q (forecast.py:9)
    double x_pred = 0.0;
    bool failed = false;
    for (int t = 0; t < N; ++t) {
        // prediction: sigma points of the augmented state (x, noise) with covariance diag(var, 1)
        double dv = var;
        if (dv <= 0.0) dv += 1e-8;
        const double sv = sqrt(dv);
        const double base = K.a * (mean - K.l) + K.l;
        const double X0 = base, X1 = K.a * (mean + phi * sv - K.l) + K.l, X2 = base + K.q * phi;
        const double X3 = K.a * (mean - phi * sv - K.l) + K.l, X4 = base + K.q * (-phi);
        x_pred = X0 * wm0 + X1 * w_rest + X2 * w_rest + X3 * w_rest + X4 * w_rest;
        const double d0 = X0 - x_pred, d1 = X1 - x_pred, d2 = X2 - x_pred, d3 = X3 - x_pred, d4 = X4 - x_pred;
        const double P = d0 * wc0 * d0 + d1 * w_rest * d1 + d2 * w_rest * d2 + d3 * w_rest * d3 + d4 * w_rest * d4;
        // update with return t on three sigma points
        const double sp = sqrt(P);
        const double Y[3] = {x_pred, x_pred + phi * sp, x_pred - phi * sp};
        const double wm2[3] = {wm0, w_rest, w_rest};
        double h[3], Z = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double eta = r[t] / exp(Y[k]);
            h[k] = (0.3989422804014327 * exp(-0.5 * eta * eta)) * fabs(eta);
            Z += wm2[k] * h[k];
        }
        if (!(Z > 0.0) || Z < 1e-10) {
            failed = true;
            break;
        }
        double m = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) m += (wm2[k] * Y[k] * h[k]) / Z;
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) v += wm2[k] * ((h[k] / Z) * ((Y[k] - m) * (Y[k] - m)));
        mean = m;
        var = v;
    }
    if (failed) atomicExch(status, 1);
    sigma_out[w * out_stride] = failed ? NAN : exp(x_pred);
}

}  // namespace cvar
