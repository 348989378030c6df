// Isolated cell loop of the Student-t power cell (sm_100a): cycles per warp-cell per SM sub-partition for several
// formulations, with the solve kernel's residency (2 CTAs x 256 threads per SM, <= 128 registers).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../copula-msm-and-copula-garch-var_b200/csrc -o cell_loop cell_loop.cu
#include <cstdio>
#include <cmath>
#include <vector>
#include "cvar_math.cuh"
using namespace cvar;

constexpr int N = 2048, OCT = 6, DEG = 5;
struct Params {
    double powc[POW_MAX_DEG + 1];
    unsigned seed_mask, seed_half;
    const double *in, *rows, *pm_pe, *utab;  // in: [N][2] (a, b); rows: [N][2] (m0, c0)
    int L, rounds, group;
};

__device__ __forceinline__ double lds64(unsigned a) { double v; asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }

// V: 0 two tables (round-1 cell)   1 combined table, masked offset   2 combined table, offset from the seed word
//    3 = 2 with polynomial coefficients in registers
template <int V>
__device__ __forceinline__ double cell(const Params& P, const double* kc, double a, double b, double m0, double c0, double acc,
                                       unsigned ptab_s, unsigned utab_s) {
    const double d = a - m0;
    const double t = fma(d, d, c0);
    const unsigned hi = (unsigned)__double2hiint(t);
    const unsigned sh = (hi & P.seed_mask) | P.seed_half;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(__hiloint2double((int)sh, 0)));
    const double f = fma(t, r, -1.0);
    double p = kc[DEG];
#pragma unroll
    for (int k = DEG - 1; k >= 0; --k) p = fma(p, f, kc[k]);
    double u;
    if (V == 0) {
        const unsigned mb = hi & (unsigned)((POW_MTAB - 1) << (20 - POW_BITS));
        const unsigned eb = hi & 0x7ff00000u;
        const double pm = lds64(ptab_s + (mb >> (20 - POW_BITS - 3)));
        const double pe = lds64(ptab_s + (unsigned)((POW_MTAB - 1023) * 8) + (eb >> 17));
        u = pm * pe;
    } else if (V == 1) {
        u = lds64(utab_s + ((hi >> 9) & ~7u));
    } else {
        u = lds64(utab_s - 4u + (sh >> 9));   // sh has bit 11 set and bits 10..0 clear: (sh >> 9) = 8 * index + 4
    }
    return fma(u * b, p, acc);
}

template <int V, int CIF, bool PREFETCH>
__global__ void __launch_bounds__(256, 2) k(Params P, double* out, long long* cyc) {
    extern __shared__ __align__(16) unsigned char sm[];
    double2* in = reinterpret_cast<double2*>(sm);
    double* ptab = reinterpret_cast<double*>(sm + N * 16);
    double* utab = ptab + POW_MTAB + POW_ETAB;
    for (int i = threadIdx.x; i < N; i += blockDim.x) in[i] = make_double2(P.in[2 * i], P.in[2 * i + 1]);
    for (int i = threadIdx.x; i < POW_MTAB + POW_ETAB; i += blockDim.x) ptab[i] = P.pm_pe[i];
    for (int i = threadIdx.x; i < OCT * POW_MTAB; i += blockDim.x) utab[i] = P.utab[i];
    __syncthreads();
    const unsigned ptab_s = (unsigned)__cvta_generic_to_shared(ptab);
    const unsigned utab_s = (unsigned)__cvta_generic_to_shared(utab) - (unsigned)((1023 << POW_BITS) * 8);
    double kc[DEG + 1];
#pragma unroll
    for (int k = 0; k <= DEG; ++k) kc[k] = P.powc[k];
    const double* kcp = (V == 3) ? kc : P.powc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double total = 0.0;
    const long long t0 = clock64();
    for (int r = 0; r < P.rounds; ++r) {
        const int i = ((r * 8 + warp) * 32 + lane) % N;
        const double m0 = P.rows[2 * i], c0 = P.rows[2 * i + 1];
        // group == 0: lane-per-row walk (adjacent lanes one column apart); else groups of `group` lanes share a column
        const int off = P.group == 0 ? lane : 9 * (lane / P.group);
        const double2* col = in + 700 - off;
        double acc[CIF];
#pragma unroll
        for (int c = 0; c < CIF; ++c) acc[c] = 0.0;
        if (!PREFETCH) {
            for (int j = 0; j < P.L; j += CIF, col += CIF) {
                double2 v[CIF];
#pragma unroll
                for (int c = 0; c < CIF; ++c) v[c] = col[c];
#pragma unroll
                for (int c = 0; c < CIF; ++c) acc[c] = cell<V == 3 ? 2 : V>(P, kcp, v[c].x, v[c].y, m0, c0, acc[c], ptab_s, utab_s);
            }
        } else {
            double2 v[CIF];
#pragma unroll
            for (int c = 0; c < CIF; ++c) v[c] = col[c];
            for (int j = 0; j < P.L; j += CIF) {
                col += CIF;
                double2 w[CIF];
#pragma unroll
                for (int c = 0; c < CIF; ++c) w[c] = col[c];   // next trip (reads up to CIF columns past the range; in bounds)
#pragma unroll
                for (int c = 0; c < CIF; ++c) acc[c] = cell<V == 3 ? 2 : V>(P, kcp, v[c].x, v[c].y, m0, c0, acc[c], ptab_s, utab_s);
#pragma unroll
                for (int c = 0; c < CIF; ++c) v[c] = w[c];
            }
        }
        double s = acc[0];
#pragma unroll
        for (int c = 1; c < CIF; ++c) s += acc[c];
        total += s;
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = total;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

double fit(double c, double fm, int D, double* coef) {   // same construction as the plan (cvar_api.cu, fit_pow_series)
    const long double PI = 3.141592653589793238462643383279502884L;
    long double g[16], a[16], mono[16] = {0}, T[16][16] = {{0}};
    for (int k = 0; k <= D; ++k) g[k] = powl(1.0L + (long double)fm * cosl(PI * (k + 0.5L) / (D + 1)), -(long double)c);
    for (int j = 0; j <= D; ++j) { long double acc = 0; for (int k = 0; k <= D; ++k) acc += g[k] * cosl(j * PI * (k + 0.5L) / (D + 1)); a[j] = acc * (j == 0 ? 1.0L : 2.0L) / (D + 1); }
    T[0][0] = 1; T[1][1] = 1;
    for (int j = 2; j <= D; ++j) for (int i = 0; i <= j; ++i) T[j][i] = (i > 0 ? 2 * T[j - 1][i - 1] : 0) - T[j - 2][i];
    for (int j = 0; j <= D; ++j) for (int i = 0; i <= j; ++i) mono[i] += a[j] * T[j][i];
    long double sc = 1; for (int i = 0; i <= D; ++i) { coef[i] = (double)(mono[i] / sc); sc *= fm; }
    return 0;
}

template <int V, int CIF, bool PF>
void run(const char* name, Params P, int group, double* out, long long* cyc, double ref[2]) {
    P.group = group;
    const size_t smem = N * 16 + (POW_MTAB + POW_ETAB + OCT * POW_MTAB) * 8;
    cudaFuncSetAttribute(k<V, CIF, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem + 49152);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<V, CIF, PF>);
    const int grid = 296;
    for (int rep = 0; rep < 2; ++rep) { k<V, CIF, PF><<<grid, 256, smem + 49152>>>(P, out, cyc); cudaDeviceSynchronize(); }
    std::vector<long long> c(grid); cudaMemcpy(c.data(), cyc, 8 * grid, cudaMemcpyDeviceToHost);
    double mean = 0; for (auto v : c) mean += v; mean /= grid;
    std::vector<double> o(grid * 256); cudaMemcpy(o.data(), out, 8 * o.size(), cudaMemcpyDeviceToHost);
    double sum = 0; for (int i = 0; i < 256; ++i) sum += o[i];
    // one SM sub-partition runs 4 warps (2 CTAs x 8 warps / 4); each warp does rounds * L warp-cells
    const double per = mean / (4.0 * P.rounds * P.L);
    if (ref[0] == 0) ref[0] = sum;
    printf("%-44s group %2d: %6.2f cycles per warp-cell per sub-partition, %3d regs, checksum rel diff %.1e (%s)\n", name, group, per,
           fa.numRegs, fabs(sum - ref[0]) / fabs(ref[0]), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const double nu = 5.3, rho = 0.6, c = 0.5 * (nu + 2);
    Params P = {};
    fit(c, POW_FMAX, DEG, P.powc);
    P.seed_mask = ~((1u << (20 - POW_BITS)) - 1u);
    P.seed_half = 1u << (19 - POW_BITS);
    std::vector<double> in(2 * N), rows(2 * N);
    const double cs = 1.0 / sqrt(nu * (1 - rho * rho));
    for (int i = 0; i < N; ++i) {
        const double y = -4.0 + 8.0 * i / (N - 1);
        in[2 * i] = cs * y; in[2 * i + 1] = exp(-0.5 * y * y) * 0.004;
        rows[2 * i] = rho * cs * y; rows[2 * i + 1] = 1 + y * y / nu;
    }
    double *d_in, *d_rows, *d_pm, *d_u, *d_out; long long* d_cyc;
    cudaMalloc(&d_in, 16 * N); cudaMalloc(&d_rows, 16 * N); cudaMalloc(&d_pm, 8 * (POW_MTAB + POW_ETAB)); cudaMalloc(&d_u, 8 * OCT * POW_MTAB * 2);
    cudaMalloc(&d_out, 8 * 296 * 256); cudaMalloc(&d_cyc, 8 * 296);
    cudaMemcpy(d_in, in.data(), 16 * N, cudaMemcpyHostToDevice); cudaMemcpy(d_rows, rows.data(), 16 * N, cudaMemcpyHostToDevice);
    powtab_build_kernel<<<1, POW_MTAB>>>(c, d_pm);
    static_assert(POW_FAST_MODE == 1, "build with -DCVAR_POW_FAST=1");
    powfast_build_kernel<<<(OCT * POW_MTAB + 255) / 256, 256>>>(c, OCT, d_u);
    P.in = d_in; P.rows = d_rows; P.pm_pe = d_pm; P.utab = d_u; P.L = 512; P.rounds = 64;
    double ref[2] = {0, 0};
    for (int g : {0, 8, 32}) {
        run<0, 4, false>("two tables, 4 cells in flight", P, g, d_out, d_cyc, ref);
        run<1, 4, false>("one table, masked offset", P, g, d_out, d_cyc, ref);
        run<2, 4, false>("one table, offset from the seed word", P, g, d_out, d_cyc, ref);
        run<3, 4, false>("  + coefficients in registers", P, g, d_out, d_cyc, ref);
        run<2, 4, true>("  + next trip's columns prefetched", P, g, d_out, d_cyc, ref);
        run<2, 8, false>("  8 cells in flight", P, g, d_out, d_cyc, ref);
        run<2, 6, false>("  6 cells in flight", P, g, d_out, d_cyc, ref);
        run<2, 2, false>("  2 cells in flight", P, g, d_out, d_cyc, ref);
    }
    return 0;
}
