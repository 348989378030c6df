// Is MUFU.RCP64H exponent-transparent?  rcp(2^e * mid_i) == 2^-e * rcp(mid_i) for all table intervals and exponents,
// and is the low word of the operand ignored?  (Both hold on B200; pow_neg_c in csrc/cvar_math.cuh relies on it.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_check tools/mufu_check.cu && ./mufu_check
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int bits, unsigned long long* bad) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (1 << bits)) return;
    const int mid_hi = 0x3ff00000 | (idx << (20 - bits)) | (1 << (19 - bits));
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(__hiloint2double(mid_hi, 0)));
    for (int e = 0; e < 64; ++e) {
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(__hiloint2double(mid_hi + (e << 20), 0)));
        const double want = __hiloint2double(__double2hiint(r0) - (e << 20), __double2loint(r0));
        if (r != want) atomicAdd(bad, 1ULL);
        // low-word insensitivity: garbage in the low word of the input must not change the result
        double r2;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r2) : "d"(__hiloint2double(mid_hi + (e << 20), 0x9e3779b9 * (idx + 1))));
        if (r2 != r) atomicAdd(bad + 1, 1ULL);
    }
}
int main() {
    unsigned long long* bad; cudaMalloc(&bad, 16); 
    for (int bits = 7; bits <= 10; ++bits) {
        cudaMemset(bad, 0, 16);
        k<<<((1 << bits) + 127) / 128, 128>>>(bits, bad);
        unsigned long long h[2]; cudaMemcpy(h, bad, 16, cudaMemcpyDeviceToHost);
        printf("bits %d: exponent mismatches %llu, low-word mismatches %llu\n", bits, h[0], h[1]);
    }
}
