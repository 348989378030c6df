// Shared-memory wavefront cost of the access patterns the cell loops use (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_patterns lds_patterns.cu && ./lds_patterns
// One CTA of 128 threads per SM (one warp per scheduler), every warp issues independent LDS in a loop; the time per
// warp-level load at SM level is the number of data-pipe cycles (wavefronts) the pattern costs.
#include <cstdio>
#include <cuda_runtime.h>

template <int BYTES>
__global__ void k(const int* __restrict__ offs, int iters, double* out, long long* cyc) {
    extern __shared__ __align__(16) unsigned char sm[];
    for (int i = threadIdx.x; i < 32768 / 8; i += blockDim.x) reinterpret_cast<double*>(sm)[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    unsigned a = (unsigned)__cvta_generic_to_shared(sm) + offs[lane];
    unsigned acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (BYTES == 16) {
                unsigned x, y, z, w;
                asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(a + 4096u * u));
                acc ^= x ^ w;
            } else {
                unsigned x, y;
                asm volatile("ld.volatile.shared.v2.u32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "r"(a + 4096u * u));
                acc ^= x ^ y;
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    if (acc == 0x12345u) out[0] = acc;
}

int main() {
    int h[32];
    int* d; double* o; long long* c;
    cudaMalloc(&d, 128); cudaMalloc(&o, 8); cudaMalloc(&c, 8);
    cudaFuncSetAttribute(k<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    struct Pat { const char* name; int bytes; int (*f)(int); };
    auto run = [&](const char* name, int bytes) {
        cudaMemcpy(d, h, 128, cudaMemcpyHostToDevice);
        const int iters = 1024;
        for (int rep = 0; rep < 2; ++rep) {
            if (bytes == 16) k<16><<<148, 1024, 65536>>>(d, iters, o, c); else k<8><<<148, 1024, 65536>>>(d, iters, o, c);
            cudaDeviceSynchronize();
        }
        long long cy; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
        printf("%-58s %2d B: %.2f cycles per warp load (32 warps per SM issuing)\n", name, bytes, (double)cy / (iters * 8.0 * 32.0));
    };
    for (int l = 0; l < 32; ++l) h[l] = 16 * l;            run("32 consecutive 16-byte words", 16);
    for (int l = 0; l < 32; ++l) h[l] = 0;                 run("one 16-byte word, all lanes (broadcast)", 16);
    for (int l = 0; l < 32; ++l) h[l] = 16 * (l / 8);      run("4 groups of 8 lanes, words in distinct banks", 16);
    for (int l = 0; l < 32; ++l) h[l] = 128 * (l / 8);     run("4 groups of 8 lanes, words in the same banks", 16);
    for (int l = 0; l < 32; ++l) h[l] = 16 * (l / 2);      run("16 consecutive words, 2 lanes each", 16);
    for (int l = 0; l < 32; ++l) h[l] = 32 * l;            run("32 words at a 32-byte stride", 16);
    for (int l = 0; l < 32; ++l) h[l] = 16 * ((l * 37) % 251);  run("32 scattered words (256-entry table)", 16);
    for (int l = 0; l < 32; ++l) h[l] = 16 * (100 + l / 3 + 40 * (l / 8));  run("4 groups, 3 neighbouring entries each", 16);
    for (int l = 0; l < 32; ++l) h[l] = 8 * l;             run("32 consecutive 8-byte words", 8);
    for (int l = 0; l < 32; ++l) h[l] = 0;                 run("one 8-byte word (broadcast)", 8);
    for (int l = 0; l < 32; ++l) h[l] = 8 * (l / 2);       run("16 consecutive words, 2 lanes each", 8);
    for (int l = 0; l < 32; ++l) h[l] = 8 * ((l * 37) % 251);   run("32 scattered words (256-entry table)", 8);
    for (int l = 0; l < 32; ++l) h[l] = 8 * ((l * 89) % 1531);  run("32 scattered words (1536-entry table)", 8);
    for (int l = 0; l < 32; ++l) h[l] = 8 * (100 + 3 * l); run("stride of 3 entries per lane", 8);
    for (int l = 0; l < 32; ++l) h[l] = 8 * (100 + l / 3 + 40 * (l / 8));   run("4 groups, 3 neighbouring entries each", 8);
    for (int l = 0; l < 32; ++l) h[l] = 8 * (100 + l / 4); run("8 neighbouring entries, 4 lanes each", 8);
    return 0;
}
