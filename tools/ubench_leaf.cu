// ubench_leaf.cu — throughput of the 32x32 carry-less leaf (clmul32_imad: 16 IMAD.WIDE + ~28 LOP3) and of the 8x8-word
// Karatsuba built on it, in isolation, at several occupancies.  Tells whether the product kernels are limited by the
// leaf itself or by what surrounds it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I homomorph_rust_b200/csrc -o tools/ubench_leaf tools/ubench_leaf.cu
#include <cstdio>
#include "kernels.cuh"

template <int CHAINS, int MINB> __global__ void __launch_bounds__(128, MINB) leaf_kernel(uint32_t *out, int iters, uint32_t seed) {
    uint32_t a[CHAINS], b[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { a[c] = seed * (threadIdx.x + 1) + c; b[c] = seed ^ (blockIdx.x * 977 + c); }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            uint32_t lo, hi;
            hmk::clmul32_imad(a[c], b[c], lo, hi);
            a[c] ^= lo; b[c] += hi;
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc ^= a[c] ^ b[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// variants of the leaf to locate its bottleneck: 0 = products only (16 IMAD.WIDE, results XORed together with 3-input
// LOP3s: 16 IMAD.WIDE + ~11 LOP3), 1 = ALU part only (splits + combine on fake products), 2 = products ordered by b class
template <int VAR> __global__ void __launch_bounds__(128, 8) leafvar_kernel(uint32_t *out, int iters, uint32_t seed) {
    uint32_t a[4], b[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { a[c] = seed * (threadIdx.x + 1) + c; b[c] = seed ^ (blockIdx.x * 977 + c); }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            uint32_t as[4], bs[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) { as[c] = a[ch] & (0x11111111u << c); bs[c] = b[ch] & (0x11111111u << c); }
            uint32_t lo = 0, hi = 0;
            if (VAR == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const unsigned long long p = (unsigned long long)as[i] * bs[j];
                        lo ^= (uint32_t)p; hi ^= (uint32_t)(p >> 32);
                    }
            } else if (VAR == 1) {
#pragma unroll
                for (int kc = 0; kc < 4; ++kc) {
                    const uint32_t m = 0x11111111u << kc;
                    lo |= (as[0] ^ bs[kc] ^ as[1] ^ bs[(kc + 1) & 3]) & m;
                    hi |= (as[2] ^ bs[(kc + 2) & 3] ^ as[3] ^ bs[(kc + 3) & 3]) & m;
                    lo ^= (as[kc] ^ hi) & m; hi ^= (bs[kc] ^ lo) & m; // pad to ~28 ALU ops
                }
            } else {
#pragma unroll
                for (int kc = 0; kc < 4; ++kc) {
                    unsigned long long p[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) p[c] = (unsigned long long)bs[c] * as[(kc - c) & 3];
                    const uint32_t m = 0x11111111u << kc;
                    lo |= ((uint32_t)p[0] ^ (uint32_t)p[1] ^ (uint32_t)p[2] ^ (uint32_t)p[3]) & m;
                    hi |= ((uint32_t)(p[0] >> 32) ^ (uint32_t)(p[1] >> 32) ^ (uint32_t)(p[2] >> 32) ^ (uint32_t)(p[3] >> 32)) & m;
                }
            }
            a[ch] ^= lo; b[ch] += hi;
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) acc ^= a[c] ^ b[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MINB> __global__ void __launch_bounds__(128, MINB) kara_kernel(uint32_t *out, int iters, uint32_t seed) {
    uint32_t a[8], b[8], r[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed * (threadIdx.x + 1) + i; b[i] = seed ^ (blockIdx.x * 977 + i); }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        hmk::clmul_kara<8>(a, b, r);
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] ^= r[i]; b[i] += r[8 + i]; }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= a[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class K> static void run(const char *name, K kern, int per_sm, int sms, double units_per_thread_iter, int iters, uint32_t *out) {
    const int blocks = sms * per_sm;
    kern<<<blocks, 128>>>(out, iters, 12345u);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<<<blocks, 128>>>(out, iters, 12345u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double units = (double)blocks * 128 * iters * units_per_thread_iter;
    printf("%-44s %2d CTAs/SM (%2d warps)  %8.2f G/s   (%.3f ms)\n", name, per_sm, per_sm * 4, units / ms / 1e6, ms);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    uint32_t *out; cudaMalloc(&out, 64 << 20);
    const int sms = p.multiProcessorCount;
    run("leaf 32x32, 1 chain/thread", leaf_kernel<1, 8>, 8, sms, 1, 20000, out);
    run("leaf 32x32, 2 chains/thread", leaf_kernel<2, 8>, 8, sms, 2, 10000, out);
    run("leaf 32x32, 4 chains/thread", leaf_kernel<4, 4>, 4, sms, 4, 5000, out);
    run("leaf 32x32, 4 chains/thread", leaf_kernel<4, 8>, 8, sms, 4, 5000, out);
    run("leaf 32x32, 4 chains/thread", leaf_kernel<4, 16>, 16, sms, 4, 5000, out);
    run("leaf variant: 16 IMAD.WIDE + xor only", leafvar_kernel<0>, 8, sms, 4, 5000, out);
    run("leaf variant: ALU part only", leafvar_kernel<1>, 8, sms, 4, 5000, out);
    run("leaf variant: products ordered by b class", leafvar_kernel<2>, 8, sms, 4, 5000, out);
    run("kara8 (27 leaves)", kara_kernel<4>, 4, sms, 1, 1000, out);
    run("kara8 (27 leaves)", kara_kernel<6>, 6, sms, 1, 1000, out);
    run("kara8 (27 leaves)", kara_kernel<8>, 8, sms, 1, 1000, out);
    return cudaDeviceSynchronize() != cudaSuccess;
}
