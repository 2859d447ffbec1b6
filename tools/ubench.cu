// ubench.cu — issue-rate microbenchmarks for the integer pipes the GF(2)[X] kernels live on.
// Prints warp-instructions per clock per SM for LOP3, SHF, IMAD, IMAD.WIDE, IMAD.HI, LDS.32 / LDS.128,
// and mixes (LOP3+IMAD) — the denominators of the "integer-logic issue rate" roofline.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu && ./ubench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define ITERS 4096
#define ILP 8

enum { K_LOP3, K_SHF, K_IMAD, K_IMADW, K_IMADHI, K_MIX, K_LDS32, K_LDS128, K_POPC, K_MIXW, K_PREDOFF, K_PREDHALF, K_MULW, K_MULW_LOP, K_MULW_2LOP, K_COUNT };
static const char *names[] = {"LOP3", "SHF", "IMAD", "IMAD.WIDE", "IMAD.HI", "LOP3+IMAD 1:1", "LDS.32", "LDS.128", "POPC", "LOP3+IMAD.WIDE 1:1", "@!p LOP3 (all off)", "@p LOP3 (half off)", "MUL.WIDE (no addend)", "LOP3+MUL.WIDE 1:1", "2 LOP3 + MUL.WIDE"};

template <int K> __global__ void __launch_bounds__(256) bench(uint32_t *out, uint32_t seed, long long *clk) {
    __shared__ uint32_t sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * seed;
    __syncthreads();
    uint32_t x[ILP], y[ILP];
    uint64_t w[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = seed + threadIdx.x * 7 + i; y[i] = seed * 3 + i; w[i] = x[i]; }
    const uint32_t c1 = seed | 1, c2 = seed ^ 0x5555;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (K == K_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(c1));
            if (K == K_SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(c1));
            if (K == K_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(c1), "r"(y[i]));
            if (K == K_IMADW) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(y[i]), "r"(c1));
            if (K == K_IMADHI) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(c1), "r"(y[i]));
            if (K == K_MIX) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(c1));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(c1), "r"(c2));
            }
            if (K == K_MIXW) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(c1));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(y[i]), "r"(c1));
            }
            if (K == K_LDS32) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(&sm[(x[i] & 4095)]))); x[i] ^= v; }
            if (K == K_LDS128) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"((uint32_t)__cvta_generic_to_shared(&sm[(threadIdx.x * 4 + i * 1024) & 4095]))); x[i] ^= v.x ^ v.w; }
            if (K == K_POPC) asm volatile("popc.b32 %0, %0;" : "+r"(x[i]));
            if (K == K_MULW) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(x[i]), "r"(y[i])); x[i] = (uint32_t)w[i]; }
            if (K == K_MULW_LOP) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(x[i]), "r"(c1)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"((uint32_t)w[i]), "r"((uint32_t)(w[i] >> 32))); }
            if (K == K_MULW_2LOP) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(x[i]), "r"(c1)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"((uint32_t)w[i]), "r"(c2)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"((uint32_t)(w[i] >> 32)), "r"(c2)); }
            if (K == K_PREDOFF) asm volatile("{ .reg .pred p; setp.eq.u32 p, %3, 0x7fffffff; @p lop3.b32 %0, %0, %1, %2, 0x96; }" : "+r"(x[i]) : "r"(y[i]), "r"(c1), "r"(c2));
            if (K == K_PREDHALF) asm volatile("{ .reg .pred p; setp.eq.u32 p, %3, 0; @p lop3.b32 %0, %0, %1, %2, 0x96; }" : "+r"(x[i]) : "r"(y[i]), "r"(c1), "r"((uint32_t)(i & 1)));
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc ^= x[i] ^ y[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int K> void run(int sms, uint32_t *out, long long *clk) {
    const int blocks = sms * 4, threads = 256; // 32 warps per SM
    bench<K><<<blocks, threads>>>(out, 12345u, clk);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    bench<K><<<blocks, threads>>>(out, 12345u, clk);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    long long h[1024]; cudaMemcpy(h, clk, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    const double per = (K == K_MIX || K == K_MIXW || K == K_MULW_LOP) ? 2.0 : (K == K_MULW_2LOP ? 3.0 : 1.0);
    const double winstr_per_sm = 32.0 * ITERS * ILP * per; // 32 warps per SM
    printf("%-20s %7.3f warp-instr/clk/SM  (%.1f lane-ops/clk/SM)  kernel %.3f ms, %.0f clk -> %.0f MHz\n", names[K],
           winstr_per_sm / avg, 32.0 * winstr_per_sm / avg, ms, avg, avg / (ms * 1e3));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    uint32_t *out; long long *clk;
    cudaMalloc(&out, 4 * 1024 * 1024); cudaMalloc(&clk, 8 * 1024);
    run<K_LOP3>(p.multiProcessorCount, out, clk);
    run<K_SHF>(p.multiProcessorCount, out, clk);
    run<K_IMAD>(p.multiProcessorCount, out, clk);
    run<K_IMADW>(p.multiProcessorCount, out, clk);
    run<K_IMADHI>(p.multiProcessorCount, out, clk);
    run<K_MIX>(p.multiProcessorCount, out, clk);
    run<K_MIXW>(p.multiProcessorCount, out, clk);
    run<K_LDS32>(p.multiProcessorCount, out, clk);
    run<K_LDS128>(p.multiProcessorCount, out, clk);
    run<K_POPC>(p.multiProcessorCount, out, clk);
    run<K_MULW>(p.multiProcessorCount, out, clk);
    run<K_MULW_LOP>(p.multiProcessorCount, out, clk);
    run<K_MULW_2LOP>(p.multiProcessorCount, out, clk);
    run<K_PREDOFF>(p.multiProcessorCount, out, clk);
    run<K_PREDHALF>(p.multiProcessorCount, out, clk);
    return cudaDeviceSynchronize() != cudaSuccess;
}
