// ubench2.cu — round-2 pipe probes: what the integer multiplier (FMA-heavy pipe) and the ALU pipe really sustain for the
// instruction FORMS the carry-less kernels execute (both multiplier operands in per-thread registers, full 32-bit values),
// and how the two pipes co-issue at the kernels' own mix (1 IMAD.WIDE : 2 LOP3).  Each probe runs >= 100 ms so that the SM
// clock has ramped; the clock it ran at is printed beside the rate (cycles from clock64, time from CUDA events).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench2 ubench2.cu && ./ubench2
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../homomorph_rust_b200/csrc/gf2_blocks.cuh"

#define ILP 8
// w = x * y with the low half written back over x (no register moves): x stays odd when x and y are odd
#define MULW_INPLACE(X, HI, Y) asm volatile("{ .reg .b64 t; mul.wide.u32 t, %0, %2; mov.b64 {%0, %1}, t; }" : "+r"(X), "=r"(HI) : "r"(Y))

enum Kind {
    MULW_RR,      // mul.wide.u32 reg x reg, full 32-bit values
    MADW_RRR,     // mad.wide.u32 reg x reg + reg64
    MULW_RR16,    // mul.wide.u32 reg x reg, both values < 2^16
    MULW_RR_A16,  // one value < 2^16
    MULW_RU,      // mul.wide.u32 reg x uniform (kernel-constant) 32-bit value
    MULW_RU16,    // reg x uniform 14-bit value (what round 1's probe did)
    MUL_LO,       // mul.lo.u32 reg x reg
    MUL_HI,       // mul.hi.u32 reg x reg
    LOP3_RRR,     // lop3 three registers
    LOP3_RRI,     // lop3 two registers + immediate
    SHF_RR,       // funnel shift
    MIX_1W_2L,    // 1 mul.wide + 2 lop3 (the Karatsuba kernels' executed mix)
    MIX_1W_1L,    // 1 mul.wide + 1 lop3
    MIX_1W_3L,    // 1 mul.wide + 3 lop3
    DFMA_RRR,     // fma.rn.f64
    DADD_RR,      // add.rn.f64
    FFMA_RRR,     // fma.rn.f32
    MIX_1W_1D,    // 1 mul.wide + 1 dfma (are the FP64 and integer-multiply pipes separate?)
    MIX_2L_1D,    // 2 lop3 + 1 dfma
    MIX_2W_1L,    // 2 mul.wide + 1 lop3
    MIX_1W_4L,    // 1 mul.wide + 4 lop3
    MIX_1W_2LI,   // 1 mul.wide + 2 lop3 with an immediate operand each (fewer register reads)
    MIX_1M_1L,    // 1 mul.lo + 1 lop3
    MIX_1M_2L,    // 1 mul.lo + 2 lop3
    MIX_1W_2L_2MOV, // 1 mul.wide + 2 lop3 + 2 register moves
    MIX_1W_1L_1S, // 1 mul.wide + 1 lop3 + 1 shf
    KCOUNT
};
static const char *names[KCOUNT] = {"IMAD.WIDE r*r (32-bit values)", "IMAD.WIDE r*r+r64", "IMAD.WIDE r*r (16-bit values)", "IMAD.WIDE r*r (one 16-bit value)",
                                     "IMAD.WIDE r*uniform32", "IMAD.WIDE r*uniform14", "IMAD lo r*r", "IMAD.HI r*r", "LOP3 r,r,r", "LOP3 r,r,imm", "SHF r,r",
                                     "1 IMAD.WIDE + 2 LOP3", "1 IMAD.WIDE + 1 LOP3", "1 IMAD.WIDE + 3 LOP3", "DFMA", "DADD", "FFMA", "1 IMAD.WIDE + 1 DFMA",
                                     "2 LOP3 + 1 DFMA", "2 IMAD.WIDE + 1 LOP3", "1 IMAD.WIDE + 4 LOP3", "1 IMAD.WIDE + 2 LOP3(imm)", "1 IMAD lo + 1 LOP3", "1 IMAD lo + 2 LOP3", "1 IMAD.WIDE + 2 LOP3 + 2 MOV", "1 IMAD.WIDE + 1 LOP3 + 1 SHF"};
static const int per_iter[KCOUNT] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 3, 2, 4, 1, 1, 1, 2, 3, 3, 5, 3, 2, 3, 5, 3};

template <int K> __global__ void __launch_bounds__(256) probe(uint32_t *out, uint32_t seed, long long *clk, int iters, uint32_t ubig, uint32_t usmall) {
    uint32_t x[ILP], y[ILP], z[ILP];
    uint64_t w[ILP];
    double d[ILP], e[ILP];
    float f[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        x[i] = (seed * 2654435761u + threadIdx.x * 40503u + i * 7919u) | 0x80000001u;
        y[i] = (seed * 40503u + (blockIdx.x * 256u + threadIdx.x) * 2654435761u + i * 104729u) | 0x40000001u;
        z[i] = x[i] ^ (y[i] >> 3);
        if (K == MULW_RR16) { x[i] = (x[i] & 0xffffu) | 1u; y[i] = (y[i] & 0xffffu) | 1u; }
        if (K == MULW_RR_A16) { y[i] = (y[i] & 0xffffu) | 1u; }
        w[i] = x[i];
        d[i] = 1.0 + 1e-9 * (double)(x[i] & 1023u);
        e[i] = 1.0 - 1e-9 * (double)(y[i] & 1023u);
        f[i] = 1.0f + 1e-6f * (float)(x[i] & 1023u);
    }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (K == MULW_RR || K == MULW_RR_A16) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"((uint32_t)w[i]), "r"(y[i])); // low half fed back (odd x odd stays odd)
            if (K == MULW_RR16) asm volatile("{ .reg .b32 t; and.b32 t, %1, 0xffff; mul.wide.u32 %0, t, %2; }" : "=l"(w[i]) : "r"((uint32_t)w[i]), "r"(y[i])); // + 1 LOP3 on the ALU pipe
            if (K == MADW_RRR) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x[i]), "r"(y[i]));
            if (K == MULW_RU) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(x[i]), "r"(ubig)); x[i] = (uint32_t)(w[i] >> 32) ^ (uint32_t)w[i]; }
            if (K == MULW_RU16) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(x[i]), "r"(usmall)); x[i] = (uint32_t)(w[i] >> 32) ^ (uint32_t)w[i]; }
            if (K == MUL_LO) asm volatile("mul.lo.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
            if (K == MUL_HI) { asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i])); }
            if (K == LOP3_RRR) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(z[i]));
            if (K == LOP3_RRI) asm volatile("lop3.b32 %0, %0, %1, 0x11111111, 0x78;" : "+r"(x[i]) : "r"(y[i]));
            if (K == SHF_RR) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(x[i]) : "r"(y[i]));
            if (K == MIX_1W_2L || K == MIX_1W_1L || K == MIX_1W_3L) {
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(x[i]), "r"(y[i]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[i]) : "r"((uint32_t)w[i]), "r"((uint32_t)(w[i] >> 32)));
                if (K != MIX_1W_1L) asm volatile("lop3.b32 %0, %0, %1, 0x11111111, 0x78;" : "+r"(x[i]) : "r"(z[i]));
                else x[i] = z[i] | 1u; // folded into the lop3 above by the compiler or one more LOP3: see the SASS
                if (K == MIX_1W_3L) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(z[i]), "r"(x[i]));
            }
            if (K == MIX_2W_1L) {
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(x[i]), "r"(y[i]));
                uint64_t w2;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w2) : "r"(z[i]), "r"(y[i]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"((uint32_t)w[i]), "r"((uint32_t)(w2 >> 32)));
                z[i] = (uint32_t)w2 | 1u;
            }
            if (K == MIX_1W_4L) {
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(x[i]), "r"(y[i]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[i]) : "r"((uint32_t)w[i]), "r"((uint32_t)(w[i] >> 32)));
                asm volatile("lop3.b32 %0, %0, %1, 0x11111111, 0x78;" : "+r"(x[i]) : "r"(z[i]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(z[i]), "r"(x[i]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[i]) : "r"(y[i]), "r"(x[i]));
            }
            if (K == MIX_1W_2LI) {
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(x[i]), "r"(y[i]));
                asm volatile("lop3.b32 %0, %0, %1, 0x22222222, 0x78;" : "+r"(z[i]) : "r"((uint32_t)w[i]));
                asm volatile("lop3.b32 %0, %0, %1, 0x11111111, 0x78;" : "+r"(x[i]) : "r"((uint32_t)(w[i] >> 32)));
            }
            if (K == MIX_1M_1L || K == MIX_1M_2L) {
                asm volatile("mul.lo.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[i]) : "r"(x[i]), "r"(y[i]));
                if (K == MIX_1M_2L) asm volatile("lop3.b32 %0, %0, %1, 0x11111111, 0x78;" : "+r"(y[i]) : "r"(z[i]));
            }
            if (K == MIX_1W_2L_2MOV) {
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(x[i]), "r"(y[i]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[i]) : "r"((uint32_t)w[i]), "r"((uint32_t)(w[i] >> 32)));
                asm volatile("lop3.b32 %0, %0, %1, 0x11111111, 0x78;" : "+r"(x[i]) : "r"(z[i]));
                uint32_t m1, m2;
                asm volatile("mov.b32 %0, %1;" : "=r"(m1) : "r"(x[i]));
                asm volatile("mov.b32 %0, %1;" : "=r"(m2) : "r"(z[i]));
                x[i] = m1; z[i] = m2;
            }
            if (K == MIX_1W_1L_1S) {
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(x[i]), "r"(y[i]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[i]) : "r"((uint32_t)w[i]), "r"((uint32_t)(w[i] >> 32)));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(x[i]) : "r"(z[i]));
            }
            if (K == DFMA_RRR) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(e[i]), "d"(e[(i + 1) % ILP]));
            if (K == DADD_RR) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(e[i]));
            if (K == FFMA_RRR) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(f[(i + 1) % ILP]), "f"(f[(i + 3) % ILP]));
            if (K == MIX_1W_1D) {
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"((uint32_t)w[i]), "r"(y[i]));
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(e[i]), "d"(e[(i + 1) % ILP]));
            }
            if (K == MIX_2L_1D) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(z[i]));
                asm volatile("lop3.b32 %0, %0, %1, 0x11111111, 0x78;" : "+r"(z[i]) : "r"(y[i]));
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(e[i]), "d"(e[(i + 1) % ILP]));
            }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc ^= x[i] ^ y[i] ^ z[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ (uint32_t)__double2loint(d[i]) ^ __float_as_uint(f[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}


// 8x8-word Karatsuba product (27 leaves on IMAD.WIDE) at different register budgets / occupancies: THREADS x MINB resident
// threads per SM, registers capped by __launch_bounds__.
template <int THREADS, int MINB> __global__ void __launch_bounds__(THREADS, MINB) kara8_probe(uint32_t *sink, int iters, uint32_t seed) {
    uint32_t a[8], b[8], r[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = seed * (threadIdx.x + 1) + i;
        b[i] = seed ^ (blockIdx.x * 977u + i);
    }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        hmk::clmul_kara<8>(a, b, r);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            a[i] ^= r[i];
            b[i] += r[8 + i];
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= a[i] ^ b[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int THREADS, int MINB> void run_kara(int sms, uint32_t *out, double target_ms) {
    int iters = 200;
    float ms = 0;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kara8_probe<THREADS, MINB>);
    for (int pass = 0; pass < 3; ++pass) {
        cudaEventRecord(a);
        kara8_probe<THREADS, MINB><<<sms * MINB, THREADS>>>(out, iters, 0x9e3779b9u + pass);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        if (pass == 0) iters = (int)(iters * (target_ms / (ms > 1e-3 ? ms : 1e-3))) + 1;
    }
    const double rate = (double)sms * MINB * THREADS * iters / (ms * 1e-3);
    printf("kara8  %3d threads x %d CTAs/SM (%2d warps/SM)  %3d regs  %4zu B local  %8.3f G products/s  %8.2f ms\n", THREADS, MINB, THREADS * MINB / 32, fa.numRegs,
           (size_t)fa.localSizeBytes, rate / 1e9, ms);
    fflush(stdout);
}

template <int K> void run(int sms, int ctas_per_sm, uint32_t *out, long long *clk, double target_ms) {
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe<K>, 256, 0);
    if (ctas_per_sm > occ) ctas_per_sm = occ; // every CTA resident, so a CTA's cycle count spans the kernel
    const int blocks = sms * ctas_per_sm, threads = 256;
    int iters = 1 << 14;
    float ms = 0;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int pass = 0; pass < 3; ++pass) { // calibrate, then the real run at >= target_ms
        cudaEventRecord(a);
        probe<K><<<blocks, threads>>>(out, 12345u + pass, clk, iters, 0x9E3779B9u, 12345u);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        if (pass == 0) iters = (int)(iters * (target_ms / (ms > 1e-3 ? ms : 1e-3))) + 1;
    }
    static long long h[8192];
    cudaMemcpy(h, clk, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += (double)h[i];
    avg /= blocks;
    const double warps_per_sm = ctas_per_sm * threads / 32.0;
    const double winstr_per_sm = warps_per_sm * (double)iters * ILP * per_iter[K];
    printf("%-36s %2.0f warps/SM  %7.3f warp-instr/clk/SM  %7.3f clk per warp-instr per SMSP   %8.2f ms  %6.0f MHz\n", names[K], warps_per_sm,
           winstr_per_sm / avg, 4.0 * avg / winstr_per_sm, ms, avg / (ms * 1e3));
    fflush(stdout);
}

int main(int argc, char **argv) {
    const double target_ms = argc > 1 ? atof(argv[1]) : 120.0;
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, target %.0f ms per probe\n", p.name, p.multiProcessorCount, target_ms);
    uint32_t *out;
    long long *clk;
    cudaMalloc(&out, 16 * 1024 * 1024);
    cudaMalloc(&clk, 8 * 8192);
    const int sms = p.multiProcessorCount;
#define RUN(K) run<K>(sms, 4, out, clk, target_ms)
    RUN(MULW_RR); RUN(MADW_RRR); RUN(MULW_RR16); RUN(MULW_RR_A16); RUN(MULW_RU); RUN(MULW_RU16); RUN(MUL_LO); RUN(MUL_HI);
    RUN(LOP3_RRR); RUN(LOP3_RRI); RUN(SHF_RR); RUN(MIX_1W_2L); RUN(MIX_1W_1L); RUN(MIX_1W_3L); RUN(DFMA_RRR); RUN(DADD_RR); RUN(FFMA_RRR);
    RUN(MIX_1W_1D); RUN(MIX_2L_1D); RUN(MIX_2W_1L); RUN(MIX_1W_4L); RUN(MIX_1W_2LI); RUN(MIX_1M_1L); RUN(MIX_1M_2L); RUN(MIX_1W_2L_2MOV); RUN(MIX_1W_1L_1S);
    run<MULW_RR>(sms, 2, out, clk, target_ms);
    run<MIX_1W_2L>(sms, 2, out, clk, target_ms);
    run<MIX_1W_2L>(sms, 8, out, clk, target_ms);
    run_kara<128, 2>(sms, out, target_ms);
    run_kara<128, 3>(sms, out, target_ms);
    run_kara<128, 4>(sms, out, target_ms);
    run_kara<128, 5>(sms, out, target_ms);
    run_kara<128, 6>(sms, out, target_ms);
    run_kara<128, 8>(sms, out, target_ms);
    run_kara<64, 7>(sms, out, target_ms);
    return cudaDeviceSynchronize() != cudaSuccess;
}
