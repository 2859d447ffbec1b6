// probes.cu — pipe-rate probes behind hm_measure_pipe_peaks: the hardware denominators of the integer rooflines.
//
// The carry-less kernels execute two instruction classes: IMAD.WIDE with both operands in per-thread registers (the
// 32x32 -> 64 products of clmul32_imad, FMA-heavy pipe) and LOP3 (ALU pipe).  Each probe issues one class (or the kernels'
// own 1 : 2 mix) from 8 independent dependency chains per thread with every launched CTA resident (grid = SMs x occupancy),
// runs for >= `min_ms` (best of three full-length launches), and reports warp-instructions per second from CUDA-event time.
// The SM clock during the probes is sampled by the caller with nvidia-smi (bench.py): the on-chip cycle counter is NOT used
// for it — measured on B200, clock64() advanced at 0.65-0.99 of the SM clock nvidia-smi and ncu report, depending on the
// instruction mix of the loop, so cycle-based "per clock" rates come out too high.  The counter value is still returned
// (cycle_counter_mhz) as a diagnostic.
#include <cuda_runtime.h>
#include <stdint.h>

#include "probes.h"

namespace hmk {
namespace {

constexpr int ILP = 8;
constexpr int UNROLL = 8; // probe instructions per loop iteration = ILP * UNROLL * per_iter: the loop counter's own ALU instructions
                          // (add, compare) stay below 4 % of the stream
constexpr int THREADS = 256;

enum { P_IMADW = 0, P_LOP3 = 1, P_MIX12 = 2 };

template <int K> __global__ void __launch_bounds__(THREADS) pipe_probe_kernel(uint32_t *sink, unsigned long long *clk, int iters, uint32_t seed) {
    uint32_t x[ILP], y[ILP], z[ILP];
    uint64_t w[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { // odd values: odd x odd stays odd, the chains never degenerate to 0
        x[i] = (seed * 2654435761u + threadIdx.x * 40503u + i * 7919u) | 0x80000001u;
        y[i] = (seed * 40503u + (blockIdx.x * THREADS + threadIdx.x) * 2654435761u + i * 104729u) | 0x40000001u;
        z[i] = x[i] ^ (y[i] >> 3);
        w[i] = (uint64_t)x[i] | ((uint64_t)y[i] << 32);
    }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i0 = 0; i0 < ILP * UNROLL; ++i0) {
            const int i = i0 % ILP;
            if (K == P_IMADW) // IMAD.WIDE.U32 Rd64, Ra, Rb, RZ — both halves of the product are the next operands, so that neither
                              // half is dead and the instruction cannot be narrowed to a 32-bit IMAD (operand values do not
                              // change the timing: tools/ubench2.cu, 16-bit vs 32-bit operands)
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"((uint32_t)w[i]), "r"((uint32_t)(w[i] >> 32)));
            if (K == P_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(z[i]));
            if (K == P_MIX12) { // one product, one 3-input XOR of its halves, one masked XOR feeding the next product
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(x[i]), "r"(y[i]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[i]) : "r"((uint32_t)w[i]), "r"((uint32_t)(w[i] >> 32)));
                asm volatile("lop3.b32 %0, %0, %1, 0x11111111, 0x78;" : "+r"(x[i]) : "r"(z[i]));
            }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc ^= x[i] ^ y[i] ^ z[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    sink[blockIdx.x * THREADS + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int K> cudaError_t run_probe(int sm_count, cudaStream_t stream, double min_ms, int per_iter, PipeProbe *out) {
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pipe_probe_kernel<K>, THREADS, 0);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    if (occ > 4) occ = 4; // 32 warps per SM are plenty; every CTA is resident, so one CTA's cycle count spans the kernel
    const int blocks = sm_count * occ;
    uint32_t *sink = nullptr;
    unsigned long long *clk = nullptr, *hclk = nullptr;
    e = cudaMalloc(&sink, (size_t)blocks * THREADS * 4);
    if (e == cudaSuccess) e = cudaMalloc(&clk, (size_t)blocks * 8);
    if (e == cudaSuccess) e = cudaMallocHost(&hclk, (size_t)blocks * 8);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    int iters = 1 << 11;
    float ms = 0;
    float best_ms = 0;
    for (int pass = 0; pass < 4 && e == cudaSuccess; ++pass) { // calibration, then three full-length launches: the fastest counts
        cudaEventRecord(e0, stream);
        pipe_probe_kernel<K><<<blocks, THREADS, 0, stream>>>(sink, clk, iters, 12345u + pass);
        cudaEventRecord(e1, stream);
        e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
        cudaEventElapsedTime(&ms, e0, e1);
        if (pass == 0) {
            const double scale = min_ms / (ms > 1e-3 ? ms : 1e-3);
            iters = (int)(iters * (scale > 1.0 ? scale : 1.0)) + 1;
        } else if (best_ms == 0 || ms < best_ms) {
            best_ms = ms;
        }
    }
    ms = best_ms;
    if (e == cudaSuccess) e = cudaMemcpy(hclk, clk, (size_t)blocks * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) {
        double cyc = 0;
        for (int i = 0; i < blocks; ++i) cyc += (double)hclk[i];
        cyc /= blocks;
        const double winstr = (double)blocks * (THREADS / 32) * (double)iters * ILP * UNROLL * per_iter;
        out->warp_instr_per_s = winstr / (ms * 1e-3);
        out->cycle_counter_mhz = cyc / (ms * 1e3);
        out->ms = ms;
        out->warps_per_sm = occ * THREADS / 32;
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (sink) cudaFree(sink);
    if (clk) cudaFree(clk);
    if (hclk) cudaFreeHost(hclk);
    return e;
}

} // namespace

cudaError_t measure_pipe_peaks(int sm_count, cudaStream_t stream, double min_ms, PipeProbe out[3]) {
    cudaError_t e = run_probe<P_IMADW>(sm_count, stream, min_ms, 1, &out[0]);
    if (e == cudaSuccess) e = run_probe<P_LOP3>(sm_count, stream, min_ms, 1, &out[1]);
    if (e == cudaSuccess) e = run_probe<P_MIX12>(sm_count, stream, min_ms, 3, &out[2]);
    return e;
}

} // namespace hmk
