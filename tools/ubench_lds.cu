// LDS.128 bank behaviour probe (B200): 1024 threads per SM read random 256-byte-spaced lines of a 128 KB table; the byte offset
// inside the 128-byte line depends on the lane through one of several mappings.  Time per LDS.128 tells how many
// shared-memory wavefronts the hardware needed (4 = conflict-free for 32 lanes x 16 B).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_lds tools/ubench_lds.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MAP> __device__ __forceinline__ uint32_t lane_off(int lane) {
    if (MAP == 0) return ((lane >> 1) & 3) * 32 + (lane & 1) * 16;          // shipped: class by ciphertext inside the quarter-warp
    if (MAP == 1) return (lane & 3) * 32 + ((lane >> 2) & 1) * 16;
    if (MAP == 2) return ((lane >> 3) & 3) * 32 + (lane & 1) * 16;          // whole quarter in one class
    if (MAP == 3) return (lane & 3) * 32 + ((lane >> 4) & 1) * 16;
    if (MAP == 4) return 0;                                                  // one bank quad for everybody
    if (MAP == 5) return (lane & 7) * 16;                                    // every quarter covers the line, lane order
    if (MAP == 6) return ((lane >> 2) & 7) * 16;                             // every aligned group of 4 lanes shares a bank quad
    return ((lane & 3) * 2 + ((lane >> 4) & 1)) * 16;
}

template <int MAP, bool SAME_E> __global__ void __launch_bounds__(1024, 1) probe(uint32_t *out, int iters) {
    extern __shared__ __align__(16) uint8_t smem[];
    for (int i = threadIdx.x; i < 131072 / 16; i += 1024) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(i, i * 3, i * 5, i * 7);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem) + lane_off<MAP>(lane);
    uint32_t x = SAME_E ? (threadIdx.x >> 5) * 2654435761u : threadIdx.x * 2654435761u + blockIdx.x;
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            x = x * 1664525u + 1013904223u;
            const uint32_t addr = base + ((x >> 15) & 0x1ff00u); // 512 lines of 256 B
            uint32_t v0, v1, v2, v3;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(addr));
            a0 ^= v0; a1 ^= v1; a2 ^= v2; a3 ^= v3;
        }
    }
    out[blockIdx.x * 1024 + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3;
}

template <int MAP, bool SAME_E> void run(const char *name, uint32_t *d, int sms) {
    auto k = probe<MAP, SAME_E>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    const int iters = 4000;
    k<<<sms, 1024, 131072>>>(d, 100);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<sms, 1024, 131072>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double instr_per_sm = 32.0 * iters * 8;  // warp-level LDS.128 per SM
    printf("%-64s %8.3f ms  %6.2f clk per LDS.128 (at 1965 MHz)\n", name, ms, ms * 1e-3 * 1.965e9 / instr_per_sm);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    uint32_t *d; cudaMalloc(&d, p.multiProcessorCount * 1024 * 4);
    const int sms = p.multiProcessorCount;
    run<0, false>("map0 class=(lane>>1)&3 half=lane&1 [shipped], random lines", d, sms);
    run<1, false>("map1 class=lane&3 half=(lane>>2)&1, random lines", d, sms);
    run<2, false>("map2 class=(lane>>3)&3 half=lane&1, random lines", d, sms);
    run<3, false>("map3 class=lane&3 half=(lane>>4)&1, random lines", d, sms);
    run<4, false>("map4 one bank quad, random lines", d, sms);
    run<5, false>("map5 off=(lane&7)*16, random lines", d, sms);
    run<6, false>("map6 off=((lane>>2)&7)*16, random lines", d, sms);
    run<7, false>("map7 off=((lane&3)*2+((lane>>4)&1))*16, random lines", d, sms);
    run<5, true>("map5 off=(lane&7)*16, one line per warp (512 contiguous... x4 dup)", d, sms);
    run<4, true>("map4 broadcast (one address per warp)", d, sms);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
