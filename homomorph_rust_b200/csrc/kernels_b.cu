// kernels_b.cu — the d = d' = 512 (D = 1024) fused mul+rem kernels, in their own translation unit (compiled in parallel
// with hmgpu.cu).  The first version, a fully unrolled 32-word Karatsuba (243 inlined leaf products, two minutes of ptxas,
// 275 M mul+rem/s), is gone; its successors are rolled.
#include "kernels.cuh"

namespace hmk {
// rolled-product kernel (mulrem_fresh32_kernel): one 512-thread CTA per SM
cudaError_t launch_mulrem_fresh_b32(const uint64_t *A, const uint64_t *B, uint64_t *O, uint64_t pairs, const uint32_t *Tg,
                                    int sm_count, cudaStream_t stream, int reduce_first) {
    constexpr int WS = 16;
    const size_t smem = mulrem32_smem_bytes<WS>();
    uint64_t blocks = (pairs + MR32_THREADS - 1) / MR32_THREADS;
    if (blocks > (uint64_t)sm_count) blocks = (uint64_t)sm_count;
    if (reduce_first == 2) { // conflict-free rotated fold table (128 KB) + one set of operand columns
        cudaError_t e = cudaFuncSetAttribute(mulrem_fresh32q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MR32Q_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        mulrem_fresh32q_kernel<<<(unsigned)blocks, MR32_THREADS, MR32Q_SMEM_BYTES, stream>>>(A, B, O, pairs, Tg);
    } else if (reduce_first) {
        auto kern = mulrem_fresh32r_kernel<WS>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned)blocks, MR32_THREADS, smem, stream>>>(A, B, O, pairs, Tg);
    } else {
        auto kern = mulrem_fresh32_kernel<WS>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned)blocks, MR32_THREADS, smem, stream>>>(A, B, O, pairs, Tg);
    }
    return cudaGetLastError();
}
} // namespace hmk
