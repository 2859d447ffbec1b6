// kernels_enc_umma.cu — config-B encryption (D = 1024, tau = 256) on the 5th-generation tensor cores.
//
// CipheredBit::cipher (reference src/cipher.rs:99-115) is C = XOR_{i in U} T_i + x: over a batch it is the GF(2) product
// masks[n x 256] * PK[256 x 1025] — the one contraction of the path with a shared right operand.  At config A the Four-Russians
// table kernel wins (tools/umma_encrypt_probe.cu: the MMA stream alone takes 4.9 clk per bit-ciphertext per SM, the table kernel
// 6.1 all included); at config B the table method is held to 4-bit windows by the 227 KB of shared memory (64 row reads of 128 B
// = 64 clk per bit-ciphertext, encrypt_tab4b_kernel runs at 77) while the MMA needs 8 x 4.9 = 39 — so here the contraction goes
// to tcgen05:
//   A[m][k] = bit k of the subset mask of ciphertext m, expanded to int8 in shared memory (K-major, no swizzle, 128 x 256),
//   B[n][k] = (bit n of T_k) * 2^(n mod 8) as int8 (-128 stands for 2^7), built once per key by the host; one launch handles
//             512 of the 1024 columns fit at a time (128 KB of shared memory): a CTA goes through its tiles twice (pass 0 / pass 1),
//   D = A * B^T in TMEM: two blocks of 256 int32 columns per tile, used as a ring (the epilogue of one overlaps the MMAs of the
//       other); the scaling puts the parity of column n at bit n mod 8 of its accumulator, so the epilogue after tcgen05.ld is
//       seven bit-select LOP3 per eight columns and three PRMT per word,
//   + x into bit 0 and the X^1024 coefficient (parity of mask AND topmask, as in encrypt_tab4b_kernel) by pass 0.
// Warp roles (672 threads, one CTA per SM): warps 0-15 epilogue (TMEM lane group = warp % 4, column quarter = warp / 4), warps
// 16-19 expand the masks (one row per thread; Philox4x32-10 drawn here when SEEDED), warp 20 lane 0 issues the MMAs.  mbarriers:
// a_full / a_empty per A buffer (2), d_full / d_empty per TMEM block (2); tcgen05.commit arrives on a_empty and d_full.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels_enc_umma.h"

namespace hmk {

namespace {

constexpr int TILE_M = 128, BLK_N = 256, K = 256;
constexpr uint32_t LBO = 128, SBO = (K / 16) * 128;                 // core matrix = 8 rows x 16 B, K-adjacent cores contiguous
constexpr uint32_t A_BYTES = TILE_M * K, B_BYTES = ENC_UMMA_PASS_COLS * K;
constexpr int A_STAGES = 2; // A tiles in flight (3 fit as well and change nothing: the producers are not what the tensor pipe waits for)
constexpr int EPI_WARPS = 16, PROD_WARPS = 4, THREADS = (EPI_WARPS + PROD_WARPS + 1) * 32;
// kind::i8: D = S32 (2 << 4), A and B signed 8 bit (1 << 7, 1 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BLK_N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(LBO >> 4) << 16) | ((uint64_t)(SBO >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_i8(uint32_t taddr, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(taddr), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, "
                 "%28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                   "=r"(v[31])
                 : "r"(taddr));
}
// 32 accumulators (column j carries its parity at bit j mod 8) -> 32 packed bits
__device__ __forceinline__ uint32_t pack32(const uint32_t (&v)[32]) {
    uint32_t byte[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        uint32_t acc = v[8 * b];
#pragma unroll
        for (int j = 1; j < 8; ++j) acc = (acc & ~(1u << j)) | (v[8 * b + j] & (1u << j)); // one LOP3 each
        byte[b] = acc;
    }
    uint32_t lo, hi, w;
    asm("prmt.b32 %0, %1, %2, 0x0040;" : "=r"(lo) : "r"(byte[0]), "r"(byte[1]));
    asm("prmt.b32 %0, %1, %2, 0x0040;" : "=r"(hi) : "r"(byte[2]), "r"(byte[3]));
    asm("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(w) : "r"(lo), "r"(hi));
    return w;
}
// Philox4x32-10 (Random123), the engine's documented mask stream (kernels.cuh)
__device__ __forceinline__ void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

template <bool SEEDED>
__global__ void __launch_bounds__(THREADS, 1) encrypt_umma_b_kernel(EncUmmaParams p, const uint4 *__restrict__ Bg) {
    extern __shared__ __align__(1024) uint8_t umma_smem[];
    uint8_t *sB = umma_smem, *sA = umma_smem + B_BYTES; // B half, then the A buffers
    __shared__ __align__(8) uint64_t a_full[A_STAGES], a_empty[A_STAGES], d_full[2], d_empty[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < A_STAGES; ++s) {
            mbar_init(&a_full[s], PROD_WARPS * 32);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&d_full[s], 1);
            mbar_init(&d_empty[s], EPI_WARPS * 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == EPI_WARPS + PROD_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t tiles = (p.units + TILE_M - 1) / TILE_M;
    const uint32_t my_tiles = blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0; // tile i of this CTA = blockIdx.x + i gridDim.x
    uint32_t it = 0; // tiles this CTA has been through (both passes counted): the same in every role, drives the barrier parities

    // Two passes over the CTA's tiles inside one launch: key bits 0..511, then 512..1023 with the other half of B loaded over the
    // first.  (`chunk_tiles` > 0 alternates the passes every that many tiles, so that a chunk's ciphertexts are still in L2 when
    // pass 1 writes their second halves and the partial 32-byte sectors at the seams are merged there instead of being read back
    // from DRAM; measured, it saves the DRAM reads and no time: 1.43 ms for 8.4 M bit-ciphertexts unchunked, 1.49 at 32 tiles.)
    const uint32_t chunk = p.chunk_tiles ? p.chunk_tiles : (my_tiles ? my_tiles : 1);
    for (uint32_t c0 = 0; c0 < my_tiles; c0 += chunk)
    for (uint32_t pass = 0; pass < 2; ++pass) {
    const uint32_t c1 = c0 + chunk < my_tiles ? c0 + chunk : my_tiles;
    {
        const uint4 *src = Bg + (size_t)pass * (B_BYTES / 16);
        for (uint32_t i = tid; i < B_BYTES / 16; i += THREADS) reinterpret_cast<uint4 *>(sB)[i] = __ldg(src + i);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // written through the generic proxy, read by the tensor core
        __syncthreads();
    }
    const uint32_t it0 = it;
    it += c1 - c0;

    if (warp < EPI_WARPS) {
        // ===== epilogue: row = TMEM lane 32 (warp % 4) + lane, columns [64 (warp / 4), +64) of the block = one u64 word =====
        const uint32_t q = warp & 3, cq = warp >> 2, row = 32 * q + lane;
        uint32_t it = it0;
        for (uint32_t ti = c0; ti < c1; ++ti, ++it) {
            const uint32_t tile = blockIdx.x + ti * gridDim.x;
            const uint32_t u = tile * TILE_M + row;
            uint32_t pbit = 0;
            if (pass == 0 && cq == 0 && u < p.units) pbit = ((uint32_t)__ldg(p.values + (u >> 3)) >> (u & 7)) & 1u;
#pragma unroll 1
            for (int blk = 0; blk < 2; ++blk) {
                mbar_wait(&d_full[blk], it & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // the 64 columns of this thread go to registers first, so that the block is handed back to the MMA issuer before
                // the packing starts
                uint32_t v0[32], v1[32];
                const uint32_t tbase = tmem + ((32u * q) << 16) + 256u * blk + 64u * cq;
                tmem_ld32(tbase, v0);
                tmem_ld32(tbase + 32, v1);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(&d_empty[blk]);
                uint32_t w0 = pack32(v0);
                const uint32_t w1 = pack32(v1);
                if (u < p.units) {
                    if (blk == 0) w0 ^= pbit; // + x (cipher.rs:112); pbit is 0 except for pass 0, column quarter 0
                    *reinterpret_cast<uint2 *>(p.out + (uint64_t)u * 17 + 8 * pass + 4 * blk + cq) = make_uint2(w0, w1);
                }
            }
        }
    } else if (warp < EPI_WARPS + PROD_WARPS) {
        // ===== producers: one row of the A tile per thread =====
        const uint32_t row = tid - EPI_WARPS * 32;
        uint32_t it = it0;
        for (uint32_t ti = c0; ti < c1; ++ti, ++it) {
            const uint32_t tile = blockIdx.x + ti * gridDim.x;
            const uint32_t s = it % A_STAGES, u = tile * TILE_M + row;
            uint32_t mw[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            if (u < p.units) {
                if constexpr (SEEDED) {
                    const uint64_t gu = p.first_unit + u;
                    uint32_t r0[4], r1[4];
                    philox((uint32_t)gu, (uint32_t)(gu >> 32), 0u, 0u, (uint32_t)p.seed, (uint32_t)(p.seed >> 32), r0);
                    philox((uint32_t)gu, (uint32_t)(gu >> 32), 1u, 0u, (uint32_t)p.seed, (uint32_t)(p.seed >> 32), r1);
#pragma unroll
                    for (int w = 0; w < 4; ++w) { mw[w] = r0[w]; mw[4 + w] = r1[w]; }
                } else {
                    const uint4 m0 = __ldg(reinterpret_cast<const uint4 *>(p.masks) + 2 * (uint64_t)u);
                    const uint4 m1 = __ldg(reinterpret_cast<const uint4 *>(p.masks) + 2 * (uint64_t)u + 1);
                    mw[0] = m0.x; mw[1] = m0.y; mw[2] = m0.z; mw[3] = m0.w;
                    mw[4] = m1.x; mw[5] = m1.y; mw[6] = m1.z; mw[7] = m1.w;
                }
                if (pass == 0) { // coefficient of X^1024 = parity(mask AND topmask)
                    uint32_t t = 0;
#pragma unroll
                    for (int w = 0; w < 8; ++w) t ^= mw[w] & p.topmask[w];
                    p.out[(uint64_t)u * 17 + 16] = (uint64_t)(__popc(t) & 1);
                }
            }
            mbar_wait(&a_empty[s], ((it / A_STAGES) & 1) ^ 1); // the MMAs that read this buffer A_STAGES tiles ago are complete
            uint8_t *arow = sA + s * A_BYTES + (row >> 3) * SBO + (row & 7) * 16;
#pragma unroll
            for (int kc = 0; kc < K / 16; ++kc) { // 16 mask bits -> one 16-byte row of a core matrix
                const uint32_t bits = (mw[kc >> 1] >> (16 * (kc & 1))) & 0xFFFFu;
                uint32_t w4[4];
#pragma unroll
                for (int n4 = 0; n4 < 4; ++n4) w4[n4] = (((bits >> (4 * n4)) & 0xFu) * 0x00204081u) & 0x01010101u; // nibble -> four 0/1 bytes
                *reinterpret_cast<uint4 *>(arow + kc * LBO) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy writes -> visible to the tensor core
            mbar_arrive(&a_full[s]);
        }
    } else if (lane == 0) {
        // ===== MMA issuer =====
        const uint64_t bdesc = make_desc(smem_u32(sB));
        uint32_t it = it0;
        for (uint32_t ti = c0; ti < c1; ++ti, ++it) {
            const uint32_t s = it % A_STAGES;
            mbar_wait(&a_full[s], (it / A_STAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t adesc = make_desc(smem_u32(sA + s * A_BYTES));
#pragma unroll 1
            for (int blk = 0; blk < 2; ++blk) {
                mbar_wait(&d_empty[blk], (it & 1) ^ 1); // the epilogue has drained this block's previous tile
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t bd = bdesc + (((uint32_t)blk * (BLK_N / 8) * SBO) >> 4);
#pragma unroll
                for (int ks = 0; ks < K / 32; ++ks) // K = 32 per instruction = two core matrices along K
                    mma_i8(tmem + 256u * blk, adesc + ((2 * ks * LBO) >> 4), bd + ((2 * ks * LBO) >> 4), ks > 0);
                mma_commit(&d_full[blk]);
            }
            mma_commit(&a_empty[s]);
        }
    }
    // end of the phase: the epilogue has seen d_full of the last tile, i.e. every MMA that reads this half of B is complete
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    } // passes, chunks
    if (warp == EPI_WARPS + PROD_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

} // namespace

size_t enc_umma_table_bytes() { return (size_t)2 * B_BYTES; }

// B for both passes: pass p covers key bits [512 p, 512 p + 512); byte (n, k) of a pass at the K-major no-swizzle offset
void enc_umma_build_table(const uint64_t *const *T, const size_t *T_words, int8_t *out) {
    for (size_t i = 0; i < enc_umma_table_bytes(); ++i) out[i] = 0;
    for (int pass = 0; pass < 2; ++pass)
        for (int n = 0; n < ENC_UMMA_PASS_COLS; ++n)
            for (int k = 0; k < K; ++k) {
                const int bitpos = ENC_UMMA_PASS_COLS * pass + n;
                const size_t w = (size_t)bitpos / 64;
                const int bit = w < T_words[k] ? (int)((T[k][w] >> (bitpos % 64)) & 1) : 0;
                if (bit) out[(size_t)pass * B_BYTES + (n / 8) * SBO + (k / 16) * LBO + (n % 8) * 16 + (k % 16)] = (int8_t)(uint8_t)(1u << (n % 8));
            }
}

cudaError_t launch_encrypt_umma_b(const EncUmmaParams &p0, bool seeded, const int8_t *d_table, int sm_count, cudaStream_t stream) {
    const size_t smem = (size_t)B_BYTES + A_STAGES * A_BYTES + 1024;
    auto kern = seeded ? encrypt_umma_b_kernel<true> : encrypt_umma_b_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const uint32_t tiles = (p0.units + TILE_M - 1) / TILE_M;
    const unsigned grid = tiles < (uint32_t)sm_count ? tiles : (unsigned)sm_count;
    kern<<<grid ? grid : 1, THREADS, smem, stream>>>(p0, reinterpret_cast<const uint4 *>(d_table));
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return cudaSuccess;
}

} // namespace hmk
