// umma_encrypt_probe.cu — the tensor-core formulation of encryption, measured (VERDICT r1 task 5b).
//
// Encryption is C = masks[n x 128] * PK[128 x 257] over GF(2) (reference src/cipher.rs:99-115): the one contraction with a
// shared right operand on the path.  This probe runs it as a tcgen05 kind::i8 GEMM for one 128-ciphertext tile per CTA:
//   A[m][k] = bit k of mask m (0/1, int8, K-major, no swizzle, expanded from the bit-packed mask inside the kernel),
//   B[n][k] = (bit n of T_k) * 2^(n mod 8) (int8; -128 stands for 2^7), built once per key,
//   D = A * B^T in TMEM (128 lanes x 256 int32 columns): D[m][n] = 2^(n mod 8) * #{k: mask bit and key bit set},
// so the parity of column n sits at bit n mod 8 of its accumulator; the epilogue merges eight columns into a byte with
// seven bit-select LOP3 and four bytes into a word with three PRMT (about one instruction per output bit after tcgen05.ld).
// The result (words 0..3 of every ciphertext; the X^256 coefficient is a parity of the mask as in encrypt_tab4_kernel) is
// checked against a CPU subset-XOR, and three things are timed on the device: the MMA stream alone, the epilogue alone,
// and the whole tile loop (expand masks -> MMA -> epilogue -> store), against encrypt_tab4_kernel's 6.1 clk per
// bit-ciphertext per SM (profiles/r02_encrypt_tab4_ncu.txt).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_encrypt_probe tools/umma_encrypt_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int TILE_M = 128, TILE_N = 256, TILE_K = 128;
constexpr uint32_t A_BYTES = TILE_M * TILE_K, B_BYTES = TILE_N * TILE_K;
constexpr uint32_t LBO = 128, SBO = (TILE_K / 16) * 128; // K-major, no swizzle: core matrix = 8 rows x 16 bytes, K-adjacent cores contiguous

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(LBO >> 4) << 16) | ((uint64_t)(SBO >> 4) << 32) | (1ull << 46);
}
// kind::i8: D = S32 (2 << 4), A and B signed 8 bit (1 << 7, 1 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TILE_N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);

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

// mode 0: whole tile loop, serial; 1: MMA stream only (masks expanded once, all tiles issued back to back); 2: epilogue only
// (accumulators computed once); 3: mask expansion only
__global__ void __launch_bounds__(128, 1) umma_encrypt_kernel(const uint4 *__restrict__ masks, const uint4 *__restrict__ Bg, uint32_t *__restrict__ out,
                                                              uint32_t tiles, int mode, unsigned long long *clk) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sA = smem, *sB = smem + A_BYTES;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid; i < B_BYTES / 16; i += 128) reinterpret_cast<uint4 *>(sB)[i] = __ldg(Bg + i);
    if (tid == 0) mbar_init(&bar, 1);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint64_t adesc = make_desc(smem_u32(sA)), bdesc = make_desc(smem_u32(sB));
    uint32_t parity = 0;
    const unsigned long long t0 = clock64();
    for (uint32_t tile = blockIdx.x, it = 0; tile < tiles; tile += gridDim.x, ++it) {
        if (mode == 0 || it == 0 || mode == 3) {
            // ---- expand the 128-bit mask of row m = tid into 128 bytes of the K-major, no-swizzle A tile ----
            const uint4 m4 = __ldg(masks + (size_t)tile * TILE_M + tid);
            const uint32_t mw[4] = {m4.x, m4.y, m4.z, m4.w};
            uint8_t *row = sA + (tid >> 3) * SBO + (tid & 7) * 16;
#pragma unroll
            for (int kc = 0; kc < 8; ++kc) { // 16 mask bits -> one 16-byte row of a core matrix
                const uint32_t bits = (mw[kc >> 1] >> (16 * (kc & 1))) & 0xFFFFu;
                uint32_t w[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) { // nibble -> four 0/1 bytes
                    const uint32_t nib = (bits >> (4 * q)) & 0xFu;
                    w[q] = (nib * 0x00204081u) & 0x01010101u;
                }
                *reinterpret_cast<uint4 *>(row + kc * LBO) = make_uint4(w[0], w[1], w[2], w[3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy writes -> visible to the tensor core
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        if (mode == 0 || it == 0 || mode == 1) {
            const bool last = tile + gridDim.x >= tiles;
            if (tid == 0) {
#pragma unroll
                for (int ks = 0; ks < TILE_K / 32; ++ks) // K = 32 per instruction = two core matrices along K
                    mma_i8(tmem, adesc + ((2 * ks * LBO) >> 4), bdesc + ((2 * ks * LBO) >> 4), ks > 0);
                if (mode != 1 || last) mma_commit(&bar); // mode 1: the whole stream is issued back to back, one commit at the end
            }
            if (mode != 1 || last) {
                mbar_wait(&bar, parity);
                parity ^= 1;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            } else {
                continue;
            }
        }
        if (mode == 3) {
            __syncthreads();
            continue;
        }
        if (mode != 1 || tile + gridDim.x >= tiles) {
            // ---- epilogue: lane (32 warp + lane) of TMEM = ciphertext row; 8 x 32 columns -> 8 packed words ----
            uint32_t words[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + 32 * c, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                words[c] = pack32(v);
            }
            uint4 *dst = reinterpret_cast<uint4 *>(out + ((size_t)tile * TILE_M + tid) * 8);
            dst[0] = make_uint4(words[0], words[1], words[2], words[3]);
            dst[1] = make_uint4(words[4], words[5], words[6], words[7]);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
        __syncthreads(); // the A tile and the accumulator are reused by the next tile
    }
    const unsigned long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0 && clk) *clk = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const uint32_t tiles_per_sm = 400, tiles = sms * tiles_per_sm; // 128 ciphertexts per tile
    const size_t n = (size_t)tiles * TILE_M;
    // key: 128 polynomials x 256 bits; masks: n x 128 bits
    std::vector<uint32_t> T(128 * 8), M(n * 4);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); };
    for (auto &x : T) x = rnd();
    for (auto &x : M) x = rnd();
    std::vector<int8_t> B(B_BYTES, 0);
    for (int nn = 0; nn < TILE_N; ++nn)
        for (int k = 0; k < TILE_K; ++k) {
            const int bit = (T[k * 8 + nn / 32] >> (nn % 32)) & 1;
            const int8_t val = (int8_t)(bit ? (uint8_t)(1u << (nn % 8)) : 0); // 2^7 is stored as -128: the parity bit is the same
            B[(nn / 8) * SBO + (k / 16) * LBO + (nn % 8) * 16 + (k % 16)] = val;
        }
    uint4 *d_masks, *d_B;
    uint32_t *d_out;
    unsigned long long *d_clk;
    CK(cudaMalloc(&d_masks, n * 16));
    CK(cudaMalloc(&d_B, B_BYTES));
    CK(cudaMalloc(&d_out, n * 32));
    CK(cudaMalloc(&d_clk, 8));
    CK(cudaMemcpy(d_masks, M.data(), n * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_B, B.data(), B_BYTES, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_out, 0xFF, n * 32));
    const size_t smem = A_BYTES + B_BYTES + 1024;
    CK(cudaFuncSetAttribute(umma_encrypt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const char *names[4] = {"whole tile loop, serial (expand masks -> 4 MMA -> epilogue -> store)", "MMA stream only (4 x M128 N256 K32 kind::i8 per tile)",
                            "epilogue only (8 x tcgen05.ld.32x32b.x32 + pack + store per tile)", "mask expansion only (128 bits -> 128 bytes of the A tile)"};
    for (int mode = 0; mode < 4; ++mode) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        umma_encrypt_kernel<<<sms, 128, smem>>>(d_masks, d_B, d_out, tiles, mode, d_clk);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        umma_encrypt_kernel<<<sms, 128, smem>>>(d_masks, d_B, d_out, tiles, mode, d_clk);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double clk_per_ct = ms * 1e-3 * 1.965e9 / ((double)tiles_per_sm * TILE_M);
        printf("mode %d  %-70s %8.3f ms  = %6.2f clk per bit-ciphertext per SM (at 1965 MHz), %6.1f G bit-ciphertexts/s on %d SMs\n", mode, names[mode],
               ms, clk_per_ct, n / (ms * 1e-3) / 1e9, sms);
        if (mode == 0) { // every ciphertext against the CPU subset XOR
            std::vector<uint32_t> got(n * 8);
            CK(cudaMemcpy(got.data(), d_out, n * 32, cudaMemcpyDeviceToHost));
            size_t bad = 0;
            for (size_t m = 0; m < n; ++m) {
                uint32_t want[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (int k = 0; k < 128; ++k)
                    if ((M[m * 4 + k / 32] >> (k % 32)) & 1)
                        for (int w = 0; w < 8; ++w) want[w] ^= T[k * 8 + w];
                if (memcmp(want, &got[m * 8], 32) != 0) {
                    if (bad < 3) printf("  mismatch at ciphertext %zu: got %08x %08x.. want %08x %08x..\n", m, got[m * 8], got[m * 8 + 1], want[0], want[1]);
                    ++bad;
                }
            }
            printf("  parity vs CPU subset-XOR: %zu of %zu ciphertexts differ%s\n", bad, n, bad ? "" : " (bit-exact)");
        }
    }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
