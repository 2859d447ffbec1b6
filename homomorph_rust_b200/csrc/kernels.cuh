// kernels.cuh — hand-written sm_100a kernels of the batched GF(2)[X] ciphertext engine.
//
// Data model: a batch is n values x L bit-ciphertext slots; slot k of every
// value has the same fixed width in 64-bit words; value v occupies words
// [v*value_words, (v+1)*value_words) of one HBM allocation ("padded layout", value-major).
// Words are the reference's coefficient words: X^i is bit i%64 of word i/64
// (reference src/polynomial.rs:144,172).  Kernels work on the same memory as 32-bit words
// (little endian), since the SM integer datapath is 32 bits wide.
//
// Kernel -> reference map
//   encrypt_tab6 / encrypt_tab / encrypt_generic   CipheredBit::cipher     src/cipher.rs:99-115 (+ loop :180-185)
//   mask_fill_kernel (Philox4x32-10)               CipheredBit::part       src/cipher.rs:92-97 (seeded replacement)
//   decrypt_uniform / decrypt_value_tma / decrypt_slots   CipheredBit::decipher   src/cipher.rs:119-122 (+ packing :227-237)
//   xor_* / not_kernel                             Polynomial::add, gate_xor/not   src/polynomial.rs:190-243, common.rs:21-35
//   mul_small / mul_thread32 / mul_thread / mul_warp / mul_views   Polynomial::mul   src/polynomial.rs:252-310
//   rem_fold / rem_generic                         Polynomial::rem         src/polynomial.rs:316-365
//   mulrem_fresh_kernel                            mul followed by rem (the BASELINE "mul+rem" unit)
//   adder_thread_smem / adder_thread / adder_fused  add_internal           src/impls/numbers/common.rs:37-56
//   prefix_xor_kernel / xor_ops_kernel             the XOR halves of mul_unsigned_internal (common.rs:66-105) and of the
//                                                  generic adder; both circuits are host-planned batches of launches: hmgpu.cu
//
// Building blocks: clmul32_imad (32x32 carry-less product on the integer multiplier), clmul_kara<N> (Karatsuba over
// words), mul24_acc / mul32_acc (24x24- / 32x32-word products as six / nine 8x8-word Karatsubas), clmul_regs (shift/mask schoolbook, ALU only),
// fold_word (CRC-style remainder step), TMA 1-D bulk copy + mbarrier helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gf2_blocks.cuh"

namespace hmk {


// ----------------------------------------------------------------------------------------
// TMA (1-D bulk copy) + mbarrier helpers
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared in pieces of at most `piece` bytes, all completing on `bar` (several bulk requests in flight
// instead of one long one).  Caller has already armed `bar` with the total byte count.
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar);
__device__ __forceinline__ void tma_load_1d_pieces(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar,
                                                   uint32_t piece) {
    for (uint32_t o = 0; o < bytes; o += piece)
        tma_load_1d(static_cast<char *>(smem_dst) + o, static_cast<const char *>(gsrc) + o,
                    (bytes - o < piece) ? (bytes - o) : piece, bar);
}
// global -> shared, completion counted in bytes on `bar`.  16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk store (bulk async-group completion).
__device__ __forceinline__ void tma_store_1d(void *gdst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ----------------------------------------------------------------------------------------
// Subset masks on the device: the reference draws tau.div_ceil(8) bytes per bit from getrandom
// (src/cipher.rs:92-97); here they can instead come from Philox4x32-10 (Random123), a counter-based generator, so
// the same stream is reproducible on the host for parity checks and no mask crosses PCIe.
//   bytes [16 b, 16 b + 16) of the mask of bit-ciphertext u = Philox(counter = (u_lo, u_hi, b, 0), key = seed)
// ----------------------------------------------------------------------------------------
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
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

static __global__ void __launch_bounds__(256) mask_fill_kernel(uint8_t *__restrict__ masks, uint64_t units, uint32_t mask_bytes,
                                                               uint64_t seed, uint64_t first_unit) {
    const uint32_t blocks = (mask_bytes + 15) / 16;
    const uint64_t total = units * blocks;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t u = i / blocks;
        const uint32_t b = (uint32_t)(i % blocks);
        const uint64_t gu = first_unit + u; // position in the stream: shards of one logical batch continue each other
        uint32_t r[4];
        philox4x32_10((uint32_t)gu, (uint32_t)(gu >> 32), b, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
        uint8_t *dst = masks + u * mask_bytes + 16ull * b;
        const uint32_t nb = (mask_bytes - 16 * b < 16) ? (mask_bytes - 16 * b) : 16;
        if (nb == 16 && (mask_bytes % 16) == 0) {
            *reinterpret_cast<uint4 *>(dst) = make_uint4(r[0], r[1], r[2], r[3]);
        } else {
            for (uint32_t q = 0; q < nb; ++q) dst[q] = (uint8_t)(r[q >> 2] >> (8 * (q & 3)));
        }
    }
}

// ----------------------------------------------------------------------------------------
// Seeded key generation, reference src/context.rs:249-261: T_i = S * Q_i + X * R_i with Q_i, R_i drawn by
// Polynomial::random (src/polynomial.rs:73-96: random bytes -> little-endian words, bits above the degree cleared, the
// leading coefficient set).  The random bytes are a documented Philox4x32-10 stream (hm_key_stream_host computes the same):
//   bytes [16 j, 16 j + 16) of stream s = Philox(counter = (j_lo, j_hi, s, "KEYS"), key = seed),  s = 0: S, s = 1: Q_0 R_0 Q_1 R_1 ...
// keygen_fill_kernel lays out Q_i (wq words) and X * R_i (wr words) for every i; S * Q_i is one launch of the product kernels
// with S broadcast (a view with stride 0) and the sum one XOR launch.
// ----------------------------------------------------------------------------------------
constexpr uint32_t KEY_STREAM_TAG = 0x5359454Bu; // "KEYS"
__host__ __device__ inline uint64_t key_stream_word(uint64_t seed, uint32_t stream, uint64_t word_index) { // u64 word `word_index` of the byte stream
    uint32_t r[4];
    const uint64_t blk = word_index >> 1;
    philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), stream, KEY_STREAM_TAG, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    return (word_index & 1) ? ((uint64_t)r[2] | ((uint64_t)r[3] << 32)) : ((uint64_t)r[0] | ((uint64_t)r[1] << 32));
}
static __global__ void __launch_bounds__(128) keygen_fill_kernel(uint64_t *__restrict__ Q, uint64_t *__restrict__ RX, uint32_t tau, uint32_t dp, uint32_t delta,
                                                                 uint64_t seed) {
    const uint32_t wq = dp / 64 + 1, wr0 = delta / 64 + 1, wr = (delta + 1) / 64 + 1; // words of Q_i, R_i, X * R_i
    const uint32_t i = blockIdx.x;
    if (i >= tau) return;
    const uint64_t base = (uint64_t)i * (wq + wr0); // u64 word offset of (Q_i, R_i) in stream 1
    for (uint32_t j = threadIdx.x; j < wq; j += blockDim.x) {
        uint64_t w = key_stream_word(seed, 1, base + j);
        if (j == wq - 1) { // src/polynomial.rs:89-90
            w &= ((uint64_t)1 << (dp % 64)) - 1;
            w |= (uint64_t)1 << (dp % 64);
        }
        Q[(size_t)i * wq + j] = w;
    }
    for (uint32_t j = threadIdx.x; j < wr; j += blockDim.x) { // (X * R_i)[j] = R_i[j] << 1 | R_i[j-1] >> 63
        auto rword = [&](uint32_t t) -> uint64_t {
            if (t >= wr0) return 0;
            uint64_t w = key_stream_word(seed, 1, base + wq + t);
            if (t == wr0 - 1) {
                w &= ((uint64_t)1 << (delta % 64)) - 1;
                w |= (uint64_t)1 << (delta % 64);
            }
            return w;
        };
        const uint64_t cur = rword(j), prev = j ? rword(j - 1) : 0;
        RX[(size_t)i * wr + j] = (cur << 1) | (prev >> 63);
    }
}

// ----------------------------------------------------------------------------------------
// K2  subset-XOR encryption                         reference src/cipher.rs:99-115
//
//   C = XOR_{i : mask bit i} T_i  XOR  x
// The tau-bit subset mask is cut into windows of WB bits; for every window the XOR of every
// subset of its WB public-key polynomials is tabulated once per key (Method of Four
// Russians): table[g][e][0..WF).  A bit-ciphertext is then tau/WB row XORs.  The table
// (160 KB at d=d'=128, tau=128, WB=8) lives in shared memory of persistent CTAs.
// ----------------------------------------------------------------------------------------
struct EncParams {
    const uint8_t *values; // n * L/8 bytes
    const uint8_t *masks;  // units * mask_bytes
    uint64_t *out;         // units * wf words
    const uint64_t *table; // groups * 2^wb * wf words (global copy)
    uint64_t units;        // n * L bit-ciphertexts
    uint32_t wf;           // u64 words per fresh slot
    uint32_t mask_bytes;   // ceil(tau/8)
    uint32_t wb;           // window bits (1,2,4,8)
    uint32_t groups;       // ceil(tau/wb)
    uint32_t table_words;  // groups << wb * wf
    uint32_t table_in_smem;
};

constexpr int ENC_THREADS = 512;

// Fast path: WF words per slot, MW 32-bit mask words (tau == 32*MW), window WB.
template <int WF, int MW, int WB, int TH = ENC_THREADS>
__global__ void __launch_bounds__(TH, 1) encrypt_tab_kernel(EncParams p) {
    extern __shared__ __align__(16) uint64_t smem64[];
    uint64_t *tab = smem64;                              // table_words
    uint64_t *stage = smem64 + ((p.table_words + 1) & ~1u); // 2 x TH*WF, 16-byte aligned
    const int tid = threadIdx.x;
    // table: global -> shared, 128-bit coalesced
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.table);
        uint4 *dst = reinterpret_cast<uint4 *>(tab);
        const uint32_t n16 = p.table_words / 2;
        for (uint32_t i = tid; i < n16; i += TH) dst[i] = __ldg(src + i);
        if ((p.table_words & 1) && tid == 0) tab[p.table_words - 1] = p.table[p.table_words - 1];
    }
    __syncthreads();
    constexpr int GROUPS = MW * 32 / WB;
    const uint64_t ntiles = (p.units + TH - 1) / TH;
    uint32_t it = 0;
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        uint64_t *st = stage + (size_t)(it & 1) * TH * WF;
        // the bulk store issued two iterations ago from this buffer must have read it
        if (tid == 0) tma_store_wait_read<1>();
        __syncthreads();
        const uint64_t u = t * TH + tid;
        uint64_t acc[WF];
#pragma unroll
        for (int j = 0; j < WF; ++j) acc[j] = 0;
        if (u < p.units) {
            uint32_t mk[MW];
            const uint4 *mp = reinterpret_cast<const uint4 *>(p.masks + u * (uint64_t)(MW * 4));
#pragma unroll
            for (int q = 0; q < MW / 4; ++q) {
                uint4 m4 = __ldg(mp + q);
                mk[4 * q + 0] = m4.x;
                mk[4 * q + 1] = m4.y;
                mk[4 * q + 2] = m4.z;
                mk[4 * q + 3] = m4.w;
            }
#pragma unroll
            for (int g = 0; g < GROUPS; ++g) {
                const uint32_t e = (mk[(g * WB) / 32] >> ((g * WB) % 32)) & ((1u << WB) - 1u);
                const uint64_t *row = tab + ((size_t)((g << WB) + e)) * WF;
#pragma unroll
                for (int j = 0; j < WF; ++j) acc[j] ^= row[j];
            }
            const uint32_t bit = (__ldg(p.values + (u >> 3)) >> (u & 7)) & 1u; // src/cipher.rs:180-185
            acc[0] ^= bit;                                                      // add_bool_assign, polynomial.rs:238-243
        }
#pragma unroll
        for (int j = 0; j < WF; ++j) st[(size_t)tid * WF + j] = acc[j];
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            const uint64_t first = t * TH;
            const uint64_t cnt = (p.units - first < TH) ? (p.units - first) : TH;
            const uint32_t bytes = (uint32_t)(cnt * WF * 8);
            if ((bytes & 15u) == 0) {
                tma_store_1d(p.out + first * WF, st, bytes);
            } else { // ragged tail: plain stores by one thread (at most one tile per launch)
                for (uint32_t i = 0; i < cnt * WF; ++i) p.out[first * WF + i] = st[i];
            }
            tma_store_commit();
        }
    }
    if (tid == 0) tma_store_wait_all();
}

// Config-A fast path (tau = 128, 8-bit windows, 40-byte rows): 5 lanes per bit-ciphertext (one 64-bit word of the
// row each), 6 bit-ciphertexts per warp iteration, and a BANK-PARTITIONED table: window groups are split in three
// classes (g mod 3); the rows of class r live at byte offset 40 r of a 128-byte line, i.e. in shared-memory banks
// [10 r, 10 r + 10).  The three ciphertexts that share a half-warp always read three different classes (rotation
// by their position), so a warp-wide LDS.64 is conflict-free whatever the (random) row indices are.
//   line(t, e) = 128 bytes: [ row(group 3t, e) | row(group 3t+1, e) | row(group 3t+2, e) | 8 B pad ],  t = 0..5
// (groups 16, 17 do not exist: zero rows).  192 KB of shared memory, one persistent 1024-thread CTA per SM.
constexpr int ENC6_THREADS = 1024;
constexpr int ENC6_SLOTS = 6;
static __global__ void __launch_bounds__(ENC6_THREADS, 1) encrypt_tab6_kernel(EncParams p, const uint64_t *__restrict__ table6) {
    extern __shared__ __align__(16) uint64_t smem64[];
    uint64_t *tab = smem64; // ENC6_SLOTS * 256 * 16 words
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(table6);
        uint4 *dst = reinterpret_cast<uint4 *>(tab);
        for (uint32_t i = tid; i < ENC6_SLOTS * 256 * 8; i += ENC6_THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    // lanes 0-14: ciphertexts 0,1,2 of the group; lanes 16-30: ciphertexts 3,4,5; lanes 15 and 31 idle
    const uint32_t hl = lane & 15, r = hl / 5, j = hl - r * 5, c = r + 3 * (lane >> 4);
    const bool active = hl < 15;
    // 64-bit word offset inside a line for rotation step k: class (r + k) % 3 at words 5*class, plus my word j
    const uint32_t o0 = 5 * ((r + 0) % 3) + j, o1 = 5 * ((r + 1) % 3) + j, o2 = 5 * ((r + 2) % 3) + j;
    const uint32_t rot = 8 * r;
    const uint64_t ngroups = (p.units + 5) / 6;
    const uint64_t wstride = (uint64_t)gridDim.x * (ENC6_THREADS / 32);
    uint64_t g6 = (uint64_t)blockIdx.x * (ENC6_THREADS / 32) + warp;
    uint4 m_nxt = make_uint4(0, 0, 0, 0);
    uint32_t b_nxt = 0;
    if (g6 < ngroups) {
        const uint64_t u = g6 * 6 + c;
        if (active && u < p.units) {
            m_nxt = __ldg(reinterpret_cast<const uint4 *>(p.masks) + u);
            b_nxt = __ldg(p.values + (u >> 3));
        }
    }
    for (; g6 < ngroups; g6 += wstride) {
        const uint64_t u = g6 * 6 + c;
        const uint4 m4 = m_nxt;
        const uint32_t pb = b_nxt;
        {
            const uint64_t un = (g6 + wstride) * 6 + c;
            if (g6 + wstride < ngroups && active && un < p.units) {
                m_nxt = __ldg(reinterpret_cast<const uint4 *>(p.masks) + un);
                b_nxt = __ldg(p.values + (un >> 3));
            }
        }
        if (!active || u >= p.units) continue;
        const uint32_t mk[5] = {m4.x, m4.y, m4.z, m4.w, 0u};
        uint64_t acc = 0;
#pragma unroll
        for (int t = 0; t < ENC6_SLOTS; ++t) {
            // F = mask bytes 3t, 3t+1, 3t+2 (24 bits), rotated left by my position so that byte k is group 3t + (r+k)%3
            const int bit = 24 * t, w = bit >> 5, sh = bit & 31;
            uint32_t F = __funnelshift_r(mk[w], mk[w + 1 > 4 ? 4 : w + 1], sh) & 0xffffffu;
            F = ((F >> rot) | (F << (24 - rot))) & 0xffffffu; // rotate right by 8r: byte k <- byte (k + r) % 3
            const uint64_t *line = tab + (size_t)t * 256 * 16;
            acc ^= line[(F & 255u) * 16 + o0];
            acc ^= line[((F >> 8) & 255u) * 16 + o1];
            acc ^= line[(F >> 16) * 16 + o2];
        }
        if (j == 0) acc ^= (uint64_t)((pb >> (u & 7)) & 1u);
        p.out[u * 5 + j] = acc;
    }
}

// Same kernel with the index arithmetic done by hand: one funnel shift aligns the three mask bytes of a slot, PRMT with a
// per-lane selector picks byte (k + r) % 3 (no rotate, no mask), one IMAD turns it into a 32-bit shared-memory address
// (row * 128 + per-lane base; the slot's 32 KB offset is the LDS immediate), and the three rows are folded into the
// accumulator with two 3-input LOP3 per half.  ~14 instructions per slot instead of ~21.
template <int OFF> __device__ __forceinline__ void lds64_off(uint32_t addr, uint32_t &x, uint32_t &y) {
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2 + %3];" : "=r"(x), "=r"(y) : "r"(addr), "n"(OFF));
}
template <int T>
__device__ __forceinline__ void enc6_slot(const uint32_t (&mk)[5], const uint32_t (&base)[3], const uint32_t (&sel)[3], uint32_t &lo, uint32_t &hi) {
    constexpr int bit = 24 * T, w = bit >> 5, sh = bit & 31, w1 = (w + 1 > 4) ? 4 : w + 1;
    const uint32_t F = sh ? __funnelshift_r(mk[w], mk[w1], sh) : mk[w]; // mask bytes 3T, 3T+1, 3T+2 in bytes 0..2
    uint32_t xl[3], xh[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        uint32_t idx;
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(idx) : "r"(F), "r"(0u), "r"(sel[k]));
        lds64_off<T * 256 * 128>(idx * 128u + base[k], xl[k], xh[k]);
    }
    lo ^= xl[0] ^ xl[1] ^ xl[2];
    hi ^= xh[0] ^ xh[1] ^ xh[2];
}
static __global__ void __launch_bounds__(ENC6_THREADS, 1) encrypt_tab6b_kernel(EncParams p, const uint64_t *__restrict__ table6) {
    extern __shared__ __align__(16) uint64_t smem64[];
    uint64_t *tab = smem64;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(table6);
        uint4 *dst = reinterpret_cast<uint4 *>(tab);
        for (uint32_t i = tid; i < ENC6_SLOTS * 256 * 8; i += ENC6_THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const uint32_t hl = lane & 15, r = hl / 5, j = hl - r * 5, c = r + 3 * (lane >> 4);
    const bool active = hl < 15;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(tab);
    uint32_t base[3], sel[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const uint32_t cls = (r + k) % 3;    // byte `cls` of the slot's three mask bytes = window group 3t + cls
        base[k] = sbase + (5 * cls + j) * 8; // rows of class cls live at words [5 cls, 5 cls + 5) of every 128-byte line
        sel[k] = 0x4440u | cls;              // PRMT: that byte -> byte 0, zeros (second source) above
    }
    // 32-bit loop arithmetic (the host takes this kernel only for units < 2^31)
    const uint32_t units = (uint32_t)p.units;
    const uint32_t ngroups = (units + 5u) / 6u;
    const uint32_t wstride = gridDim.x * (ENC6_THREADS / 32);
    uint32_t g6 = blockIdx.x * (ENC6_THREADS / 32) + warp;
    uint4 m_nxt = make_uint4(0, 0, 0, 0);
    uint32_t b_nxt = 0;
    if (g6 < ngroups) {
        const uint32_t u = g6 * 6 + c;
        if (active && u < units) {
            m_nxt = __ldg(reinterpret_cast<const uint4 *>(p.masks) + u);
            b_nxt = __ldg(p.values + (u >> 3));
        }
    }
    for (; g6 < ngroups; g6 += wstride) {
        const uint32_t u = g6 * 6 + c;
        const uint4 m4 = m_nxt;
        const uint32_t pb = b_nxt;
        {
            const uint32_t un = u + wstride * 6;
            if (active && un < units) { // un < units implies g6 + wstride < ngroups
                m_nxt = __ldg(reinterpret_cast<const uint4 *>(p.masks) + un);
                b_nxt = __ldg(p.values + (un >> 3));
            }
        }
        if (!active || u >= units) continue;
        const uint32_t mk[5] = {m4.x, m4.y, m4.z, m4.w, 0u};
        uint32_t lo = 0, hi = 0;
        enc6_slot<0>(mk, base, sel, lo, hi);
        enc6_slot<1>(mk, base, sel, lo, hi);
        enc6_slot<2>(mk, base, sel, lo, hi);
        enc6_slot<3>(mk, base, sel, lo, hi);
        enc6_slot<4>(mk, base, sel, lo, hi);
        { // slot 5 holds one real window group (15); its row is stored in all three classes, so one conflict-free lookup
            uint32_t xl, xh;
            lds64_off<5 * 256 * 128>((mk[3] >> 24) * 128u + base[0], xl, xh);
            lo ^= xl;
            hi ^= xh;
        }
        if (j == 0) lo ^= (pb >> (u & 7)) & 1u;
        *reinterpret_cast<uint2 *>(p.out + (uint64_t)u * 5 + j) = make_uint2(lo, hi);
    }
}

// Round-2 successor (config A shape: tau = 128, D = 256): TWO lanes per bit-ciphertext (16 bytes of the row each, LDS.128),
// 16 bit-ciphertexts per warp pass, and rows of 32 bytes.  Word 4 of a fresh ciphertext holds only the coefficient of X^256,
// which is parity(mask AND topmask) with topmask_i = [X^256] T_i — so the table keeps words 0..3 only: 16 groups x 256 rows x
// 32 B = 128 KB with no padding, in FOUR bank classes:
//   line(t, e) = 128 bytes: [ row(group 4t, e) | row(4t+1, e) | row(4t+2, e) | row(4t+3, e) ],   t = 0..3 = mask word
// A quarter-warp (8 lanes = 4 ciphertexts x 2 halves, the unit an LDS.128 is served in) reads four different classes in every
// step (class = (position + step) % 4), i.e. all 32 banks once: conflict-free for random row indices.  4 shared-memory
// wavefronts per bit-ciphertext instead of 5.6, and 16 ciphertexts per 16 LDS instead of 6.
// The address of a lookup is ONE instruction: lines are 256 bytes apart (the lines of mask words t and t+1 interleave) and the
// table starts on a 64 KB boundary of the shared window, so PRMT drops the mask byte into bits 8..15 of a per-lane base
// address (class and half offsets in bits 0..7, table base above); t picks the LDS immediate.
// SEEDED: the masks come from Philox4x32-10 inside the kernel (same stream as mask_fill_kernel / hm_masks_generate_host):
// every lane draws the mask of one of the 32 ciphertexts of a double pass, the pass fetches its four words by shuffle — no
// mask buffer in HBM (16 B written + 16 B read per bit-ciphertext before).
constexpr int ENC4_THREADS = 1024;
constexpr int ENC4_TABLE_BYTES = 4 * 256 * 128;
struct Enc4Params {
    const uint8_t *values;
    const uint8_t *masks; // unused when SEEDED
    uint64_t *out;
    uint32_t units;
    uint32_t topmask[4];
    uint64_t seed, first_unit;
};
template <int IMM> __device__ __forceinline__ uint4 lds128_off(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4 + %5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr), "n"(IMM));
    return v;
}
template <int T> __device__ __forceinline__ void enc4_word(uint32_t m, const uint32_t (&bb)[4], const uint32_t (&sel)[4], uint32_t (&acc)[4]) {
    uint4 x[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        uint32_t addr;
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(addr) : "r"(m), "r"(bb[s]), "r"(sel[s]));
        x[s] = lds128_off<(T >> 1) * 65536 + (T & 1) * 128>(addr);
    }
    acc[0] ^= x[0].x ^ x[1].x; acc[0] ^= x[2].x ^ x[3].x;
    acc[1] ^= x[0].y ^ x[1].y; acc[1] ^= x[2].y ^ x[3].y;
    acc[2] ^= x[0].z ^ x[1].z; acc[2] ^= x[2].z ^ x[3].z;
    acc[3] ^= x[0].w ^ x[1].w; acc[3] ^= x[2].w ^ x[3].w;
}
template <bool SEEDED>
static __global__ void __launch_bounds__(ENC4_THREADS, 1) encrypt_tab4_kernel(Enc4Params p, const uint4 *__restrict__ table4) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t tbase = (s0 + 0xffffu) & ~0xffffu; // the host sizes the allocation for the worst-case alignment gap
    {
        uint32_t dyn;
        asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
        if (tbase - s0 + ENC4_TABLE_BYTES > dyn) __trap();
    }
    {
        uint4 *dst = reinterpret_cast<uint4 *>(smem_raw + (tbase - s0));
        for (uint32_t i = tid; i < ENC4_TABLE_BYTES / 16; i += ENC4_THREADS) dst[i] = __ldg(table4 + i);
    }
    __syncthreads();
    const uint32_t h = lane & 1, c = lane >> 1, r = c & 3;
    uint32_t bb[4], sel[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const uint32_t k = (r + s) & 3;      // class = byte k of the mask word = window group 4t + k
        bb[s] = tbase + k * 32 + h * 16;     // bits 8..15 are zero: the mask byte goes there
        sel[s] = 0x7604u | (k << 4);         // PRMT: byte 0, 2, 3 of the base, byte 1 <- byte k of the mask word
    }
    const uint32_t units = p.units;
    const uint32_t nsuper = (units + 31u) / 32u;
    const uint32_t gstride = gridDim.x * (ENC4_THREADS / 32);
    for (uint32_t G = blockIdx.x * (ENC4_THREADS / 32) + warp; G < nsuper; G += gstride) {
        uint32_t rnd[4] = {0u, 0u, 0u, 0u};
        uint4 mg[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
        if constexpr (SEEDED) {
            const uint64_t gu = p.first_unit + (uint64_t)G * 32u + lane;
            philox4x32_10((uint32_t)gu, (uint32_t)(gu >> 32), 0u, 0u, (uint32_t)p.seed, (uint32_t)(p.seed >> 32), rnd);
        } else {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const uint32_t u = G * 32u + 16u * i + c;
                if (u < units) mg[i] = __ldg(reinterpret_cast<const uint4 *>(p.masks) + u);
            }
        }
        uint32_t pb[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const uint32_t u = G * 32u + 16u * i + c;
            pb[i] = (u < units) ? (uint32_t)__ldg(p.values + (u >> 3)) : 0u;
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const uint32_t u = G * 32u + 16u * i + c;
            uint32_t mk[4];
            if constexpr (SEEDED) {
#pragma unroll
                for (int w = 0; w < 4; ++w) mk[w] = __shfl_sync(FULL, rnd[w], 16 * i + (int)c);
            } else {
                mk[0] = mg[i].x; mk[1] = mg[i].y; mk[2] = mg[i].z; mk[3] = mg[i].w;
            }
            if (u >= units) continue;
            uint32_t acc[4] = {0u, 0u, 0u, 0u};
            enc4_word<0>(mk[0], bb, sel, acc);
            enc4_word<1>(mk[1], bb, sel, acc);
            enc4_word<2>(mk[2], bb, sel, acc);
            enc4_word<3>(mk[3], bb, sel, acc);
            uint64_t *dst = p.out + (uint64_t)u * 5 + 2 * h;
            if (h == 0) {
                acc[0] ^= (pb[i] >> (u & 7)) & 1u; // + x (cipher.rs:112, polynomial.rs:238-243)
            } else {
                const uint32_t t = (mk[0] & p.topmask[0]) ^ (mk[1] & p.topmask[1]) ^ (mk[2] & p.topmask[2]) ^ (mk[3] & p.topmask[3]);
                dst[2] = (uint64_t)(__popc(t) & 1);   // coefficient of X^256
            }
            *reinterpret_cast<uint2 *>(dst) = make_uint2(acc[0], acc[1]);
            *reinterpret_cast<uint2 *>(dst + 1) = make_uint2(acc[2], acc[3]);
        }
    }
}

// The same design for the config-B shape (D = 1024, tau = 256): 128-byte rows (the coefficient of X^1024 is again computed as
// parity(mask AND topmask)), 4-bit windows: 64 groups x 16 rows x 128 B = 128 KB.  EIGHT lanes per bit-ciphertext (16 bytes of the
// row each): a quarter-warp reads one whole row per LDS.128 = all 32 banks once, conflict-free for any row; 4 ciphertexts per
// warp pass, 64 lookups each.  Lines are again 256 B apart (the rows of an even group and the next odd group share a line)
// and the table starts on a 64 KB boundary, so the address of a lookup is one PRMT on the mask word's even (or odd) nibbles.
// SEEDED: every lane draws one 16-byte Philox block for the 16 ciphertexts of four passes; a pass fetches its 8 words by SHFL.
// Replaces encrypt_tab_kernel<17,8,4,256> (thread per ciphertext, 136-byte rows at a 136-byte stride: bank conflicts on
// every LDS.128, and a 32 B/bit mask buffer written and re-read).
constexpr int ENC4B_THREADS = 1024;
constexpr int ENC4B_TABLE_BYTES = 64 * 16 * 128;
struct Enc4bParams {
    const uint8_t *values;
    const uint8_t *masks; // unused when SEEDED
    uint64_t *out;
    uint32_t units;
    uint32_t topmask[8];
    uint64_t seed, first_unit;
};
template <int GP> __device__ __forceinline__ void enc4b_pair(uint32_t ev, uint32_t od, uint32_t base, uint32_t sel, uint32_t (&acc)[4]) {
    constexpr int IMM = (GP >> 4) * 65536 + (GP & 15) * 4096;
    uint32_t ae, ao;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(ae) : "r"(ev), "r"(base), "r"(sel));
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(ao) : "r"(od), "r"(base), "r"(sel));
    const uint4 x = lds128_off<IMM>(ae), y = lds128_off<IMM + 128>(ao);
    acc[0] ^= x.x ^ y.x;
    acc[1] ^= x.y ^ y.y;
    acc[2] ^= x.z ^ y.z;
    acc[3] ^= x.w ^ y.w;
}
template <int W> __device__ __forceinline__ void enc4b_word(uint32_t m, uint32_t base, const uint32_t (&sel)[4], uint32_t (&acc)[4]) {
    const uint32_t ev = m & 0x0f0f0f0fu, od = (m >> 4) & 0x0f0f0f0fu; // even / odd window groups of this mask word, one per byte
    enc4b_pair<4 * W + 0>(ev, od, base, sel[0], acc);
    enc4b_pair<4 * W + 1>(ev, od, base, sel[1], acc);
    enc4b_pair<4 * W + 2>(ev, od, base, sel[2], acc);
    enc4b_pair<4 * W + 3>(ev, od, base, sel[3], acc);
}
template <bool SEEDED>
static __global__ void __launch_bounds__(ENC4B_THREADS, 1) encrypt_tab4b_kernel(Enc4bParams p, const uint4 *__restrict__ table4b) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t tbase = (s0 + 0xffffu) & ~0xffffu;
    {
        uint32_t dyn;
        asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
        if (tbase - s0 + ENC4B_TABLE_BYTES > dyn) __trap();
    }
    {
        uint4 *dst = reinterpret_cast<uint4 *>(smem_raw + (tbase - s0));
        for (uint32_t i = tid; i < ENC4B_TABLE_BYTES / 16; i += ENC4B_THREADS) dst[i] = __ldg(table4b + i);
    }
    __syncthreads();
    const uint32_t j = lane & 7, q = lane >> 3;
    const uint32_t base = tbase + 16 * j; // bits 8..15 are zero: the window value goes there
    uint32_t sel[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) sel[k] = 0x7604u | (k << 4);
    const uint32_t units = p.units;
    const uint32_t nsuper = (units + 15u) / 16u; // 16 ciphertexts = 4 passes per super-iteration
    const uint32_t gstride = gridDim.x * (ENC4B_THREADS / 32);
    for (uint32_t S = blockIdx.x * (ENC4B_THREADS / 32) + warp; S < nsuper; S += gstride) {
        uint32_t rnd[4] = {0u, 0u, 0u, 0u};
        if constexpr (SEEDED) { // lane l: block (l & 1) of ciphertext 16 S + (l >> 1)
            const uint64_t gu = p.first_unit + (uint64_t)S * 16u + (lane >> 1);
            philox4x32_10((uint32_t)gu, (uint32_t)(gu >> 32), lane & 1, 0u, (uint32_t)p.seed, (uint32_t)(p.seed >> 32), rnd);
        }
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
            const uint32_t u = S * 16u + 4u * i + q;
            uint32_t mk[8];
            if constexpr (SEEDED) {
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    mk[w] = __shfl_sync(FULL, rnd[w], 8 * i + 2 * (int)q);
                    mk[4 + w] = __shfl_sync(FULL, rnd[w], 8 * i + 2 * (int)q + 1);
                }
            } else {
                uint4 m0 = make_uint4(0, 0, 0, 0), m1 = m0;
                if (u < units) {
                    m0 = __ldg(reinterpret_cast<const uint4 *>(p.masks) + 2 * (uint64_t)u);
                    m1 = __ldg(reinterpret_cast<const uint4 *>(p.masks) + 2 * (uint64_t)u + 1);
                }
                mk[0] = m0.x; mk[1] = m0.y; mk[2] = m0.z; mk[3] = m0.w;
                mk[4] = m1.x; mk[5] = m1.y; mk[6] = m1.z; mk[7] = m1.w;
            }
            if (u >= units) continue;
            const uint32_t pb = (uint32_t)__ldg(p.values + (u >> 3));
            uint32_t acc[4] = {0u, 0u, 0u, 0u};
            enc4b_word<0>(mk[0], base, sel, acc);
            enc4b_word<1>(mk[1], base, sel, acc);
            enc4b_word<2>(mk[2], base, sel, acc);
            enc4b_word<3>(mk[3], base, sel, acc);
            enc4b_word<4>(mk[4], base, sel, acc);
            enc4b_word<5>(mk[5], base, sel, acc);
            enc4b_word<6>(mk[6], base, sel, acc);
            enc4b_word<7>(mk[7], base, sel, acc);
            uint64_t *dst = p.out + (uint64_t)u * 17 + 2 * j;
            if (j == 0) acc[0] ^= (pb >> (u & 7)) & 1u; // + x (cipher.rs:112)
            if (j == 7) {
                uint32_t t = 0;
#pragma unroll
                for (int w = 0; w < 8; ++w) t ^= mk[w] & p.topmask[w];
                dst[2] = (uint64_t)(__popc(t) & 1); // coefficient of X^1024
            }
            *reinterpret_cast<uint2 *>(dst) = make_uint2(acc[0], acc[1]);
            *reinterpret_cast<uint2 *>(dst + 1) = make_uint2(acc[2], acc[3]);
        }
    }
}

// Generic path: any tau / D / window, table read from shared memory if it fits, else from L2.
static __global__ void __launch_bounds__(256) encrypt_generic_kernel(EncParams p) {
    extern __shared__ __align__(16) uint64_t smem64[];
    const uint64_t *tab = p.table;
    if (p.table_in_smem) {
        for (uint32_t i = threadIdx.x; i < p.table_words; i += blockDim.x) smem64[i] = p.table[i];
        __syncthreads();
        tab = smem64;
    }
    const uint32_t emask = (1u << p.wb) - 1u;
    for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < p.units; u += (uint64_t)gridDim.x * blockDim.x) {
        const uint8_t *mk = p.masks + u * p.mask_bytes;
        const uint32_t bit = (p.values[u >> 3] >> (u & 7)) & 1u;
        for (uint32_t j = 0; j < p.wf; ++j) {
            uint64_t acc = (j == 0) ? bit : 0;
            for (uint32_t g = 0; g < p.groups; ++g) {
                const uint32_t bitpos = g * p.wb;
                const uint32_t e = (mk[bitpos >> 3] >> (bitpos & 7)) & emask;
                acc ^= tab[((size_t)((g << p.wb) + e)) * p.wf + j];
            }
            p.out[u * p.wf + j] = acc;
        }
    }
}

// ----------------------------------------------------------------------------------------
// K3  decryption = parity(popcount(C AND v)),  v_k = (X^k mod S)(0)
//     identical to rem + evaluate(false)            reference src/cipher.rs:119-122
// ----------------------------------------------------------------------------------------
constexpr int DEC_THREADS = 256;

// Uniform slot width w (fresh ciphertexts): tiles of 256 slots staged in shared memory by TMA,
// double buffered; thread per slot; result bits packed by ballot (bit k of the value is slot k,
// src/cipher.rs:227-237, so 32 consecutive slots are 4 consecutive output bytes).
static __global__ void __launch_bounds__(DEC_THREADS) decrypt_uniform_kernel(const uint64_t *__restrict__ ct,
                                                                      const uint64_t *__restrict__ v,
                                                                      uint8_t *__restrict__ out, uint64_t units,
                                                                      uint32_t w) {
    extern __shared__ __align__(16) uint64_t smem64[];
    __shared__ __align__(8) uint64_t bars[2];
    uint64_t *sv = smem64;                       // w words (padded to even)
    uint64_t *tiles = smem64 + ((w + 1) & ~1u);  // 2 x 256*w
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t tile_words = DEC_THREADS * w;
    for (uint32_t i = tid; i < w; i += DEC_THREADS) sv[i] = v[i];
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const uint64_t nfull = units / DEC_THREADS;
    uint32_t it = 0;
    if (tid == 0 && blockIdx.x < nfull) {
        mbar_expect_tx(&bars[0], tile_words * 8);
        tma_load_1d(tiles, ct + (uint64_t)blockIdx.x * tile_words, tile_words * 8, &bars[0]);
    }
    for (uint64_t t = blockIdx.x; t < nfull; t += gridDim.x, ++it) {
        const uint32_t buf = it & 1;
        const uint64_t tn = t + gridDim.x;
        if (tid == 0 && tn < nfull) { // prefetch next tile into the other buffer (freed by the barrier below)
            mbar_expect_tx(&bars[buf ^ 1], tile_words * 8);
            tma_load_1d(tiles + (size_t)(buf ^ 1) * tile_words, ct + tn * tile_words, tile_words * 8, &bars[buf ^ 1]);
        }
        mbar_wait(&bars[buf], (it >> 1) & 1);
        const uint64_t *c = tiles + (size_t)buf * tile_words + (size_t)tid * w;
        uint64_t acc = 0;
        for (uint32_t j = 0; j < w; ++j) acc ^= c[j] & sv[j];
        const uint32_t bit = __popcll(acc) & 1u;
        const uint32_t word = __ballot_sync(FULL, bit);
        if (lane == 0) reinterpret_cast<uint32_t *>(out)[(t * DEC_THREADS + tid) >> 5] = word;
        __syncthreads();
    }
    // ragged tail (< 256 slots): straight from global memory, handled by block 0
    if (blockIdx.x == 0 && nfull * DEC_THREADS < units) {
        const uint64_t u = nfull * DEC_THREADS + tid;
        uint32_t bit = 0;
        if (u < units) {
            const uint64_t *c = ct + u * w;
            uint64_t acc = 0;
            for (uint32_t j = 0; j < w; ++j) acc ^= c[j] & sv[j];
            bit = __popcll(acc) & 1u;
        }
        const uint32_t word = __ballot_sync(FULL, bit);
        if ((lane & 7) == 0 && u < units) out[u >> 3] = (uint8_t)(word >> lane);
    }
}

// Any slot layout: one warp per slot, lanes stride the slot's words (coalesced), shuffle parity.
// vv is v laid out like one value: vv[off[k] + j] = v[j].
static __global__ void __launch_bounds__(256) decrypt_slots_kernel(const uint64_t *__restrict__ ct,
                                                            const uint64_t *__restrict__ vv,
                                                            uint8_t *__restrict__ out, uint64_t units, Layout lay) {
    __shared__ uint32_t bits;
    if (threadIdx.x == 0) bits = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t u = (uint64_t)blockIdx.x * 8 + warp;
    if (u < units) {
        const uint64_t val = u / lay.L;
        const uint32_t k = (uint32_t)(u % lay.L);
        const uint32_t o = lay.off[k], w = lay.off[k + 1] - o;
        const uint64_t *c = ct + val * lay.value_words + o;
        const uint64_t *vk = vv + o;
        uint64_t acc = 0;
        uint32_t j = lane;
        for (; j + 96 < w; j += 128) {
            const uint64_t c0 = __ldg(c + j), c1 = __ldg(c + j + 32), c2 = __ldg(c + j + 64), c3 = __ldg(c + j + 96);
            acc ^= (c0 & __ldg(vk + j)) ^ (c1 & __ldg(vk + j + 32)) ^ (c2 & __ldg(vk + j + 64)) ^ (c3 & __ldg(vk + j + 96));
        }
        for (; j < w; j += 32) acc ^= __ldg(c + j) & __ldg(vk + j);
        uint32_t par = __popcll(acc) & 1u;
        par ^= __shfl_xor_sync(FULL, par, 16);
        par ^= __shfl_xor_sync(FULL, par, 8);
        par ^= __shfl_xor_sync(FULL, par, 4);
        par ^= __shfl_xor_sync(FULL, par, 2);
        par ^= __shfl_xor_sync(FULL, par, 1);
        if (lane == 0 && par) atomicOr(&bits, 1u << warp);
    }
    __syncthreads();
    if (threadIdx.x == 0 && (uint64_t)blockIdx.x * 8 < units) out[blockIdx.x] = (uint8_t)bits;
}

// Ragged layouts whose whole value fits in shared memory a few times over (results of the adder: 46.9 KB per
// u32): one persistent 512-thread CTA per SM streams whole values through a STAGES-deep ring of TMA bulk loads
// (the copy engine keeps (STAGES-1) x value_bytes in flight per SM, no thread waits on a global load), the
// decrypt vector v sits in shared memory, warp w reduces slots w, w+16, ... and the L plaintext bits are
// assembled in shared memory.
template <int STAGES, int DECV_THREADS>
__global__ void __launch_bounds__(DECV_THREADS) decrypt_value_tma_kernel(const uint64_t *__restrict__ ct,
                                                                           const uint64_t *__restrict__ v, uint32_t vwords,
                                                                           uint8_t *__restrict__ out, uint64_t n, Layout lay) {
    extern __shared__ __align__(16) uint64_t smem64[];
    __shared__ __align__(8) uint64_t bars[STAGES];
    __shared__ uint32_t sbits[2][MAX_SLOTS / 32];
    const uint32_t VW = lay.value_words;
    uint64_t *bufs = smem64;                        // STAGES x VW words (VW*8 is a multiple of 16)
    uint64_t *sv = smem64 + (size_t)STAGES * VW;    // decrypt vector, vwords = widest slot
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t i = tid; i < vwords; i += DECV_THREADS) sv[i] = v[i];
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const uint64_t first = blockIdx.x, stride = gridDim.x;
    const uint64_t mine = (n > first) ? (n - first + stride - 1) / stride : 0;
    const uint32_t vbytes = VW * 8;
    if (tid == 0)
        for (uint32_t j = 0; j + 1 < (uint32_t)STAGES && j < mine; ++j) {
            mbar_expect_tx(&bars[j], vbytes);
            tma_load_1d(bufs + (size_t)j * VW, ct + (first + j * stride) * VW, vbytes, &bars[j]);
        }
    const uint32_t nbytes = lay.L / 8;
    for (uint64_t i = 0; i < mine; ++i) {
        const uint32_t st = (uint32_t)(i % STAGES);
        const uint64_t ahead = i + STAGES - 1;
        if (tid == 0 && ahead < mine) { // refill the buffer consumed in the previous iteration
            const uint32_t sa = (uint32_t)(ahead % STAGES);
            mbar_expect_tx(&bars[sa], vbytes);
            tma_load_1d(bufs + (size_t)sa * VW, ct + (first + ahead * stride) * VW, vbytes, &bars[sa]);
        }
        uint32_t *bits = sbits[i & 1];
        if (tid < MAX_SLOTS / 32) bits[tid] = 0;
        __syncthreads();
        mbar_wait(&bars[st], (uint32_t)((i / STAGES) & 1));
        const uint64_t *c = bufs + (size_t)st * VW;
        for (uint32_t k = warp; k < lay.L; k += DECV_THREADS / 32) {
            const uint32_t o = lay.off[k], w = lay.off[k + 1] - o;
            uint64_t acc = 0;
            for (uint32_t j = lane; j < w; j += 32) acc ^= c[o + j] & sv[j];
            uint32_t par = __popcll(acc) & 1u;
            par ^= __shfl_xor_sync(FULL, par, 16);
            par ^= __shfl_xor_sync(FULL, par, 8);
            par ^= __shfl_xor_sync(FULL, par, 4);
            par ^= __shfl_xor_sync(FULL, par, 2);
            par ^= __shfl_xor_sync(FULL, par, 1);
            if (lane == 0 && par) atomicOr(&bits[k >> 5], 1u << (k & 31));
        }
        __syncthreads();
        if ((uint32_t)tid < nbytes) out[(first + i * stride) * nbytes + tid] = (uint8_t)(bits[tid >> 2] >> (8 * (tid & 3)));
    }
}

// ----------------------------------------------------------------------------------------
// K1  XOR / NOT                                     reference src/polynomial.rs:190-243
// ----------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) xor_flat_kernel(const uint4 *__restrict__ a, const uint4 *__restrict__ b,
                                                       uint4 *__restrict__ o, uint64_t n16) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 x = __ldg(a + i), y = __ldg(b + i);
        o[i] = make_uint4(x.x ^ y.x, x.y ^ y.y, x.z ^ y.z, x.w ^ y.w);
    }
}
static __global__ void __launch_bounds__(256) xor_flat_tail_kernel(const uint64_t *a, const uint64_t *b, uint64_t *o,
                                                            uint64_t from, uint64_t to) {
    const uint64_t i = from + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < to) o[i] = a[i] ^ b[i];
}
// Different layouts: out slot k = a slot k XOR b slot k, each zero-extended to the output width.
static __global__ void __launch_bounds__(256) xor_layout_kernel(const uint64_t *__restrict__ a, Layout la,
                                                         const uint64_t *__restrict__ b, Layout lb,
                                                         uint64_t *__restrict__ o, Layout lo, uint64_t n) {
    const uint64_t total = n * lo.value_words;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t v = i / lo.value_words;
        const uint32_t p = (uint32_t)(i % lo.value_words);
        uint32_t lo_k = 0, hi_k = lo.L; // slot k with off[k] <= p < off[k+1]
        while (hi_k - lo_k > 1) {
            const uint32_t mid = (lo_k + hi_k) >> 1;
            if (lo.off[mid] <= p) lo_k = mid; else hi_k = mid;
        }
        const uint32_t k = lo_k, j = p - lo.off[k];
        uint64_t x = 0;
        if (j < la.off[k + 1] - la.off[k]) x ^= a[v * la.value_words + la.off[k] + j];
        if (j < lb.off[k + 1] - lb.off[k]) x ^= b[v * lb.value_words + lb.off[k] + j];
        o[i] = x;
    }
}
// gate_not: a + 1, flips the constant term of every slot (common.rs:29-35).
static __global__ void __launch_bounds__(256) not_kernel(uint64_t *d, Layout lay, uint64_t n) {
    const uint64_t total = n * lay.L;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t v = i / lay.L;
        const uint32_t k = (uint32_t)(i % lay.L);
        d[v * lay.value_words + lay.off[k]] ^= 1ull;
    }
}
// o[v][0..o.w) = (zero-extended) a[v] XOR b[v];  b.base may be null (copy / widen).  In place allowed
// when o aliases a or b slot-for-slot.
static __global__ void __launch_bounds__(256) xor_views_kernel(View o, View a, View b, uint64_t n) {
    const uint64_t total = n * o.w;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t v = i / o.w;
        const uint32_t j = (uint32_t)(i % o.w);
        uint64_t x = 0;
        if (j < a.w) x ^= a.base[v * a.stride + a.off + j];
        if (b.base && j < b.w) x ^= b.base[v * b.stride + b.off + j];
        o.base[v * o.stride + o.off + j] = x;
    }
}

// Batched form of xor_views_kernel: ops[blockIdx.y].o = ops[..].a ^ ops[..].b (b may be a null view; o may alias a).
static __global__ void __launch_bounds__(256) xor_ops_kernel(const MulOp *__restrict__ ops, uint64_t n) {
    const MulOp op = ops[blockIdx.y];
    const View o = op.o, a = op.a, b = op.b;
    const uint64_t total = n * o.w;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t v = i / o.w;
        const uint32_t j = (uint32_t)(i % o.w);
        uint64_t x = 0;
        if (j < a.w) x ^= a.base[v * a.stride + a.off + j];
        if (b.base && j < b.w) x ^= b.base[v * b.stride + b.off + j];
        o.base[v * o.stride + o.off + j] = x;
    }
}

// Running XOR over a list of slots: ops[t].a is the t-th item x_t, ops[t].o (when its base is not null) receives the
// prefix x_0 ^ ... ^ x_t.  One thread per (value, word); `width` is the widest destination.  Used by the multiplier
// circuit (reference src/impls/numbers/common.rs:78-101): the carries of one column are x_t * (x_0 ^ ... ^ x_{t-1}),
// so with the prefixes materialised all of a column's products are independent and go out in one launch.
constexpr uint32_t PREFIX_MAX_ITEMS = 160; // L = 8: at most 36 items per column, L = 16: 136 (16 partial products + 120 carries)
static __global__ void __launch_bounds__(256) prefix_xor_kernel(const MulOp *__restrict__ ops, uint32_t cnt, uint32_t width, uint64_t n) {
    // descriptors once per CTA into shared memory (cnt <= PREFIX_MAX_ITEMS, checked by the host)
    __shared__ View s_a[PREFIX_MAX_ITEMS], s_o[PREFIX_MAX_ITEMS];
    for (uint32_t t = threadIdx.x; t < cnt; t += blockDim.x) {
        s_a[t] = ops[t].a;
        s_o[t] = ops[t].o;
    }
    __syncthreads();
    const uint64_t total = n * width;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t v = i / width;
        const uint32_t j = (uint32_t)(i % width);
        uint64_t acc = 0;
        for (uint32_t t0 = 0; t0 < cnt; t0 += 4) { // four items in flight: the loads of a group do not depend on each other
            uint64_t x[4];
#pragma unroll
            for (uint32_t u = 0; u < 4; ++u) {
                const uint32_t t = t0 + u;
                x[u] = (t < cnt && j < s_a[t].w) ? __ldg(s_a[t].base + v * s_a[t].stride + s_a[t].off + j) : 0;
            }
#pragma unroll
            for (uint32_t u = 0; u < 4; ++u) {
                const uint32_t t = t0 + u;
                acc ^= x[u];
                if (t < cnt && s_o[t].base && j < s_o[t].w) s_o[t].base[v * s_o[t].stride + s_o[t].off + j] = acc;
            }
        }
    }
}

// ----------------------------------------------------------------------------------------
// K4  generic carry-less multiply, one warp per product     reference src/polynomial.rs:252-310
//
// Operands staged in shared memory; output words sliced over the lanes; the shorter operand is
// scanned word by word and its set bits are warp-uniform, so zero bits cost nothing.
// ----------------------------------------------------------------------------------------
constexpr int MUL_WARPS = 4;

static __global__ void __launch_bounds__(MUL_WARPS * 32) mul_views_kernel(const MulOp *__restrict__ ops, uint64_t n,
                                                                   uint32_t smem_words_per_warp) {
    extern __shared__ __align__(16) uint32_t smem32[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t v = (uint64_t)blockIdx.x * MUL_WARPS + warp;
    if (v >= n) return;
    const MulOp op = ops[blockIdx.y];
    const uint32_t *ga = reinterpret_cast<const uint32_t *>(op.a.base + v * op.a.stride + op.a.off);
    const uint32_t *gb = reinterpret_cast<const uint32_t *>(op.b.base + v * op.b.stride + op.b.off);
    uint32_t na = 2 * op.a.w, nb = 2 * op.b.w;
    if (na > nb) { // scan the shorter operand
        const uint32_t *tp = ga; ga = gb; gb = tp;
        const uint32_t tn = na; na = nb; nb = tn;
    }
    uint32_t *sa = smem32 + (size_t)warp * smem_words_per_warp;
    uint32_t *sb = sa + na;
    for (uint32_t i = lane; i < na; i += 32) sa[i] = ga[i];
    for (uint32_t i = lane; i < nb; i += 32) sb[i] = gb[i];
    __syncwarp();
    uint32_t *go = reinterpret_cast<uint32_t *>(op.o.base + v * op.o.stride + op.o.off);
    const uint32_t no = 2 * op.o.w;
    for (uint32_t r0 = 0; r0 < no; r0 += 32) {
        const uint32_t r = r0 + lane;
        uint32_t acc = 0;
        // a[j] * b[r-j] (low half) and a[j] * b[r-j-1] (high half); j such that some lane has 0 <= r-j <= nb
        const uint32_t jlo = (r0 > nb) ? (r0 - nb) : 0;
        const uint32_t jhi = (r0 + 31 < na - 1) ? (r0 + 31) : (na - 1);
        for (uint32_t j = jlo; j <= jhi && j < na; ++j) {
            uint32_t aw = sa[j];
            if (aw == 0) continue; // warp-uniform
            const int bi = (int)r - (int)j;
            const uint32_t hi = (bi >= 0 && bi < (int)nb) ? sb[bi] : 0u;
            const uint32_t lo = (bi >= 1 && bi - 1 < (int)nb) ? sb[bi - 1] : 0u;
            while (aw) { // warp-uniform trip count
                const int s = __ffs(aw) - 1;
                aw &= aw - 1;
                acc ^= __funnelshift_l(lo, hi, s);
            }
        }
        if (r < no) go[r] = acc;
    }
}

// ----------------------------------------------------------------------------------------
// register-resident schoolbook product of NA x NB 32-bit words (per-thread operands)
// Horner over the bit index of a: r = r*X ^ sum_j (bit s of a[j]) * b * X^(32 j)
// ----------------------------------------------------------------------------------------
template <int NA, int NB>
__device__ __forceinline__ void clmul_regs(const uint32_t (&a)[NA], const uint32_t (&b)[NB], uint32_t (&r)[NA + NB]) {
#pragma unroll
    for (int i = 0; i < NA + NB; ++i) r[i] = 0;
    uint32_t at[NA];
#pragma unroll
    for (int j = 0; j < NA; ++j) at[j] = a[j];
#pragma unroll 1
    for (int s = 0; s < 32; ++s) {
#pragma unroll
        for (int i = NA + NB - 1; i > 0; --i) r[i] = __funnelshift_l(r[i - 1], r[i], 1);
        r[0] <<= 1;
#pragma unroll
        for (int j = 0; j < NA; ++j) {
            const uint32_t m = (uint32_t)((int32_t)at[j] >> 31); // current top bit of a[j], as a mask
            at[j] <<= 1;
#pragma unroll
            for (int k = 0; k < NB; ++k) r[j + k] ^= b[k] & m;
        }
    }
}


// ----------------------------------------------------------------------------------------
// K4a  small balanced products, one THREAD per product: operands of NX / NY low words plus the single coefficient
// X^(32 NX) / X^(32 NY) (fresh ciphertexts and their first products: degree bound an exact multiple of 256).
// Karatsuba on the multiplier in 8- or 16-word blocks.  Used for gate_and / gate_or and for the partial products and
// early carries of the multiplier circuit.
// ----------------------------------------------------------------------------------------
template <int NX, int NY>
__global__ void __launch_bounds__(128) mul_small_kernel(const MulOp *__restrict__ ops, uint64_t n, int fuse_or = 0) {
    constexpr int K = (NX == 8) ? 8 : 16; // Karatsuba block
    static_assert(NX % K == 0 && NY % K == 0, "operands are whole blocks");
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const MulOp op = ops[blockIdx.y];
    const uint64_t *gx = op.a.base + v * op.a.stride + op.a.off;
    const uint64_t *gy = op.b.base + v * op.b.stride + op.b.off;
    uint32_t x[NX], y[NY];
#pragma unroll
    for (int j = 0; j < NX / 2; ++j) {
        const uint64_t t = gx[j];
        x[2 * j] = (uint32_t)t;
        x[2 * j + 1] = (uint32_t)(t >> 32);
    }
#pragma unroll
    for (int j = 0; j < NY / 2; ++j) {
        const uint64_t t = gy[j];
        y[2 * j] = (uint32_t)t;
        y[2 * j + 1] = (uint32_t)(t >> 32);
    }
    const uint32_t xt = (uint32_t)gx[NX / 2] & 1u, yt = (uint32_t)gy[NY / 2] & 1u;
    uint32_t r[NX + NY + 2];
#pragma unroll
    for (int i = 0; i < NX + NY + 2; ++i) r[i] = 0;
#pragma unroll
    for (int bx = 0; bx < NX / K; ++bx) {
#pragma unroll
        for (int by = 0; by < NY / K; ++by) {
            uint32_t xa[K], yb[K], t[2 * K];
#pragma unroll
            for (int i = 0; i < K; ++i) {
                xa[i] = x[bx * K + i];
                yb[i] = y[by * K + i];
            }
            clmul_kara<K>(xa, yb, t);
#pragma unroll
            for (int i = 0; i < 2 * K; ++i) r[(bx + by) * K + i] ^= t[i];
        }
    }
    const uint32_t mx = 0u - xt, my = 0u - yt;
#pragma unroll
    for (int j = 0; j < NY; ++j) r[NX + j] ^= y[j] & mx;
#pragma unroll
    for (int j = 0; j < NX; ++j) r[NY + j] ^= x[j] & my;
    r[NX + NY] ^= xt & yt;
    if (fuse_or) { // gate_or = a + b + a*b (src/cipher.rs:76-83): the operands are still in registers
#pragma unroll
        for (int j = 0; j < NX; ++j) r[j] ^= x[j];
#pragma unroll
        for (int j = 0; j < NY; ++j) r[j] ^= y[j];
        r[NX] ^= xt;
        r[NY] ^= yt;
    }
    uint64_t *go = op.o.base + v * op.o.stride + op.o.off;
#pragma unroll
    for (int j = 0; j < (NX + NY) / 2 + 1; ++j)
        if ((uint32_t)j < op.o.w) go[j] = (uint64_t)r[2 * j] | ((uint64_t)r[2 * j + 1] << 32);
    for (uint32_t j = (NX + NY) / 2 + 1; j < op.o.w; ++j) go[j] = 0;
}

// ----------------------------------------------------------------------------------------
// K4b  general products, one WARP per product (the carry-chain machinery of the fused adder, for any two widths).
// The shorter operand a is the multiplier: it is cut into chunks of 24 words whose bits are transposed once
// (Bt[chunk][s] = which words of the chunk have bit s set) so that every lane sees them as warp-uniform masks; the
// longer operand c sits in shared memory; each lane owns TQ output words per pass and, per chunk, runs the Horner
// recurrence over s with one halo word (see adder_step_pass), skipping the zero bits of a.
// ----------------------------------------------------------------------------------------
constexpr int MW_J = 24;

template <int TQ>
__device__ __forceinline__ void mul_warp_pass(const uint32_t *__restrict__ sc, int nc, const uint32_t *__restrict__ sBt,
                                              int nchunks, int w0, int no, uint32_t *__restrict__ gout, int lane) {
    uint32_t tot[TQ];
#pragma unroll
    for (int i = 0; i < TQ; ++i) tot[i] = 0;
    const int t0 = w0 + lane * TQ;
    for (int ch = 0; ch < nchunks; ++ch) {
        const int j0 = ch * MW_J;
        // warp-uniform skip: the union of this pass's windows is [w0 - j0 - J, w0 + 32 TQ - j0)
        if (w0 + 32 * TQ - j0 <= 0 || w0 - j0 - MW_J >= nc) continue;
        const uint32_t Bmine = sBt[ch * 32 + lane];
        if (__ballot_sync(FULL, Bmine != 0u) == 0u) continue;
        uint32_t win[MW_J + TQ], acc[TQ + 1];
        const int base = t0 - j0 - MW_J;
#pragma unroll
        for (int y = 0; y < MW_J + TQ; ++y) {
            const int idx = base + y;
            win[y] = (idx >= 0 && idx < nc) ? sc[idx] : 0u;
        }
#pragma unroll
        for (int i = 0; i <= TQ; ++i) acc[i] = 0;
#pragma unroll 1
        for (int s = 31; s >= 0; --s) {
            const uint32_t Bs = __shfl_sync(FULL, Bmine, s);
#pragma unroll
            for (int i = TQ; i > 0; --i) acc[i] = __funnelshift_l(acc[i - 1], acc[i], 1);
            acc[0] <<= 1;
#pragma unroll
            for (int j = 0; j < MW_J; ++j) {
                if ((Bs >> j) & 1u) {
#pragma unroll
                    for (int i = 0; i <= TQ; ++i) acc[i] ^= win[i + MW_J - 1 - j];
                }
            }
        }
#pragma unroll
        for (int i = 0; i < TQ; ++i) tot[i] ^= acc[i + 1];
    }
#pragma unroll
    for (int i = 0; i < TQ; ++i)
        if (t0 + i < no) gout[t0 + i] = tot[i];
}

static __global__ void __launch_bounds__(128, 4) mul_warp_kernel(const MulOp *__restrict__ ops, uint64_t n, uint32_t smem_words_per_warp) {
    extern __shared__ __align__(16) uint32_t smem32[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t v = (uint64_t)blockIdx.x * 4 + warp;
    if (v >= n) return;
    const MulOp op = ops[blockIdx.y];
    const uint32_t *ga = reinterpret_cast<const uint32_t *>(op.a.base + v * op.a.stride + op.a.off);
    const uint32_t *gc = reinterpret_cast<const uint32_t *>(op.b.base + v * op.b.stride + op.b.off);
    int na = 2 * (int)op.a.w, nc = 2 * (int)op.b.w;
    if (na > nc) { // the shorter operand is the multiplier
        const uint32_t *tp = ga; ga = gc; gc = tp;
        const int tn = na; na = nc; nc = tn;
    }
    const int nchunks = (na + MW_J - 1) / MW_J;
    uint32_t *sBt = smem32 + (size_t)warp * smem_words_per_warp; // nchunks x 32
    uint32_t *sa = sBt + nchunks * 32;                           // nchunks x MW_J (zero padded)
    uint32_t *sc = sa + nchunks * MW_J;                          // nc
    for (int i = lane; i < nchunks * MW_J; i += 32) sa[i] = (i < na) ? ga[i] : 0u;
    for (int i = lane; i < nc; i += 32) sc[i] = gc[i];
    __syncwarp();
    for (int ch = 0; ch < nchunks; ++ch) { // transpose: lane s collects bit s of the chunk's words
        uint32_t m = 0;
#pragma unroll
        for (int j = 0; j < MW_J; ++j) m |= ((sa[ch * MW_J + j] >> lane) & 1u) << j;
        sBt[ch * 32 + lane] = m;
    }
    __syncwarp();
    uint32_t *gout = reinterpret_cast<uint32_t *>(op.o.base + v * op.o.stride + op.o.off);
    const int no = 2 * (int)op.o.w;
    // Passes of 32 lanes x TQ words.  All warps of a launch run the same shape, so many tile widths cost no
    // instruction-cache pressure here (unlike the fused adder); up to MW_TAIL trailing words (typically the word that
    // only holds the leading coefficient) are left to the scalar tail below instead of forcing a wider tile.
    constexpr int MW_TAIL = 3;
    int w0 = 0;
    while (no - w0 > MW_TAIL) {
        const int left = no - w0 - MW_TAIL;
        if (left > 32 * 20) {
            mul_warp_pass<24>(sc, nc, sBt, nchunks, w0, no, gout, lane);
            w0 += 32 * 24;
        } else if (left > 32 * 16) {
            mul_warp_pass<20>(sc, nc, sBt, nchunks, w0, no, gout, lane);
            w0 += 32 * 20;
        } else if (left > 32 * 12) {
            mul_warp_pass<16>(sc, nc, sBt, nchunks, w0, no, gout, lane);
            w0 += 32 * 16;
        } else if (left > 32 * 8) {
            mul_warp_pass<12>(sc, nc, sBt, nchunks, w0, no, gout, lane);
            w0 += 32 * 12;
        } else if (left > 32 * 6) {
            mul_warp_pass<8>(sc, nc, sBt, nchunks, w0, no, gout, lane);
            w0 += 32 * 8;
        } else if (left > 32 * 4) {
            mul_warp_pass<6>(sc, nc, sBt, nchunks, w0, no, gout, lane);
            w0 += 32 * 6;
        } else if (left > 32 * 2) {
            mul_warp_pass<4>(sc, nc, sBt, nchunks, w0, no, gout, lane);
            w0 += 32 * 4;
        } else {
            mul_warp_pass<2>(sc, nc, sBt, nchunks, w0, no, gout, lane);
            w0 += 32 * 2;
        }
    }
    // scalar tail: output word r = sum_j a[j] * c[r-j] (low half) + a[j] * c[r-j-1] (high half); for the top words
    // only the top few words of a contribute, so the j range is tiny
    if (w0 < no) {
        const int r = w0 + lane;
        if (r < no) {
            uint32_t acc = 0;
            int jlo = r - nc;
            if (jlo < 0) jlo = 0;
            for (int j = jlo; j < na && j <= r; ++j) {
                uint32_t aw = sa[j];
                const int bi = r - j;
                const uint32_t hi = (bi < nc) ? sc[bi] : 0u;
                const uint32_t lo = (bi >= 1 && bi - 1 < nc) ? sc[bi - 1] : 0u;
                while (aw) {
                    const int sft = __ffs(aw) - 1;
                    aw &= aw - 1;
                    acc ^= __funnelshift_l(lo, hi, sft);
                }
            }
            gout[r] = acc;
        }
    }
}

// ----------------------------------------------------------------------------------------
// K5  remainder by the secret key S                  reference src/polynomial.rs:316-365
//
// d % 32 == 0: word folding, T[b][x] = (x * X^(8b) * X^d) mod S (gf2host.hpp), i.e. a CRC with
// four 256-entry tables; WS = d/32 words of state in registers.
// ----------------------------------------------------------------------------------------
// REP > 1: every table row is stored REP times, copy `rep` of a row starts (rep * WS) words after copy 0, so that
// the lanes of a quarter-warp (rep = lane % 8) read from disjoint bank groups whatever their indices are.
// RS = row stride in words (>= WS): a stride that is an odd number of 16-byte groups spreads the rows of a
// quarter-warp over all eight bank groups (with RS = WS = 16 every row starts in group 0 or 4: 4-way conflicts).
template <int WS, int REP = 1, int RS = WS>
__device__ __forceinline__ void fold_word(uint32_t (&dst)[WS], uint32_t t, const uint32_t *__restrict__ T, uint32_t rep = 0) {
    const uint32_t *r0 = T + ((size_t)(0 * 256 + (t & 255u)) * REP + rep) * RS;
    const uint32_t *r1 = T + ((size_t)(1 * 256 + ((t >> 8) & 255u)) * REP + rep) * RS;
    const uint32_t *r2 = T + ((size_t)(2 * 256 + ((t >> 16) & 255u)) * REP + rep) * RS;
    const uint32_t *r3 = T + ((size_t)(3 * 256 + (t >> 24)) * REP + rep) * RS;
    if constexpr (WS % 4 == 0) {
#pragma unroll
        for (int q = 0; q < WS / 4; ++q) {
            const uint4 x0 = reinterpret_cast<const uint4 *>(r0)[q], x1 = reinterpret_cast<const uint4 *>(r1)[q];
            const uint4 x2 = reinterpret_cast<const uint4 *>(r2)[q], x3 = reinterpret_cast<const uint4 *>(r3)[q];
            dst[4 * q + 0] ^= x0.x ^ x1.x ^ x2.x ^ x3.x;
            dst[4 * q + 1] ^= x0.y ^ x1.y ^ x2.y ^ x3.y;
            dst[4 * q + 2] ^= x0.z ^ x1.z ^ x2.z ^ x3.z;
            dst[4 * q + 3] ^= x0.w ^ x1.w ^ x2.w ^ x3.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < WS; ++q) dst[q] ^= r0[q] ^ r1[q] ^ r2[q] ^ r3[q];
    }
}

template <int WS> __host__ __device__ constexpr int rem_fold_row_stride() { return WS >= 8 ? WS + 4 : WS; } // see fold_word
template <int WS>
__global__ void __launch_bounds__(128) rem_fold_kernel(View a, View o, uint64_t n, const uint32_t *__restrict__ Tg) {
    extern __shared__ __align__(16) uint32_t smem32[];
    constexpr int RS = rem_fold_row_stride<WS>();
    for (uint32_t i = threadIdx.x; i < 4u * 256u * WS; i += blockDim.x) smem32[(i / WS) * RS + (i % WS)] = Tg[i];
    __syncthreads();
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const uint32_t *p = reinterpret_cast<const uint32_t *>(a.base + v * a.stride + a.off);
    const int na = 2 * (int)a.w;
    uint32_t r[WS];
#pragma unroll
    for (int q = 0; q < WS; ++q) r[q] = 0;
    for (int i = na - 1; i >= 0; --i) { // r = (r * X^32 + word_i) mod S
        const uint32_t t = r[WS - 1];
#pragma unroll
        for (int q = WS - 1; q > 0; --q) r[q] = r[q - 1];
        r[0] = p[i];
        if (t) fold_word<WS, 1, RS>(r, t, smem32);
    }
    uint32_t *out = reinterpret_cast<uint32_t *>(o.base + v * o.stride + o.off);
#pragma unroll
    for (int q = 0; q < WS; ++q) out[q] = r[q];
    for (uint32_t q = WS; q < 2 * o.w; ++q) out[q] = 0;
}

// Any d >= 1: warp-cooperative long division in shared memory (slow; small/odd parameter sets only).
static __global__ void __launch_bounds__(MUL_WARPS * 32) rem_generic_kernel(View a, View o, uint64_t n,
                                                                     const uint32_t *__restrict__ Sg, uint32_t d,
                                                                     uint32_t smem_words_per_warp) {
    extern __shared__ __align__(16) uint32_t smem32[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t v = (uint64_t)blockIdx.x * MUL_WARPS + warp;
    if (v >= n) return;
    uint32_t *sp = smem32 + (size_t)warp * smem_words_per_warp;
    const uint32_t na = 2 * a.w, ns = d / 32 + 1;
    const uint32_t *ga = reinterpret_cast<const uint32_t *>(a.base + v * a.stride + a.off);
    for (uint32_t i = lane; i < na; i += 32) sp[i] = ga[i];
    if (lane == 0) sp[na] = 0; // spill word for the shifted divisor
    __syncwarp();
    for (int pos = (int)na * 32 - 1; pos >= (int)d; --pos) {
        const uint32_t wv = sp[pos >> 5];
        if ((wv >> (pos & 31)) & 1u) { // warp-uniform
            const uint32_t sh = (uint32_t)pos - d, ws = sh >> 5, bs = sh & 31;
            for (uint32_t i = lane; i <= ns; i += 32) { // S << sh covers words ws .. ws+ns
                const uint32_t hi = (i < ns) ? Sg[i] : 0u, lo = (i >= 1) ? Sg[i - 1] : 0u;
                const uint32_t x = __funnelshift_l(lo, hi, bs);
                if (x) sp[ws + i] ^= x;
            }
        }
        __syncwarp();
    }
    uint32_t *out = reinterpret_cast<uint32_t *>(o.base + v * o.stride + o.off);
    for (uint32_t q = lane; q < 2 * o.w; q += 32) out[q] = (q < na) ? sp[q] : 0u;
}

// ----------------------------------------------------------------------------------------
// Fused (a*b) mod S for fresh ciphertext pairs, thread per pair       BASELINE "mul+rem"
//   D = 32*WD (D % 64 == 0): operands are WD words + the X^D bit; d = 32*WS.
//   Operand tiles are staged in shared memory by TMA (two buffers); the 2D+1-bit product stays
//   in registers and is folded down to d bits with the tables above; only d bits are written.
// ----------------------------------------------------------------------------------------
template <int WD, int WS, int MODE, int MR_THREADS, int REP, int MINB = (WD >= 32 ? 2 : (MR_THREADS >= 512 ? 1 : 4))>
__global__ void __launch_bounds__(MR_THREADS, MINB) mulrem_fresh_kernel(const uint64_t *__restrict__ A,
                                                                  const uint64_t *__restrict__ B,
                                                                  uint64_t *__restrict__ O, uint64_t n,
                                                                  const uint32_t *__restrict__ Tg) {
    constexpr int WF = WD / 2 + 1;   // u64 words per operand slot
    constexpr int NP = 2 * WD + 1;   // product words (bit 2D lives in the last one)
    extern __shared__ __align__(16) uint32_t smem32[];
    __shared__ __align__(8) uint64_t bars[2];
    uint32_t *T = smem32;                                                     // 4*256*REP*WS
    uint64_t *tiles = reinterpret_cast<uint64_t *>(smem32 + 4 * 256 * REP * WS); // [2 bufs][2 operands][MR_THREADS*WF]
    const int tid = threadIdx.x;
    const uint32_t rep = (REP > 1) ? (uint32_t)(tid % REP) : 0u;
    for (uint32_t i = tid; i < 4u * 256u * REP * WS; i += MR_THREADS) T[i] = Tg[((i / WS) / REP) * WS + (i % WS)];
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    constexpr uint32_t TILE_WORDS = MR_THREADS * WF;
    const uint64_t nfull = n / MR_THREADS;
    const uint64_t ntiles = (n + MR_THREADS - 1) / MR_THREADS;
    if (tid == 0 && blockIdx.x < nfull) {
        mbar_expect_tx(&bars[0], 2 * TILE_WORDS * 8);
        tma_load_1d(tiles, A + (uint64_t)blockIdx.x * TILE_WORDS, TILE_WORDS * 8, &bars[0]);
        tma_load_1d(tiles + TILE_WORDS, B + (uint64_t)blockIdx.x * TILE_WORDS, TILE_WORDS * 8, &bars[0]);
    }
    uint32_t it = 0;
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const uint32_t buf = it & 1;
        const uint64_t tn = t + gridDim.x;
        if (tid == 0 && tn < nfull) {
            uint64_t *dst = tiles + (size_t)(buf ^ 1) * 2 * TILE_WORDS;
            mbar_expect_tx(&bars[buf ^ 1], 2 * TILE_WORDS * 8);
            tma_load_1d(dst, A + tn * TILE_WORDS, TILE_WORDS * 8, &bars[buf ^ 1]);
            tma_load_1d(dst + TILE_WORDS, B + tn * TILE_WORDS, TILE_WORDS * 8, &bars[buf ^ 1]);
        }
        uint32_t a[WD], b[WD], atop, btop;
        const uint64_t u = t * MR_THREADS + tid;
        if (t < nfull) {
            mbar_wait(&bars[buf], (it >> 1) & 1);
            const uint64_t *sa = tiles + (size_t)buf * 2 * TILE_WORDS + (size_t)tid * WF;
            const uint64_t *sb = sa + TILE_WORDS;
#pragma unroll
            for (int j = 0; j < WD / 2; ++j) {
                const uint64_t x = sa[j], y = sb[j];
                a[2 * j] = (uint32_t)x; a[2 * j + 1] = (uint32_t)(x >> 32);
                b[2 * j] = (uint32_t)y; b[2 * j + 1] = (uint32_t)(y >> 32);
            }
            atop = (uint32_t)sa[WD / 2] & 1u;
            btop = (uint32_t)sb[WD / 2] & 1u;
        } else { // ragged last tile: straight from global
#pragma unroll
            for (int j = 0; j < WD; ++j) a[j] = b[j] = 0;
            atop = btop = 0;
            if (u < n) {
                const uint64_t *ga = A + u * WF, *gb = B + u * WF;
#pragma unroll
                for (int j = 0; j < WD / 2; ++j) {
                    const uint64_t x = ga[j], y = gb[j];
                    a[2 * j] = (uint32_t)x; a[2 * j + 1] = (uint32_t)(x >> 32);
                    b[2 * j] = (uint32_t)y; b[2 * j + 1] = (uint32_t)(y >> 32);
                }
                atop = (uint32_t)ga[WD / 2] & 1u;
                btop = (uint32_t)gb[WD / 2] & 1u;
            }
        }
        if constexpr (MODE == 3) {
            // (a b) mod S = ((a mod S)(b mod S)) mod S: the remainder is unique, so reducing the operands first gives the
            // reference's rem(mul(a, b)) bit for bit with a WS-word product (9 leaf products at WS = 4) instead of a
            // WD-word one (27).  Fold order: word i into words [i - WS, i), top down.
            uint32_t xa[WD + 1], xb[WD + 1];
#pragma unroll
            for (int j = 0; j < WD; ++j) {
                xa[j] = a[j];
                xb[j] = b[j];
            }
            xa[WD] = atop;
            xb[WD] = btop;
#pragma unroll
            for (int i = WD; i >= WS; --i) {
                uint32_t da[WS], db[WS];
#pragma unroll
                for (int q = 0; q < WS; ++q) {
                    da[q] = xa[i - WS + q];
                    db[q] = xb[i - WS + q];
                }
                fold_word<WS, REP>(da, xa[i], T, rep);
                fold_word<WS, REP>(db, xb[i], T, rep);
#pragma unroll
                for (int q = 0; q < WS; ++q) {
                    xa[i - WS + q] = da[q];
                    xb[i - WS + q] = db[q];
                }
            }
            uint32_t ra[WS], rb[WS], pq[2 * WS];
#pragma unroll
            for (int q = 0; q < WS; ++q) {
                ra[q] = xa[q];
                rb[q] = xb[q];
            }
            clmul_kara<WS>(ra, rb, pq); // degree <= 2 (32 WS - 1): 2 WS words
#pragma unroll
            for (int i = 2 * WS - 1; i >= WS; --i) {
                uint32_t dst[WS];
#pragma unroll
                for (int q = 0; q < WS; ++q) dst[q] = pq[i - WS + q];
                fold_word<WS, REP>(dst, pq[i], T, rep);
#pragma unroll
                for (int q = 0; q < WS; ++q) pq[i - WS + q] = dst[q];
            }
            if (u < n) {
                uint32_t *out = reinterpret_cast<uint32_t *>(O + u * (WS / 2));
                if constexpr (WS % 4 == 0) {
#pragma unroll
                    for (int q = 0; q < WS / 4; ++q)
                        reinterpret_cast<uint4 *>(out)[q] = make_uint4(pq[4 * q], pq[4 * q + 1], pq[4 * q + 2], pq[4 * q + 3]);
                } else {
#pragma unroll
                    for (int q = 0; q < WS; ++q) out[q] = pq[q];
                }
            }
            __syncthreads(); // everyone is done with tiles[buf] before it is refilled
            continue;
        }
        uint32_t lowp[2 * WD];
        if constexpr (MODE == 0) clmul_regs<WD, WD>(a, b, lowp);
        else if constexpr (MODE == 1) clmul_imad<WD, WD>(a, b, lowp);
        else clmul_kara<WD>(a, b, lowp);
        uint32_t pr[NP];
#pragma unroll
        for (int i = 0; i < 2 * WD; ++i) pr[i] = lowp[i];
        const uint32_t ma = 0u - atop, mb = 0u - btop;
#pragma unroll
        for (int j = 0; j < WD; ++j) pr[WD + j] ^= (b[j] & ma) ^ (a[j] & mb);
        pr[2 * WD] = atop & btop;
        // fold words NP-1 .. WS down into the low WS words
#pragma unroll
        for (int i = NP - 1; i >= WS; --i) {
            uint32_t dst[WS];
#pragma unroll
            for (int q = 0; q < WS; ++q) dst[q] = pr[i - WS + q];
            fold_word<WS, REP>(dst, pr[i], T, rep);
#pragma unroll
            for (int q = 0; q < WS; ++q) pr[i - WS + q] = dst[q];
        }
        if (u < n) {
            uint32_t *out = reinterpret_cast<uint32_t *>(O + u * (WS / 2));
            if constexpr (WS % 4 == 0) {
#pragma unroll
                for (int q = 0; q < WS / 4; ++q)
                    reinterpret_cast<uint4 *>(out)[q] = make_uint4(pr[4 * q], pr[4 * q + 1], pr[4 * q + 2], pr[4 * q + 3]);
            } else {
#pragma unroll
                for (int q = 0; q < WS; ++q) out[q] = pr[q];
            }
        }
        __syncthreads(); // everyone is done with tiles[buf] before it is refilled
    }
}

// ----------------------------------------------------------------------------------------
// K5c  config A (D = 256, d = 128) fused mul+rem, operands reduced first (MODE 3 above) with a conflict-free fold table
// that needs no 8-way replication: a fold is four lookups (one per byte k of the folded word) whose order is free, so
// lane l does them in the order k = (q + l) % 4 and the 16-byte row of (table k, copy c, byte e) sits at
// ((e * 4 + k) * 2 + c) * 16 bytes, i.e. in bank group 2 k + c whatever e is.  With c = (l >> 2) & 1 the eight lanes of a
// quarter-warp read eight different bank groups in every LDS.128.  32 KB of tables instead of 128 KB, which leaves room
// for 1024-thread CTAs (32 warps per SM) with double-buffered TMA tiles.
// ----------------------------------------------------------------------------------------
struct FoldRot {
    uint32_t sel[4]; // PRMT selectors: byte (q + l) % 4 of the folded word -> byte 0
    uint32_t off[4]; // shared-memory byte address of row (e = 0, k = (q + l) % 4, c)
};
__device__ __forceinline__ FoldRot fold_rot_init(const uint32_t *T2, int lane) {
    FoldRot f;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(T2), c = (lane >> 2) & 1;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t k = (q + lane) & 3;
        f.sel[q] = 0x4440u | k;
        f.off[q] = sbase + (k * 2 + c) * 16;
    }
    return f;
}
__device__ __forceinline__ void fold_word_rot(uint32_t (&dst)[4], uint32_t t, const FoldRot &f) {
    uint32_t x[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t e;
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(e) : "r"(t), "r"(0u), "r"(f.sel[q]));
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x[q][0]), "=r"(x[q][1]), "=r"(x[q][2]), "=r"(x[q][3]) : "r"(e * 128u + f.off[q]));
    }
#pragma unroll
    for (int w = 0; w < 4; ++w) dst[w] ^= x[0][w] ^ x[1][w] ^ x[2][w] ^ x[3][w];
}

template <int MR_THREADS>
__global__ void __launch_bounds__(MR_THREADS, 1) mulrem_fresh_a_kernel(const uint64_t *__restrict__ A, const uint64_t *__restrict__ B,
                                                                    uint64_t *__restrict__ O, uint64_t n,
                                                                    const uint32_t *__restrict__ Tg) {
    constexpr int WD = 8, WS = 4, WF = WD / 2 + 1;
    extern __shared__ __align__(16) uint32_t smem32[];
    __shared__ __align__(8) uint64_t bars[2];
    uint32_t *T2 = smem32;                                                   // 256 * 4 * 2 rows of 4 words
    uint64_t *tiles = reinterpret_cast<uint64_t *>(smem32 + 256 * 4 * 2 * 4); // [2 bufs][2 operands][MR_THREADS*WF]
    const int tid = threadIdx.x;
    for (uint32_t i = tid; i < 256u * 4u * 2u * 4u; i += MR_THREADS) {
        const uint32_t w = i & 3, c = (i >> 2) & 1, k = (i >> 3) & 3, e = i >> 5;
        (void)c;
        T2[i] = Tg[(k * 256 + e) * WS + w];
    }
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const FoldRot fr = fold_rot_init(T2, tid & 31);
    const uint32_t krow = (uint32_t)__cvta_generic_to_shared(T2) + ((1 * 4 + 0) * 2 + 0) * 16; // row (byte 1, table 0, copy 0)
    constexpr uint32_t TILE_WORDS = MR_THREADS * WF;
    const uint64_t nfull = n / MR_THREADS;
    const uint64_t ntiles = (n + MR_THREADS - 1) / MR_THREADS;
    if (tid == 0 && blockIdx.x < nfull) {
        mbar_expect_tx(&bars[0], 2 * TILE_WORDS * 8);
        tma_load_1d(tiles, A + (uint64_t)blockIdx.x * TILE_WORDS, TILE_WORDS * 8, &bars[0]);
        tma_load_1d(tiles + TILE_WORDS, B + (uint64_t)blockIdx.x * TILE_WORDS, TILE_WORDS * 8, &bars[0]);
    }
    uint32_t it = 0;
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const uint32_t buf = it & 1;
        const uint64_t tn = t + gridDim.x;
        if (tid == 0 && tn < nfull) {
            uint64_t *dst = tiles + (size_t)(buf ^ 1) * 2 * TILE_WORDS;
            mbar_expect_tx(&bars[buf ^ 1], 2 * TILE_WORDS * 8);
            tma_load_1d(dst, A + tn * TILE_WORDS, TILE_WORDS * 8, &bars[buf ^ 1]);
            tma_load_1d(dst + TILE_WORDS, B + tn * TILE_WORDS, TILE_WORDS * 8, &bars[buf ^ 1]);
        }
        uint32_t xa[WD + 1], xb[WD + 1];
        const uint64_t u = t * MR_THREADS + tid;
        if (t < nfull) {
            mbar_wait(&bars[buf], (it >> 1) & 1);
            const uint64_t *sa = tiles + (size_t)buf * 2 * TILE_WORDS + (size_t)tid * WF;
            const uint64_t *sb = sa + TILE_WORDS;
#pragma unroll
            for (int j = 0; j < WD / 2; ++j) {
                const uint64_t x = sa[j], y = sb[j];
                xa[2 * j] = (uint32_t)x; xa[2 * j + 1] = (uint32_t)(x >> 32);
                xb[2 * j] = (uint32_t)y; xb[2 * j + 1] = (uint32_t)(y >> 32);
            }
            xa[WD] = (uint32_t)sa[WD / 2] & 1u;
            xb[WD] = (uint32_t)sb[WD / 2] & 1u;
        } else { // ragged last tile: straight from global
#pragma unroll
            for (int j = 0; j <= WD; ++j) xa[j] = xb[j] = 0;
            if (u < n) {
                const uint64_t *ga = A + u * WF, *gb = B + u * WF;
#pragma unroll
                for (int j = 0; j < WD / 2; ++j) {
                    const uint64_t x = ga[j], y = gb[j];
                    xa[2 * j] = (uint32_t)x; xa[2 * j + 1] = (uint32_t)(x >> 32);
                    xb[2 * j] = (uint32_t)y; xb[2 * j + 1] = (uint32_t)(y >> 32);
                }
                xa[WD] = (uint32_t)ga[WD / 2] & 1u;
                xb[WD] = (uint32_t)gb[WD / 2] & 1u;
            }
        }
        // a mod S, b mod S (two independent fold chains), top down.  The top "word" is the single coefficient of X^256:
        // its fold is one row (table 0, byte 1 = X^256 mod S; every lane reads the same address, a broadcast) under a mask.
        {
            uint32_t k0, k1, k2, k3;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(k0), "=r"(k1), "=r"(k2), "=r"(k3) : "r"(krow));
            const uint32_t ma = 0u - xa[WD], mb = 0u - xb[WD];
            xa[WD - 4] ^= k0 & ma; xa[WD - 3] ^= k1 & ma; xa[WD - 2] ^= k2 & ma; xa[WD - 1] ^= k3 & ma;
            xb[WD - 4] ^= k0 & mb; xb[WD - 3] ^= k1 & mb; xb[WD - 2] ^= k2 & mb; xb[WD - 1] ^= k3 & mb;
        }
#pragma unroll
        for (int i = WD - 1; i >= WS; --i) {
            uint32_t da[WS], db[WS];
#pragma unroll
            for (int q = 0; q < WS; ++q) {
                da[q] = xa[i - WS + q];
                db[q] = xb[i - WS + q];
            }
            fold_word_rot(da, xa[i], fr);
            fold_word_rot(db, xb[i], fr);
#pragma unroll
            for (int q = 0; q < WS; ++q) {
                xa[i - WS + q] = da[q];
                xb[i - WS + q] = db[q];
            }
        }
        uint32_t ra[WS], rb[WS], pq[2 * WS];
#pragma unroll
        for (int q = 0; q < WS; ++q) {
            ra[q] = xa[q];
            rb[q] = xb[q];
        }
        clmul_kara<WS>(ra, rb, pq);
#pragma unroll
        for (int i = 2 * WS - 1; i >= WS; --i) {
            uint32_t dst[WS];
#pragma unroll
            for (int q = 0; q < WS; ++q) dst[q] = pq[i - WS + q];
            fold_word_rot(dst, pq[i], fr);
#pragma unroll
            for (int q = 0; q < WS; ++q) pq[i - WS + q] = dst[q];
        }
        if (u < n) *reinterpret_cast<uint4 *>(O + u * (WS / 2)) = make_uint4(pq[0], pq[1], pq[2], pq[3]);
        __syncthreads(); // everyone is done with tiles[buf] before it is refilled
    }
}

// ----------------------------------------------------------------------------------------
// K6  fused ripple-carry adder, one warp per value        reference common.rs:37-56
//
//   p_k = a_k + b_k, g_k = a_k * b_k, m_k = (1 + g_k) * p_k           (lane k, registers)
//   c_0 = 0, c_{k+1} = m_k * c_k + g_k,  s_k = p_k + c_k             (serial over k)
// which is the reference's  carry' = p*c + (a*b)*(p*c + 1)  regrouped with ring identities.
// The long carry polynomial c_k lives in shared memory (two buffers); every lane owns tiles of
// 8 output words of c_{k+1} and computes them completely (no write conflicts).  m_k is the
// same for all lanes, so its bits are warp-uniform: zero bits are skipped by uniform branches.
// Horner over the bit index s of m_k: acc = acc*X ^ sum_{j in B_s} c[r - j], with one halo
// word below the tile so that the shift never needs a neighbour.
// ----------------------------------------------------------------------------------------
template <int WD> struct AdderCfg {
    static constexpr int NP = WD + 1;      // words of p_k (bit D in the last)
    static constexpr int NG = 2 * WD + 1;  // words of g_k
    static constexpr int NM = 3 * WD + 1;  // words of m_k
    static constexpr int TQMAX = 24;       // widest output tile (words per lane per pass)
    static constexpr int PAD = NM + 7;     // zero words in front of each carry buffer (window underflow)
    __host__ __device__ static constexpr uint32_t carry_cap(uint32_t L) { // words
        return (((3 * (L - 1) - 1) * WD + 1 + TQMAX - 1) / TQMAX) * TQMAX + TQMAX;
    }
    __host__ __device__ static constexpr uint32_t warp_words(uint32_t L) {
        return ((L * (NP + NG + NM) + (PAD + carry_cap(L)) + 3) / 4) * 4;
    }
};

// One pass of a step: lane l computes output words [w0 + l*TQ, w0 + (l+1)*TQ) of c_{k+1} = m_k*c_k + g_k.
template <int WD, int TQ, int MODE>
__device__ __forceinline__ void adder_step_pass(uint32_t *cbuf, const uint32_t *__restrict__ gk, uint32_t Bmine,
                                                uint32_t w0, uint32_t len_next, int lane) {
    using C = AdderCfg<WD>;
    constexpr int NM = C::NM, NW = NM + TQ; // window words
    static_assert(NM <= 32, "the multiplier's word mask must fit one 32-bit register (D <= 320)");
    uint32_t win[NW], acc[TQ + 1];
    const uint32_t t0 = w0 + (uint32_t)lane * TQ;
    const bool live = t0 < len_next;
    {
        // win[y] = c[t0 - NM + y]; PAD >= NM zero words sit in front of the buffer, so no lower bound check;
        // lanes past the end of the polynomial read the (zero) front of the buffer instead
        const uint32_t *src = cbuf + (live ? (int)t0 : 0) - NM;
#pragma unroll
        for (int y = 0; y < NW; ++y) win[y] = src[y];
#pragma unroll
        for (int i = 0; i <= TQ; ++i) acc[i] = 0;
    }
    __syncwarp(); // the carry is updated in place: every lane holds its window before anyone stores
#pragma unroll 1
    for (int s = 31; s >= 0; --s) {
        // bit j set <=> bit s of m_k[j] set (warp-uniform)
        // (a REDUX-to-uniform-register variant that moves the bit tests to the uniform datapath measured 2 % slower)
        const uint32_t Bs = __shfl_sync(FULL, Bmine, s);
#pragma unroll
        for (int i = TQ; i > 0; --i) acc[i] = __funnelshift_l(acc[i - 1], acc[i], 1);
        acc[0] <<= 1;
        // acc[i] is output word t0-1+i (acc[0] = halo); it receives c[t0-1+i-j] = win[i + NM-1 - j]
        if constexpr (MODE == 0) { // pairs of multiplier words: one 3-input LOP3 when both bits are set
#pragma unroll
            for (int jp = 0; jp + 1 < NM; jp += 2) {
                const uint32_t sel = (Bs >> jp) & 3u;
                if (sel == 3u) {
#pragma unroll
                    for (int i = 0; i <= TQ; ++i) acc[i] ^= win[i + NM - 1 - jp] ^ win[i + NM - 2 - jp];
                } else if (sel == 1u) {
#pragma unroll
                    for (int i = 0; i <= TQ; ++i) acc[i] ^= win[i + NM - 1 - jp];
                } else if (sel == 2u) {
#pragma unroll
                    for (int i = 0; i <= TQ; ++i) acc[i] ^= win[i + NM - 2 - jp];
                }
            }
            if constexpr (NM & 1) {
                if ((Bs >> (NM - 1)) & 1u) {
#pragma unroll
                    for (int i = 0; i <= TQ; ++i) acc[i] ^= win[i];
                }
            }
        } else { // one uniform branch per multiplier word (smaller code, fewer taken branches, more LOP3)
#pragma unroll
            for (int j = 0; j < NM; ++j) {
                if ((Bs >> j) & 1u) {
#pragma unroll
                    for (int i = 0; i <= TQ; ++i) acc[i] ^= win[i + NM - 1 - j];
                }
            }
        }
    }
    if (live) {
        uint32_t o[TQ];
#pragma unroll
        for (int i = 0; i < TQ; ++i) {
            const uint32_t x = t0 + i;
            o[i] = acc[i + 1] ^ ((x < (uint32_t)C::NG) ? gk[x] : 0u);
        }
        uint4 *dst = reinterpret_cast<uint4 *>(cbuf + t0);
#pragma unroll
        for (int q = 0; q < TQ / 4; ++q) dst[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
    }
}

template <int WD, int MODE, int TS>
__global__ void __launch_bounds__(128, 5) adder_fused_kernel(const uint64_t *__restrict__ A, const uint64_t *__restrict__ B,
                                                             uint64_t *__restrict__ O, uint64_t n, uint32_t L, Layout lo) {
    using C = AdderCfg<WD>;
    constexpr int WF = WD / 2 + 1;
    constexpr int NP = C::NP, NG = C::NG, NM = C::NM;
    extern __shared__ __align__(16) uint32_t smem32[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t v = (uint64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (v >= n) return;
    const uint32_t cap = C::carry_cap(L);
    uint32_t *base = smem32 + (size_t)warp * C::warp_words(L);
    uint32_t *cb = base + C::PAD;                 // carry c_k: PAD zero words in front, cap words (16-byte aligned)
    uint32_t *P = base + (C::PAD + cap);          // L x NP
    uint32_t *G = P + L * NP;                     // L x NG
    uint32_t *M = G + L * NG;                     // L x NM
    static_assert((C::PAD % 4) == 0 && (C::TQMAX % 4) == 0, "carry buffers must stay 16-byte aligned");

    for (uint32_t i = lane; i < C::PAD + cap; i += 32) base[i] = 0;

    // ---- stage 1: per-bit products, lane k owns bit k ------------------------------------
    const uint64_t *Av = A + v * (uint64_t)L * WF, *Bv = B + v * (uint64_t)L * WF;
    for (uint32_t k = lane; k < ((L + 31) & ~31u); k += 32) {
        uint32_t a[WD], b[WD], p[WD], atop = 0, btop = 0;
        const bool live = k < L;
#pragma unroll
        for (int j = 0; j < WD / 2; ++j) {
            const uint64_t x = live ? __ldg(Av + (size_t)k * WF + j) : 0ull;
            const uint64_t y = live ? __ldg(Bv + (size_t)k * WF + j) : 0ull;
            a[2 * j] = (uint32_t)x; a[2 * j + 1] = (uint32_t)(x >> 32);
            b[2 * j] = (uint32_t)y; b[2 * j + 1] = (uint32_t)(y >> 32);
        }
        if (live) {
            atop = (uint32_t)__ldg(Av + (size_t)k * WF + WD / 2) & 1u;
            btop = (uint32_t)__ldg(Bv + (size_t)k * WF + WD / 2) & 1u;
        }
#pragma unroll
        for (int j = 0; j < WD; ++j) p[j] = a[j] ^ b[j];
        const uint32_t ptop = atop ^ btop;
        uint32_t g[2 * WD];
        clmul_regs<WD, WD>(a, b, g);
        const uint32_t ma = 0u - atop, mb = 0u - btop;
#pragma unroll
        for (int j = 0; j < WD; ++j) g[WD + j] ^= (b[j] & ma) ^ (a[j] & mb);
        const uint32_t gtop = atop & btop; // coefficient of X^(2D)
        uint32_t gp[3 * WD];
        clmul_regs<2 * WD, WD>(g, p, gp);
        const uint32_t mp = 0u - ptop, mg = 0u - gtop;
#pragma unroll
        for (int j = 0; j < 2 * WD; ++j) gp[WD + j] ^= g[j] & mp;
#pragma unroll
        for (int j = 0; j < WD; ++j) gp[2 * WD + j] ^= p[j] & mg;
#pragma unroll
        for (int j = 0; j < WD; ++j) gp[j] ^= p[j]; // m = p + g*p
        gp[WD] ^= ptop;
        if (live) {
#pragma unroll
            for (int j = 0; j < WD; ++j) P[k * NP + j] = p[j];
            P[k * NP + WD] = ptop;
#pragma unroll
            for (int j = 0; j < 2 * WD; ++j) G[k * NG + j] = g[j];
            G[k * NG + 2 * WD] = gtop;
#pragma unroll
            for (int j = 0; j < 3 * WD; ++j) M[k * NM + j] = gp[j];
            M[k * NM + 3 * WD] = gtop & ptop; // coefficient of X^(3D)
        }
    }
    __syncwarp();

    // ---- s_0 = p_0 ; c_1 = g_0 ------------------------------------------------------------
    uint64_t *Ov = O + v * (uint64_t)lo.value_words;
    for (uint32_t j = lane; j < lo.off[1] - lo.off[0]; j += 32) {
        const uint32_t x0 = (2 * j < (uint32_t)NP) ? P[2 * j] : 0u, x1 = (2 * j + 1 < (uint32_t)NP) ? P[2 * j + 1] : 0u;
        Ov[lo.off[0] + j] = (uint64_t)x0 | ((uint64_t)x1 << 32);
    }
    for (uint32_t j = lane; j < (uint32_t)NG; j += 32) cb[j] = G[j];
    __syncwarp();

    // ---- serial chain ---------------------------------------------------------------------
    for (uint32_t k = 1; k < L; ++k) {
        // s_k = p_k + c_k, coalesced 64-bit stores
        {
            const uint32_t wo = lo.off[k + 1] - lo.off[k];
            uint64_t *dst = Ov + lo.off[k];
            const uint32_t *pk = P + k * NP;
            for (uint32_t j = lane; j < wo; j += 32) {
                uint32_t x0 = cb[2 * j], x1 = cb[2 * j + 1];
                if (2 * j < (uint32_t)NP) x0 ^= pk[2 * j];
                if (2 * j + 1 < (uint32_t)NP) x1 ^= pk[2 * j + 1];
                dst[j] = (uint64_t)x0 | ((uint64_t)x1 << 32);
            }
        }
        if (k + 1 == L) break; // no carry out of the last bit (common.rs:47-49)
        // transpose m_k: lane s gets the mask of words j whose bit s is set
        uint32_t Bmine = 0;
        {
            const uint32_t *mk = M + k * NM;
#pragma unroll
            for (int j = 0; j < NM; ++j) Bmine |= ((mk[j] >> lane) & 1u) << j;
        }
        const uint32_t len_next = (3 * k + 2) * WD + 1; // words of c_{k+1}
        const uint32_t *gk = G + k * NG;
        __syncwarp(); // s_k has been read out of the buffer
        // passes of 32 lanes x TQ words, highest words first: a pass only reads words below its own top, so the
        // in-place update never clobbers what a later (lower) pass still needs
        constexpr uint32_t FULLW = 32 * 24;
        const uint32_t nfull = (len_next - 1) / FULLW;
        {
            // tile widths per pass: {8,16,24} (TS=0), {12,24} (TS=1) or {8,24} (TS=2).  Finer sets do less work per step
            // (work ~ TQ+1) but every extra instantiation is ~8 KB of hot code and the loop is instruction-fetch
            // sensitive (profiles/README.md): {4,...,24 step 4} measured 18 % SLOWER than {8,16,24}.
            const uint32_t w0 = nfull * FULLW, left = len_next - w0;
            if constexpr (TS == 0) {
                if (left > 32 * 16) adder_step_pass<WD, 24, MODE>(cb, gk, Bmine, w0, len_next, lane);
                else if (left > 32 * 8) adder_step_pass<WD, 16, MODE>(cb, gk, Bmine, w0, len_next, lane);
                else adder_step_pass<WD, 8, MODE>(cb, gk, Bmine, w0, len_next, lane);
            } else if constexpr (TS == 1) {
                if (left > 32 * 12) adder_step_pass<WD, 24, MODE>(cb, gk, Bmine, w0, len_next, lane);
                else adder_step_pass<WD, 12, MODE>(cb, gk, Bmine, w0, len_next, lane);
            } else {
                if (left > 32 * 8) adder_step_pass<WD, 24, MODE>(cb, gk, Bmine, w0, len_next, lane);
                else adder_step_pass<WD, 8, MODE>(cb, gk, Bmine, w0, len_next, lane);
            }
        }
        for (uint32_t f = nfull; f-- > 0;) {
            __syncwarp();
            adder_step_pass<WD, 24, MODE>(cb, gk, Bmine, f * FULLW, len_next, lane);
        }
        __syncwarp();
    }
}

// ----------------------------------------------------------------------------------------
// K6b  ripple-carry adder, one THREAD per value (D = 256): the chain c_{k+1} = m_k c_k + g_k evaluated with
// Karatsuba on the integer multiplier instead of the warp-uniform comb.  c_k is cut into 24-word chunks; every chunk
// times the 24 low words of m_k is a 3-way Karatsuba of six 8x8-word products (clmul_kara<8>: 27 single-word leaves
// each, 16 IMAD.WIDE + 20 LOP3 per leaf), accumulated into a 48-word register window whose upper half is the carry
// into the next chunk.  There is no carry buffer: c_k is read back from the result itself (slot k holds s_k = c_k + p_k,
// and p_k only touches its first 9 words) while s_{k+1} is written to slot k+1, so the only HBM traffic besides the
// algorithmic bytes is that one re-read.  No warp cooperation, no shared memory, no idle lanes in the short early steps.
// ----------------------------------------------------------------------------------------
template <int O> __device__ __forceinline__ void xor16_at(uint32_t (&t)[48], const uint32_t (&r)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) t[O + i] ^= r[i];
}

// t ^= m * c for 24-word m and c (3-way Karatsuba over 8-word blocks: 6 block products).  m and c are read from
// memory (per-thread scratch, L1 hits) block by block so that only the 48-word accumulator stays live, and the six
// products share ONE copy of the 8x8-word Karatsuba code (rolled loop): the kernel is instruction-fetch sensitive.
__device__ __forceinline__ void mul24_acc(const uint32_t *__restrict__ m, const uint32_t *__restrict__ c, uint32_t (&t)[48]) {
    auto blk = [](uint32_t (&dst)[8], const uint32_t *src) { // 8-byte aligned (slot offsets are in u64 words)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint2 w = *reinterpret_cast<const uint2 *>(src + 2 * q);
            dst[2 * q] = w.x;
            dst[2 * q + 1] = w.y;
        }
    };
#pragma unroll 1
    for (int i = 0; i < 6; ++i) {
        // i: 0 -> P0 = m0 c0, 1 -> P1 = m1 c1, 2 -> P2 = m2 c2, 3 -> (m0+m1)(c0+c1), 4 -> (m0+m2)(c0+c2), 5 -> (m1+m2)(c1+c2)
        const int o1 = (i < 3) ? 8 * i : (i == 5 ? 8 : 0);
        const int o2 = (i < 3) ? -1 : (i == 3 ? 8 : 16);
        uint32_t x[8], y[8], r[16];
        blk(x, m + o1);
        blk(y, c + o1);
        if (o2 >= 0) {
            uint32_t u[8], w[8];
            blk(u, m + o2);
            blk(w, c + o2);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                x[q] ^= u[q];
                y[q] ^= w[q];
            }
        }
        clmul_kara<8>(x, y, r);
        switch (i) {
            case 0: xor16_at<0>(t, r); xor16_at<8>(t, r); xor16_at<16>(t, r); break;   // P0 at 0, cross terms at 8, 16
            case 1: xor16_at<16>(t, r); xor16_at<8>(t, r); xor16_at<24>(t, r); break;  // P1 at 16, cross terms at 8, 24
            case 2: xor16_at<32>(t, r); xor16_at<16>(t, r); xor16_at<24>(t, r); break; // P2 at 32, cross terms at 16, 24
            case 3: xor16_at<8>(t, r); break;
            case 4: xor16_at<16>(t, r); break;
            default: xor16_at<24>(t, r); break;
        }
    }
}

// out-of-line 8x8-word product for the once-per-bit quantities (g_k, m_k): keeps their three uses from tripling the code
static __device__ __noinline__ void kara8_call(const uint32_t *x, const uint32_t *y, uint32_t *r) {
    uint32_t a[8], b[8], o[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = x[i];
        b[i] = y[i];
    }
    clmul_kara<8>(a, b, o);
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = o[i];
}

constexpr int ADT_THREAD_WORDS = 64; // per-thread scratch: m_k (24 words) + one staged chunk (24 words)

template <int MINB>
__global__ void __launch_bounds__(128, MINB) adder_thread_kernel(const uint64_t *__restrict__ A, const uint64_t *__restrict__ B,
                                                              uint64_t *__restrict__ O, uint64_t n, uint32_t L, Layout lo,
                                                              uint32_t *__restrict__ scratch) {
    constexpr int WD = 8, WF = 5;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (uint64_t)gridDim.x * blockDim.x;
    uint32_t *mbuf = scratch + tid * ADT_THREAD_WORDS, *tmp = mbuf + 32;
    for (uint64_t v = tid; v < n; v += nthreads) {
        const uint64_t *Av = A + v * (uint64_t)L * WF, *Bv = B + v * (uint64_t)L * WF;
        uint32_t *Ov = reinterpret_cast<uint32_t *>(O + v * (uint64_t)lo.value_words);
        uint32_t len = 0; // words of c_k (c_0 = 0)
        for (uint32_t k = 0; k < L; ++k) {
            // ---- per-bit quantities: p = a + b, g = a * b, m = (1 + g) * p -----------------------------------------
            uint32_t a[WD], b[WD], p[WD + 1];
#pragma unroll
            for (int j = 0; j < WD / 2; ++j) {
                const uint64_t x = __ldg(Av + (size_t)k * WF + j), y = __ldg(Bv + (size_t)k * WF + j);
                a[2 * j] = (uint32_t)x; a[2 * j + 1] = (uint32_t)(x >> 32);
                b[2 * j] = (uint32_t)y; b[2 * j + 1] = (uint32_t)(y >> 32);
            }
            const uint32_t atop = (uint32_t)__ldg(Av + (size_t)k * WF + WD / 2) & 1u, btop = (uint32_t)__ldg(Bv + (size_t)k * WF + WD / 2) & 1u;
#pragma unroll
            for (int j = 0; j < WD; ++j) p[j] = a[j] ^ b[j];
            const uint32_t ptop = atop ^ btop;
            p[WD] = ptop;
            if (k == 0) { // s_0 = p_0
                uint32_t *dst = Ov + 2 * lo.off[0];
                const uint32_t wo = 2 * (lo.off[1] - lo.off[0]);
#pragma unroll
                for (int j = 0; j <= WD; ++j) dst[j] = p[j];
                for (uint32_t j = WD + 1; j < wo; ++j) dst[j] = 0;
            }
            if (k + 1 == L) break; // s_k (k >= 1) was written while c_k was produced; no carry out of the last bit
            uint32_t g[2 * WD];
            kara8_call(a, b, g);
            const uint32_t ma = 0u - atop, mb = 0u - btop;
#pragma unroll
            for (int j = 0; j < WD; ++j) g[WD + j] ^= (b[j] & ma) ^ (a[j] & mb);
            const uint32_t gtop = atop & btop;
            // next bit's p (s_{k+1} = c_{k+1} + p_{k+1} is emitted on the fly)
            uint32_t pn[WD + 1];
#pragma unroll
            for (int j = 0; j < WD / 2; ++j) {
                const uint64_t x = __ldg(Av + (size_t)(k + 1) * WF + j) ^ __ldg(Bv + (size_t)(k + 1) * WF + j);
                pn[2 * j] = (uint32_t)x; pn[2 * j + 1] = (uint32_t)(x >> 32);
            }
            pn[WD] = (uint32_t)(__ldg(Av + (size_t)(k + 1) * WF + WD / 2) ^ __ldg(Bv + (size_t)(k + 1) * WF + WD / 2)) & 1u;
            uint32_t *sdst = Ov + 2 * lo.off[k + 1];
            const uint32_t swo = 2 * (lo.off[k + 2] - lo.off[k + 1]);
            if (k == 0) { // c_1 = g_0
#pragma unroll
                for (int j = 0; j < 2 * WD; ++j) sdst[j] = g[j] ^ (j <= WD ? pn[j] : 0u);
                sdst[2 * WD] = gtop;
                for (uint32_t j = 2 * WD + 1; j < swo; ++j) sdst[j] = 0;
                len = 2 * WD + 1;
                continue;
            }
            // m = p + g * p : 24 low words (to scratch: read block-wise by mul24_acc) + the coefficient of X^768
            {
                uint32_t m[24], glo[WD], ghi[WD], q0[2 * WD], q1[2 * WD];
#pragma unroll
                for (int j = 0; j < WD; ++j) { glo[j] = g[j]; ghi[j] = g[WD + j]; }
                kara8_call(glo, p, q0);
                kara8_call(ghi, p, q1);
#pragma unroll
                for (int j = 0; j < WD; ++j) {
                    m[j] = q0[j] ^ p[j];
                    m[WD + j] = q0[WD + j] ^ q1[j];
                    m[2 * WD + j] = q1[WD + j];
                }
                const uint32_t mp = 0u - ptop, mg = 0u - gtop;
#pragma unroll
                for (int j = 0; j < 2 * WD; ++j) m[WD + j] ^= g[j] & mp;
#pragma unroll
                for (int j = 0; j < WD; ++j) m[2 * WD + j] ^= p[j] & mg;
                m[WD] ^= ptop;
#pragma unroll
                for (int q = 0; q < 12; ++q) reinterpret_cast<uint2 *>(mbuf)[q] = make_uint2(m[2 * q], m[2 * q + 1]);
            }
            const uint32_t mtop = gtop & ptop;
            // ---- c_{k+1} = m c_k + g_k, chunk by chunk; t[0..24) carries the upper half of the previous chunk ---------
            const uint32_t *cslot = Ov + 2 * lo.off[k]; // s_k = c_k + p_k
            uint32_t t[48];
#pragma unroll
            for (int i = 0; i < 48; ++i) t[i] = 0;
            const uint32_t nchunks = (len + 23) / 24;
            for (uint32_t j = 0; j <= nchunks; ++j) {
                if (j < nchunks) {
                    const uint32_t *cp = cslot + 24 * j;
                    if (j == 0 || 24 * j + 24 > len) { // chunk touches p_k or the end of the polynomial: stage a clean copy
                        for (uint32_t i = 0; i < 24; ++i) {
                            const uint32_t w = 24 * j + i;
                            uint32_t val = (w < len) ? cslot[w] : 0u;
                            if (w <= (uint32_t)WD) {
                                uint32_t pw = p[0];
#pragma unroll
                                for (int q = 1; q <= WD; ++q) pw = (w == (uint32_t)q) ? p[q] : pw;
                                val ^= pw;
                            }
                            tmp[i] = val;
                        }
                        cp = tmp;
                    }
                    mul24_acc(mbuf, cp, t);
                    if (mtop) { // X^768 * c: chunk j lands in chunk j+1
#pragma unroll
                        for (int q = 0; q < 12; ++q) {
                            const uint2 w = *reinterpret_cast<const uint2 *>(cp + 2 * q);
                            t[24 + 2 * q] ^= w.x;
                            t[24 + 2 * q + 1] ^= w.y;
                        }
                    }
                }
                if (j == 0) {
#pragma unroll
                    for (int i = 0; i < 2 * WD; ++i) t[i] ^= g[i];
                    t[2 * WD] ^= gtop;
#pragma unroll
                    for (int i = 0; i <= WD; ++i) t[i] ^= pn[i]; // s_{k+1} = c_{k+1} + p_{k+1}
                }
#pragma unroll
                for (int q = 0; q < 12; ++q)
                    if (24 * j + 2 * q < swo) *reinterpret_cast<uint2 *>(sdst + 24 * j + 2 * q) = make_uint2(t[2 * q], t[2 * q + 1]);
#pragma unroll
                for (int i = 0; i < 24; ++i) {
                    t[i] = t[24 + i];
                    t[24 + i] = 0;
                }
            }
            len += 3 * WD; // deg c_{k+1} = deg c_k + 3D
        }
    }
}

// ----------------------------------------------------------------------------------------
// K6c  the same thread-per-value chain with its working set in shared memory: m_k and a double-buffered 24-word chunk
// of c_k live in a [word pair][thread] array (one 8-byte column per thread: conflict-free LDS.64/STS.64), and chunk j+1
// is fetched from the result slot with cp.async (LDGSTS, no registers) while chunk j is being multiplied, so the
// global-memory latency of the operand reads is off the critical path of the four resident warps per scheduler.
// ----------------------------------------------------------------------------------------
constexpr int ADS_PAIRS = 36;                          // 12 (m_k) + 2 x 12 (chunk double buffer) uint2 per thread
constexpr int ADS_SMEM_BYTES = ADS_PAIRS * 128 * 8;    // per 128-thread CTA

// t ^= m * c, both operands in the interleaved shared layout (pair q of this thread at base[q * 128])
__device__ __forceinline__ void mul24_acc_s(const uint2 *__restrict__ m, const uint2 *__restrict__ c, uint32_t (&t)[48]) {
    auto blk = [](uint32_t (&dst)[8], const uint2 *src) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint2 w = src[q * 128];
            dst[2 * q] = w.x;
            dst[2 * q + 1] = w.y;
        }
    };
#pragma unroll 1
    for (int i = 0; i < 6; ++i) {
        const int o1 = (i < 3) ? 4 * i : (i == 5 ? 4 : 0); // in pairs
        const int o2 = (i < 3) ? -1 : (i == 3 ? 4 : 8);
        uint32_t x[8], y[8], r[16];
        blk(x, m + o1 * 128);
        blk(y, c + o1 * 128);
        if (o2 >= 0) {
            uint32_t u[8], w[8];
            blk(u, m + o2 * 128);
            blk(w, c + o2 * 128);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                x[q] ^= u[q];
                y[q] ^= w[q];
            }
        }
        clmul_kara<8>(x, y, r);
        switch (i) {
            case 0: xor16_at<0>(t, r); xor16_at<8>(t, r); xor16_at<16>(t, r); break;
            case 1: xor16_at<16>(t, r); xor16_at<8>(t, r); xor16_at<24>(t, r); break;
            case 2: xor16_at<32>(t, r); xor16_at<16>(t, r); xor16_at<24>(t, r); break;
            case 3: xor16_at<8>(t, r); break;
            case 4: xor16_at<16>(t, r); break;
            default: xor16_at<24>(t, r); break;
        }
    }
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB) adder_thread_smem_kernel(const uint64_t *__restrict__ A, const uint64_t *__restrict__ B,
                                                                   uint64_t *__restrict__ O, uint64_t n, uint32_t L, Layout lo) {
    constexpr int WD = 8, WF = 5;
    extern __shared__ __align__(16) uint2 ads_smem[];
    uint2 *mb = ads_smem + threadIdx.x, *cb0 = mb + 12 * 128, *cb1 = mb + 24 * 128;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = tid; v < n; v += nthreads) {
        const uint64_t *Av = A + v * (uint64_t)L * WF, *Bv = B + v * (uint64_t)L * WF;
        uint32_t *Ov = reinterpret_cast<uint32_t *>(O + v * (uint64_t)lo.value_words);
        uint32_t len = 0; // words of c_k (c_0 = 0)
        for (uint32_t k = 0; k < L; ++k) {
            uint32_t a[WD], b[WD], p[WD + 1];
#pragma unroll
            for (int j = 0; j < WD / 2; ++j) {
                const uint64_t x = __ldg(Av + (size_t)k * WF + j), y = __ldg(Bv + (size_t)k * WF + j);
                a[2 * j] = (uint32_t)x; a[2 * j + 1] = (uint32_t)(x >> 32);
                b[2 * j] = (uint32_t)y; b[2 * j + 1] = (uint32_t)(y >> 32);
            }
            const uint32_t atop = (uint32_t)__ldg(Av + (size_t)k * WF + WD / 2) & 1u, btop = (uint32_t)__ldg(Bv + (size_t)k * WF + WD / 2) & 1u;
#pragma unroll
            for (int j = 0; j < WD; ++j) p[j] = a[j] ^ b[j];
            const uint32_t ptop = atop ^ btop;
            p[WD] = ptop;
            if (k == 0) { // s_0 = p_0
                uint32_t *dst = Ov + 2 * lo.off[0];
                const uint32_t wo = 2 * (lo.off[1] - lo.off[0]);
#pragma unroll
                for (int j = 0; j <= WD; ++j) dst[j] = p[j];
                for (uint32_t j = WD + 1; j < wo; ++j) dst[j] = 0;
            }
            if (k + 1 == L) break;
            uint32_t g[2 * WD];
            kara8_call(a, b, g);
            const uint32_t ma = 0u - atop, mbm = 0u - btop;
#pragma unroll
            for (int j = 0; j < WD; ++j) g[WD + j] ^= (b[j] & ma) ^ (a[j] & mbm);
            const uint32_t gtop = atop & btop;
            uint32_t pn[WD + 1];
#pragma unroll
            for (int j = 0; j < WD / 2; ++j) {
                const uint64_t x = __ldg(Av + (size_t)(k + 1) * WF + j) ^ __ldg(Bv + (size_t)(k + 1) * WF + j);
                pn[2 * j] = (uint32_t)x; pn[2 * j + 1] = (uint32_t)(x >> 32);
            }
            pn[WD] = (uint32_t)(__ldg(Av + (size_t)(k + 1) * WF + WD / 2) ^ __ldg(Bv + (size_t)(k + 1) * WF + WD / 2)) & 1u;
            uint32_t *sdst = Ov + 2 * lo.off[k + 1];
            const uint32_t swo = 2 * (lo.off[k + 2] - lo.off[k + 1]);
            if (k == 0) { // c_1 = g_0
#pragma unroll
                for (int j = 0; j < 2 * WD; ++j) sdst[j] = g[j] ^ (j <= WD ? pn[j] : 0u);
                sdst[2 * WD] = gtop;
                for (uint32_t j = 2 * WD + 1; j < swo; ++j) sdst[j] = 0;
                len = 2 * WD + 1;
                continue;
            }
            { // m = p + g * p : 24 low words into shared memory + the coefficient of X^768
                uint32_t m[24], glo[WD], ghi[WD], q0[2 * WD], q1[2 * WD];
#pragma unroll
                for (int j = 0; j < WD; ++j) { glo[j] = g[j]; ghi[j] = g[WD + j]; }
                kara8_call(glo, p, q0);
                kara8_call(ghi, p, q1);
#pragma unroll
                for (int j = 0; j < WD; ++j) {
                    m[j] = q0[j] ^ p[j];
                    m[WD + j] = q0[WD + j] ^ q1[j];
                    m[2 * WD + j] = q1[WD + j];
                }
                const uint32_t mp = 0u - ptop, mg = 0u - gtop;
#pragma unroll
                for (int j = 0; j < 2 * WD; ++j) m[WD + j] ^= g[j] & mp;
#pragma unroll
                for (int j = 0; j < WD; ++j) m[2 * WD + j] ^= p[j] & mg;
                m[WD] ^= ptop;
#pragma unroll
                for (int q = 0; q < 12; ++q) mb[q * 128] = make_uint2(m[2 * q], m[2 * q + 1]);
            }
            const uint32_t mtop = gtop & ptop;
            const uint32_t *cslot = Ov + 2 * lo.off[k]; // s_k = c_k + p_k
            const uint32_t nchunks = (len + 23) / 24;
            // first chunk (p_k mixed in) and a ragged last chunk are staged by hand; full chunks come by cp.async
            auto fill_sync = [&](uint2 *dstb, uint32_t j) {
                for (uint32_t q = 0; q < 12; ++q) {
                    uint32_t v2[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t w = 24 * j + 2 * q + h;
                        uint32_t val = (w < len) ? cslot[w] : 0u;
                        if (w <= (uint32_t)WD) {
                            uint32_t pw = p[0];
#pragma unroll
                            for (int qq = 1; qq <= WD; ++qq) pw = (w == (uint32_t)qq) ? p[qq] : pw;
                            val ^= pw;
                        }
                        v2[h] = val;
                    }
                    dstb[q * 128] = make_uint2(v2[0], v2[1]);
                }
            };
            auto fill_async = [&](uint2 *dstb, uint32_t j) {
                const uint32_t sa = (uint32_t)__cvta_generic_to_shared(dstb);
                const uint32_t *src = cslot + 24 * j;
#pragma unroll
                for (int q = 0; q < 12; ++q)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa + q * 128 * 8), "l"(src + 2 * q) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
            fill_sync(cb0, 0);
            uint32_t t[48];
#pragma unroll
            for (int i = 0; i < 48; ++i) t[i] = 0;
            for (uint32_t j = 0; j <= nchunks; ++j) {
                uint2 *cur = (j & 1) ? cb1 : cb0, *nxt = (j & 1) ? cb0 : cb1;
                if (j < nchunks) {
                    bool pref = false;
                    if (j + 1 < nchunks) {
                        if (24 * (j + 1) + 24 <= len) {
                            fill_async(nxt, j + 1);
                            pref = true;
                        } else {
                            fill_sync(nxt, j + 1);
                        }
                    }
                    if (pref) asm volatile("cp.async.wait_group 1;" ::: "memory");
                    else asm volatile("cp.async.wait_group 0;" ::: "memory");
                    mul24_acc_s(mb, cur, t);
                    if (mtop) { // X^768 * c: chunk j lands in chunk j+1
#pragma unroll
                        for (int q = 0; q < 12; ++q) {
                            const uint2 w = cur[q * 128];
                            t[24 + 2 * q] ^= w.x;
                            t[24 + 2 * q + 1] ^= w.y;
                        }
                    }
                }
                if (j == 0) {
#pragma unroll
                    for (int i = 0; i < 2 * WD; ++i) t[i] ^= g[i];
                    t[2 * WD] ^= gtop;
#pragma unroll
                    for (int i = 0; i <= WD; ++i) t[i] ^= pn[i];
                }
#pragma unroll
                for (int q = 0; q < 12; ++q)
                    if (24 * j + 2 * q < swo) *reinterpret_cast<uint2 *>(sdst + 24 * j + 2 * q) = make_uint2(t[2 * q], t[2 * q + 1]);
#pragma unroll
                for (int i = 0; i < 24; ++i) {
                    t[i] = t[24 + i];
                    t[24 + i] = 0;
                }
            }
            len += 3 * WD;
        }
    }
}

// ----------------------------------------------------------------------------------------
// K4c  general products with the Karatsuba-on-the-multiplier machinery of the thread adder: one THREAD per
// (value, product, 24-word chunk xi of the shorter operand).  The thread multiplies its chunk by the whole longer
// operand, 24 words at a time (mul24_acc, carry in the upper half of the 48-word window) and XORs the result into the
// zero-initialised output with 32-bit atomics (threads of different xi overlap by one chunk).  Used for the large
// products of the multiplier circuit when there are enough (value, chunk) pairs to fill the GPU.
// ----------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(128, 4) mul_thread_kernel(const MulOp *__restrict__ ops, uint64_t n, uint32_t *__restrict__ scratch) {
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const MulOp op = ops[blockIdx.y];
    const uint32_t xi = blockIdx.z;
    const uint32_t *gx = reinterpret_cast<const uint32_t *>(op.a.base + v * op.a.stride + op.a.off);
    const uint32_t *gy = reinterpret_cast<const uint32_t *>(op.b.base + v * op.b.stride + op.b.off);
    uint32_t nx = 2 * op.a.w, ny = 2 * op.b.w;
    if (nx > ny) { // chunks of the shorter operand are spread over blockIdx.z
        const uint32_t *tp = gx; gx = gy; gy = tp;
        const uint32_t tn = nx; nx = ny; ny = tn;
    }
    if (24 * xi >= nx) return;
    const uint64_t slot = ((uint64_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    uint32_t *mbuf = scratch + slot * ADT_THREAD_WORDS, *tmp = mbuf + 32;
    for (uint32_t i = 0; i < 24; ++i) mbuf[i] = (24 * xi + i < nx) ? gx[24 * xi + i] : 0u;
    uint32_t *go = reinterpret_cast<uint32_t *>(op.o.base + v * op.o.stride + op.o.off);
    const uint32_t no = 2 * op.o.w;
    const uint32_t nyc = (ny + 23) / 24;
    uint32_t t[48];
#pragma unroll
    for (int i = 0; i < 48; ++i) t[i] = 0;
    for (uint32_t j = 0; j <= nyc; ++j) {
        if (j < nyc) {
            const uint32_t *cp = gy + 24 * j;
            if (24 * j + 24 > ny) {
                for (uint32_t i = 0; i < 24; ++i) tmp[i] = (24 * j + i < ny) ? gy[24 * j + i] : 0u;
                cp = tmp;
            }
            mul24_acc(mbuf, cp, t);
        }
        const uint32_t w0 = 24 * (xi + j);
#pragma unroll
        for (int i = 0; i < 24; ++i)
            if (t[i] && w0 + i < no) atomicXor(go + w0 + i, t[i]);
#pragma unroll
        for (int i = 0; i < 24; ++i) {
            t[i] = t[24 + i];
            t[24 + i] = 0;
        }
    }
}

// t ^= m * c for 32-word m and c: two Karatsuba levels over 8-word blocks = nine 8x8-word products (mul24_acc's
// 3-way split needs six for 24 words: 96 word pairs per product there, 114 here).  Rolled like mul24_acc: one copy
// of the 8x8-word code; leaf i = 3 * top + sub with top/sub in {low, high, middle}.
template <int O> __device__ __forceinline__ void xor16_at64(uint32_t (&t)[64], const uint32_t (&r)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) t[O + i] ^= r[i];
}
__device__ __forceinline__ void mul32_acc(const uint32_t *__restrict__ m, const uint32_t *__restrict__ c, uint32_t (&t)[64]) {
#pragma unroll 1
    for (int i = 0; i < 9; ++i) {
        // blocks XORed into the leaf operand: bit b = 8-word block b.  top: ll {0,1}, hh {2,3}, mid {0^2, 1^3}
        const uint32_t sel = (uint32_t)(0xFA5C84321ull >> (4 * i)) & 0xFu; // leaves 0..8: 1, 2, 3, 4, 8, C, 5, A, F
        uint32_t x[8], y[8], r[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) x[q] = y[q] = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (sel >> b & 1) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint2 u = *reinterpret_cast<const uint2 *>(m + 8 * b + 2 * q);
                    const uint2 w = *reinterpret_cast<const uint2 *>(c + 8 * b + 2 * q);
                    x[2 * q] ^= u.x; x[2 * q + 1] ^= u.y;
                    y[2 * q] ^= w.x; y[2 * q + 1] ^= w.y;
                }
            }
        }
        clmul_kara<8>(x, y, r);
        switch (i) { // destination offsets = {top bases} + {sub offsets}: ll {0,16}, hh {32,16}, mid {16}; lo {0,8}, hi {16,8}, mid {8}
            case 0: xor16_at64<0>(t, r); xor16_at64<8>(t, r); xor16_at64<16>(t, r); xor16_at64<24>(t, r); break;
            case 1: xor16_at64<16>(t, r); xor16_at64<8>(t, r); xor16_at64<32>(t, r); xor16_at64<24>(t, r); break;
            case 2: xor16_at64<8>(t, r); xor16_at64<24>(t, r); break;
            case 3: xor16_at64<32>(t, r); xor16_at64<40>(t, r); xor16_at64<16>(t, r); xor16_at64<24>(t, r); break;
            case 4: xor16_at64<48>(t, r); xor16_at64<40>(t, r); xor16_at64<32>(t, r); xor16_at64<24>(t, r); break;
            case 5: xor16_at64<40>(t, r); xor16_at64<24>(t, r); break;
            case 6: xor16_at64<16>(t, r); xor16_at64<24>(t, r); break;
            case 7: xor16_at64<32>(t, r); xor16_at64<24>(t, r); break;
            default: xor16_at64<24>(t, r); break;
        }
    }
}

// K4d  same thread-per-(value, product, chunk) scheme with 32-word chunks (mul32_acc).  An operand whose degree bound is
// exactly 64 (w - 1) — every carry of the multiplier circuit: bounds are sums of multiples of 256 — is taken as its
// w - 1 low words plus the single coefficient of X^(64 (w-1)), so that 2048-bit operands are 2 chunks and not 3;
// the coefficient's terms are word-aligned shifted XORs.
struct ThreadMulShape {
    uint32_t nx, ny; // 32-bit words of the low parts
    bool xt, yt;     // operand has the separate top coefficient
    bool swapped;    // x is op.b
};
__host__ __device__ inline ThreadMulShape thread_mul_shape(const MulOp &op) {
    ThreadMulShape s;
    s.nx = 2 * op.a.w;
    s.ny = 2 * op.b.w;
    s.xt = op.a.w >= 2 && op.a.deg == (uint64_t)64 * (op.a.w - 1);
    s.yt = op.b.w >= 2 && op.b.deg == (uint64_t)64 * (op.b.w - 1);
    if (s.xt) s.nx -= 2;
    if (s.yt) s.ny -= 2;
    s.swapped = s.nx > s.ny;
    if (s.swapped) {
        const uint32_t tn = s.nx; s.nx = s.ny; s.ny = tn;
        const bool tt = s.xt; s.xt = s.yt; s.yt = tt;
    }
    return s;
}

static __global__ void __launch_bounds__(128, 4) mul_thread32_kernel(const MulOp *__restrict__ ops, uint64_t n, uint32_t *__restrict__ scratch) {
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    // CTAs are dispatched x-fastest, z-slowest: walk products and chunks backwards so that the long threads (high chunks
    // exist only for the big products, which come last in a column) start first and the short ones fill the tail
    const MulOp op = ops[gridDim.y - 1 - blockIdx.y];
    const uint32_t xi = gridDim.z - 1 - blockIdx.z;
    const ThreadMulShape sh = thread_mul_shape(op);
    if (32 * xi >= sh.nx) return;
    const View &vx = sh.swapped ? op.b : op.a, &vy = sh.swapped ? op.a : op.b;
    const uint32_t *gx = reinterpret_cast<const uint32_t *>(vx.base + v * vx.stride + vx.off);
    const uint32_t *gy = reinterpret_cast<const uint32_t *>(vy.base + v * vy.stride + vy.off);
    const uint32_t nx = sh.nx, ny = sh.ny;
    const uint32_t xtop = sh.xt ? (gx[nx] & 1u) : 0u, ytop = sh.yt ? (gy[ny] & 1u) : 0u;
    const uint64_t slot = ((uint64_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    uint32_t *mbuf = scratch + slot * ADT_THREAD_WORDS, *tmp = mbuf + 32;
    for (uint32_t i = 0; i < 32; ++i) mbuf[i] = (32 * xi + i < nx) ? gx[32 * xi + i] : 0u;
    uint32_t *go = reinterpret_cast<uint32_t *>(op.o.base + v * op.o.stride + op.o.off);
    const uint32_t no = 2 * op.o.w;
    const uint32_t nyc = (ny + 31) / 32;
    uint32_t t[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) t[i] = 0;
    for (uint32_t j = 0; j <= nyc; ++j) {
        if (j < nyc) {
            const uint32_t *cp = gy + 32 * j;
            if (32 * j + 32 > ny) {
                for (uint32_t i = 0; i < 32; ++i) tmp[i] = (32 * j + i < ny) ? gy[32 * j + i] : 0u;
                cp = tmp;
            }
            mul32_acc(mbuf, cp, t);
            if (xi == 0 && xtop) { // X^(32 nx) * y: chunk j of y lands nx words up
                for (uint32_t i = 0; i < 32; ++i) {
                    const uint32_t w = cp[i];
                    if (w && nx + 32 * j + i < no) atomicXor(go + nx + 32 * j + i, w);
                }
            }
        }
        const uint32_t w0 = 32 * (xi + j);
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (t[i] && w0 + i < no) atomicXor(go + w0 + i, t[i]);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            t[i] = t[32 + i];
            t[32 + i] = 0;
        }
    }
    if (ytop) { // X^(32 ny) * (this chunk of x)
        for (uint32_t i = 0; i < 32; ++i) {
            const uint32_t w = mbuf[i];
            if (w && ny + 32 * xi + i < no) atomicXor(go + ny + 32 * xi + i, w);
        }
    }
    if (xi == 0 && xtop && ytop && nx + ny < no) atomicXor(go + nx + ny, 1u);
}

// ----------------------------------------------------------------------------------------
// K5b  fused (a*b) mod S for fresh pairs at D = 1024 (32 words + the X^1024 coefficient), d = 32*WS — config B.
// The fully unrolled 32-word Karatsuba of mulrem_fresh_kernel<32,...> is 243 inlined leaves (instruction-fetch bound,
// 255 registers, spills); here the product is the rolled mul32_acc (nine 8x8-word products, one code copy) with both
// operands in per-thread shared-memory columns ([word pair][thread], fetched by cp.async), the 2049-bit product goes back
// to the same columns and is folded down by a rolled sliding-window loop (16-word state in registers, words read
// top-down from shared memory, 4 x 256-entry tables).  One thread per pair, 512 threads per CTA, one CTA per SM.
// ----------------------------------------------------------------------------------------
template <int STRIDE>
__device__ __forceinline__ void mul32_acc_ss(const uint2 *__restrict__ m, const uint2 *__restrict__ c, uint32_t (&t)[64]) {
#pragma unroll 1
    for (int i = 0; i < 9; ++i) {
        const uint32_t sel = (uint32_t)(0xFA5C84321ull >> (4 * i)) & 0xFu; // as in mul32_acc
        uint32_t x[8], y[8], r[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) x[q] = y[q] = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (sel >> b & 1) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint2 u = m[(4 * b + q) * STRIDE];
                    const uint2 w = c[(4 * b + q) * STRIDE];
                    x[2 * q] ^= u.x; x[2 * q + 1] ^= u.y;
                    y[2 * q] ^= w.x; y[2 * q + 1] ^= w.y;
                }
            }
        }
        clmul_kara<8>(x, y, r);
        switch (i) {
            case 0: xor16_at64<0>(t, r); xor16_at64<8>(t, r); xor16_at64<16>(t, r); xor16_at64<24>(t, r); break;
            case 1: xor16_at64<16>(t, r); xor16_at64<8>(t, r); xor16_at64<32>(t, r); xor16_at64<24>(t, r); break;
            case 2: xor16_at64<8>(t, r); xor16_at64<24>(t, r); break;
            case 3: xor16_at64<32>(t, r); xor16_at64<40>(t, r); xor16_at64<16>(t, r); xor16_at64<24>(t, r); break;
            case 4: xor16_at64<48>(t, r); xor16_at64<40>(t, r); xor16_at64<32>(t, r); xor16_at64<24>(t, r); break;
            case 5: xor16_at64<40>(t, r); xor16_at64<24>(t, r); break;
            case 6: xor16_at64<16>(t, r); xor16_at64<24>(t, r); break;
            case 7: xor16_at64<32>(t, r); xor16_at64<24>(t, r); break;
            default: xor16_at64<24>(t, r); break;
        }
    }
}

constexpr int MR32_THREADS = 512;
constexpr int MR32_PAIRS = 34; // per thread: a = pairs 0..16 (16 = the top coefficient's word), b = pairs 17..33
template <int WS> __host__ __device__ constexpr int mulrem32_row_stride() { return WS + 4; } // words: 5 bank groups of 16 B for WS = 16
template <int WS> constexpr size_t mulrem32_smem_bytes() { return (size_t)4 * 256 * mulrem32_row_stride<WS>() * 4 + (size_t)MR32_PAIRS * MR32_THREADS * 8; }

template <int WS>
__global__ void __launch_bounds__(MR32_THREADS, 1) mulrem_fresh32_kernel(const uint64_t *__restrict__ A, const uint64_t *__restrict__ B,
                                                                      uint64_t *__restrict__ O, uint64_t n,
                                                                      const uint32_t *__restrict__ Tg) {
    constexpr int WD = 32, WF = WD / 2 + 1, TH = MR32_THREADS;
    extern __shared__ __align__(16) uint32_t smem32[];
    constexpr int RS = mulrem32_row_stride<WS>();
    uint32_t *T = smem32;
    uint2 *col = reinterpret_cast<uint2 *>(smem32 + 4 * 256 * RS) + threadIdx.x; // pair q of this thread at col[q * TH]
    for (uint32_t i = threadIdx.x; i < 4u * 256u * WS; i += TH) T[(i / WS) * RS + (i % WS)] = Tg[i];
    __syncthreads();
    const uint32_t scol = (uint32_t)__cvta_generic_to_shared(col);
    for (uint64_t u = (uint64_t)blockIdx.x * TH + threadIdx.x; u < n; u += (uint64_t)gridDim.x * TH) {
        const uint64_t *ga = A + u * WF, *gb = B + u * WF;
#pragma unroll
        for (int q = 0; q < WF; ++q) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(scol + q * TH * 8), "l"(ga + q) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(scol + (WF + q) * TH * 8), "l"(gb + q) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        uint32_t t[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) t[i] = 0;
        mul32_acc_ss<TH>(col, col + WF * TH, t);
        // the X^1024 coefficients: (a' + at X^1024)(b' + bt X^1024) = a'b' + X^1024 (at b' + bt a') + at bt X^2048
        const uint32_t atop = col[16 * TH].x & 1u, btop = col[(WF + 16) * TH].x & 1u;
        const uint32_t ma = 0u - atop, mb = 0u - btop;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const uint2 x = col[q * TH], y = col[(WF + q) * TH];
            t[32 + 2 * q] ^= (y.x & ma) ^ (x.x & mb);
            t[32 + 2 * q + 1] ^= (y.y & ma) ^ (x.y & mb);
        }
        // product words 0..63 back into the columns (the operands are dead now); word 64 = at & bt stays in a register
#pragma unroll
        for (int q = 0; q < 32; ++q) col[q * TH] = make_uint2(t[2 * q], t[2 * q + 1]);
        const uint32_t *pw = reinterpret_cast<const uint32_t *>(col); // word i at pw[(i / 2) * TH * 2 + (i & 1)]
        // sliding-window fold: r = product words [i - WS, i), top = word i, for i = 64 down to WS
        uint32_t r[WS];
#pragma unroll
        for (int q = 0; q < WS; ++q) r[q] = t[64 - WS + q];
        uint32_t top = atop & btop;
        if (top) fold_word<WS, 1, RS>(r, top, T);
#pragma unroll 16
        for (int i = 63; i >= WS; --i) {
            top = r[WS - 1];
#pragma unroll
            for (int q = WS - 1; q > 0; --q) r[q] = r[q - 1];
            const int w = i - WS;
            r[0] = pw[(w >> 1) * TH * 2 + (w & 1)];
            if (top) fold_word<WS, 1, RS>(r, top, T);
        }
        uint32_t *out = reinterpret_cast<uint32_t *>(O + u * (WS / 2));
#pragma unroll
        for (int q = 0; q < WS / 4; ++q) reinterpret_cast<uint4 *>(out)[q] = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
    }
}

// K5d  config B fused mul+rem with the operands reduced first: (a b) mod S = ((a mod S)(b mod S)) mod S, so the product is
// 16 x 16 words (three 8x8-word Karatsubas) instead of 32 x 32 (nine).  Same shared-memory columns and sliding-window fold
// as mulrem_fresh32_kernel; the fold code exists once and runs three times (a, b, the product) from a phase loop.
template <int STRIDE>
__device__ __forceinline__ void mul16_acc_ss(const uint2 *__restrict__ m, const uint2 *__restrict__ c, uint32_t (&t)[32]) {
#pragma unroll 1
    for (int i = 0; i < 3; ++i) { // 0: low blocks, 1: high blocks, 2: (low + high) blocks
        uint32_t x[8], y[8], r[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint2 u = make_uint2(0, 0), w = make_uint2(0, 0);
            if (i != 1) {
                const uint2 u0 = m[q * STRIDE], w0 = c[q * STRIDE];
                u.x ^= u0.x; u.y ^= u0.y; w.x ^= w0.x; w.y ^= w0.y;
            }
            if (i != 0) {
                const uint2 u1 = m[(4 + q) * STRIDE], w1 = c[(4 + q) * STRIDE];
                u.x ^= u1.x; u.y ^= u1.y; w.x ^= w1.x; w.y ^= w1.y;
            }
            x[2 * q] = u.x; x[2 * q + 1] = u.y;
            y[2 * q] = w.x; y[2 * q + 1] = w.y;
        }
        clmul_kara<8>(x, y, r);
        if (i == 0) {
#pragma unroll
            for (int q = 0; q < 16; ++q) { t[q] ^= r[q]; t[8 + q] ^= r[q]; }
        } else if (i == 1) {
#pragma unroll
            for (int q = 0; q < 16; ++q) { t[16 + q] ^= r[q]; t[8 + q] ^= r[q]; }
        } else {
#pragma unroll
            for (int q = 0; q < 16; ++q) t[8 + q] ^= r[q];
        }
    }
}

template <int WS>
__global__ void __launch_bounds__(MR32_THREADS, 1) mulrem_fresh32r_kernel(const uint64_t *__restrict__ A, const uint64_t *__restrict__ B,
                                                                       uint64_t *__restrict__ O, uint64_t n,
                                                                       const uint32_t *__restrict__ Tg) {
    static_assert(WS == 16, "the product stage is written for d = 512");
    constexpr int WD = 32, WF = WD / 2 + 1, TH = MR32_THREADS;
    constexpr int RS = mulrem32_row_stride<WS>();
    extern __shared__ __align__(16) uint32_t smem32[];
    uint32_t *T = smem32;
    uint2 *col = reinterpret_cast<uint2 *>(smem32 + 4 * 256 * RS) + threadIdx.x; // pair q of this thread at col[q * TH]
    for (uint32_t i = threadIdx.x; i < 4u * 256u * WS; i += TH) T[(i / WS) * RS + (i % WS)] = Tg[i];
    __syncthreads();
    const uint32_t scol = (uint32_t)__cvta_generic_to_shared(col);
    for (uint64_t u = (uint64_t)blockIdx.x * TH + threadIdx.x; u < n; u += (uint64_t)gridDim.x * TH) {
        const uint64_t *ga = A + u * WF, *gb = B + u * WF;
#pragma unroll
        for (int q = 0; q < WF; ++q) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(scol + q * TH * 8), "l"(ga + q) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(scol + (WF + q) * TH * 8), "l"(gb + q) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        uint32_t r[WS];
#pragma unroll 1
        for (int ph = 0; ph < 3; ++ph) {
            // phase 0: a (33 words from pair 0), 1: b (33 words from pair 17), 2: the product (32 words from pair 0)
            const uint32_t *pw = reinterpret_cast<const uint32_t *>(col + (ph == 1 ? WF * TH : 0)); // word i at pw[(i / 2) * TH * 2 + (i & 1)]
            const int N = ph == 2 ? 2 * WS : WD + 1;
#pragma unroll
            for (int q = 0; q < WS; ++q) {
                const int w = N - 1 - WS + q;
                r[q] = pw[(w >> 1) * TH * 2 + (w & 1)];
            }
            uint32_t top = pw[((N - 1) >> 1) * TH * 2 + ((N - 1) & 1)];
            if (ph != 2) top &= 1u; // the X^1024 coefficient
            if (top) fold_word<WS, 1, RS>(r, top, T);
#pragma unroll 16
            for (int i = N - 2; i >= WS; --i) {
                top = r[WS - 1];
#pragma unroll
                for (int q = WS - 1; q > 0; --q) r[q] = r[q - 1];
                const int w = i - WS;
                r[0] = pw[(w >> 1) * TH * 2 + (w & 1)];
                if (top) fold_word<WS, 1, RS>(r, top, T);
            }
            if (ph == 2) break;
            uint2 *dst = col + (ph == 1 ? WF * TH : 0); // the reduced operand replaces the low words of the operand
#pragma unroll
            for (int q = 0; q < WS / 2; ++q) dst[q * TH] = make_uint2(r[2 * q], r[2 * q + 1]);
            if (ph == 1) {
                uint32_t t[2 * WS];
#pragma unroll
                for (int i = 0; i < 2 * WS; ++i) t[i] = 0;
                mul16_acc_ss<TH>(col, col + WF * TH, t);
#pragma unroll
                for (int q = 0; q < WS; ++q) col[q * TH] = make_uint2(t[2 * q], t[2 * q + 1]);
            }
        }
        uint32_t *out = reinterpret_cast<uint32_t *>(O + u * (WS / 2));
#pragma unroll
        for (int q = 0; q < WS / 4; ++q) reinterpret_cast<uint4 *>(out)[q] = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
    }
}

// K5e  mulrem_fresh32r_kernel with a conflict-free fold table (ncu on K5d: shared-memory wavefronts 88 % busy, 54 % of them
// bank conflicts).  Same rotation as config A: lane l does the four byte lookups of a fold in the order k = (q + l) % 4 and
// chunk j of row (byte e, table k, copy c) sits at (((e * 4 + j) * 4 + k) * 2 + c) * 16 bytes, bank group 2 k + c for every
// e and j; c = bit 2 of the lane.  The table is 128 KB, which fits because the operands are now loaded one after the other
// into ONE set of 17 column pairs (a is reduced to 16 registers before b arrives).
constexpr int MR32Q_PAIRS = 17;
constexpr size_t MR32Q_SMEM_BYTES = (size_t)256 * 4 * 4 * 2 * 16 + (size_t)MR32Q_PAIRS * MR32_THREADS * 8;

__device__ __forceinline__ void fold_word_rot16(uint32_t (&dst)[16], uint32_t t, const FoldRot &f) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t x[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t e;
            asm("prmt.b32 %0, %1, %2, %3;" : "=r"(e) : "r"(t), "r"(0u), "r"(f.sel[q]));
            const uint32_t addr = e * 512u + f.off[q] + j * 128;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x[q][0]), "=r"(x[q][1]), "=r"(x[q][2]), "=r"(x[q][3]) : "r"(addr));
        }
#pragma unroll
        for (int w = 0; w < 4; ++w) dst[4 * j + w] ^= x[0][w] ^ x[1][w] ^ x[2][w] ^ x[3][w];
    }
}

static __global__ void __launch_bounds__(MR32_THREADS, 1) mulrem_fresh32q_kernel(const uint64_t *__restrict__ A, const uint64_t *__restrict__ B,
                                                                              uint64_t *__restrict__ O, uint64_t n,
                                                                              const uint32_t *__restrict__ Tg) {
    constexpr int WS = 16, WD = 32, WF = WD / 2 + 1, TH = MR32_THREADS;
    extern __shared__ __align__(16) uint32_t smem32[];
    uint32_t *T2 = smem32; // 256 * 4 * 4 * 2 rows of 4 words
    uint2 *col = reinterpret_cast<uint2 *>(smem32 + 256 * 4 * 4 * 2 * 4) + threadIdx.x; // pair q of this thread at col[q * TH]
    for (uint32_t i = threadIdx.x; i < 256u * 4u * 4u * 2u * 4u; i += TH) {
        const uint32_t w = i & 3, k = (i >> 3) & 3, j = (i >> 5) & 3, e = i >> 7; // bit 2 = copy
        T2[i] = Tg[(k * 256 + e) * WS + 4 * j + w];
    }
    __syncthreads();
    const FoldRot fr = fold_rot_init(T2, threadIdx.x & 31);
    const uint32_t krow = (uint32_t)__cvta_generic_to_shared(T2) + 512; // row (byte 1, chunk 0, table 0, copy 0); chunk j is 128 j further
    const uint32_t scol = (uint32_t)__cvta_generic_to_shared(col);
    for (uint64_t u = (uint64_t)blockIdx.x * TH + threadIdx.x; u < n; u += (uint64_t)gridDim.x * TH) {
        uint32_t r[WS], ra[WS];
#pragma unroll 1
        for (int ph = 0; ph < 3; ++ph) {
            // phase 0: a (33 words), 1: b (33 words), 2: the product (32 words); all of them in column pairs 0..16
            if (ph < 2) {
                const uint64_t *g = (ph == 0 ? A : B) + u * WF;
#pragma unroll
                for (int q = 0; q < WF; ++q)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(scol + q * TH * 8), "l"(g + q) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            const uint32_t *pw = reinterpret_cast<const uint32_t *>(col); // word i at pw[(i / 2) * TH * 2 + (i & 1)]
            const int N = ph == 2 ? 2 * WS : WD + 1;
#pragma unroll
            for (int q = 0; q < WS; ++q) {
                const int w = N - 1 - WS + q;
                r[q] = pw[(w >> 1) * TH * 2 + (w & 1)];
            }
            uint32_t top = pw[((N - 1) >> 1) * TH * 2 + ((N - 1) & 1)];
            if (ph != 2) { // the X^1024 coefficient: one row (byte 1 of table 0 = X^1024 mod S, same address in every lane) under a mask
                const uint32_t mask = 0u - (top & 1u);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    uint32_t k0, k1, k2, k3;
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(k0), "=r"(k1), "=r"(k2), "=r"(k3) : "r"(krow + jj * 128));
                    r[4 * jj] ^= k0 & mask; r[4 * jj + 1] ^= k1 & mask; r[4 * jj + 2] ^= k2 & mask; r[4 * jj + 3] ^= k3 & mask;
                }
            } else {
                fold_word_rot16(r, top, fr);
            }
#pragma unroll 16
            for (int i = N - 2; i >= WS; --i) {
                top = r[WS - 1];
#pragma unroll
                for (int q = WS - 1; q > 0; --q) r[q] = r[q - 1];
                const int w = i - WS;
                r[0] = pw[(w >> 1) * TH * 2 + (w & 1)];
                fold_word_rot16(r, top, fr);
            }
            if (ph == 0) {
#pragma unroll
                for (int q = 0; q < WS; ++q) ra[q] = r[q];
            } else if (ph == 1) {
#pragma unroll
                for (int q = 0; q < WS / 2; ++q) {
                    col[q * TH] = make_uint2(ra[2 * q], ra[2 * q + 1]);
                    col[(8 + q) * TH] = make_uint2(r[2 * q], r[2 * q + 1]);
                }
                uint32_t t[2 * WS];
#pragma unroll
                for (int i = 0; i < 2 * WS; ++i) t[i] = 0;
                mul16_acc_ss<TH>(col, col + 8 * TH, t);
#pragma unroll
                for (int q = 0; q < WS; ++q) col[q * TH] = make_uint2(t[2 * q], t[2 * q + 1]);
            }
        }
        uint32_t *out = reinterpret_cast<uint32_t *>(O + u * (WS / 2));
#pragma unroll
        for (int q = 0; q < WS / 4; ++q) reinterpret_cast<uint4 *>(out)[q] = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
    }
}

// ----------------------------------------------------------------------------------------
// LOP3 issue-rate probe: the measured denominator of the integer-logic roofline.
// 8 independent dependency chains per thread, 8 warps per CTA, 8 CTAs per SM.
// ----------------------------------------------------------------------------------------
constexpr int PEAK_ILP = 8;
static __global__ void __launch_bounds__(256) lop3_peak_kernel(uint32_t *sink, unsigned long long *clk, int iters, uint32_t seed) {
    uint32_t x[PEAK_ILP], y[PEAK_ILP];
#pragma unroll
    for (int i = 0; i < PEAK_ILP; ++i) {
        x[i] = seed + threadIdx.x * 7u + i;
        y[i] = seed * 3u + i + blockIdx.x;
    }
    const uint32_t c = seed | 1u;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < PEAK_ILP; ++i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(c));
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < PEAK_ILP; ++i) acc ^= x[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = (unsigned long long)(t1 - t0);
}

// 8x8-word Karatsuba product rate probe: the denominator of the roofline of the thread-per-value adder and of the fused
// mul+rem, both of which are chains of these products (bound by the FMA-heavy pipe: 432 IMAD.WIDE per product).
static __global__ void __launch_bounds__(128, 4) kara8_peak_kernel(uint32_t *sink, int iters, uint32_t seed) {
    uint32_t a[8], b[8], r[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = seed * (threadIdx.x + 1) + i;
        b[i] = seed ^ (blockIdx.x * 977u + i);
    }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        clmul_kara<8>(a, b, r);
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

} // namespace hmk
