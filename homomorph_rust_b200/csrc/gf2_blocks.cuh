// gf2_blocks.cuh — device building blocks shared by every translation unit of the engine: the batch layout structs and
// the carry-less products on the integer multiplier (clmul_imad / clmul32_imad / clmul_kara).  Split out of kernels.cuh so
// that kernels living in their own .cu files (kernels_adder.cu) do not recompile every kernel of the header.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hmk {

constexpr int MAX_SLOTS = 128;
constexpr unsigned FULL = 0xffffffffu;

struct Layout { // slot layout of a batch, in u64 words
    uint32_t L;
    uint32_t value_words;
    uint32_t off[MAX_SLOTS + 1];
};

struct View { // one slot of every value of some buffer
    uint64_t *base;
    uint64_t stride; // u64 words between consecutive values
    uint32_t off;    // u64 word offset of the slot inside a value
    uint32_t w;      // slot width in u64 words
    uint64_t deg;    // degree bound of the polynomials in this slot (host bookkeeping; picks the kernel)
};

struct MulOp {
    View a, b, o;
};

// ----------------------------------------------------------------------------------------
// register-resident carry-less product on the integer MULTIPLIER (FMA pipe), NA x NB 32-bit words.
//
// Spread trick: keep only every 4th bit of each operand word (class c = bit index mod 4).  The
// integer product of two such words has, in every 4-bit group, the NUMBER of coefficient pairs that
// land there (at most 8 < 16, so groups never carry into each other); its low bit is their XOR, i.e.
// the carry-less product restricted to class (c + c') mod 4.  Products of one output class are
// XOR-accumulated unmasked (XOR does not carry either) and masked once at the end.  A 32x32 -> 64
// clmul is 16 IMAD.WIDE + 16 three-input LOP3, against ~48 ALU instructions for the shift/mask
// schoolbook above — and the multiplies issue on the otherwise idle FMA pipe (tools/ubench.cu).
// ----------------------------------------------------------------------------------------
template <int NA, int NB>
__device__ __forceinline__ void clmul_imad(const uint32_t (&a)[NA], const uint32_t (&b)[NB], uint32_t (&r)[NA + NB]) {
    uint32_t as[4][NA], bs[4][NB];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int j = 0; j < NA; ++j) as[c][j] = a[j] & (0x11111111u << c);
#pragma unroll
        for (int k = 0; k < NB; ++k) bs[c][k] = b[k] & (0x11111111u << c);
    }
#pragma unroll
    for (int i = 0; i < NA + NB; ++i) r[i] = 0;
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
        uint32_t acc[NA + NB];
#pragma unroll
        for (int i = 0; i < NA + NB; ++i) acc[i] = 0;
#pragma unroll
        for (int j = 0; j < NA; ++j) {
#pragma unroll
            for (int k = 0; k < NB; ++k) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const unsigned long long p = (unsigned long long)as[c][j] * (unsigned long long)bs[(kc - c) & 3][k];
                    acc[j + k] ^= (uint32_t)p;
                    acc[j + k + 1] ^= (uint32_t)(p >> 32);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < NA + NB; ++i) r[i] |= acc[i] & (0x11111111u << kc);
    }
}

// 32 x 32 -> 64 carry-less product on the multiplier (see clmul_imad): 16 IMAD.WIDE + 20 LOP3.
__device__ __forceinline__ void clmul32_imad(uint32_t a, uint32_t b, uint32_t &lo, uint32_t &hi) {
    uint32_t as[4], bs[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        as[c] = a & (0x11111111u << c);
        bs[c] = b & (0x11111111u << c);
    }
    lo = 0;
    hi = 0;
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
        const unsigned long long p0 = (unsigned long long)as[0] * bs[kc & 3];
        const unsigned long long p1 = (unsigned long long)as[1] * bs[(kc - 1) & 3];
        const unsigned long long p2 = (unsigned long long)as[2] * bs[(kc - 2) & 3];
        const unsigned long long p3 = (unsigned long long)as[3] * bs[(kc - 3) & 3];
        const uint32_t m = 0x11111111u << kc;
        lo |= ((uint32_t)p0 ^ (uint32_t)p1 ^ (uint32_t)p2 ^ (uint32_t)p3) & m;
        hi |= ((uint32_t)(p0 >> 32) ^ (uint32_t)(p1 >> 32) ^ (uint32_t)(p2 >> 32) ^ (uint32_t)(p3 >> 32)) & m;
    }
}

// Karatsuba over 32-bit words down to single-word leaves on the multiplier: N = 8 needs 27 leaf products
// (432 IMAD.WIDE on the FMA pipe + ~900 LOP3 on the ALU pipe, the two pipes issue concurrently) instead of
// ~3 000 ALU instructions for the shift/mask schoolbook.
template <int N>
__device__ __forceinline__ void clmul_kara(const uint32_t (&a)[N], const uint32_t (&b)[N], uint32_t (&r)[2 * N]) {
    if constexpr (N == 1) {
        clmul32_imad(a[0], b[0], r[0], r[1]);
    } else {
        constexpr int H = N / 2;
        static_assert(N % 2 == 0, "power-of-two word counts only");
        uint32_t a0[H], a1[H], b0[H], b1[H], sa[H], sb[H];
#pragma unroll
        for (int i = 0; i < H; ++i) {
            a0[i] = a[i]; a1[i] = a[H + i]; b0[i] = b[i]; b1[i] = b[H + i];
            sa[i] = a0[i] ^ a1[i];
            sb[i] = b0[i] ^ b1[i];
        }
        uint32_t p0[N], p1[N], p2[N];
        clmul_kara<H>(a0, b0, p0);
        clmul_kara<H>(a1, b1, p2);
        clmul_kara<H>(sa, sb, p1);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            r[i] = p0[i];
            r[N + i] = p2[i];
        }
#pragma unroll
        for (int i = 0; i < N; ++i) r[H + i] ^= p1[i] ^ p0[i] ^ p2[i];
    }
}

} // namespace hmk
