// gf2host.hpp — small host-side GF(2)[X] helpers used ONLY for key-derived constants
// (decrypt vector, remainder folding tables, encrypt subset tables).  They run once per
// key on the host; all batch arithmetic is done by the CUDA kernels in kernels.cuh.
//
// Bit order is the reference's: coefficient of X^i is bit i%64 of word i/64
// (reference src/polynomial.rs:144,172).
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace gf2 {

using words = std::vector<uint64_t>;

// Highest set bit of a zero-padded word run; 0 for the null polynomial
// (reference src/polynomial.rs:35-42).
inline size_t degree(const uint64_t *w, size_t n) {
    for (size_t i = n; i-- > 0;)
        if (w[i]) return 64 * i + 63 - (size_t)__builtin_clzll(w[i]);
    return 0;
}
inline bool is_zero(const uint64_t *w, size_t n) {
    for (size_t i = 0; i < n; i++)
        if (w[i]) return false;
    return true;
}
inline bool get_bit(const words &w, size_t i) { return i / 64 < w.size() && ((w[i / 64] >> (i % 64)) & 1); }
inline void flip_bit(words &w, size_t i) { w[i / 64] ^= (uint64_t)1 << (i % 64); }

// r = (r * X) mod S for deg r < d = deg S.  r has ceil(d/64)+? words; S has d/64+1 words.
inline void mulx_mod(words &r, const words &S, size_t d) {
    uint64_t carry = 0;
    for (size_t i = 0; i < r.size(); i++) {
        uint64_t nc = r[i] >> 63;
        r[i] = (r[i] << 1) | carry;
        carry = nc;
    }
    if (get_bit(r, d))
        for (size_t i = 0; i < r.size() && i < S.size(); i++) r[i] ^= S[i];
}

// v[k] = (X^k mod S)(0) for k < nbits, packed LSB first.  (C mod S)(0) == parity(popcount(C & v)):
// the remainder map is linear, so reference src/cipher.rs:119-122 (rem, then evaluate(false))
// collapses to one inner product per ciphertext bit.
inline words decrypt_vector(const words &S, size_t d, size_t nbits) {
    words v((nbits + 63) / 64 + 1, 0);
    words r(d / 64 + 1, 0);
    r[0] = 1; // X^0
    for (size_t k = 0; k < nbits; k++) {
        if (r[0] & 1) v[k / 64] |= (uint64_t)1 << (k % 64);
        mulx_mod(r, S, d);
    }
    return v;
}

// Folding tables for remainder by S when d % 32 == 0 and d >= 32:
//   T[b][x] = (x(X) * X^(8b) * X^d) mod S,  b in 0..3, x in 0..255, each d/32 u32 words.
// A 32-bit word t sitting at bit offset 32*i (i >= d/32) is congruent to
// X^(32*(i - d/32)) * (T[0][t0] ^ T[1][t1] ^ T[2][t2] ^ T[3][t3]).
inline std::vector<uint32_t> rem_fold_tables(const words &S, size_t d) {
    const size_t wd = d / 32;
    std::vector<uint32_t> T(4 * 256 * wd, 0);
    // basis[j] = X^(d+j) mod S for j in 0..31
    std::vector<words> basis(32);
    words r(d / 64 + 1, 0);
    // X^d mod S = S without its leading term
    for (size_t i = 0; i < r.size() && i < S.size(); i++) r[i] = S[i];
    flip_bit(r, d);
    for (int j = 0; j < 32; j++) {
        basis[j] = r;
        mulx_mod(r, S, d);
    }
    for (int b = 0; b < 4; b++)
        for (int x = 0; x < 256; x++) {
            uint32_t *row = &T[((size_t)b * 256 + x) * wd];
            for (int bit = 0; bit < 8; bit++)
                if ((x >> bit) & 1) {
                    const words &p = basis[8 * b + bit];
                    for (size_t k = 0; k < wd; k++) row[k] ^= (uint32_t)(p[k / 2] >> (32 * (k % 2)));
                }
        }
    return T;
}

} // namespace gf2
