// kernels_enc_umma.h — host interface of kernels_enc_umma.cu: config-B encryption (D = 1024, tau = 256) as a tcgen05 kind::i8 GEMM.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace hmk {

constexpr int ENC_UMMA_PASS_COLS = 512; // key bits (columns of B) resident at a time: two passes over a CTA's tiles

struct EncUmmaParams {
    const uint8_t *values; // n * L / 8 bytes of plaintext
    const uint8_t *masks;  // units * 32 bytes (unused when seeded)
    uint64_t *out;         // units * 17 words
    uint32_t units;
    uint32_t chunk_tiles;  // tiles per CTA between two swaps of the B half (0 = all of the CTA's tiles per pass)
    uint32_t topmask[8];   // bit i = coefficient of X^1024 of T_i
    uint64_t seed, first_unit;
};

size_t enc_umma_table_bytes();
// T[k] = words of public-key polynomial k (k < 256), T_words[k] its length; out = enc_umma_table_bytes() bytes
void enc_umma_build_table(const uint64_t *const *T, const size_t *T_words, int8_t *out);
// one launch; masks drawn from Philox4x32-10 inside the kernel when seeded
cudaError_t launch_encrypt_umma_b(const EncUmmaParams &p, bool seeded, const int8_t *d_table, int sm_count, cudaStream_t stream);

} // namespace hmk
