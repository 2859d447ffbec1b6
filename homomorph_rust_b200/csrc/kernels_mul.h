// kernels_mul.h — host interface of kernels_mul.cu: the fused column multiplier, one launch for the whole
// circuit of reference src/impls/numbers/common.rs:66-105 on fresh u8 operands at D = d + d' = 256.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <vector>

#include "gf2_blocks.cuh"

namespace hmk {

constexpr uint32_t K7_L = 8;       // bits of the integers
constexpr uint32_t K7_D = 256;     // degree bound of a fresh operand slot
constexpr uint32_t K7_PPS = K7_L * (K7_L + 1) / 2;

// An object of a value's arena: `nw` low 32-bit words at word offset `off`, then one word holding the coefficient of
// X^(32 nw) in bit 0 (every degree bound of the circuit is a multiple of 256), then one padding word.
struct K7Item {
    uint16_t off, nw;
};
struct K7Prod { // carry = x * p
    K7Item x, p;
    uint16_t out_off, pad;
};
struct K7Unit { // out[out_off ..] ^= x[x_off .. x_off + xw) * p[p_off .. p_off + pw), xw and pw multiples of 8 words
    uint16_t x_off, p_off, out_off;
    uint8_t xw, pw;
};
struct K7Col {
    uint16_t item_first, item_count; // the column's items in the reference's order (partial products, then incoming carries)
    uint16_t prefix_first;           // item_count - 2 prefix objects P_2 .. P_{m-1} (columns that emit carries)
    uint16_t prod_first, prod_count; // item_count - 1 carry products (0 in the last column)
    uint16_t unit_first, unit_count, unit_words; // block products of unit_words x unit_words words (32, 16 or 8)
    uint16_t zero_off, zero_words;   // the carries this column emits (cleared before accumulation)
    uint16_t res_words;              // 32-bit words of the result slot
    uint32_t res_off;                // u64 word offset of the result slot in an output value
};
struct K7Plan {
    K7Col col[K7_L];
    uint16_t pp_off[K7_PPS];
    uint8_t pp_j[K7_PPS], pp_k[K7_PPS]; // partial product idx = a_j * b_k
    uint32_t arena_words;
};
struct K7Host {
    K7Plan plan;
    std::vector<K7Item> items;
    std::vector<K7Prod> prods;
    std::vector<K7Unit> units;
};

// Builds the static plan for L = 8 fresh operands of degree bound 256; `out` = layout and per-slot degree bounds of the result
// batch (hm result_bounds).  Returns false when the bounds do not match the plan's own (the caller falls back).
// nacc = accumulators the carries of the last emitting column are summed into (1..8).
bool k7_build_plan(const Layout &out, const uint64_t *out_degb, uint32_t nacc, K7Host *h);
size_t k7_smem_bytes(const K7Plan &plan, int warps);
// warps = values in flight per CTA (one warp each; 8, 10, 12, 14 or 16), one CTA per SM
cudaError_t launch_mul_circuit_fused(int warps, const uint64_t *A, const uint64_t *B, uint64_t *O, uint64_t n, uint32_t out_value_words, const K7Plan &plan,
                                     const K7Item *d_items, const K7Prod *d_prods, const K7Unit *d_units, int sm_count, cudaStream_t stream);

} // namespace hmk
