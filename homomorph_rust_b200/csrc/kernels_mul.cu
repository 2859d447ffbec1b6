// kernels_mul.cu — the fused column multiplier: reference src/impls/numbers/common.rs:66-105
// (mul_unsigned_internal) for u8 operands of fresh ciphertexts at D = d + d' = 256, ONE launch for the whole circuit.
//
// One WARP per value; everything the circuit produces on the way — the 36 partial products, the column prefixes and the
// 56 carries — lives in the warp's 25 KB slice of shared memory, so HBM sees the 2 x 320 B of operands and the 4.2 KB
// result and nothing else (the column-batched plan of hmgpu.cu moves a 32 KB arena per value through HBM in 24 launches).
//
// The circuit, column by column (the same regrouping as the host plan, so the same canonical polynomials): in column i
// the reference XORs the items x_1..x_m (partial products a_j b_{i-j}, then the carries of column i-1 in push order)
// into result[i] one at a time and pushes the carry x_t * result[i] before each XOR; result[i] at that moment is the prefix
// P_{t-1} = x_1 + ... + x_{t-1}.  Per column the warp
//   1. clears the carries it is about to emit, runs the prefix pass (lanes over words; writes P_2..P_{m-1} to shared memory
//      and P_m = result[i] to the output slot),
//   2. computes the m-1 carry products x_t * P_{t-1} as independent block products (32x32 words = nine 8x8-word Karatsubas
//      on IMAD.WIDE, or 16x16 = three, whichever fills the 32 lanes better in that column), one block product per lane
//      per round, accumulated into the carry with shared-memory atomics (the blocks of one product overlap),
//   3. adds the word-aligned terms of the top coefficients (every degree bound is a multiple of 256, so an object is its
//      low words plus the single coefficient of X^deg).
// The carries of even and odd columns alternate between two regions, the prefixes reuse one.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "gf2_blocks.cuh"
#include "kernels_mul.h"

namespace hmk {

namespace {

template <int O> __device__ __forceinline__ void xor16(uint32_t (&t)[64], const uint32_t (&r)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) t[O + i] ^= r[i];
}

// t ^= x * p through ONE rolled copy of the 8x8-word product.  Schedule entries [s0, s1): 0..8 = the two Karatsuba levels of a
// 32x32-word product over 8-word blocks (leaf operand = XOR of the blocks in `sel`), 9..11 = one level for 16x16 words,
// 12 = a single 8x8 block.  Blocks at or beyond xw / pw words are zero (operands are not padded in memory).
__device__ __forceinline__ void unit_mul(const uint32_t *__restrict__ x, uint32_t xw, const uint32_t *__restrict__ p, uint32_t pw, int s0, int s1,
                                         uint32_t (&t)[64]) {
#pragma unroll 1
    for (int s = s0; s < s1; ++s) {
        const uint32_t sel = (uint32_t)(0x1321FA5C84321ull >> (4 * s)) & 0xFu; // 1 2 3 4 8 C 5 A F | 1 2 3 | 1
        uint32_t a[8], b[8], r[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) a[q] = b[q] = 0;
#pragma unroll
        for (int blk = 0; blk < 4; ++blk) {
            if (sel >> blk & 1) {
                if (8u * blk < xw) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint2 u = *reinterpret_cast<const uint2 *>(x + 8 * blk + 2 * q);
                        a[2 * q] ^= u.x; a[2 * q + 1] ^= u.y;
                    }
                }
                if (8u * blk < pw) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint2 w = *reinterpret_cast<const uint2 *>(p + 8 * blk + 2 * q);
                        b[2 * q] ^= w.x; b[2 * q + 1] ^= w.y;
                    }
                }
            }
        }
        clmul_kara<8>(a, b, r);
        switch (s) { // 0..8 as mul32_acc (kernels.cuh); 9..11: lo {0,8}, hi {16,8}, mid {8}
            case 0: xor16<0>(t, r); xor16<8>(t, r); xor16<16>(t, r); xor16<24>(t, r); break;
            case 1: xor16<16>(t, r); xor16<8>(t, r); xor16<32>(t, r); xor16<24>(t, r); break;
            case 2: xor16<8>(t, r); xor16<24>(t, r); break;
            case 3: xor16<32>(t, r); xor16<40>(t, r); xor16<16>(t, r); xor16<24>(t, r); break;
            case 4: xor16<48>(t, r); xor16<40>(t, r); xor16<32>(t, r); xor16<24>(t, r); break;
            case 5: xor16<40>(t, r); xor16<24>(t, r); break;
            case 6: xor16<16>(t, r); xor16<24>(t, r); break;
            case 7: xor16<32>(t, r); xor16<24>(t, r); break;
            case 8: xor16<24>(t, r); break;
            case 9: xor16<0>(t, r); xor16<8>(t, r); break;
            case 10: xor16<16>(t, r); xor16<8>(t, r); break;
            case 11: xor16<8>(t, r); break;
            default: xor16<0>(t, r); break;
        }
    }
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
mul_circuit_fused_kernel(const uint64_t *__restrict__ A, const uint64_t *__restrict__ B, uint64_t *__restrict__ O, uint64_t n, uint32_t out_value_words,
                         const K7Plan plan, const K7Item *__restrict__ items, const K7Prod *__restrict__ prods, const K7Unit *__restrict__ units) {
    extern __shared__ __align__(16) uint32_t k7_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *ar = k7_smem + (size_t)warp * plan.arena_words;
    constexpr uint32_t WF = K7_D / 64 + 1; // u64 words of a fresh slot
    // values are dealt CTA-minor, so that a batch smaller than one wave spreads over all SMs instead of filling a few of them
    for (uint64_t v = (uint64_t)warp * gridDim.x + blockIdx.x; v < n; v += (uint64_t)gridDim.x * WARPS) {
        const uint64_t *av = A + v * (K7_L * WF), *bv = B + v * (K7_L * WF);
        uint32_t *ov = reinterpret_cast<uint32_t *>(O + v * (uint64_t)out_value_words);
        // ---- the 36 partial products a_j * b_k (j + k < 8), one lane each, two rounds ----
        for (uint32_t idx = lane; idx < K7_PPS; idx += 32) {
            uint32_t *slot = ar + plan.pp_off[idx];
            const uint64_t *aj = av + plan.pp_j[idx] * WF, *bk = bv + plan.pp_k[idx] * WF;
#pragma unroll
            for (int q = 0; q < 4; ++q) { // operands parked in the slot itself: words 0..7 = a', 8..15 = b'
                const uint64_t x = __ldg(aj + q), y = __ldg(bk + q);
                slot[2 * q] = (uint32_t)x; slot[2 * q + 1] = (uint32_t)(x >> 32);
                slot[8 + 2 * q] = (uint32_t)y; slot[8 + 2 * q + 1] = (uint32_t)(y >> 32);
            }
            const uint32_t atop = (uint32_t)__ldg(aj + 4) & 1u, btop = (uint32_t)__ldg(bk + 4) & 1u;
            uint32_t t[64];
#pragma unroll
            for (int i = 0; i < 64; ++i) t[i] = 0;
            unit_mul(slot, 8, slot + 8, 8, 12, 13, t);
            const uint32_t ma = 0u - atop, mb = 0u - btop; // (a' + at X^256)(b' + bt X^256)
#pragma unroll
            for (int i = 0; i < 8; ++i) t[8 + i] ^= (slot[8 + i] & ma) ^ (slot[i] & mb);
#pragma unroll
            for (int i = 0; i < 16; ++i) slot[i] = t[i];
            slot[16] = atop & btop;
            slot[17] = 0;
        }
        __syncwarp();
        for (uint32_t i = 0; i < K7_L; ++i) {
            const K7Col c = plan.col[i];
            for (uint32_t w = lane; w < c.zero_words; w += 32) ar[c.zero_off + w] = 0;
            // ---- prefix pass: lanes over words (word lane + 32 q in register q), items in the reference's order ----
            const bool emits = c.prod_count != 0;
            {
                constexpr int MAXQ = 15; // the widest result slot (bit 7: degree 14 336) is 450 words
                uint32_t run[MAXQ];
#pragma unroll
                for (int q = 0; q < MAXQ; ++q) run[q] = 0;
                for (uint32_t t = 0; t < c.item_count; ++t) {
                    const K7Item it = items[c.item_first + t];
#pragma unroll
                    for (int q = 0; q < MAXQ; ++q)
                        if ((uint32_t)(lane + 32 * q) <= it.nw) run[q] ^= ar[it.off + lane + 32 * q];
                    if (emits && t >= 1 && t + 1 < c.item_count) { // P_{t+1}
                        const K7Item pf = items[c.prefix_first + t - 1];
#pragma unroll
                        for (int q = 0; q < MAXQ; ++q)
                            if ((uint32_t)(lane + 32 * q) <= pf.nw) ar[pf.off + lane + 32 * q] = run[q];
                    }
                }
#pragma unroll
                for (int q = 0; q < MAXQ; ++q)
                    if ((uint32_t)(lane + 32 * q) < c.res_words) ov[2 * c.res_off + lane + 32 * q] = run[q];
            }
            __syncwarp();
            if (!emits) continue;
            // ---- carry products: one block product per lane per round ----
            const int s0 = c.unit_words == 32 ? 0 : (c.unit_words == 16 ? 9 : 12), s1 = c.unit_words == 32 ? 9 : (c.unit_words == 16 ? 12 : 13);
            for (uint32_t base = 0; base < c.unit_count; base += 32) {
                if (base + lane < c.unit_count) {
                    const K7Unit u = units[c.unit_first + base + lane];
                    uint32_t t[64];
#pragma unroll
                    for (int q = 0; q < 64; ++q) t[q] = 0;
                    unit_mul(ar + u.x_off, u.xw, ar + u.p_off, u.pw, s0, s1, t);
                    uint32_t *out = ar + u.out_off;
#pragma unroll
                    for (int q = 0; q < 64; ++q)
                        if (t[q]) atomicXor(out + q, t[q]);
                }
            }
            // ---- top coefficients: (x' + xt X^a)(p' + pt X^b) = x'p' + xt X^a p' + pt X^b x' + xt pt X^(a+b) ----
            for (uint32_t t = 0; t < c.prod_count; ++t) {
                const K7Prod pr = prods[c.prod_first + t];
                const uint32_t xt = ar[pr.x.off + pr.x.nw] & 1u, pt = ar[pr.p.off + pr.p.nw] & 1u;
                uint32_t *out = ar + pr.out_off;
                if (xt)
                    for (uint32_t w = lane; w < pr.p.nw; w += 32) {
                        const uint32_t val = ar[pr.p.off + w];
                        if (val) atomicXor(out + pr.x.nw + w, val);
                    }
                if (pt)
                    for (uint32_t w = lane; w < pr.x.nw; w += 32) {
                        const uint32_t val = ar[pr.x.off + w];
                        if (val) atomicXor(out + pr.p.nw + w, val);
                    }
                if (lane == 0 && (xt & pt)) atomicXor(out + pr.x.nw + pr.p.nw, 1u);
            }
            __syncwarp();
        }
    }
}

} // namespace

bool k7_build_plan(const Layout &out, const uint64_t *out_degb, uint32_t nacc, K7Host *h) {
    if (nacc < 1 || nacc > 8) return false;
    constexpr uint32_t ACC_SKEW = 8; // words between accumulators beyond their size: successive accumulators start 12 banks apart
    if (out.L != K7_L) return false;
    struct Obj {
        uint32_t off;
        uint64_t deg;
    };
    auto words = [](uint64_t deg) { return (uint32_t)(2 * (deg / 64 + 1)); }; // storage, 32-bit words
    K7Plan &pl = h->plan;
    h->items.clear();
    h->prods.clear();
    h->units.clear();
    // first pass: sizes of the regions (prefixes of one column; carries of even / odd columns)
    uint32_t pps_words = K7_PPS * words(2 * K7_D), prefix_max = 0, carry_max[2] = {0, 0}, prefix_words[K7_L] = {};
    {
        std::vector<uint64_t> incoming;
        for (uint32_t i = 0; i < K7_L; ++i) {
            std::vector<uint64_t> degs(i + 1, 2 * K7_D);
            degs.insert(degs.end(), incoming.begin(), incoming.end());
            incoming.clear();
            uint64_t dres = degs[0];
            uint32_t pw = 0, cw = 0;
            for (size_t t = 1; t < degs.size(); ++t) {
                if (i + 1 < K7_L) {
                    if (t >= 2) pw += words(dres);
                    cw += words(degs[t] + dres);
                    incoming.push_back(degs[t] + dres);
                }
                dres = std::max(dres, degs[t]);
            }
            if (dres != out_degb[i]) return false;
            if (i + 2 == K7_L && !incoming.empty()) { // the last carries are only ever summed (see below): nacc accumulators
                const uint64_t dmax = *std::max_element(incoming.begin(), incoming.end());
                incoming.assign(nacc, dmax);
                cw = nacc * (words(dmax) + ACC_SKEW);
            }
            prefix_words[i] = pw;
            prefix_max = std::max(prefix_max, pw);
            carry_max[i & 1] = std::max(carry_max[i & 1], cw);
        }
    }
    // Partial products of the early columns are dead before the late columns' prefixes need the tail of the prefix region, so
    // they live there; only the late columns' partial products get space of their own.
    const uint32_t ppw = words(2 * K7_D);
    uint32_t early_cols = 0; // columns 0 .. early_cols - 1 keep their partial products in the prefix region's tail
    for (uint32_t e = K7_L; e > 0; --e) {
        uint32_t early = 0, need = 0;
        for (uint32_t c = 0; c < e; ++c) {
            early += (c + 1) * ppw;
            need = std::max(need, prefix_words[c]);
        }
        if (need + early <= prefix_max) {
            early_cols = e;
            break;
        }
    }
    uint32_t early_words = 0, late_words = 0;
    for (uint32_t c = 0; c < K7_L; ++c) (c < early_cols ? early_words : late_words) += (c + 1) * ppw;
    (void)pps_words;
    const uint32_t prefix_base = late_words, carry_base[2] = {prefix_base + prefix_max, prefix_base + prefix_max + carry_max[0]};
    pl.arena_words = carry_base[1] + carry_max[1];
    if (pl.arena_words >= 65536) return false;
    // partial products: idx in (j, k) order
    std::vector<std::vector<Obj>> pp(K7_L, std::vector<Obj>(K7_L));
    {
        uint32_t idx = 0, early_cur = prefix_base + prefix_max - early_words, late_cur = 0;
        for (uint32_t j = 0; j < K7_L; ++j)
            for (uint32_t k = 0; j + k < K7_L; ++k, ++idx) {
                uint32_t &cur = (j + k < early_cols) ? early_cur : late_cur;
                pp[j][k] = Obj{cur, 2 * K7_D};
                cur += ppw;
                pl.pp_off[idx] = (uint16_t)pp[j][k].off;
                pl.pp_j[idx] = (uint8_t)j;
                pl.pp_k[idx] = (uint8_t)k;
            }
    }
    auto item = [](const Obj &o) { return K7Item{(uint16_t)o.off, (uint16_t)(o.deg / 32)}; };
    std::vector<Obj> incoming;
    for (uint32_t i = 0; i < K7_L; ++i) {
        K7Col &c = pl.col[i];
        std::vector<Obj> its;
        for (uint32_t j = 0; j <= i; ++j) its.push_back(pp[j][i - j]);
        its.insert(its.end(), incoming.begin(), incoming.end());
        incoming.clear();
        c.item_first = (uint16_t)h->items.size();
        c.item_count = (uint16_t)its.size();
        for (const Obj &o : its) h->items.push_back(item(o));
        const bool emits = i + 1 < K7_L;
        uint32_t pcur = prefix_base, ccur = carry_base[i & 1];
        std::vector<Obj> prefixes, carries;
        uint64_t dres = its[0].deg;
        // The column after the last emitting one has no products: its result is the plain sum of its items, so the carries of
        // column L - 2 are never needed one by one: they accumulate into `nacc` objects (2.6 K words of shared memory otherwise).
        // Every block of every product lands a multiple of 32 words above its object's base, so with ONE accumulator all lanes
        // of a round would hit the same bank in each of their 64 atomics; several accumulators whose bases differ by 12 banks
        // (blocks are dealt to them round robin) spread that.
        const bool summed = i + 2 == K7_L;
        uint64_t dsum = 0;
        for (size_t t = 1; t < its.size(); ++t) {
            if (emits) {
                if (t >= 2) {
                    prefixes.push_back(Obj{pcur, dres});
                    pcur += words(dres);
                }
                carries.push_back(Obj{ccur, its[t].deg + dres});
                dsum = std::max(dsum, its[t].deg + dres);
                if (!summed) ccur += words(its[t].deg + dres);
            }
            dres = std::max(dres, its[t].deg);
        }
        if (summed && emits) ccur += nacc * (words(dsum) + ACC_SKEW);
        c.prefix_first = (uint16_t)h->items.size();
        for (const Obj &o : prefixes) h->items.push_back(item(o));
        c.prod_first = (uint16_t)h->prods.size();
        c.prod_count = (uint16_t)carries.size();
        c.zero_off = (uint16_t)carry_base[i & 1];
        c.zero_words = (uint16_t)(ccur - carry_base[i & 1]);
        c.res_off = out.off[i];
        c.res_words = (uint16_t)(2 * (out.off[i + 1] - out.off[i]));
        std::vector<K7Unit> cand[3]; // blocks of 32, 16 and 8 words
        const uint32_t usize[3] = {32, 16, 8}, ucost[3] = {9, 3, 1}; // a round costs this many 8x8-word products
        for (size_t t = 1; t < its.size() && emits; ++t) {
            const Obj &x = its[t];
            Obj p = t == 1 ? its[0] : prefixes[t - 2];
            const K7Prod pr{item(x), item(p), (uint16_t)(summed ? carry_base[i & 1] : carries[t - 1].off), 0};
            h->prods.push_back(pr);
            for (int k = 0; k < 3; ++k) {
                const uint32_t U = usize[k];
                std::vector<K7Unit> &dst = cand[k];
                for (uint32_t ci = 0; ci < pr.x.nw; ci += U)
                    for (uint32_t cj = 0; cj < pr.p.nw; cj += U) {
                        const uint32_t obase = summed ? carry_base[i & 1] + (uint32_t)(dst.size() % nacc) * (words(dsum) + ACC_SKEW) : pr.out_off;
                        dst.push_back(K7Unit{(uint16_t)(pr.x.off + ci), (uint16_t)(pr.p.off + cj), (uint16_t)(obase + ci + cj),
                                             (uint8_t)std::min(U, pr.x.nw - ci), (uint8_t)std::min(U, pr.p.nw - cj)});
                    }
            }
        }
        int best = 0;
        for (int k = 1; k < 3; ++k)
            if (ucost[k] * ((cand[k].size() + 31) / 32) < ucost[best] * ((cand[best].size() + 31) / 32)) best = k;
        c.unit_words = (uint16_t)usize[best];
        c.unit_first = (uint16_t)h->units.size();
        c.unit_count = (uint16_t)cand[best].size();
        h->units.insert(h->units.end(), cand[best].begin(), cand[best].end());
        if (summed && emits) {
            incoming.clear();
            for (uint32_t a = 0; a < nacc; ++a) incoming.push_back(Obj{carry_base[i & 1] + a * (words(dsum) + ACC_SKEW), dsum});
        } else {
            incoming = carries;
        }
    }
    return h->items.size() < 65536 && h->prods.size() < 65536 && h->units.size() < 65536;
}

size_t k7_smem_bytes(const K7Plan &plan, int warps) { return (size_t)warps * plan.arena_words * 4; }

template <int WARPS>
static cudaError_t launch_k7(const uint64_t *A, const uint64_t *B, uint64_t *O, uint64_t n, uint32_t out_value_words, const K7Plan &plan,
                             const K7Item *d_items, const K7Prod *d_prods, const K7Unit *d_units, int sm_count, cudaStream_t stream) {
    const size_t smem = k7_smem_bytes(plan, WARPS);
    auto kern = mul_circuit_fused_kernel<WARPS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    uint64_t blocks = n;
    if (blocks > (uint64_t)sm_count) blocks = sm_count;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, WARPS * 32, smem, stream>>>(A, B, O, n, out_value_words, plan, d_items, d_prods, d_units);
    return cudaGetLastError();
}

cudaError_t launch_mul_circuit_fused(int warps, const uint64_t *A, const uint64_t *B, uint64_t *O, uint64_t n, uint32_t out_value_words, const K7Plan &plan,
                                     const K7Item *d_items, const K7Prod *d_prods, const K7Unit *d_units, int sm_count, cudaStream_t stream) {
    if (warps == 14) return launch_k7<14>(A, B, O, n, out_value_words, plan, d_items, d_prods, d_units, sm_count, stream);
    if (warps == 10) return launch_k7<10>(A, B, O, n, out_value_words, plan, d_items, d_prods, d_units, sm_count, stream);
    if (warps == 16) return launch_k7<16>(A, B, O, n, out_value_words, plan, d_items, d_prods, d_units, sm_count, stream);
    if (warps == 12) return launch_k7<12>(A, B, O, n, out_value_words, plan, d_items, d_prods, d_units, sm_count, stream);
    if (warps == 8) return launch_k7<8>(A, B, O, n, out_value_words, plan, d_items, d_prods, d_units, sm_count, stream);
    return cudaErrorInvalidValue;
}

} // namespace hmk
