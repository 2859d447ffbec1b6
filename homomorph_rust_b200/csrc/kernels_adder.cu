// kernels_adder.cu — the ripple-carry adder chain (reference src/impls/numbers/common.rs:37-56), one THREAD per value,
// scheduled dynamically.  Round-2 successor of adder_thread_smem_kernel (kernels.cuh), same arithmetic:
//
//   p_k = a_k + b_k,  g_k = a_k * b_k,  m_k = (1 + g_k) * p_k        (per bit k, three WD x WD-word Karatsuba products)
//   c_0 = 0,  c_{k+1} = m_k * c_k + g_k,  s_k = p_k + c_k             (serial over k)
//
// which is the reference's carry' = p c + (a b)(p c + 1) regrouped with ring identities, so every s_k is the
// same canonical polynomial.  c_k is cut in chunks of 3 WD words; chunk x m_k is a 3-way Karatsuba of six WD x WD-word
// products (clmul_kara<WD>, leaves on IMAD.WIDE), accumulated in a window of 6 WD words whose upper half carries into the
// next chunk.  c_k is re-read from the result itself (slot k holds s_k = c_k + p_k).
//
// What is new:
//  * PERSISTENT warps pull work units from an atomic counter.  A unit is (group of 32 consecutive values, phase); a phase is
//    a range of bits [kb[p], kb[p+1]) holding about 1/P of the chain's work (the work of bit k grows like k).  Units are
//    numbered phase-major, so when the counter reaches phase p+1 every phase-p unit has already been claimed by a running
//    warp; a per-group flag (release / acquire) orders the hand-over of c_k through global memory.  With one unit = a whole
//    value, 2^18 values over 75 776 resident threads are 3.46 waves that cost 4; with P = 4 they are 13.84 quarter-waves that
//    cost 14 — and warps that finish early no longer leave their scheduler slot idle (ncu, round 1: 12.3 of 16 warps active).
//  * WD is a template parameter: D = 256 (WD = 8, the tuned configuration) and D = 128 (WD = 4).
//  * the window can live in shared memory (TSM = 1) instead of 6 WD registers.
#include <cuda_runtime.h>
#include <stdint.h>

#include "gf2_blocks.cuh"
#include "kernels_adder.h"

namespace hmk {

namespace {

// out-of-line WD x WD-word product for the once-per-bit quantities (g_k, m_k)
template <int WD> __device__ __noinline__ void kara_call(const uint32_t *x, const uint32_t *y, uint32_t *r) {
    uint32_t a[WD], b[WD], o[2 * WD];
#pragma unroll
    for (int i = 0; i < WD; ++i) {
        a[i] = x[i];
        b[i] = y[i];
    }
    clmul_kara<WD>(a, b, o);
#pragma unroll
    for (int i = 0; i < 2 * WD; ++i) r[i] = o[i];
}

// this thread's column of a [pair][thread] shared array (128 threads per CTA): pair q at col[q * 128]
constexpr int CTA = 128;

template <int WD> __device__ __forceinline__ void load_block(uint32_t (&dst)[WD], const uint2 *src) {
#pragma unroll
    for (int q = 0; q < WD / 2; ++q) {
        const uint2 w = src[q * CTA];
        dst[2 * q] = w.x;
        dst[2 * q + 1] = w.y;
    }
}

// t[O .. O + WD) ^= v
template <int WD, int O, int NT> __device__ __forceinline__ void xor_block(uint32_t (&t)[NT], const uint32_t (&v)[WD]) {
#pragma unroll
    for (int i = 0; i < WD; ++i) t[O + i] ^= v[i];
}

// window ^= m * c for one chunk (3 blocks of WD words each); window in registers.
// 3-way Karatsuba: P0 = m0 c0, P1 = m1 c1, P2 = m2 c2, P01 = (m0+m1)(c0+c1), P02, P12 and
//   m c = P0 + (P01+P0+P1) X + (P02+P0+P2+P1) X^2 + (P12+P1+P2) X^3 + P2 X^4      (X = 32 WD bits)
// so with (lo, hi) the halves of a product and x = lo + hi:  P_i placed at block offsets {A, A+1, A+2} contributes
// lo, x, x, hi to blocks A .. A+3; the mixed products contribute lo, hi to two blocks.
template <int WD> __device__ __forceinline__ void mul_chunk_acc(const uint2 *__restrict__ m, const uint2 *__restrict__ c, uint32_t (&t)[6 * WD]) {
#pragma unroll 1
    for (int i = 0; i < 6; ++i) {
        const int o1 = (i < 3) ? i : (i == 5 ? 1 : 0); // in blocks
        const int o2 = (i < 3) ? -1 : (i == 3 ? 1 : 2);
        uint32_t x[WD], y[WD], r[2 * WD];
        load_block<WD>(x, m + o1 * (WD / 2) * CTA);
        load_block<WD>(y, c + o1 * (WD / 2) * CTA);
        if (o2 >= 0) {
            uint32_t u[WD], w[WD];
            load_block<WD>(u, m + o2 * (WD / 2) * CTA);
            load_block<WD>(w, c + o2 * (WD / 2) * CTA);
#pragma unroll
            for (int q = 0; q < WD; ++q) {
                x[q] ^= u[q];
                y[q] ^= w[q];
            }
        }
        clmul_kara<WD>(x, y, r);
        uint32_t lo[WD], hi[WD], xs[WD];
#pragma unroll
        for (int q = 0; q < WD; ++q) {
            lo[q] = r[q];
            hi[q] = r[WD + q];
            xs[q] = lo[q] ^ hi[q];
        }
        switch (i) {
            case 0: xor_block<WD, 0 * WD>(t, lo); xor_block<WD, 1 * WD>(t, xs); xor_block<WD, 2 * WD>(t, xs); xor_block<WD, 3 * WD>(t, hi); break;
            case 1: xor_block<WD, 1 * WD>(t, lo); xor_block<WD, 2 * WD>(t, xs); xor_block<WD, 3 * WD>(t, xs); xor_block<WD, 4 * WD>(t, hi); break;
            case 2: xor_block<WD, 2 * WD>(t, lo); xor_block<WD, 3 * WD>(t, xs); xor_block<WD, 4 * WD>(t, xs); xor_block<WD, 5 * WD>(t, hi); break;
            case 3: xor_block<WD, 1 * WD>(t, lo); xor_block<WD, 2 * WD>(t, hi); break;
            case 4: xor_block<WD, 2 * WD>(t, lo); xor_block<WD, 3 * WD>(t, hi); break;
            default: xor_block<WD, 3 * WD>(t, lo); xor_block<WD, 4 * WD>(t, hi); break;
        }
    }
}

// same with the window in shared memory: blocks 0..2 at wl (this chunk's output), blocks 3..5 at wh (the carry into the next
// chunk; its first touch of a block is a plain store, so wh needs no clearing).
template <int WD, bool STORE> __device__ __forceinline__ void smem_block(uint2 *p, const uint32_t (&v)[WD]) {
#pragma unroll
    for (int q = 0; q < WD / 2; ++q) {
        uint2 w = make_uint2(v[2 * q], v[2 * q + 1]);
        if (!STORE) {
            const uint2 o = p[q * CTA];
            w.x ^= o.x;
            w.y ^= o.y;
        }
        p[q * CTA] = w;
    }
}
template <int WD> __device__ __forceinline__ void mul_chunk_acc_s(const uint2 *__restrict__ m, const uint2 *__restrict__ c, uint2 *wl, uint2 *wh) {
    constexpr int BP = (WD / 2) * CTA; // uint2 elements per block
#pragma unroll 1
    for (int i = 0; i < 6; ++i) {
        const int o1 = (i < 3) ? i : (i == 5 ? 1 : 0);
        const int o2 = (i < 3) ? -1 : (i == 3 ? 1 : 2);
        uint32_t x[WD], y[WD], r[2 * WD];
        load_block<WD>(x, m + o1 * BP);
        load_block<WD>(y, c + o1 * BP);
        if (o2 >= 0) {
            uint32_t u[WD], w[WD];
            load_block<WD>(u, m + o2 * BP);
            load_block<WD>(w, c + o2 * BP);
#pragma unroll
            for (int q = 0; q < WD; ++q) {
                x[q] ^= u[q];
                y[q] ^= w[q];
            }
        }
        clmul_kara<WD>(x, y, r);
        uint32_t lo[WD], hi[WD], xs[WD];
#pragma unroll
        for (int q = 0; q < WD; ++q) {
            lo[q] = r[q];
            hi[q] = r[WD + q];
            xs[q] = lo[q] ^ hi[q];
        }
        switch (i) { // order 0..5 fixes who touches a carry block first: P0 -> block 3, P1 -> block 4, P2 -> block 5
            case 0: smem_block<WD, false>(wl, lo); smem_block<WD, false>(wl + BP, xs); smem_block<WD, false>(wl + 2 * BP, xs); smem_block<WD, true>(wh, hi); break;
            case 1: smem_block<WD, false>(wl + BP, lo); smem_block<WD, false>(wl + 2 * BP, xs); smem_block<WD, false>(wh, xs); smem_block<WD, true>(wh + BP, hi); break;
            case 2: smem_block<WD, false>(wl + 2 * BP, lo); smem_block<WD, false>(wh, xs); smem_block<WD, false>(wh + BP, xs); smem_block<WD, true>(wh + 2 * BP, hi); break;
            case 3: smem_block<WD, false>(wl + BP, lo); smem_block<WD, false>(wl + 2 * BP, hi); break;
            case 4: smem_block<WD, false>(wl + 2 * BP, lo); smem_block<WD, false>(wh, hi); break;
            default: smem_block<WD, false>(wh, lo); smem_block<WD, false>(wh + BP, hi); break;
        }
    }
}

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// bits [k0, k1) of the chain for one value.  Iteration k writes slot k + 1 (and slot 0 at k = 0); it reads slot k.
template <int WD, int TSM>
__device__ __forceinline__ void chain_range(const uint64_t *__restrict__ Av, const uint64_t *__restrict__ Bv, uint32_t *__restrict__ Ov,
                                            const Layout &lo, uint32_t L, uint32_t k0, uint32_t k1, uint2 *mb, uint2 *cb0, uint2 *cb1, uint2 *wb) {
    constexpr int WF = WD / 2 + 1;  // u64 words of a fresh slot
    constexpr int CH = 3 * WD;      // 32-bit words per chunk = words of m_k
    constexpr int NPR = CH / 2;     // pairs per chunk
    for (uint32_t k = k0; k < k1; ++k) {
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
        if (k + 1 == L) break; // no carry out of the last bit (common.rs:47-49)
        uint32_t g[2 * WD];
        kara_call<WD>(a, b, g);
        const uint32_t ma = 0u - atop, mbm = 0u - btop;
#pragma unroll
        for (int j = 0; j < WD; ++j) g[WD + j] ^= (b[j] & ma) ^ (a[j] & mbm);
        const uint32_t gtop = atop & btop; // coefficient of X^(2D)
        uint32_t pn[WD + 1];               // p_{k+1}: s_{k+1} = c_{k+1} + p_{k+1} is emitted on the fly
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
            continue;
        }
        { // m = p + g p: 3 WD words into shared memory.  (Its X^(3D) coefficient gtop & ptop is identically 0.)
            uint32_t m[CH], glo[WD], ghi[WD], q0[2 * WD], q1[2 * WD];
#pragma unroll
            for (int j = 0; j < WD; ++j) { glo[j] = g[j]; ghi[j] = g[WD + j]; }
            kara_call<WD>(glo, p, q0);
            kara_call<WD>(ghi, p, q1);
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
            for (int q = 0; q < NPR; ++q) mb[q * CTA] = make_uint2(m[2 * q], m[2 * q + 1]);
        }
        const uint32_t *cslot = Ov + 2 * lo.off[k]; // s_k = c_k + p_k
        const uint32_t len = CH * k - (WD - 1);     // words of c_k; exactly k chunks, the last one has 2 WD + 1 valid words
        // chunk 0 (p_k mixed in) and the ragged last chunk are staged by hand; full chunks come by cp.async
        auto fill_first = [&](uint2 *dstb) {
#pragma unroll
            for (int q = 0; q < NPR; ++q) {
                uint2 w = make_uint2(0u, 0u);
                if ((uint32_t)(2 * q) < len) w.x = cslot[2 * q];
                if ((uint32_t)(2 * q + 1) < len) w.y = cslot[2 * q + 1];
                if (2 * q <= WD) w.x ^= p[2 * q <= WD ? 2 * q : 0];
                if (2 * q + 1 <= WD) w.y ^= p[2 * q + 1 <= WD ? 2 * q + 1 : 0];
                dstb[q * CTA] = w;
            }
        };
        auto fill_last = [&](uint2 *dstb, uint32_t j) { // words [CH j, CH j + 2 WD + 1) are valid, the rest of the chunk is 0
            const uint32_t *src = cslot + CH * j;
#pragma unroll
            for (int q = 0; q < NPR; ++q) {
                uint2 w = make_uint2(0u, 0u);
                if (2 * q < 2 * WD + 1) w.x = src[2 * q];
                if (2 * q + 1 < 2 * WD + 1) w.y = src[2 * q + 1];
                dstb[q * CTA] = w;
            }
        };
        auto fill_async = [&](uint2 *dstb, uint32_t j) {
            const uint32_t sa = (uint32_t)__cvta_generic_to_shared(dstb);
            const uint32_t *src = cslot + CH * j;
#pragma unroll
            for (int q = 0; q < NPR; ++q)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa + q * CTA * 8), "l"(src + 2 * q) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        fill_first(cb0);
        if constexpr (TSM == 0) {
            // the window starts as g_k + p_{k+1} (chunk 0 of c_{k+1} = m c_k + g_k, emitted as s_{k+1} = c_{k+1} + p_{k+1}), so that
            // g and p_{k+1} are dead before the chunk loop begins: nothing but the window and one product lives across it
            uint32_t t[2 * CH];
#pragma unroll
            for (int i = 0; i < 2 * CH; ++i) t[i] = 0;
#pragma unroll
            for (int i = 0; i < 2 * WD; ++i) t[i] = g[i];
            t[2 * WD] = gtop;
#pragma unroll
            for (int i = 0; i <= WD; ++i) t[i] ^= pn[i];
            for (uint32_t j = 0; j <= k; ++j) {
                uint2 *cur = (j & 1) ? cb1 : cb0, *nxt = (j & 1) ? cb0 : cb1;
                if (j < k) {
                    bool pref = false;
                    if (j + 1 < k) {
                        if (j + 2 < k) {
                            fill_async(nxt, j + 1);
                            pref = true;
                        } else {
                            fill_last(nxt, j + 1);
                        }
                    }
                    if (pref) asm volatile("cp.async.wait_group 1;" ::: "memory");
                    else asm volatile("cp.async.wait_group 0;" ::: "memory");
                    mul_chunk_acc<WD>(mb, cur, t);
                }
#pragma unroll
                for (int q = 0; q < NPR; ++q)
                    if (CH * j + 2 * q < swo) *reinterpret_cast<uint2 *>(sdst + CH * j + 2 * q) = make_uint2(t[2 * q], t[2 * q + 1]);
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    t[i] = t[CH + i];
                    t[CH + i] = 0;
                }
            }
        } else {
            // window halves in shared memory, swapped after every chunk: wl = this chunk's output, wh = carry
            uint2 *wl = wb, *wh = wb + NPR * CTA;
#pragma unroll
            for (int q = 0; q < NPR; ++q) { // chunk 0's output starts as g_k + p_{k+1}
                const uint32_t i0 = 2 * q, i1 = 2 * q + 1;
                uint32_t x0 = 0, x1 = 0;
                if (i0 < 2 * WD) x0 = g[i0 < 2 * WD ? i0 : 0];
                if (i1 < 2 * WD) x1 = g[i1 < 2 * WD ? i1 : 0];
                if (i0 == 2 * WD) x0 = gtop;
                if (i1 == 2 * WD) x1 = gtop;
                if (i0 <= WD) x0 ^= pn[i0 <= WD ? i0 : 0];
                if (i1 <= WD) x1 ^= pn[i1 <= WD ? i1 : 0];
                wl[q * CTA] = make_uint2(x0, x1);
            }
            for (uint32_t j = 0; j <= k; ++j) {
                uint2 *cur = (j & 1) ? cb1 : cb0, *nxt = (j & 1) ? cb0 : cb1;
                if (j < k) {
                    bool pref = false;
                    if (j + 1 < k) {
                        if (j + 2 < k) {
                            fill_async(nxt, j + 1);
                            pref = true;
                        } else {
                            fill_last(nxt, j + 1);
                        }
                    }
                    if (pref) asm volatile("cp.async.wait_group 1;" ::: "memory");
                    else asm volatile("cp.async.wait_group 0;" ::: "memory");
                    mul_chunk_acc_s<WD>(mb, cur, wl, wh);
                }
#pragma unroll
                for (int q = 0; q < NPR; ++q)
                    if (CH * j + 2 * q < swo) *reinterpret_cast<uint2 *>(sdst + CH * j + 2 * q) = wl[q * CTA];
                uint2 *tp = wl; wl = wh; wh = tp; // the carry becomes the next output; the old output half is overwritten by stores
            }
        }
    }
}

template <int WD, int TSM, int MINB>
__global__ void __launch_bounds__(CTA, MINB) adder_chain_kernel(const uint64_t *__restrict__ A, const uint64_t *__restrict__ B,
                                                               uint64_t *__restrict__ O, uint64_t n, uint32_t L, Layout lo, AdderSched sc) {
    constexpr int WF = WD / 2 + 1, NPR = 3 * WD / 2;
    extern __shared__ __align__(16) uint2 chain_smem[];
    uint2 *mb = chain_smem + threadIdx.x, *cb0 = mb + NPR * CTA, *cb1 = mb + 2 * NPR * CTA, *wb = mb + 3 * NPR * CTA;
    const int lane = threadIdx.x & 31;
    const uint32_t total = sc.ngroups * sc.nphases;
    for (;;) {
        uint32_t unit = 0;
        if (lane == 0) unit = atomicAdd(sc.counter, 1u);
        unit = __shfl_sync(FULL, unit, 0);
        if (unit >= total) break;
        const uint32_t ph = unit / sc.ngroups, grp = unit - ph * sc.ngroups;
        if (ph) { // c_k of this group comes from the unit (grp, ph - 1), claimed earlier by a warp that is running or done
            while (ld_acquire_u32(sc.done + grp) < ph) __nanosleep(256);
        }
        const uint64_t v = (uint64_t)grp * 32 + lane;
        if (v < n)
            chain_range<WD, TSM>(A + v * (uint64_t)L * WF, B + v * (uint64_t)L * WF, reinterpret_cast<uint32_t *>(O + v * (uint64_t)lo.value_words), lo, L,
                                 sc.kb[ph], sc.kb[ph + 1], mb, cb0, cb1, wb);
        if (ph + 1 < sc.nphases) {
            __threadfence();
            __syncwarp();
            if (lane == 0) st_release_u32(sc.done + grp, ph + 1);
        }
    }
}

// ----------------------------------------------------------------------------------------------------------------------
// D = 1024 (config B: d = d' = 512; WD = 32 words per fresh operand).  Same chain, same 24-word chunks of c_k and the same
// 8x8-word products, but m_k has 96 words = FOUR sub-multipliers M_0..M_3 of 24 words:
//     c_{k+1}[chunk o] = sum_q low(M_q * C_{o-q}) + high(M_q * C_{o-q-1})  (+ g_k)
// so every output chunk takes four chunk products (24 8x8-word products) from the four most recent chunks of c_k, which sit
// in a ring of five per-thread shared-memory chunk buffers (the fifth receives the cp.async prefetch).  Per thread: m_k 96
// words + ring 120 words = 864 B of shared memory, 128 threads per CTA, 2 CTAs per SM (up to 255 registers); the
// 8x8-word product runs at the same rate with 8 warps per SM as with 32 (profiles/r02_ubench2_pipes.txt).
// The once-per-bit products (g = a b: 32 x 32 words, m = (1 + g) p: 64 x 32 words) go block by block through shared memory.
// ----------------------------------------------------------------------------------------------------------------------
constexpr int WIDE_M_PAIRS = 48;     // m_k: 96 words
constexpr int WIDE_RING = 5;         // chunk buffers
constexpr int WIDE_PAIRS = WIDE_M_PAIRS + WIDE_RING * 12;

// out[(bi + bj) blocks ...] ^= x * y, operands and result in this thread's shared-memory columns (8-word blocks = 4 pairs)
__device__ __forceinline__ void smem_mul_acc8(const uint2 *x, int nbx, const uint2 *y, int nby, uint2 *out) {
#pragma unroll 1
    for (int bi = 0; bi < nbx; ++bi) {
#pragma unroll 1
        for (int bj = 0; bj < nby; ++bj) {
            uint32_t a[8], b[8], r[16];
            load_block<8>(a, x + bi * 4 * CTA);
            load_block<8>(b, y + bj * 4 * CTA);
            clmul_kara<8>(a, b, r);
            uint2 *o = out + (bi + bj) * 4 * CTA;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                uint2 w = o[q * CTA];
                w.x ^= r[2 * q];
                w.y ^= r[2 * q + 1];
                o[q * CTA] = w;
            }
        }
    }
}

__device__ __forceinline__ void chain_range_wide(const uint64_t *__restrict__ Av, const uint64_t *__restrict__ Bv, uint32_t *__restrict__ Ov,
                                                 const Layout &lo, uint32_t L, uint32_t k0, uint32_t k1, uint2 *mb, uint2 *ring) {
    constexpr int WD = 32, WF = WD / 2 + 1, CH = 24, NPR = 12;
    uint32_t *mbw = reinterpret_cast<uint32_t *>(mb);     // word i of a column: w[(i / 2) * CTA * 2 + (i & 1)]
    uint32_t *ringw = reinterpret_cast<uint32_t *>(ring);
    auto W = [](uint32_t *col, int i) -> uint32_t & { return col[(i >> 1) * CTA * 2 + (i & 1)]; };
    for (uint32_t k = k0; k < k1; ++k) {
        const uint64_t *ak = Av + (size_t)k * WF, *bk = Bv + (size_t)k * WF;
        const uint32_t atop = (uint32_t)__ldg(ak + WD / 2) & 1u, btop = (uint32_t)__ldg(bk + WD / 2) & 1u;
        const uint32_t ptop = atop ^ btop, gtop = atop & btop;
        if (k == 0) { // s_0 = p_0
            uint32_t *dst = Ov + 2 * lo.off[0];
            const uint32_t wo = 2 * (lo.off[1] - lo.off[0]);
            for (int j = 0; j < WD / 2; ++j) {
                const uint64_t x = __ldg(ak + j) ^ __ldg(bk + j);
                dst[2 * j] = (uint32_t)x;
                dst[2 * j + 1] = (uint32_t)(x >> 32);
            }
            dst[WD] = ptop;
            for (uint32_t j = WD + 1; j < wo; ++j) dst[j] = 0;
        }
        if (k + 1 == L) break; // no carry out of the last bit (common.rs:47-49)
        // ---- g = a b into mb[0 .. 64) (+ gtop), operands staged in the ring: a' at pairs [0, 16), b' at [16, 32) ----
        for (int j = 0; j < WD / 2; ++j) {
            const uint64_t x = __ldg(ak + j), y = __ldg(bk + j);
            ring[j * CTA] = make_uint2((uint32_t)x, (uint32_t)(x >> 32));
            ring[(16 + j) * CTA] = make_uint2((uint32_t)y, (uint32_t)(y >> 32));
        }
        for (int q = 0; q < WIDE_M_PAIRS; ++q) mb[q * CTA] = make_uint2(0u, 0u);
        smem_mul_acc8(ring, 4, ring + 16 * CTA, 4, mb);
        {
            const uint32_t ma = 0u - atop, mbm = 0u - btop; // X^1024 (at b' + bt a') into words [32, 64)
            for (int j = 0; j < WD / 2; ++j) {
                const uint2 xa = ring[j * CTA], xb = ring[(16 + j) * CTA];
                uint2 w = mb[(16 + j) * CTA];
                w.x ^= (xb.x & ma) ^ (xa.x & mbm);
                w.y ^= (xb.y & ma) ^ (xa.y & mbm);
                mb[(16 + j) * CTA] = w;
            }
        }
        // ---- slot k + 1 starts as g_k + p_{k+1}: the chain adds m_k c_k chunk by chunk on top of it (read back below) ----
        uint32_t *sdst = Ov + 2 * lo.off[k + 1];
        const uint32_t swo = 2 * (lo.off[k + 2] - lo.off[k + 1]);
        {
            const uint64_t *an = Av + (size_t)(k + 1) * WF, *bn = Bv + (size_t)(k + 1) * WF;
            for (int j = 0; j < 32; ++j) { // 64 words of g; the first 32 (+1) also take p_{k+1}
                uint2 w = mb[j * CTA];
                if (j < WD / 2) {
                    const uint64_t pn = __ldg(an + j) ^ __ldg(bn + j);
                    w.x ^= (uint32_t)pn;
                    w.y ^= (uint32_t)(pn >> 32);
                }
                if (j == WD / 2) w.x ^= (uint32_t)(__ldg(an + WD / 2) ^ __ldg(bn + WD / 2)) & 1u;
                *reinterpret_cast<uint2 *>(sdst + 2 * j) = w;
            }
            sdst[64] = gtop;
            const uint32_t init_end = (k == 0) ? swo : 72u; // k >= 1: the chain reads back chunks 0..2 (72 words) only
            for (uint32_t j = 65; j < init_end && j < swo; ++j) sdst[j] = 0;
        }
        if (k == 0) continue; // c_1 = g_0
        // ---- m = p + g p (96 words) into mb; g' moves to ring pairs [0, 32), p' to ring pairs [32, 48) ----
        for (int j = 0; j < 32; ++j) ring[j * CTA] = mb[j * CTA];
        for (int j = 0; j < WD / 2; ++j) {
            const uint64_t x = __ldg(ak + j) ^ __ldg(bk + j);
            ring[(32 + j) * CTA] = make_uint2((uint32_t)x, (uint32_t)(x >> 32));
        }
        {
            const uint32_t mp = 0u - ptop, mg = 0u - gtop;
            for (int j = 0; j < WIDE_M_PAIRS; ++j) { // p' + X^1024 (pt g' + pt) + X^2048 gt p'
                uint2 w = make_uint2(0u, 0u);
                if (j < 16) w = ring[(32 + j) * CTA];
                if (j >= 16) {
                    const uint2 gq = ring[(j - 16) * CTA];
                    w.x ^= gq.x & mp;
                    w.y ^= gq.y & mp;
                }
                if (j >= 32) {
                    const uint2 pq = ring[(32 + j - 32) * CTA];
                    w.x ^= pq.x & mg;
                    w.y ^= pq.y & mg;
                }
                if (j == 16) w.x ^= ptop;
                mb[j * CTA] = w;
            }
        }
        smem_mul_acc8(ring, 8, ring + 32 * CTA, 4, mb); // + g' p'
        // ---- chain ----
        const uint32_t *cslot = Ov + 2 * lo.off[k]; // s_k = c_k + p_k
        const uint32_t len = 96 * k - 31;           // words of c_k
        const uint32_t nch = 4 * k - 1;             // chunks of c_k; the last one holds 17 valid words
        auto ring_slot = [&](uint32_t j) -> uint2 * { return ring + (j % WIDE_RING) * NPR * CTA; };
        auto fill_sync = [&](uint32_t j) { // chunks 0 and 1 carry p_k (words 0..32), the last chunk is ragged
            uint2 *dstb = ring_slot(j);
            for (int q = 0; q < NPR; ++q) {
                uint32_t v2[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t w = CH * j + 2 * q + h;
                    uint32_t val = (w < len) ? cslot[w] : 0u;
                    if (w < (uint32_t)WD) {
                        const uint64_t x = __ldg(ak + (w >> 1)) ^ __ldg(bk + (w >> 1));
                        val ^= (w & 1) ? (uint32_t)(x >> 32) : (uint32_t)x;
                    } else if (w == (uint32_t)WD) {
                        val ^= ptop;
                    }
                    v2[h] = val;
                }
                dstb[q * CTA] = make_uint2(v2[0], v2[1]);
            }
        };
        auto fill_async = [&](uint32_t j) {
            const uint32_t sa = (uint32_t)__cvta_generic_to_shared(ring_slot(j));
            const uint32_t *src = cslot + CH * j;
#pragma unroll
            for (int q = 0; q < NPR; ++q)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa + q * CTA * 8), "l"(src + 2 * q) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        fill_sync(0);
        if (nch > 1) fill_sync(1);
        uint32_t t[2 * CH];
#pragma unroll
        for (int i = 0; i < 2 * CH; ++i) t[i] = 0;
        const uint32_t nout = nch + 4; // output chunks 0 .. nch + 3
        for (uint32_t o = 0; o < nout; ++o) {
            bool pref = false;
            if (o >= 1 && o + 1 < nch) { // chunk o + 1 replaces chunk o - 4 in the ring (chunks 0 and 1 are already there)
                if (o + 2 < nch) {
                    fill_async(o + 1);
                    pref = true;
                } else {
                    fill_sync(o + 1);
                }
            }
            if (pref) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll 1
            for (uint32_t q = 0; q < 4; ++q) {
                if (o >= q && o - q < nch) mul_chunk_acc<8>(mb + q * NPR * CTA, ring_slot(o - q), t);
            }
            if (o < 3) { // + g_k + p_{k+1}, parked in the slot itself
#pragma unroll
                for (int q = 0; q < NPR; ++q) {
                    const uint2 w = *reinterpret_cast<const uint2 *>(sdst + CH * o + 2 * q);
                    t[2 * q] ^= w.x;
                    t[2 * q + 1] ^= w.y;
                }
            }
#pragma unroll
            for (int q = 0; q < NPR; ++q)
                if (CH * o + 2 * q < swo) *reinterpret_cast<uint2 *>(sdst + CH * o + 2 * q) = make_uint2(t[2 * q], t[2 * q + 1]);
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                t[i] = t[CH + i];
                t[CH + i] = 0;
            }
        }
    }
}

template <int MINB>
__global__ void __launch_bounds__(CTA, MINB) adder_chain_wide_kernel(const uint64_t *__restrict__ A, const uint64_t *__restrict__ B,
                                                                    uint64_t *__restrict__ O, uint64_t n, uint32_t L, Layout lo, AdderSched sc) {
    constexpr int WF = 17;
    extern __shared__ __align__(16) uint2 chain_smem[];
    uint2 *mb = chain_smem + threadIdx.x, *ring = mb + WIDE_M_PAIRS * CTA;
    const int lane = threadIdx.x & 31;
    const uint32_t total = sc.ngroups * sc.nphases;
    for (;;) {
        uint32_t unit = 0;
        if (lane == 0) unit = atomicAdd(sc.counter, 1u);
        unit = __shfl_sync(FULL, unit, 0);
        if (unit >= total) break;
        const uint32_t ph = unit / sc.ngroups, grp = unit - ph * sc.ngroups;
        if (ph) {
            while (ld_acquire_u32(sc.done + grp) < ph) __nanosleep(256);
        }
        const uint64_t v = (uint64_t)grp * 32 + lane;
        if (v < n)
            chain_range_wide(A + v * (uint64_t)L * WF, B + v * (uint64_t)L * WF, reinterpret_cast<uint32_t *>(O + v * (uint64_t)lo.value_words), lo, L,
                             sc.kb[ph], sc.kb[ph + 1], mb, ring);
        if (ph + 1 < sc.nphases) {
            __threadfence();
            __syncwarp();
            if (lane == 0) st_release_u32(sc.done + grp, ph + 1);
        }
    }
}

} // namespace

size_t adder_chain_sched_words(uint64_t n) { return 1 + (size_t)((n + 31) / 32); }

void adder_chain_plan(uint32_t L, int wd, uint32_t want_phases, AdderSched *sc) {
    // work of iteration k: 1 product at k = 0, 3 + 6 k products for 1 <= k <= L - 2 (k = L - 1 only emits nothing new)
    (void)wd;
    if (want_phases < 1) want_phases = 1;
    if (want_phases > ADDER_MAX_PHASES) want_phases = ADDER_MAX_PHASES;
    double total = 0;
    for (uint32_t k = 0; k + 1 < L; ++k) total += (k == 0) ? 1.0 : 3.0 + 6.0 * k;
    uint32_t np = 0;
    sc->kb[0] = 0;
    double acc = 0;
    for (uint32_t k = 0; k < L; ++k) {
        acc += (k + 1 < L) ? ((k == 0) ? 1.0 : 3.0 + 6.0 * k) : 0.0;
        // close a phase after iteration k once it holds its share; a phase starts at k >= 1 only (slot k must exist)
        if (np + 1 < want_phases && k + 1 < L && k >= 1 && acc >= total * (np + 1) / want_phases) sc->kb[++np] = k + 1;
    }
    sc->kb[++np] = L;
    sc->nphases = np;
}

template <int WD, int TSM, int MINB>
static cudaError_t launch_one(const uint64_t *A, const uint64_t *B, uint64_t *O, uint64_t n, uint32_t L, const Layout &lo, const AdderSched &sc,
                              int sm_count, cudaStream_t stream) {
    auto kern = adder_chain_kernel<WD, TSM, MINB>;
    const size_t smem = (size_t)(TSM ? 5 : 3) * (3 * WD / 2) * CTA * 8;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const uint64_t warps = (uint64_t)sc.ngroups; // never more warps than groups
    uint64_t blocks = (uint64_t)sm_count * MINB;
    if (blocks * (CTA / 32) > warps) blocks = (warps + CTA / 32 - 1) / (CTA / 32);
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, CTA, smem, stream>>>(A, B, O, n, L, lo, sc);
    return cudaGetLastError();
}

cudaError_t launch_adder_chain(int wd, int variant, const uint64_t *A, const uint64_t *B, uint64_t *O, uint64_t n, uint32_t L, const Layout &lo,
                               const AdderSched &sc, int sm_count, cudaStream_t stream) {
    // variant: tens digit = window in shared memory (TSM), units digit = CTAs of 128 threads per SM
    const int tsm = variant / 10, minb = variant % 10;
    if (wd == 8) {
        if (tsm == 0 && minb == 4) return launch_one<8, 0, 4>(A, B, O, n, L, lo, sc, sm_count, stream);
        if (tsm == 0 && minb == 3) return launch_one<8, 0, 3>(A, B, O, n, L, lo, sc, sm_count, stream);
        if (tsm == 1 && minb == 3) return launch_one<8, 1, 3>(A, B, O, n, L, lo, sc, sm_count, stream);
        if (tsm == 1 && minb == 2) return launch_one<8, 1, 2>(A, B, O, n, L, lo, sc, sm_count, stream);
    } else if (wd == 4) {
        if (tsm == 0) return launch_one<4, 0, 4>(A, B, O, n, L, lo, sc, sm_count, stream);
        return launch_one<4, 1, 4>(A, B, O, n, L, lo, sc, sm_count, stream);
    } else if (wd == 32) { // D = 1024: four sub-multipliers, 2 CTAs of 128 threads per SM
        auto kern = adder_chain_wide_kernel<2>;
        const size_t smem = (size_t)WIDE_PAIRS * CTA * 8;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        uint64_t blocks = (uint64_t)sm_count * 2;
        if (blocks * (CTA / 32) > sc.ngroups) blocks = ((uint64_t)sc.ngroups + CTA / 32 - 1) / (CTA / 32);
        if (blocks < 1) blocks = 1;
        kern<<<(unsigned)blocks, CTA, smem, stream>>>(A, B, O, n, L, lo, sc);
        return cudaGetLastError();
    }
    return cudaErrorInvalidValue;
}

} // namespace hmk
