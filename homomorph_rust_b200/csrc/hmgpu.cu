// hmgpu.cu — C ABI (include/hmgpu.h) over the sm_100a kernels in kernels.cuh.
//
// Host-side responsibilities, mirroring the reference's thin layers L1-L3:
//   - Parameters / key containers and their byte formats      reference src/context.rs
//   - per-key constants (subset tables, decrypt vector, remainder folding tables)
//   - batch bookkeeping (slot widths + degree bounds) and circuit planning
// All polynomial arithmetic on batches runs on the GPU; there is no CPU fallback.
#include "../../include/hmgpu.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <set>
#include <string>
#include <vector>

#include "gf2host.hpp"
#include "kernels.cuh"
#include "kernels_adder.h"
#include "kernels_mul.h"
#include "kernels_enc_umma.h"
#include "probes.h"

namespace hmk {
cudaError_t launch_mulrem_fresh_b32(const uint64_t *A, const uint64_t *B, uint64_t *O, uint64_t pairs, const uint32_t *Tg,
                                    int sm_count, cudaStream_t stream, int reduce_first); // kernels_b.cu
}

// tuning knobs (hm_set_tuning): minimum batch size for the thread-per-value adder; < 0 = default (96 values per SM)
static long g_adder_thread_min = -1;
static long g_mul_thread_min = -1; // minimum (values x chunks) for the thread-per-chunk multiply; < 0 = default (128 per SM)
static long g_mul_thread_chunk = 32; // 24 = the first thread-per-chunk kernel (3-way Karatsuba chunks)
static long g_mul_circuit_seq = 0;
static long g_mul_circuit_fused = getenv("HM_MUL_FUSED") ? atol(getenv("HM_MUL_FUSED")) : 1; // 1 = one-launch fused column multiplier where it applies (u8, D = 256)
static long g_adder_chain = getenv("HM_ADDER_CHAIN") ? atol(getenv("HM_ADDER_CHAIN")) : 4;   // 0 = round-1 thread kernels; else 10 * (window in smem) + CTAs per SM
static long g_adder_wide_min = getenv("HM_ADDER_WIDE_MIN") ? atol(getenv("HM_ADDER_WIDE_MIN")) : 0; // D = 1024 chain kernel from this many values on (0 = 256 per SM, < 0 = never)
static long g_adder_phases = getenv("HM_ADDER_PHASES") ? atol(getenv("HM_ADDER_PHASES")) : 0; // work units per value of the scheduled adder chain; 0 = by batch size
static long g_host_chunk_mb = getenv("HM_HOST_CHUNK_MB") ? atol(getenv("HM_HOST_CHUNK_MB")) : 96; // per-stage bytes of the host-buffer pipeline
static long g_pool_max_mb = getenv("HM_POOL_MAX_MB") ? atol(getenv("HM_POOL_MAX_MB")) : 16384; // larger batches use plain cudaMalloc
static long g_adder_generic_seq = 0; // 1 = the generic adder evaluates the reference's formula literally (two long products per bit)  // 1 = launch the multiplier circuit's carry products one by one (the first plan)

using hmk::Layout;
using hmk::MulOp;
using hmk::View;

// ------------------------------------------------------------------------------------------
struct hm_batch {
    hm_context *ctx = nullptr;
    size_t n = 0;
    uint32_t L = 0;
    std::vector<uint32_t> w;    // slot widths, u64 words
    std::vector<uint32_t> off;  // L+1 prefix sums
    std::vector<uint64_t> degb; // per-slot degree bound (>= true degree of every polynomial in the slot)
    size_t value_words = 0;
    uint64_t *d = nullptr;
    bool pooled = false; // allocated from the stream-ordered pool (else plain cudaMalloc)
};

struct hm_context {
    uint16_t d = 0, dp = 0, delta = 0, tau = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t side_stream = nullptr; // runs the small products of a multiplier column beside the big ones
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    uint64_t launches = 0;
    std::string last_error;
    int sm_count = 148;
    size_t smem_optin = 0;

    // secret key and what is derived from it
    bool has_sk = false;
    gf2::words S;
    size_t ds = 0; // true degree of S
    gf2::words v_host;
    size_t v_bits = 0;
    uint64_t *d_v = nullptr; // decrypt vector on device, v_bits bits
    std::vector<uint32_t> vv_layout_key;
    uint64_t *d_vv = nullptr;
    size_t d_vv_words = 0;
    uint32_t *d_remT = nullptr; // folding tables (ds % 32 == 0)
    uint32_t *d_S32 = nullptr;  // S as u32 words (generic remainder)

    // public key and what is derived from it
    bool has_pk = false;
    std::vector<gf2::words> T_host; // the tau public polynomials as set (for hm_public_key_bytes)
    size_t fresh_deg = 0; // max degree over the T_i (== d+dp for generated keys)
    uint32_t wf = 0;      // u64 words of a fresh slot
    uint32_t enc_wb = 0, enc_groups = 0, enc_table_words = 0;
    bool enc_table_in_smem = false;
    uint64_t *d_enc_table = nullptr;
    uint64_t *d_enc_table6 = nullptr; // bank-partitioned copy for encrypt_tab6_kernel (config A shape only)
    uint64_t *d_enc_table4 = nullptr; // four-class 32-byte-row copy for encrypt_tab4_kernel (config A shape only)
    uint32_t enc_topmask[8] = {0, 0, 0, 0, 0, 0, 0, 0}; // bit i = leading (X^D) coefficient of T_i: the last word of a row is computed, not looked up
    uint64_t *d_enc_table4b = nullptr; // 128-byte-row copy for encrypt_tab4b_kernel (config B shape only)
    int8_t *d_enc_umma = nullptr;      // B operand of the tensor-core encrypt kernel (config B shape only)

    // ring of op descriptors: pinned host staging + device copy, so that launches need no host synchronisation
    MulOp *d_ops = nullptr;
    MulOp *h_ops = nullptr;
    size_t d_ops_cap = 0;
    size_t ops_cursor = 0;

    // work queue of the dynamically scheduled adder chain (kernels_adder.cu): counter + one flag per 32 values
    uint32_t *d_sched = nullptr;
    size_t sched_words = 0;

    // fused column multiplier (kernels_mul.cu): the static plan for fresh u8 operands at D = 256, uploaded once
    hmk::K7Plan k7_plan;
    bool k7_ready = false;
    hmk::K7Item *d_k7_items = nullptr;
    hmk::K7Prod *d_k7_prods = nullptr;
    hmk::K7Unit *d_k7_units = nullptr;

    cudaMemPool_t pool = nullptr;  // private stream-ordered pool: nothing process-wide is reconfigured
    cudaMemPool_t pool_big = nullptr; // second pool for batches above 1 GiB (results of circuits on big batches), so that small
                                      // allocations cannot be carved out of a freed multi-GiB block and force the next big one
                                      // to map fresh memory (seen: 350 ms instead of 62 ms per 11.45 GiB result)
    std::set<hm_batch *> live;     // batches still owned by callers; orphaned (device memory released) by hm_context_destroy
};

namespace {

int fail_cuda(hm_context *ctx, cudaError_t e, const char *what) {
    if (ctx) {
        char buf[256];
        snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
        ctx->last_error = buf;
    }
    return HM_ERR_CUDA;
}
#define CK(call)                                                   \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return fail_cuda(ctx, e__, #call); \
    } while (0)

// stream-ordered allocation from the context's own pool, in ctx->stream order
cudaError_t pool_alloc(hm_context *ctx, void **p, size_t bytes) {
    if (ctx->pool) return cudaMallocFromPoolAsync(p, bytes, ctx->pool, ctx->stream);
    return cudaMallocAsync(p, bytes, ctx->stream);
}
template <typename T> cudaError_t pool_alloc(hm_context *ctx, T **p, size_t bytes) { return pool_alloc(ctx, reinterpret_cast<void **>(p), bytes); }

// frees a pool allocation on every exit path of the enclosing scope (in the order of the stream it was allocated on)
struct PoolGuard {
    hm_context *ctx;
    cudaStream_t stream;
    void *p = nullptr;
    explicit PoolGuard(hm_context *c) : ctx(c), stream(c->stream) {}
    PoolGuard(const PoolGuard &) = delete;
    PoolGuard &operator=(const PoolGuard &) = delete;
    ~PoolGuard() {
        if (p) cudaFreeAsync(p, stream);
    }
};

int use_device(hm_context *ctx) {
    CK(cudaSetDevice(ctx->device));
    return HM_OK;
}
#define USE_DEV(ctx)                    \
    do {                                \
        int rc__ = use_device(ctx);     \
        if (rc__ != HM_OK) return rc__; \
    } while (0)

void zero_free_device(hm_context *ctx, void *p, size_t bytes) {
    if (!p) return;
    if (bytes) cudaMemsetAsync(p, 0, bytes, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(p);
}

Layout make_layout(const hm_batch *b) {
    Layout l;
    l.L = b->L;
    l.value_words = (uint32_t)b->value_words;
    for (uint32_t k = 0; k <= b->L; ++k) l.off[k] = b->off[k];
    return l;
}

// Largest degree bound a slot may carry (result_bounds refuses anything above it too) and the widest value the 32-bit
// slot offsets of the kernels' Layout can describe.
constexpr uint64_t MAX_DEGREE_BOUND = (uint64_t)1 << 31;
constexpr uint64_t MAX_VALUE_WORDS = ((uint64_t)1 << 32) - 1;

// Sum of the slot widths of `degb`, or 0 if a bound or the sum is out of range.
uint64_t checked_value_words(uint32_t L, const uint64_t *degb) {
    uint64_t o = 0;
    for (uint32_t k = 0; k < L; ++k) {
        if (degb[k] > MAX_DEGREE_BOUND) return 0;
        o += degb[k] / 64 + 1;
        if (o > MAX_VALUE_WORDS) return 0;
    }
    return o;
}

// *status: HM_ERR_INVALID_ARGUMENT (bounds / sizes out of range) or HM_ERR_OUT_OF_MEMORY (host allocation failed)
hm_batch *new_batch(hm_context *ctx, size_t n, uint32_t L, const uint64_t *degb, int *status = nullptr) {
    int dummy;
    int &st = status ? *status : dummy;
    st = HM_ERR_INVALID_ARGUMENT;
    if (L == 0 || L > (uint32_t)hmk::MAX_SLOTS) return nullptr;
    const uint64_t vw = checked_value_words(L, degb);
    if (vw == 0) return nullptr;
    if (n > (SIZE_MAX / 8) / vw) return nullptr; // n * value_words * 8 must fit a size_t
    st = HM_ERR_OUT_OF_MEMORY;
    hm_batch *b = new (std::nothrow) hm_batch;
    if (!b) return nullptr;
    try {
        b->w.resize(L);
        b->off.resize(L + 1);
        b->degb.assign(degb, degb + L);
        ctx->live.insert(b);
    } catch (const std::bad_alloc &) {
        delete b;
        return nullptr;
    }
    b->ctx = ctx;
    b->n = n;
    b->L = L;
    uint64_t o = 0;
    for (uint32_t k = 0; k < L; ++k) {
        b->off[k] = (uint32_t)o;
        b->w[k] = (uint32_t)(degb[k] / 64 + 1);
        o += b->w[k];
    }
    b->off[L] = (uint32_t)o;
    b->value_words = o;
    st = HM_OK;
    return b;
}

// drops a batch that never got (or no longer has) device memory
void discard_batch(hm_batch *b) {
    if (!b) return;
    if (b->ctx) b->ctx->live.erase(b);
    delete b;
}

int alloc_batch(hm_context *ctx, hm_batch *b) {
    const size_t bytes = std::max<size_t>(b->n * b->value_words * 8, 16);
    // Stream-ordered pool for ordinary batches (no device-wide sync, blocks are recycled).  Multi-GB results go through
    // plain cudaMalloc: returning such a block to the pool costs ~0.5 s of unmapping at the next synchronisation.
    cudaError_t e;
    if (bytes <= ((size_t)(g_pool_max_mb > 0 ? g_pool_max_mb : 0) << 20)) {
        if (bytes > ((size_t)1 << 30) && ctx->pool_big) e = cudaMallocFromPoolAsync(reinterpret_cast<void **>(&b->d), bytes, ctx->pool_big, ctx->stream);
        else e = pool_alloc(ctx, &b->d, bytes);
        b->pooled = true;
    } else {
        e = cudaMalloc(&b->d, bytes);
        b->pooled = false;
    }
    if (e != cudaSuccess) {
        b->d = nullptr;
        fail_cuda(ctx, e, "batch allocation");
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? HM_ERR_OUT_OF_MEMORY : HM_ERR_CUDA;
    }
    return HM_OK;
}

int grid_for(hm_context *ctx, uint64_t work_items, int threads, int per_sm) {
    uint64_t blocks = (work_items + threads - 1) / threads;
    const uint64_t cap = (uint64_t)ctx->sm_count * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int post_launch(hm_context *ctx, const char *what) {
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail_cuda(ctx, e, what);
    return HM_OK;
}
#define LAUNCHED(what)                        \
    do {                                      \
        int rc__ = post_launch(ctx, what);    \
        if (rc__ != HM_OK) return rc__;       \
    } while (0)

// ---- per-key constants ---------------------------------------------------------------------

int ensure_decrypt_vector(hm_context *ctx, size_t nbits) {
    if (ctx->d_v && ctx->v_bits >= nbits) return HM_OK;
    size_t want = std::max<size_t>(nbits, 1024);
    ctx->v_host = gf2::decrypt_vector(ctx->S, ctx->ds, want);
    if (ctx->d_v) zero_free_device(ctx, ctx->d_v, (ctx->v_bits + 63) / 64 * 8);
    ctx->d_v = nullptr;
    CK(cudaMalloc(&ctx->d_v, ctx->v_host.size() * 8));
    CK(cudaMemcpyAsync(ctx->d_v, ctx->v_host.data(), ctx->v_host.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->v_bits = want;
    ctx->vv_layout_key.clear();
    return HM_OK;
}

// vv: the decrypt vector laid out like one value of the batch.
int ensure_vv(hm_context *ctx, const hm_batch *b) {
    size_t maxbits = 0;
    for (uint32_t k = 0; k < b->L; ++k) maxbits = std::max<size_t>(maxbits, (size_t)b->w[k] * 64);
    int rc = ensure_decrypt_vector(ctx, maxbits);
    if (rc != HM_OK) return rc;
    if (ctx->d_vv && ctx->vv_layout_key == b->w) return HM_OK;
    std::vector<uint64_t> vv(b->value_words);
    for (uint32_t k = 0; k < b->L; ++k)
        for (uint32_t j = 0; j < b->w[k]; ++j) vv[b->off[k] + j] = ctx->v_host[j];
    if (ctx->d_vv_words < vv.size()) {
        if (ctx->d_vv) zero_free_device(ctx, ctx->d_vv, ctx->d_vv_words * 8);
        ctx->d_vv = nullptr;
        CK(cudaMalloc(&ctx->d_vv, vv.size() * 8));
        ctx->d_vv_words = vv.size();
    }
    CK(cudaMemcpyAsync(ctx->d_vv, vv.data(), vv.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::fill(vv.begin(), vv.end(), 0);
    ctx->vv_layout_key = b->w;
    return HM_OK;
}

void clear_secret(hm_context *ctx) {
    if (ctx->d_v) zero_free_device(ctx, ctx->d_v, (ctx->v_bits + 63) / 64 * 8 + 8);
    if (ctx->d_vv) zero_free_device(ctx, ctx->d_vv, ctx->d_vv_words * 8);
    if (ctx->d_remT) zero_free_device(ctx, ctx->d_remT, 4 * 256 * (ctx->ds / 32) * 4);
    if (ctx->d_S32) zero_free_device(ctx, ctx->d_S32, (ctx->ds / 32 + 2) * 4);
    ctx->d_v = ctx->d_vv = nullptr;
    ctx->d_remT = ctx->d_S32 = nullptr;
    ctx->d_vv_words = 0;
    ctx->v_bits = 0;
    // volatile-style wipe of host copies (reference src/polynomial.rs:379-401)
    volatile uint64_t *p = ctx->S.data();
    for (size_t i = 0; i < ctx->S.size(); ++i) p[i] = 0;
    volatile uint64_t *q = ctx->v_host.data();
    for (size_t i = 0; i < ctx->v_host.size(); ++i) q[i] = 0;
    ctx->S.clear();
    ctx->v_host.clear();
    ctx->vv_layout_key.clear();
    ctx->has_sk = false;
    ctx->ds = 0;
}

void clear_public(hm_context *ctx) {
    if (ctx->d_enc_table) cudaFree(ctx->d_enc_table);
    if (ctx->d_enc_table6) cudaFree(ctx->d_enc_table6);
    if (ctx->d_enc_table4) cudaFree(ctx->d_enc_table4);
    ctx->d_enc_table4 = nullptr;
    if (ctx->d_enc_table4b) cudaFree(ctx->d_enc_table4b);
    ctx->d_enc_table4b = nullptr;
    if (ctx->d_enc_umma) cudaFree(ctx->d_enc_umma);
    ctx->d_enc_umma = nullptr;
    for (uint32_t &m : ctx->enc_topmask) m = 0;
    ctx->d_enc_table = nullptr;
    ctx->d_enc_table6 = nullptr;
    ctx->T_host.clear();
    ctx->has_pk = false;
    ctx->enc_table_words = 0;
}

// Reserves n_ops consecutive slots of the descriptor ring and returns the first slot.  Slots are reused only after a
// wrap-around, which synchronises the stream (every earlier kernel has then consumed its descriptors).
int ensure_ops(hm_context *ctx, size_t n_ops, size_t *first_slot) {
    constexpr size_t RING = 16384;
    if (!ctx->d_ops || ctx->d_ops_cap < n_ops) {
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaStreamSynchronize(ctx->side_stream));
        if (ctx->d_ops) cudaFree(ctx->d_ops);
        if (ctx->h_ops) cudaFreeHost(ctx->h_ops);
        ctx->d_ops = nullptr;
        ctx->h_ops = nullptr;
        const size_t cap = std::max<size_t>(n_ops, RING);
        CK(cudaMalloc(&ctx->d_ops, cap * sizeof(MulOp)));
        CK(cudaHostAlloc(&ctx->h_ops, cap * sizeof(MulOp), cudaHostAllocDefault));
        ctx->d_ops_cap = cap;
        ctx->ops_cursor = 0;
    }
    if (ctx->ops_cursor + n_ops > ctx->d_ops_cap) {
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaStreamSynchronize(ctx->side_stream));
        ctx->ops_cursor = 0;
    }
    *first_slot = ctx->ops_cursor;
    ctx->ops_cursor += n_ops;
    return HM_OK;
}

// ---- kernel launch wrappers --------------------------------------------------------------------

// Shape class of a product: 0 = general (warp per product); otherwise NX*100+NY (32-bit words) for the
// thread-per-product Karatsuba kernels, which need "whole blocks + the single top coefficient" operands.
int mul_shape_class(const MulOp &o) {
    auto blocks = [](const View &v) -> int { // low words if deg == 64 * (w - 1) and that is 256, 512 or 1024
        if (v.deg != (uint64_t)64 * (v.w - 1)) return 0;
        if (v.deg == 256) return 8;
        if (v.deg == 512) return 16;
        if (v.deg == 1024) return 32;
        return 0;
    };
    const int nx = blocks(o.a), ny = blocks(o.b);
    if (!nx || !ny) return 0;
    if (o.o.w < (uint32_t)((nx + ny) / 2 + 1)) return 0;
    if ((nx == 8) != (ny == 8)) return 0; // 8-word blocks only pair with 8-word blocks
    if (nx == 32 && ny == 32) return 0;    // too many live registers for one thread: warp kernel
    return nx * 100 + ny;
}

int launch_xor_views(hm_context *ctx, View o, View a, View b, size_t n);
View null_view();

int launch_mul_class(hm_context *ctx, int cls, const MulOp *d_ops, const MulOp *h_ops, size_t cnt, size_t n, size_t smem_general,
                     uint32_t per_warp_general, bool outputs_zeroed, int fuse_or) {
    // thread-per-chunk kernels: chunks of the shorter operand (32-word chunks see "low words + top coefficient" operands)
    const bool chunk32 = g_mul_thread_chunk != 24;
    uint32_t xchunks_max = 0;
    uint64_t xchunks_sum = 0;
    for (size_t i = 0; i < cnt; ++i) {
        const uint32_t xc = chunk32 ? (hmk::thread_mul_shape(h_ops[i]).nx + 31) / 32 : (2 * std::min(h_ops[i].a.w, h_ops[i].b.w) + 23) / 24;
        xchunks_max = std::max(xchunks_max, xc);
        xchunks_sum += xc;
    }
    const dim3 grid_t((unsigned)((n + 127) / 128), (unsigned)cnt);
    switch (cls) {
        case 808: hmk::mul_small_kernel<8, 8><<<grid_t, 128, 0, ctx->stream>>>(d_ops, n, fuse_or); break;
        case 1616: hmk::mul_small_kernel<16, 16><<<grid_t, 128, 0, ctx->stream>>>(d_ops, n, fuse_or); break;
        case 1632: hmk::mul_small_kernel<16, 32><<<grid_t, 128, 0, ctx->stream>>>(d_ops, n, fuse_or); break;
        case 3216: hmk::mul_small_kernel<32, 16><<<grid_t, 128, 0, ctx->stream>>>(d_ops, n, fuse_or); break;
        default: {
            // enough (value, chunk) pairs to fill the GPU with one thread each?  -> Karatsuba thread kernel
            static const int no_thread = getenv("HM_MUL_NO_THREAD") ? atoi(getenv("HM_MUL_NO_THREAD")) : 0;
            const bool must_thread = smem_general == 0; // the warp kernel cannot stage these operands
            auto xchunks_of = [&](size_t i) -> uint32_t {
                return chunk32 ? (hmk::thread_mul_shape(h_ops[i]).nx + 31) / 32 : (2 * std::min(h_ops[i].a.w, h_ops[i].b.w) + 23) / 24;
            };
            const uint64_t rows = (uint64_t)((n + 127) / 128) * 128; // threads along x
            const uint64_t max_threads = (uint64_t)1 << 22;          // 1 GiB of per-thread scratch at most per launch
            const bool fits_one = (uint64_t)n * cnt * xchunks_max <= max_threads;
            if ((must_thread || (fits_one && !no_thread && (uint64_t)n * xchunks_sum >= (g_mul_thread_min >= 0 ? (uint64_t)g_mul_thread_min : (uint64_t)ctx->sm_count * 128))) &&
                rows * xchunks_max <= max_threads && cnt <= 65535 && xchunks_max <= 65535) {
                // outputs are accumulated with atomics: zero them first (unless the caller already did)
                for (size_t i = 0; i < cnt && !outputs_zeroed; ++i) {
                    const MulOp &o = h_ops[i];
                    if (o.o.stride == o.o.w) {
                        CK(cudaMemsetAsync(o.o.base + o.o.off, 0, (size_t)n * o.o.w * 8, ctx->stream));
                    } else {
                        int zrc = launch_xor_views(ctx, o.o, null_view(), null_view(), n);
                        if (zrc != HM_OK) return zrc;
                    }
                }
                // one launch per run of products whose (rows x products x widest chunk count) stays within the scratch budget
                for (size_t first = 0; first < cnt;) {
                    size_t last = first;
                    uint32_t xm = 0;
                    while (last < cnt) {
                        const uint32_t xn = std::max(xm, xchunks_of(last));
                        if (last > first && rows * (last - first + 1) * xn > max_threads) break;
                        xm = xn;
                        ++last;
                    }
                    const size_t threads = (size_t)rows * (last - first) * xm;
                    PoolGuard scratch_guard(ctx);
                    CK(pool_alloc(ctx, &scratch_guard.p, threads * hmk::ADT_THREAD_WORDS * 4));
                    uint32_t *scratch = static_cast<uint32_t *>(scratch_guard.p);
                    const dim3 grid_t3((unsigned)(rows / 128), (unsigned)(last - first), (unsigned)xm);
                    if (chunk32) hmk::mul_thread32_kernel<<<grid_t3, 128, 0, ctx->stream>>>(d_ops + first, n, scratch);
                    else hmk::mul_thread_kernel<<<grid_t3, 128, 0, ctx->stream>>>(d_ops + first, n, scratch);
                    if (last < cnt) {
                        int prc = post_launch(ctx, "mul_thread kernel");
                        if (prc != HM_OK) return prc;
                    }
                    first = last;
                }
                break;
            }
            if (must_thread) return HM_ERR_UNSUPPORTED;
            const dim3 grid_w((unsigned)((n + 3) / 4), (unsigned)cnt);
            static const int old_generic = getenv("HM_MUL_OLD") ? atoi(getenv("HM_MUL_OLD")) : 0;
            if (old_generic) {
                CK(cudaFuncSetAttribute(hmk::mul_views_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_general));
                hmk::mul_views_kernel<<<grid_w, hmk::MUL_WARPS * 32, smem_general, ctx->stream>>>(d_ops, n, per_warp_general);
            } else {
                CK(cudaFuncSetAttribute(hmk::mul_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_general));
                hmk::mul_warp_kernel<<<grid_w, 128, smem_general, ctx->stream>>>(d_ops, n, per_warp_general);
            }
        }
    }
    return post_launch(ctx, "mul kernel");
}

// fused_or: in/out.  On entry non-zero asks for a + b + a*b; on return it says whether every product went through a kernel
// that fused it (the caller adds a and b itself otherwise).
int launch_mul_ops(hm_context *ctx, const std::vector<MulOp> &ops_in, size_t n, bool outputs_zeroed = false, int *fused_or = nullptr) {
    if (ops_in.empty() || n == 0) return HM_OK;
    // group by shape class (stable), one launch per class
    std::vector<MulOp> ops(ops_in);
    static const int no_small = getenv("HM_MUL_NO_SMALL") ? atoi(getenv("HM_MUL_NO_SMALL")) : 0;
    std::vector<int> cls(ops.size());
    for (size_t i = 0; i < ops.size(); ++i) cls[i] = no_small ? 0 : mul_shape_class(ops[i]);
    std::vector<size_t> order(ops.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return cls[x] < cls[y]; });
    std::vector<MulOp> sorted(ops.size());
    std::vector<int> scls(ops.size());
    for (size_t i = 0; i < order.size(); ++i) {
        sorted[i] = ops[order[i]];
        scls[i] = cls[order[i]];
    }
    int fuse_or = 0;
    if (fused_or && *fused_or) { // only the register-resident kernels can fuse the two extra XORs
        fuse_or = 1;
        for (int c : scls) fuse_or &= (c != 0);
        *fused_or = fuse_or;
    }
    size_t slot = 0;
    int rc = ensure_ops(ctx, sorted.size(), &slot);
    if (rc != HM_OK) return rc;
    // stream-ordered and asynchronous: descriptors go through a pinned ring, so neither the copy nor the kernels
    // that read them need the host to wait
    memcpy(ctx->h_ops + slot, sorted.data(), sorted.size() * sizeof(MulOp));
    CK(cudaMemcpyAsync(ctx->d_ops + slot, ctx->h_ops + slot, sorted.size() * sizeof(MulOp), cudaMemcpyHostToDevice, ctx->stream));
    for (size_t first = 0; first < sorted.size();) {
        size_t last = first;
        while (last < sorted.size() && scls[last] == scls[first] && last - first < 65535) ++last;
        uint32_t per_warp = 0;
        size_t smem = 0;
        if (scls[first] == 0) {
            for (size_t i = first; i < last; ++i) {
                const uint32_t na = 2 * std::min(sorted[i].a.w, sorted[i].b.w), nc = 2 * std::max(sorted[i].a.w, sorted[i].b.w);
                const uint32_t nch = (na + hmk::MW_J - 1) / hmk::MW_J;
                per_warp = std::max(per_warp, std::max(nch * 32 + nch * hmk::MW_J + nc, na + nc));
            }
            per_warp = (per_warp + 3) & ~3u;
            smem = (size_t)per_warp * 4 * 4;
            // operands too long for the warp kernel's shared-memory staging (u16 multiplier carries: tens of thousands of words):
            // smem = 0 tells launch_mul_class to take the thread-per-chunk kernel whatever the batch size
            if (smem > ctx->smem_optin) smem = 0;
        }
        rc = launch_mul_class(ctx, scls[first], ctx->d_ops + slot + first, sorted.data() + first, last - first, n, smem, per_warp,
                              outputs_zeroed, fuse_or);
        if (rc != HM_OK) return rc;
        first = last;
    }
    return HM_OK;
}

int launch_xor_views(hm_context *ctx, View o, View a, View b, size_t n) {
    if (n == 0) return HM_OK;
    const int grid = grid_for(ctx, (uint64_t)n * o.w, 256, 16);
    hmk::xor_views_kernel<<<grid, 256, 0, ctx->stream>>>(o, a, b, n);
    LAUNCHED("xor_views_kernel");
    return HM_OK;
}

int launch_rem(hm_context *ctx, View a, View o, size_t n) {
    if (n == 0) return HM_OK;
    const size_t ds = ctx->ds;
    const bool fold = (ds % 32 == 0) && (ds / 32 == 2 || ds / 32 == 4 || ds / 32 == 8 || ds / 32 == 16);
    if (fold) {
        const int ws = (int)(ds / 32);
        const size_t smem = (size_t)4 * 256 * (ws >= 8 ? ws + 4 : ws) * 4; // rem_fold_row_stride
        const unsigned grid = (unsigned)((n + 127) / 128);
#define REM_CASE(WS)                                                                                                  \
    case WS:                                                                                                          \
        CK(cudaFuncSetAttribute(hmk::rem_fold_kernel<WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
        hmk::rem_fold_kernel<WS><<<grid, 128, smem, ctx->stream>>>(a, o, n, ctx->d_remT);                             \
        break;
        switch (ws) {
            REM_CASE(2)
            REM_CASE(4)
            REM_CASE(8)
            REM_CASE(16)
        }
#undef REM_CASE
        LAUNCHED("rem_fold_kernel");
        return HM_OK;
    }
    uint32_t per_warp = (2 * a.w + 1 + 3) & ~3u;
    const size_t smem = (size_t)per_warp * 4 * hmk::MUL_WARPS;
    if (smem > ctx->smem_optin) return HM_ERR_UNSUPPORTED;
    CK(cudaFuncSetAttribute(hmk::rem_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((n + hmk::MUL_WARPS - 1) / hmk::MUL_WARPS);
    hmk::rem_generic_kernel<<<grid, hmk::MUL_WARPS * 32, smem, ctx->stream>>>(a, o, n, ctx->d_S32, (uint32_t)ds, per_warp);
    LAUNCHED("rem_generic_kernel");
    return HM_OK;
}

View slot_view(const hm_batch *b, uint32_t k) {
    View v;
    v.base = b->d;
    v.stride = b->value_words;
    v.off = b->off[k];
    v.w = b->w[k];
    v.deg = b->degb[k];
    return v;
}
View null_view() {
    View v;
    v.base = nullptr;
    v.stride = 0;
    v.off = 0;
    v.w = 0;
    v.deg = 0;
    return v;
}

bool same_layout(const hm_batch *a, const hm_batch *b) { return a->L == b->L && a->w == b->w; }

int xor_batches(hm_context *ctx, const hm_batch *a, const hm_batch *b, hm_batch *o) {
    if (a->n == 0) return HM_OK;
    if (same_layout(a, b) && same_layout(a, o)) {
        const uint64_t total = (uint64_t)a->n * a->value_words;
        const uint64_t n16 = total / 2;
        if (n16) {
            const int grid = grid_for(ctx, n16, 256, 16);
            hmk::xor_flat_kernel<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<const uint4 *>(a->d),
                                                                 reinterpret_cast<const uint4 *>(b->d),
                                                                 reinterpret_cast<uint4 *>(o->d), n16);
            LAUNCHED("xor_flat_kernel");
        }
        if (total & 1) {
            hmk::xor_flat_tail_kernel<<<1, 32, 0, ctx->stream>>>(a->d, b->d, o->d, n16 * 2, total);
            LAUNCHED("xor_flat_tail_kernel");
        }
        return HM_OK;
    }
    const int grid = grid_for(ctx, (uint64_t)o->n * o->value_words, 256, 16);
    hmk::xor_layout_kernel<<<grid, 256, 0, ctx->stream>>>(a->d, make_layout(a), b->d, make_layout(b), o->d, make_layout(o),
                                                           o->n);
    LAUNCHED("xor_layout_kernel");
    return HM_OK;
}

// ---- circuit planning on an arena of per-value polynomials ------------------------------------
struct PolyRef {
    uint64_t *base;
    uint64_t stride;
    uint32_t off;
    uint32_t w_alloc;
    uint64_t degb;
    bool zero; // known to be the null polynomial for every value
    View view() const {
        View v;
        v.base = base;
        v.stride = stride;
        v.off = off;
        v.w = zero ? 1u : (uint32_t)std::min<uint64_t>(w_alloc, degb / 64 + 1);
        v.deg = zero ? 0 : degb;
        return v;
    }
};

} // namespace

// =============================================================================================
extern "C" {

const char *hm_status_string(int s) {
    switch (s) {
        case HM_OK: return "ok";
        case HM_ERR_INVALID_PARAMETERS: return "invalid parameters (d, dp, delta, tau must be > 0 and delta < d)";
        case HM_ERR_PUBLIC_KEY_UNSET: return "public key unset";
        case HM_ERR_SECRET_KEY_UNSET: return "secret key unset";
        case HM_ERR_OPERATION_REQUIREMENT: return "operation requirement not met (d < MIN_D_OVER_DELTA * delta)";
        case HM_ERR_INVALID_LENGTH: return "invalid ciphered length (not a multiple of 8 bits)";
        case HM_ERR_CUDA: return "CUDA error or no CUDA device";
        case HM_ERR_UNSUPPORTED: return "shape not supported by the kernels";
        case HM_ERR_INVALID_ARGUMENT: return "invalid argument";
        case HM_ERR_DIVIDE_BY_ZERO: return "attempt to divide by zero";
        case HM_ERR_OUT_OF_MEMORY: return "out of host or device memory";
        default: return "unknown status";
    }
}

const char *hm_last_error(const hm_context *ctx) { return ctx ? ctx->last_error.c_str() : ""; }

int hm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int hm_context_create(uint16_t d, uint16_t dp, uint16_t delta, uint16_t tau, int device, hm_context **out) {
    if (!out) return HM_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    // Parameters::new asserts, reference src/context.rs:87-94
    if (d == 0 || dp == 0 || delta == 0 || tau == 0 || delta >= d) return HM_ERR_INVALID_PARAMETERS;
    if (device < 0 || device >= hm_device_count()) return HM_ERR_CUDA;
    hm_context *ctx = new (std::nothrow) hm_context;
    if (!ctx) return HM_ERR_INVALID_ARGUMENT;
    ctx->d = d;
    ctx->dp = dp;
    ctx->delta = delta;
    ctx->tau = tau;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return HM_ERR_CUDA;
    }
    ctx->own_stream = true;
    if (cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        delete ctx;
        return HM_ERR_CUDA;
    }
    {   // a private stream-ordered pool that keeps freed blocks (no trimming at every synchronisation); the device's default
        // pool, which the rest of the process may be using, is left alone
        cudaMemPoolProps props;
        memset(&props, 0, sizeof props);
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        static const int private_pool = getenv("HM_PRIVATE_POOL") ? atoi(getenv("HM_PRIVATE_POOL")) : 1; // 0 = round-1 behaviour (A/B only)
        if (private_pool && cudaMemPoolCreate(&ctx->pool, &props) == cudaSuccess && ctx->pool) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep);
            if (cudaMemPoolCreate(&ctx->pool_big, &props) == cudaSuccess && ctx->pool_big) {
                cudaMemPoolSetAttribute(ctx->pool_big, cudaMemPoolAttrReleaseThreshold, &keep);
            } else {
                ctx->pool_big = nullptr;
                cudaGetLastError();
            }
        } else {
            ctx->pool = nullptr; // the device's default pool
            cudaGetLastError();
            if (!private_pool) {
                cudaMemPool_t dpool = nullptr;
                if (cudaDeviceGetDefaultMemPool(&dpool, device) == cudaSuccess && dpool) {
                    uint64_t keep = UINT64_MAX;
                    cudaMemPoolSetAttribute(dpool, cudaMemPoolAttrReleaseThreshold, &keep);
                }
            }
        }
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
        ctx->sm_count = prop.multiProcessorCount;
        ctx->smem_optin = prop.sharedMemPerBlockOptin;
    }
    *out = ctx;
    return HM_OK;
}

void hm_context_destroy(hm_context *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    // Batches the caller still holds lose their device memory here and become orphans: hm_batch_free(NULL or any ctx, b)
    // on an orphan only releases the host-side handle, so the two destruction orders are both safe.
    for (hm_batch *b : ctx->live) {
        if (b->d) {
            if (b->pooled) cudaFreeAsync(b->d, ctx->stream);
            else cudaFree(b->d);
        }
        b->d = nullptr;
        b->ctx = nullptr;
    }
    ctx->live.clear();
    cudaStreamSynchronize(ctx->stream);
    clear_secret(ctx);
    clear_public(ctx);
    if (ctx->d_ops) cudaFree(ctx->d_ops);
    if (ctx->h_ops) cudaFreeHost(ctx->h_ops);
    if (ctx->d_sched) cudaFree(ctx->d_sched);
    if (ctx->side_stream) {
        cudaStreamSynchronize(ctx->side_stream);
        cudaStreamDestroy(ctx->side_stream);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->d_k7_items) cudaFree(ctx->d_k7_items);
    if (ctx->d_k7_prods) cudaFree(ctx->d_k7_prods);
    if (ctx->d_k7_units) cudaFree(ctx->d_k7_units);
    if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
    if (ctx->pool_big) cudaMemPoolDestroy(ctx->pool_big);
    delete ctx;
}

int hm_context_parameters(const hm_context *ctx, uint16_t *d, uint16_t *dp, uint16_t *delta, uint16_t *tau) {
    if (!ctx) return HM_ERR_INVALID_ARGUMENT;
    if (d) *d = ctx->d;
    if (dp) *dp = ctx->dp;
    if (delta) *delta = ctx->delta;
    if (tau) *tau = ctx->tau;
    return HM_OK;
}

int hm_context_set_stream(hm_context *ctx, void *cuda_stream) {
    if (!ctx) return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    if (cuda_stream) {
        ctx->stream = (cudaStream_t)cuda_stream;
        ctx->own_stream = false;
    } else {
        CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    return HM_OK;
}
void *hm_context_stream(const hm_context *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int hm_context_synchronize(hm_context *ctx) {
    if (!ctx) return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    return HM_OK;
}
uint64_t hm_context_kernel_launches(const hm_context *ctx) { return ctx ? ctx->launches : 0; }
int hm_context_device(const hm_context *ctx) { return ctx ? ctx->device : -1; }

int hm_set_tuning(const char *key, long value) {
    if (!key) return HM_ERR_INVALID_ARGUMENT;
    if (strcmp(key, "adder_thread_min") == 0) {
        g_adder_thread_min = value;
        return HM_OK;
    }
    if (strcmp(key, "adder_chain") == 0) {
        if (value != 0 && value != 3 && value != 4 && value != 12 && value != 13) return HM_ERR_INVALID_ARGUMENT;
        g_adder_chain = value;
        return HM_OK;
    }
    if (strcmp(key, "mul_circuit_fused") == 0) {
        if (value != 0 && value != 1) return HM_ERR_INVALID_ARGUMENT;
        g_mul_circuit_fused = value;
        return HM_OK;
    }
    if (strcmp(key, "adder_wide_min") == 0) {
        g_adder_wide_min = value;
        return HM_OK;
    }
    if (strcmp(key, "adder_phases") == 0) {
        if (value < 0 || value > (long)hmk::ADDER_MAX_PHASES) return HM_ERR_INVALID_ARGUMENT;
        g_adder_phases = value;
        return HM_OK;
    }
    if (strcmp(key, "pool_max_mb") == 0) {
        if (value < 0 || value > (1 << 20)) return HM_ERR_INVALID_ARGUMENT;
        g_pool_max_mb = value;
        return HM_OK;
    }
    if (strcmp(key, "host_chunk_mb") == 0) {
        if (value < 1 || value > 4096) return HM_ERR_INVALID_ARGUMENT;
        g_host_chunk_mb = value;
        return HM_OK;
    }
    if (strcmp(key, "mul_thread_min") == 0) {
        g_mul_thread_min = value;
        return HM_OK;
    }
    if (strcmp(key, "mul_thread_chunk") == 0) {
        if (value != 24 && value != 32) return HM_ERR_INVALID_ARGUMENT;
        g_mul_thread_chunk = value;
        return HM_OK;
    }
    if (strcmp(key, "adder_generic_sequential") == 0) {
        g_adder_generic_seq = value;
        return HM_OK;
    }
    if (strcmp(key, "mul_circuit_sequential") == 0) {
        g_mul_circuit_seq = value;
        return HM_OK;
    }
    return HM_ERR_INVALID_ARGUMENT;
}

// ---- keys -------------------------------------------------------------------------------------
static gf2::words words_from_bytes(const uint8_t *bytes, size_t len) { // Polynomial::from_bytes, polynomial.rs:108-122
    gf2::words w((len + 7) / 8, 0);
    for (size_t i = 0; i < len; ++i) w[i / 8] |= (uint64_t)bytes[i] << (8 * (i % 8));
    return w;
}

int hm_set_secret_key(hm_context *ctx, const uint8_t *bytes, size_t len) {
    if (!ctx || !bytes || len == 0) return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    gf2::words S = words_from_bytes(bytes, len);
    if (gf2::is_zero(S.data(), S.size())) return HM_ERR_DIVIDE_BY_ZERO; // rem would panic, polynomial.rs:319-322
    const size_t ds = gf2::degree(S.data(), S.size());
    if (ds == 0) return HM_ERR_INVALID_ARGUMENT; // S = 1: the reference's rem never terminates
    clear_secret(ctx);
    clear_public(ctx); // Context::set_secret_key clears the public key, context.rs:568-571
    S.resize(ds / 64 + 1);
    ctx->S = S;
    ctx->ds = ds;
    ctx->has_sk = true;
    // S as u32 words for the generic remainder kernel
    std::vector<uint32_t> s32(ds / 32 + 2, 0);
    for (size_t i = 0; i < ds / 32 + 1; ++i) s32[i] = (uint32_t)(S[i / 2] >> (32 * (i % 2)));
    CK(cudaMalloc(&ctx->d_S32, s32.size() * 4));
    CK(cudaMemcpyAsync(ctx->d_S32, s32.data(), s32.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (ds % 32 == 0) {
        std::vector<uint32_t> T = gf2::rem_fold_tables(S, ds);
        CK(cudaMalloc(&ctx->d_remT, T.size() * 4));
        CK(cudaMemcpyAsync(ctx->d_remT, T.data(), T.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        std::fill(T.begin(), T.end(), 0);
    }
    CK(cudaStreamSynchronize(ctx->stream));
    std::fill(s32.begin(), s32.end(), 0);
    return HM_OK;
}

int hm_set_public_key(hm_context *ctx, const uint8_t *const *polys, const size_t *lens, size_t n_polys) {
    if (!ctx || !polys || !lens) return HM_ERR_INVALID_ARGUMENT;
    if (n_polys != ctx->tau) return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    std::vector<gf2::words> T(n_polys);
    size_t maxdeg = 0;
    for (size_t i = 0; i < n_polys; ++i) {
        if (!polys[i] || lens[i] == 0) return HM_ERR_INVALID_ARGUMENT;
        T[i] = words_from_bytes(polys[i], lens[i]);
        maxdeg = std::max(maxdeg, gf2::degree(T[i].data(), T[i].size()));
    }
    clear_public(ctx);
    ctx->T_host = T;
    ctx->fresh_deg = maxdeg;
    ctx->wf = (uint32_t)(maxdeg / 64 + 1);
    const uint32_t wf = ctx->wf, tau = ctx->tau;
    // window: largest of 8/4/2/1 whose table (+ staging) fits in shared memory
    const size_t budget = ctx->smem_optin > 48 * 1024 ? ctx->smem_optin - 44 * 1024 : 0;
    uint32_t wb = 1;
    bool in_smem = false;
    for (uint32_t cand : {8u, 4u, 2u, 1u}) {
        const size_t groups = (tau + cand - 1) / cand;
        const size_t bytes = groups * ((size_t)1 << cand) * wf * 8;
        if (bytes <= budget) {
            wb = cand;
            in_smem = true;
            break;
        }
    }
    if (!in_smem) wb = 4; // table stays in global memory / L2
    const uint32_t groups = (tau + wb - 1) / wb;
    std::vector<uint64_t> tab((size_t)groups * ((size_t)1 << wb) * wf, 0);
    for (uint32_t g = 0; g < groups; ++g)
        for (uint32_t e = 1; e < (1u << wb); ++e) {
            const uint32_t low = e & (~e + 1u);      // lowest set bit
            const uint32_t bit = __builtin_ctz(low); // its index
            const uint32_t i = g * wb + bit;         // public polynomial index
            uint64_t *row = &tab[((size_t)(g << wb) + e) * wf];
            const uint64_t *prev = &tab[((size_t)(g << wb) + (e ^ low)) * wf];
            for (uint32_t j = 0; j < wf; ++j) row[j] = prev[j];
            if (i < tau) // mask bits >= tau are never looked at by the reference (cipher.rs:105)
                for (uint32_t j = 0; j < wf && j < T[i].size(); ++j) row[j] ^= T[i][j];
        }
    CK(cudaMalloc(&ctx->d_enc_table, tab.size() * 8));
    CK(cudaMemcpyAsync(ctx->d_enc_table, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (in_smem && wb == 8 && wf == 5 && tau == 128 && (size_t)hmk::ENC6_SLOTS * 256 * 128 + 1024 <= ctx->smem_optin) {
        // line(t, e): rows of groups 3t, 3t+1, 3t+2 at 64-bit word offsets 0, 5, 10 of a 16-word line
        std::vector<uint64_t> t6((size_t)hmk::ENC6_SLOTS * 256 * 16, 0);
        for (uint32_t g = 0; g < groups; ++g)
            for (uint32_t e = 0; e < 256; ++e)
                for (uint32_t j = 0; j < 5; ++j)
                    t6[((size_t)(g / 3) * 256 + e) * 16 + 5 * (g % 3) + j] = tab[((size_t)(g << 8) + e) * 5 + j];
        // the last line set has one real group (15): repeat its rows in the two spare classes, so that encrypt_tab6b_kernel
        // can fetch it with one rotation-free lookup (index 0 of the spare classes stays the zero row: e = 0 is the empty subset)
        for (uint32_t e = 0; e < 256; ++e)
            for (uint32_t cls = 1; cls < 3; ++cls)
                for (uint32_t j = 0; j < 5; ++j) t6[((size_t)5 * 256 + e) * 16 + 5 * cls + j] = tab[((size_t)(15u << 8) + e) * 5 + j];
        CK(cudaMalloc(&ctx->d_enc_table6, t6.size() * 8));
        CK(cudaMemcpyAsync(ctx->d_enc_table6, t6.data(), t6.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (in_smem && wb == 8 && wf == 5 && maxdeg == 256 && tau == 128 && (size_t)hmk::ENC4_TABLE_BYTES + 65536 + 1024 <= ctx->smem_optin) {
        // encrypt_tab4_kernel: words 0..3 of row (g, e) at byte (t >> 1) * 65536 + e * 256 + (t & 1) * 128 + k * 32, g = 4 t + k
        std::vector<uint64_t> t4((size_t)hmk::ENC4_TABLE_BYTES / 8, 0);
        for (uint32_t g = 0; g < groups; ++g)
            for (uint32_t e = 0; e < 256; ++e) {
                const uint32_t t = g / 4, k = g % 4;
                const size_t byte = (size_t)(t >> 1) * 65536 + (size_t)e * 256 + (t & 1) * 128 + k * 32;
                for (uint32_t j = 0; j < 4; ++j) t4[byte / 8 + j] = tab[((size_t)(g << 8) + e) * 5 + j];
            }
        for (uint32_t i = 0; i < 128; ++i)
            if (T[i].size() > 4 && (T[i][4] & 1)) ctx->enc_topmask[i / 32] |= 1u << (i % 32);
        CK(cudaMalloc(&ctx->d_enc_table4, t4.size() * 8));
        CK(cudaMemcpyAsync(ctx->d_enc_table4, t4.data(), t4.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (in_smem && wb == 4 && wf == 17 && maxdeg == 1024 && tau == 256 && (size_t)hmk::ENC4B_TABLE_BYTES + 65536 + 1024 <= ctx->smem_optin) {
        // encrypt_tab4b_kernel: words 0..15 of row (g, e) at byte (gp >> 4) * 65536 + (gp & 15) * 4096 + e * 256 + (g & 1) * 128, gp = g / 2
        std::vector<uint64_t> t4((size_t)hmk::ENC4B_TABLE_BYTES / 8, 0);
        for (uint32_t g = 0; g < groups; ++g)
            for (uint32_t e = 0; e < 16; ++e) {
                const uint32_t gp = g / 2;
                const size_t byte = (size_t)(gp >> 4) * 65536 + (size_t)(gp & 15) * 4096 + (size_t)e * 256 + (g & 1) * 128;
                for (uint32_t j = 0; j < 16; ++j) t4[byte / 8 + j] = tab[((size_t)(g << 4) + e) * 17 + j];
            }
        for (uint32_t i = 0; i < 256; ++i)
            if (T[i].size() > 16 && (T[i][16] & 1)) ctx->enc_topmask[i / 32] |= 1u << (i % 32);
        CK(cudaMalloc(&ctx->d_enc_table4b, t4.size() * 8));
        CK(cudaMemcpyAsync(ctx->d_enc_table4b, t4.data(), t4.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        { // the same key as the int8 B operand of encrypt_umma_b_kernel (2 x 128 KB)
            std::vector<const uint64_t *> tp(256);
            std::vector<size_t> tw(256);
            for (uint32_t i = 0; i < 256; ++i) {
                tp[i] = T[i].data();
                tw[i] = T[i].size();
            }
            std::vector<int8_t> bt(hmk::enc_umma_table_bytes());
            hmk::enc_umma_build_table(tp.data(), tw.data(), bt.data());
            CK(cudaMalloc(&ctx->d_enc_umma, bt.size()));
            CK(cudaMemcpy(ctx->d_enc_umma, bt.data(), bt.size(), cudaMemcpyHostToDevice));
        }
    }
    ctx->enc_wb = wb;
    ctx->enc_groups = groups;
    ctx->enc_table_words = (uint32_t)tab.size();
    ctx->enc_table_in_smem = in_smem;
    ctx->has_pk = true;
    return HM_OK;
}

// ---- seeded key generation on the device ----------------------------------------------------------
int hm_key_stream_host(uint64_t seed, uint32_t stream, size_t nbytes, uint8_t *out) {
    if (!out && nbytes) return HM_ERR_INVALID_ARGUMENT;
    for (size_t b = 0; b < nbytes; ++b) out[b] = (uint8_t)(hmk::key_stream_word(seed, stream, b / 8) >> (8 * (b % 8)));
    return HM_OK;
}

int hm_generate_keys_seeded(hm_context *ctx, uint64_t seed) {
    if (!ctx) return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    const uint32_t d = ctx->d, dp = ctx->dp, delta = ctx->delta, tau = ctx->tau;
    // S = Polynomial::random(d) from stream 0 (src/context.rs:160-162, src/polynomial.rs:73-96).  It is needed on the host
    // anyway: the decrypt vector and the fold tables are built there.
    const uint32_t ws = d / 64 + 1;
    std::vector<uint64_t> S(ws);
    for (uint32_t j = 0; j < ws; ++j) S[j] = hmk::key_stream_word(seed, 0, j);
    S[ws - 1] &= ((uint64_t)1 << (d % 64)) - 1;
    S[ws - 1] |= (uint64_t)1 << (d % 64);
    int rc = hm_set_secret_key(ctx, reinterpret_cast<const uint8_t *>(S.data()), (size_t)ws * 8);
    if (rc != HM_OK) return rc;
    // T_i = S * Q_i + X * R_i on the device (src/context.rs:249-261)
    const uint32_t wq = dp / 64 + 1, wr = (delta + 1) / 64 + 1, wt = (d + dp) / 64 + 1;
    PoolGuard gS(ctx), gQ(ctx), gR(ctx), gT(ctx);
    CK(pool_alloc(ctx, &gS.p, (size_t)ws * 8));
    CK(pool_alloc(ctx, &gQ.p, (size_t)tau * wq * 8));
    CK(pool_alloc(ctx, &gR.p, (size_t)tau * wr * 8));
    CK(pool_alloc(ctx, &gT.p, (size_t)tau * wt * 8));
    CK(cudaMemcpyAsync(gS.p, S.data(), (size_t)ws * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(gT.p, 0, (size_t)tau * wt * 8, ctx->stream));
    hmk::keygen_fill_kernel<<<tau, 128, 0, ctx->stream>>>(static_cast<uint64_t *>(gQ.p), static_cast<uint64_t *>(gR.p), tau, dp, delta, seed);
    LAUNCHED("keygen_fill_kernel");
    auto view = [](void *base, uint64_t stride, uint32_t w, uint64_t deg) {
        View v;
        v.base = static_cast<uint64_t *>(base);
        v.stride = stride;
        v.off = 0;
        v.w = w;
        v.deg = deg;
        return v;
    };
    View vS = view(gS.p, 0, ws, d), vQ = view(gQ.p, wq, wq, dp), vR = view(gR.p, wr, wr, (uint64_t)delta + 1), vT = view(gT.p, wt, wt, (uint64_t)d + dp);
    std::vector<MulOp> ops(1, MulOp{vS, vQ, vT}); // S is the same for every i: a view with stride 0
    rc = launch_mul_ops(ctx, ops, tau, true);
    if (rc != HM_OK) return rc;
    View vTr = vT;
    vTr.w = std::min(wt, wr);
    rc = launch_xor_views(ctx, vTr, vTr, vR, tau); // + X * R_i over its own width (deg X R_i = delta + 1 <= d + dp)
    if (rc != HM_OK) return rc;
    std::vector<uint64_t> T((size_t)tau * wt);
    CK(cudaMemcpyAsync(T.data(), gT.p, T.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::fill(S.begin(), S.end(), 0);
    // PublicKey::to_bytes format: degree/64+1 little-endian words per polynomial (src/polynomial.rs:99-105)
    std::vector<const uint8_t *> ptrs(tau);
    std::vector<size_t> lens(tau);
    for (uint32_t i = 0; i < tau; ++i) {
        const uint64_t *w = &T[(size_t)i * wt];
        ptrs[i] = reinterpret_cast<const uint8_t *>(w);
        lens[i] = (gf2::degree(w, wt) / 64 + 1) * 8;
    }
    return hm_set_public_key(ctx, ptrs.data(), lens.data(), tau);
}

// SecretKey::to_bytes / PublicKey::to_bytes()[i] of the keys the context holds (src/context.rs:192-194, :291-297):
// degree/64+1 little-endian u64 words.  *len = bytes needed; out may be NULL to query.
int hm_secret_key_bytes(const hm_context *ctx, uint8_t *out, size_t capacity, size_t *len) {
    if (!ctx || !len) return HM_ERR_INVALID_ARGUMENT;
    if (!ctx->has_sk) return HM_ERR_SECRET_KEY_UNSET;
    *len = (ctx->ds / 64 + 1) * 8;
    if (!out) return HM_OK;
    if (capacity < *len) return HM_ERR_INVALID_ARGUMENT;
    memcpy(out, ctx->S.data(), *len);
    return HM_OK;
}
int hm_public_key_bytes(const hm_context *ctx, size_t i, uint8_t *out, size_t capacity, size_t *len) {
    if (!ctx || !len) return HM_ERR_INVALID_ARGUMENT;
    if (!ctx->has_pk) return HM_ERR_PUBLIC_KEY_UNSET;
    if (i >= ctx->T_host.size()) return HM_ERR_INVALID_ARGUMENT;
    const gf2::words &t = ctx->T_host[i];
    *len = (gf2::degree(t.data(), t.size()) / 64 + 1) * 8;
    if (!out) return HM_OK;
    if (capacity < *len) return HM_ERR_INVALID_ARGUMENT;
    memcpy(out, t.data(), *len);
    return HM_OK;
}

int hm_has_secret_key(const hm_context *ctx) { return ctx && ctx->has_sk; }
int hm_has_public_key(const hm_context *ctx) { return ctx && ctx->has_pk; }

// ---- batches ----------------------------------------------------------------------------------
size_t hm_batch_len(const hm_batch *b) { return b ? b->n : 0; }
uint32_t hm_batch_bits(const hm_batch *b) { return b ? b->L : 0; }
size_t hm_batch_value_words(const hm_batch *b) { return b ? b->value_words : 0; }
int hm_batch_slot_words(const hm_batch *b, uint32_t *widths_out) {
    if (!b || !widths_out) return HM_ERR_INVALID_ARGUMENT;
    for (uint32_t k = 0; k < b->L; ++k) widths_out[k] = b->w[k];
    return HM_OK;
}
int hm_batch_slot_degree_bounds(const hm_batch *b, uint64_t *bounds_out) {
    if (!b || !bounds_out) return HM_ERR_INVALID_ARGUMENT;
    for (uint32_t k = 0; k < b->L; ++k) bounds_out[k] = b->degb[k];
    return HM_OK;
}
void *hm_batch_device_ptr(const hm_batch *b) { return b ? (void *)b->d : nullptr; }

void hm_batch_free(hm_context *ctx, hm_batch *b) {
    if (!b) return;
    // The owning context is the one recorded in the batch; it is NULL once that context has been destroyed (the batch is
    // then an orphan without device memory).  The `ctx` argument is accepted for symmetry with the other calls only.
    (void)ctx;
    hm_context *owner = b->ctx;
    if (owner) {
        cudaSetDevice(owner->device);
        if (b->d) {
            if (b->pooled) {
                cudaFreeAsync(b->d, owner->stream); // ordered after every kernel of this context that used it
            } else {
                cudaStreamSynchronize(owner->stream);
                cudaFree(b->d);
            }
        }
        owner->live.erase(b);
    }
    delete b;
}

int hm_batch_upload_bounded(hm_context *ctx, size_t n, uint32_t L, const uint64_t *degree_bounds, const uint64_t *host,
                            hm_batch **out) {
    if (!ctx || !out || !degree_bounds || (!host && n) || L == 0 || L > (uint32_t)hmk::MAX_SLOTS)
        return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    int nb_status = HM_OK;
    hm_batch *b = new_batch(ctx, n, L, degree_bounds, &nb_status);
    if (!b) return nb_status;
    int rc = alloc_batch(ctx, b);
    if (rc != HM_OK) {
        discard_batch(b);
        return rc;
    }
    if (n) {
        cudaError_t e = cudaMemcpyAsync(b->d, host, n * b->value_words * 8, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            hm_batch_free(ctx, b);
            return fail_cuda(ctx, e, "upload");
        }
    }
    *out = b;
    return HM_OK;
}

int hm_batch_upload(hm_context *ctx, size_t n, uint32_t L, const uint32_t *slot_words, const uint64_t *host,
                    hm_batch **out) {
    if (!slot_words || L == 0 || L > (uint32_t)hmk::MAX_SLOTS) return HM_ERR_INVALID_ARGUMENT;
    std::vector<uint64_t> degb(L);
    for (uint32_t k = 0; k < L; ++k) {
        if (slot_words[k] == 0) return HM_ERR_INVALID_ARGUMENT; // empty polynomial, polynomial.rs:54-57
        degb[k] = (uint64_t)slot_words[k] * 64 - 1;
    }
    return hm_batch_upload_bounded(ctx, n, L, degb.data(), host, out);
}

int hm_batch_download(hm_context *ctx, const hm_batch *b, uint64_t *host) {
    if (!ctx || !b || (!host && b->n)) return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    if (b->n) CK(cudaMemcpyAsync(host, b->d, b->n * b->value_words * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return HM_OK;
}

int hm_batch_download_range(hm_context *ctx, const hm_batch *b, size_t first, size_t count, uint64_t *host) {
    if (!ctx || !b || (!host && count)) return HM_ERR_INVALID_ARGUMENT;
    if (first > b->n || count > b->n - first) return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    if (count) CK(cudaMemcpyAsync(host, b->d + first * b->value_words, count * b->value_words * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return HM_OK;
}

int hm_batch_clone(hm_context *ctx, const hm_batch *b, hm_batch **out) {
    if (!ctx || !b || !out) return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    int nb_status = HM_OK;
    hm_batch *c = new_batch(ctx, b->n, b->L, b->degb.data(), &nb_status);
    if (!c) return nb_status;
    int rc = alloc_batch(ctx, c);
    if (rc != HM_OK) {
        discard_batch(c);
        return rc;
    }
    if (b->n) {
        cudaError_t e = cudaMemcpyAsync(c->d, b->d, b->n * b->value_words * 8, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e != cudaSuccess) {
            hm_batch_free(ctx, c);
            return fail_cuda(ctx, e, "clone");
        }
    }
    *out = c;
    return HM_OK;
}

// ---- wire format ------------------------------------------------------------------------------------
// The reference has no ciphertext serialisation (CipheredBit's field is private, src/cipher.rs:30); this is the
// engine's own, little endian:
//   "HMB1" | u16 d, dp, delta, tau | u32 L | u64 n | L x u64 degree bound | n * value_words x u64 (padded layout)
namespace {
struct WireHeader {
    char magic[4];
    uint16_t d, dp, delta, tau;
    uint32_t L;
    uint64_t n;
};
constexpr size_t WIRE_HEADER_BYTES = 4 + 8 + 4 + 8;
void put(uint8_t *&p, const void *src, size_t n) {
    memcpy(p, src, n);
    p += n;
}
void get(const uint8_t *&p, void *dst, size_t n) {
    memcpy(dst, p, n);
    p += n;
}
} // namespace

size_t hm_batch_serialized_size(const hm_batch *b) {
    if (!b) return 0;
    return WIRE_HEADER_BYTES + (size_t)b->L * 8 + b->n * b->value_words * 8;
}

int hm_batch_serialize(hm_context *ctx, const hm_batch *b, uint8_t *out, size_t capacity) {
    if (!ctx || !b || !out) return HM_ERR_INVALID_ARGUMENT;
    if (capacity < hm_batch_serialized_size(b)) return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    uint8_t *p = out;
    put(p, "HMB1", 4);
    put(p, &ctx->d, 2);
    put(p, &ctx->dp, 2);
    put(p, &ctx->delta, 2);
    put(p, &ctx->tau, 2);
    put(p, &b->L, 4);
    const uint64_t n = b->n;
    put(p, &n, 8);
    put(p, b->degb.data(), (size_t)b->L * 8);
    if (b->n) CK(cudaMemcpyAsync(p, b->d, b->n * b->value_words * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return HM_OK;
}

// Validates a wire header against its own length without trusting any field: checked arithmetic only.  The parser sees
// ciphertexts produced by other parties.
int hm_batch_wire_inspect(const uint8_t *in, size_t len, uint16_t params_out[4], uint32_t *L_out, uint64_t *n_out,
                          uint64_t *value_words_out) {
    if (!in) return HM_ERR_INVALID_ARGUMENT;
    if (len < WIRE_HEADER_BYTES || memcmp(in, "HMB1", 4) != 0) return HM_ERR_INVALID_ARGUMENT;
    const uint8_t *p = in + 4;
    uint16_t prm[4];
    uint32_t L;
    uint64_t n;
    get(p, prm, 8);
    get(p, &L, 4);
    get(p, &n, 8);
    if (L == 0 || L > (uint32_t)hmk::MAX_SLOTS) return HM_ERR_INVALID_ARGUMENT;
    if (len - WIRE_HEADER_BYTES < (size_t)L * 8) return HM_ERR_INVALID_LENGTH;
    std::vector<uint64_t> degb(L);
    get(p, degb.data(), (size_t)L * 8);
    const uint64_t vw = checked_value_words(L, degb.data()); // every bound <= 2^31, sum of widths < 2^32
    if (vw == 0) return HM_ERR_INVALID_ARGUMENT;
    const size_t body = len - WIRE_HEADER_BYTES - (size_t)L * 8;
    // body must be exactly n * vw * 8 bytes: compare by division, never by a product that could wrap
    if (body % (vw * 8) != 0 || (uint64_t)(body / (vw * 8)) != n) return HM_ERR_INVALID_LENGTH;
    if (params_out) memcpy(params_out, prm, 8);
    if (L_out) *L_out = L;
    if (n_out) *n_out = n;
    if (value_words_out) *value_words_out = vw;
    return HM_OK;
}

int hm_batch_deserialize(hm_context *ctx, const uint8_t *in, size_t len, hm_batch **out) {
    if (!ctx || !in || !out) return HM_ERR_INVALID_ARGUMENT;
    uint16_t prm[4];
    uint32_t L = 0;
    uint64_t n = 0, vw = 0;
    int rc = hm_batch_wire_inspect(in, len, prm, &L, &n, &vw);
    if (rc != HM_OK) return rc;
    if (prm[0] != ctx->d || prm[1] != ctx->dp || prm[2] != ctx->delta || prm[3] != ctx->tau) return HM_ERR_INVALID_PARAMETERS;
    std::vector<uint64_t> degb(L);
    memcpy(degb.data(), in + WIRE_HEADER_BYTES, (size_t)L * 8);
    const uint8_t *p = in + WIRE_HEADER_BYTES + (size_t)L * 8;
    // the body may be unaligned inside the caller's buffer: upload byte-wise
    USE_DEV(ctx);
    int nb_status = HM_OK;
    hm_batch *b = new_batch(ctx, (size_t)n, L, degb.data(), &nb_status);
    if (!b) return nb_status;
    rc = alloc_batch(ctx, b);
    if (rc != HM_OK) {
        discard_batch(b);
        return rc;
    }
    if (n) {
        cudaError_t e = cudaMemcpyAsync(b->d, p, (size_t)n * vw * 8, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            hm_batch_free(ctx, b);
            return fail_cuda(ctx, e, "deserialize upload");
        }
    }
    *out = b;
    return HM_OK;
}

// Canonical export: for every polynomial (value-major, slot-minor) `u64 degree` then degree/64+1 words — exactly the
// (degree, coefficients) pair the reference's Polynomial holds (src/polynomial.rs:22-26, :404-426).  Returns the
// number of u64 written in *written; with out == NULL only counts.
int hm_batch_download_canonical(hm_context *ctx, const hm_batch *b, uint64_t *out, size_t capacity_words, size_t *written) {
    if (!ctx || !b || !written) return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    std::vector<uint64_t> host(b->n * b->value_words);
    if (b->n) CK(cudaMemcpyAsync(host.data(), b->d, host.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    size_t pos = 0;
    for (size_t v = 0; v < b->n; ++v)
        for (uint32_t k = 0; k < b->L; ++k) {
            const uint64_t *w = &host[v * b->value_words + b->off[k]];
            const size_t deg = gf2::degree(w, b->w[k]);
            const size_t nw = deg / 64 + 1;
            if (out) {
                if (pos + 1 + nw > capacity_words) return HM_ERR_INVALID_ARGUMENT;
                out[pos] = deg;
                memcpy(out + pos + 1, w, nw * 8);
            }
            pos += 1 + nw;
        }
    *written = pos;
    return HM_OK;
}

// Inverse of hm_batch_download_canonical: per polynomial (value-major, slot-minor) `u64 degree` then degree/64+1 words,
// as a Rust build of the reference would export its Polynomial { coefficients, degree } (src/polynomial.rs:22-26).  The
// engine's slot k gets the width of the largest degree seen in slot k (or of degree_bounds[k] when given, which must then
// cover every polynomial of the slot).  Rejects a stated degree that disagrees with the words (src/polynomial.rs:35-42).
int hm_batch_upload_canonical(hm_context *ctx, size_t n, uint32_t L, const uint64_t *in, size_t in_words, const uint64_t *degree_bounds,
                              hm_batch **out) {
    if (!ctx || !out || (!in && n) || L == 0 || L > (uint32_t)hmk::MAX_SLOTS) return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    std::vector<uint64_t> degb(L, 0);
    // pass 1: validate and find the slot bounds
    size_t pos = 0;
    for (size_t v = 0; v < n; ++v)
        for (uint32_t k = 0; k < L; ++k) {
            if (pos >= in_words) return HM_ERR_INVALID_LENGTH;
            const uint64_t deg = in[pos];
            if (deg > MAX_DEGREE_BOUND) return HM_ERR_INVALID_ARGUMENT;
            const size_t nw = (size_t)(deg / 64 + 1);
            if (in_words - pos - 1 < nw) return HM_ERR_INVALID_LENGTH;
            if (gf2::degree(in + pos + 1, nw) != deg) return HM_ERR_INVALID_ARGUMENT;
            degb[k] = std::max(degb[k], deg);
            pos += 1 + nw;
        }
    if (pos != in_words) return HM_ERR_INVALID_LENGTH;
    if (degree_bounds)
        for (uint32_t k = 0; k < L; ++k) {
            if (degree_bounds[k] < degb[k]) return HM_ERR_INVALID_ARGUMENT;
            degb[k] = degree_bounds[k];
        }
    int nb_status = HM_OK;
    hm_batch *b = new_batch(ctx, n, L, degb.data(), &nb_status);
    if (!b) return nb_status;
    int rc = alloc_batch(ctx, b);
    if (rc != HM_OK) {
        discard_batch(b);
        return rc;
    }
    // pass 2: zero-padded slots
    std::vector<uint64_t> host;
    try {
        host.assign(n * b->value_words, 0);
    } catch (const std::bad_alloc &) {
        hm_batch_free(ctx, b);
        return HM_ERR_OUT_OF_MEMORY;
    }
    pos = 0;
    for (size_t v = 0; v < n; ++v)
        for (uint32_t k = 0; k < L; ++k) {
            const size_t nw = (size_t)(in[pos] / 64 + 1);
            memcpy(&host[v * b->value_words + b->off[k]], in + pos + 1, nw * 8);
            pos += 1 + nw;
        }
    if (n) {
        cudaError_t e = cudaMemcpyAsync(b->d, host.data(), host.size() * 8, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            hm_batch_free(ctx, b);
            return fail_cuda(ctx, e, "upload canonical");
        }
    }
    *out = b;
    return HM_OK;
}

void *hm_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void hm_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

// ---- encrypt / decrypt ------------------------------------------------------------------------
// d_masks == nullptr: masks drawn inside the kernel from Philox(seed) at stream position first_unit (only where
// encrypt_fuses_masks(ctx, n * L) says so)
static int encrypt_exec(hm_context *ctx, const uint8_t *d_values, size_t n, uint32_t L, const uint8_t *d_masks, hm_batch *b, uint64_t seed = 0,
                        uint64_t first_unit = 0);
static const int g_enc_mode = getenv("HM_ENC_MODE") ? atoi(getenv("HM_ENC_MODE")) : 3; // 3 = encrypt_tab4_kernel at config A, tensor-core encrypt_umma_b_kernel at config B; 2 = encrypt_tab4(b)_kernel; 1 = encrypt_tab6(b)_kernel; 0 = encrypt_tab_kernel
static bool encrypt_fuses_masks(const hm_context *ctx, uint64_t units) {
    return g_enc_mode >= 2 && (ctx->d_enc_table4 || ctx->d_enc_table4b) && units < ((uint64_t)1 << 31);
}

int hm_encrypt_device(hm_context *ctx, const uint8_t *d_values, size_t n, uint32_t L, const uint8_t *d_masks,
                      hm_batch **out) {
    if (!ctx || !out || ((!d_values || !d_masks) && n)) return HM_ERR_INVALID_ARGUMENT;
    if (!ctx->has_pk) return HM_ERR_PUBLIC_KEY_UNSET; // context.rs:463-471
    if (L == 0 || L % 8 != 0 || L > (uint32_t)hmk::MAX_SLOTS) return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    std::vector<uint64_t> degb(L, ctx->fresh_deg);
    int nb_status = HM_OK;
    hm_batch *b = new_batch(ctx, n, L, degb.data(), &nb_status);
    if (!b) return nb_status;
    int rc = alloc_batch(ctx, b);
    if (rc == HM_OK) rc = encrypt_exec(ctx, d_values, n, L, d_masks, b);
    if (rc != HM_OK) {
        hm_batch_free(ctx, b);
        return rc;
    }
    *out = b;
    return HM_OK;
}

int hm_encrypt_device_into(hm_context *ctx, const uint8_t *d_values, size_t n, uint32_t L, const uint8_t *d_masks,
                           hm_batch *out) {
    if (!ctx || !out || ((!d_values || !d_masks) && n)) return HM_ERR_INVALID_ARGUMENT;
    if (!ctx->has_pk) return HM_ERR_PUBLIC_KEY_UNSET;
    if (out->n != n || out->L != L || L % 8 != 0) return HM_ERR_INVALID_ARGUMENT;
    for (uint32_t k = 0; k < L; ++k)
        if (out->w[k] != ctx->wf) return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    out->degb.assign(L, ctx->fresh_deg);
    return encrypt_exec(ctx, d_values, n, L, d_masks, out);
}

static int encrypt_exec(hm_context *ctx, const uint8_t *d_values, size_t n, uint32_t L, const uint8_t *d_masks, hm_batch *b, uint64_t seed,
                        uint64_t first_unit) {
    if (n == 0) return HM_OK;
    if (g_enc_mode == 3 && encrypt_fuses_masks(ctx, (uint64_t)n * L) && (!d_masks || ((uintptr_t)d_masks % 16) == 0) && ctx->d_enc_umma) {
        hmk::EncUmmaParams q;
        q.values = d_values;
        q.masks = d_masks;
        q.out = b->d;
        q.units = (uint32_t)((uint64_t)n * L);
        static const int umma_chunk = getenv("HM_UMMA_CHUNK") ? atoi(getenv("HM_UMMA_CHUNK")) : 0;
        q.chunk_tiles = (uint32_t)umma_chunk;
        for (int i = 0; i < 8; ++i) q.topmask[i] = ctx->enc_topmask[i];
        q.seed = seed;
        q.first_unit = first_unit;
        CK(hmk::launch_encrypt_umma_b(q, d_masks == nullptr, ctx->d_enc_umma, ctx->sm_count, ctx->stream));
        LAUNCHED("encrypt_umma_b_kernel");
        return HM_OK;
    }
    if (encrypt_fuses_masks(ctx, (uint64_t)n * L) && (!d_masks || ((uintptr_t)d_masks % 16) == 0) && ctx->d_enc_table4b) {
        hmk::Enc4bParams q;
        q.values = d_values;
        q.masks = d_masks;
        q.out = b->d;
        q.units = (uint32_t)((uint64_t)n * L);
        for (int i = 0; i < 8; ++i) q.topmask[i] = ctx->enc_topmask[i];
        q.seed = seed;
        q.first_unit = first_unit;
        const size_t smem = (size_t)hmk::ENC4B_TABLE_BYTES + 65536;
        const int grid = grid_for(ctx, ((uint64_t)q.units + 15) / 16 * 32, hmk::ENC4B_THREADS, 1);
        const uint4 *t4 = reinterpret_cast<const uint4 *>(ctx->d_enc_table4b);
        if (d_masks) {
            CK(cudaFuncSetAttribute(hmk::encrypt_tab4b_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            hmk::encrypt_tab4b_kernel<false><<<grid, hmk::ENC4B_THREADS, smem, ctx->stream>>>(q, t4);
        } else {
            CK(cudaFuncSetAttribute(hmk::encrypt_tab4b_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            hmk::encrypt_tab4b_kernel<true><<<grid, hmk::ENC4B_THREADS, smem, ctx->stream>>>(q, t4);
        }
        LAUNCHED("encrypt_tab4b_kernel");
        return HM_OK;
    }
    if (encrypt_fuses_masks(ctx, (uint64_t)n * L) && (!d_masks || ((uintptr_t)d_masks % 16) == 0)) {
        hmk::Enc4Params q;
        q.values = d_values;
        q.masks = d_masks;
        q.out = b->d;
        q.units = (uint32_t)((uint64_t)n * L);
        for (int i = 0; i < 4; ++i) q.topmask[i] = ctx->enc_topmask[i];
        q.seed = seed;
        q.first_unit = first_unit;
        const size_t smem = (size_t)hmk::ENC4_TABLE_BYTES + 65536; // the table starts on a 64 KB boundary of the shared window
        const int grid = grid_for(ctx, ((uint64_t)q.units + 31) / 32 * 32, hmk::ENC4_THREADS, 1);
        const uint4 *t4 = reinterpret_cast<const uint4 *>(ctx->d_enc_table4);
        if (d_masks) {
            CK(cudaFuncSetAttribute(hmk::encrypt_tab4_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            hmk::encrypt_tab4_kernel<false><<<grid, hmk::ENC4_THREADS, smem, ctx->stream>>>(q, t4);
        } else {
            CK(cudaFuncSetAttribute(hmk::encrypt_tab4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            hmk::encrypt_tab4_kernel<true><<<grid, hmk::ENC4_THREADS, smem, ctx->stream>>>(q, t4);
        }
        LAUNCHED("encrypt_tab4_kernel");
        return HM_OK;
    }
    if (!d_masks) return HM_ERR_INVALID_ARGUMENT;
    hmk::EncParams p;
    p.values = d_values;
    p.masks = d_masks;
    p.out = b->d;
    p.table = ctx->d_enc_table;
    p.units = (uint64_t)n * L;
    p.wf = ctx->wf;
    p.mask_bytes = (ctx->tau + 7u) / 8u;
    p.wb = ctx->enc_wb;
    p.groups = ctx->enc_groups;
    p.table_words = ctx->enc_table_words;
    p.table_in_smem = ctx->enc_table_in_smem ? 1 : 0;
    const bool masks_aligned = ((uintptr_t)d_masks % 16) == 0;
    const size_t table_bytes = ((size_t)p.table_words + 1) / 2 * 16;
    int path = 0; // 0 generic, 1 config A tables, 2 config B tables
    if (ctx->enc_table_in_smem && masks_aligned && p.wf == 5 && ctx->tau == 128 && p.wb == 8) path = 1;
    if (ctx->enc_table_in_smem && masks_aligned && p.wf == 17 && ctx->tau == 256 && p.wb == 4 &&
        table_bytes + 2 * (size_t)256 * 17 * 8 <= ctx->smem_optin) // 256-thread CTAs: table 136 KB + 2 x 34 KB of staging
        path = 2;
    const int enc_mode = g_enc_mode;
    if (path == 1 && enc_mode >= 1 && ctx->d_enc_table6) {
        const size_t smem = (size_t)hmk::ENC6_SLOTS * 256 * 128;
        CK(cudaFuncSetAttribute(hmk::encrypt_tab6_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = grid_for(ctx, (p.units + 5) / 6 * 32, hmk::ENC6_THREADS, 1);
        static const int enc6b = getenv("HM_ENC6B") ? atoi(getenv("HM_ENC6B")) : 1; // 0 = first version (compiler-generated index arithmetic)
        if (enc6b && p.units < ((uint64_t)1 << 31) - 6 * (uint64_t)grid * (hmk::ENC6_THREADS / 32)) {
            CK(cudaFuncSetAttribute(hmk::encrypt_tab6b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            hmk::encrypt_tab6b_kernel<<<grid, hmk::ENC6_THREADS, smem, ctx->stream>>>(p, ctx->d_enc_table6);
        } else {
            hmk::encrypt_tab6_kernel<<<grid, hmk::ENC6_THREADS, smem, ctx->stream>>>(p, ctx->d_enc_table6);
        }
        LAUNCHED("encrypt_tab6_kernel");
    } else if (path == 1) {
        const size_t smem = table_bytes + 2 * (size_t)hmk::ENC_THREADS * 5 * 8;
        auto kern = hmk::encrypt_tab_kernel<5, 4, 8>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = grid_for(ctx, p.units, hmk::ENC_THREADS, 1);
        kern<<<grid, hmk::ENC_THREADS, smem, ctx->stream>>>(p);
        LAUNCHED("encrypt_tab_kernel<5,4,8>");
    } else if (path == 2) {
        const size_t smem = table_bytes + 2 * (size_t)256 * 17 * 8;
        auto kern = hmk::encrypt_tab_kernel<17, 8, 4, 256>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = grid_for(ctx, p.units, 256, 1);
        kern<<<grid, 256, smem, ctx->stream>>>(p);
        LAUNCHED("encrypt_tab_kernel<17,8,4,256>");
    } else {
        const size_t smem = ctx->enc_table_in_smem ? (size_t)p.table_words * 8 : 0;
        CK(cudaFuncSetAttribute(hmk::encrypt_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int per_sm = smem > 100 * 1024 ? 1 : (smem > 48 * 1024 ? 2 : 4);
        const int grid = grid_for(ctx, p.units, 256, per_sm);
        hmk::encrypt_generic_kernel<<<grid, 256, smem, ctx->stream>>>(p);
        LAUNCHED("encrypt_generic_kernel");
    }
    return HM_OK;
}

// Host plaintexts (and, unless seeded, host masks) -> device batch.  masks == NULL means: generate them on the device from
// `seed`, bit-ciphertext u of this call taking position first_unit + u of the Philox stream.  With sync == false the call
// returns as soon as the work is enqueued: the host buffers must stay valid until the context is synchronised.
static int masks_generate_device_at(hm_context *ctx, size_t units, uint64_t seed, uint64_t first_unit, uint8_t *d_masks_out);
static int encrypt_host_impl(hm_context *ctx, const uint8_t *values, size_t n, uint32_t L, const uint8_t *masks, uint64_t seed, uint64_t first_unit,
                             bool sync, hm_batch **out) {
    if (!ctx || !out || (!values && n)) return HM_ERR_INVALID_ARGUMENT;
    if (!ctx->has_pk) return HM_ERR_PUBLIC_KEY_UNSET;
    if (L == 0 || L % 8 != 0 || L > (uint32_t)hmk::MAX_SLOTS) return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    const size_t vbytes = n * (L / 8), mbytes = n * L * ((ctx->tau + 7u) / 8u);
    PoolGuard dv(ctx), dm(ctx);
    CK(pool_alloc(ctx, &dv.p, std::max<size_t>(vbytes, 16)));
    const bool fused_masks = !masks && encrypt_fuses_masks(ctx, (uint64_t)n * L); // Philox inside the encrypt kernel: no mask buffer
    if (!fused_masks) CK(pool_alloc(ctx, &dm.p, std::max<size_t>(mbytes, 16)));
    if (n) {
        CK(cudaMemcpyAsync(dv.p, values, vbytes, cudaMemcpyHostToDevice, ctx->stream));
        if (masks) CK(cudaMemcpyAsync(dm.p, masks, mbytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    int rc = HM_OK;
    if (!masks && !fused_masks) rc = masks_generate_device_at(ctx, n * L, seed, first_unit, static_cast<uint8_t *>(dm.p));
    if (rc == HM_OK && fused_masks) {
        std::vector<uint64_t> degb(L, ctx->fresh_deg);
        int nb_status = HM_OK;
        hm_batch *b = new_batch(ctx, n, L, degb.data(), &nb_status);
        if (!b) return nb_status;
        rc = alloc_batch(ctx, b);
        if (rc == HM_OK) rc = encrypt_exec(ctx, static_cast<const uint8_t *>(dv.p), n, L, nullptr, b, seed, first_unit);
        if (rc != HM_OK) {
            hm_batch_free(ctx, b);
            return rc;
        }
        *out = b;
    } else if (rc == HM_OK) {
        rc = hm_encrypt_device(ctx, static_cast<const uint8_t *>(dv.p), n, L, static_cast<const uint8_t *>(dm.p), out);
    }
    if (sync) {
        cudaError_t e = cudaStreamSynchronize(ctx->stream); // the caller's host buffers may be reused after return
        if (rc == HM_OK && e != cudaSuccess) {
            hm_batch_free(ctx, *out);
            *out = nullptr;
            rc = fail_cuda(ctx, e, "encrypt");
        }
    }
    return rc;
}

int hm_encrypt(hm_context *ctx, const uint8_t *values, size_t n, uint32_t L, const uint8_t *masks, hm_batch **out) {
    if (!masks && n) return HM_ERR_INVALID_ARGUMENT;
    return encrypt_host_impl(ctx, values, n, L, masks ? masks : reinterpret_cast<const uint8_t *>(""), 0, 0, true, out);
}
int hm_encrypt_async(hm_context *ctx, const uint8_t *values, size_t n, uint32_t L, const uint8_t *masks, hm_batch **out) {
    if (!masks && n) return HM_ERR_INVALID_ARGUMENT;
    return encrypt_host_impl(ctx, values, n, L, masks ? masks : reinterpret_cast<const uint8_t *>(""), 0, 0, false, out);
}

// ---- seeded subset masks (Philox4x32-10) -----------------------------------------------------------
int hm_masks_generate_host(uint16_t tau, size_t units, uint64_t seed, uint8_t *masks_out) {
    if (!masks_out || tau == 0) return HM_ERR_INVALID_ARGUMENT;
    const uint32_t mb = (tau + 7u) / 8u, blocks = (mb + 15) / 16;
    for (size_t u = 0; u < units; ++u)
        for (uint32_t b = 0; b < blocks; ++b) {
            uint32_t r[4];
            hmk::philox4x32_10((uint32_t)u, (uint32_t)((uint64_t)u >> 32), b, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
            const uint32_t nb = std::min<uint32_t>(16, mb - 16 * b);
            for (uint32_t q = 0; q < nb; ++q) masks_out[u * mb + 16 * b + q] = (uint8_t)(r[q >> 2] >> (8 * (q & 3)));
        }
    return HM_OK;
}

static int masks_generate_device_at(hm_context *ctx, size_t units, uint64_t seed, uint64_t first_unit, uint8_t *d_masks_out) {
    if (!ctx || (!d_masks_out && units)) return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    if (!units) return HM_OK;
    const uint32_t mb = (ctx->tau + 7u) / 8u;
    const int grid = grid_for(ctx, (uint64_t)units * ((mb + 15) / 16), 256, 16);
    hmk::mask_fill_kernel<<<grid, 256, 0, ctx->stream>>>(d_masks_out, units, mb, seed, first_unit);
    LAUNCHED("mask_fill_kernel");
    return HM_OK;
}
int hm_masks_generate_device(hm_context *ctx, size_t units, uint64_t seed, uint8_t *d_masks_out) {
    return masks_generate_device_at(ctx, units, seed, 0, d_masks_out);
}

int hm_encrypt_seeded(hm_context *ctx, const uint8_t *values, size_t n, uint32_t L, uint64_t seed, hm_batch **out) {
    return encrypt_host_impl(ctx, values, n, L, nullptr, seed, 0, true, out);
}
int hm_encrypt_seeded_at(hm_context *ctx, const uint8_t *values, size_t n, uint32_t L, uint64_t seed, uint64_t first_unit, int sync,
                         hm_batch **out) {
    return encrypt_host_impl(ctx, values, n, L, nullptr, seed, first_unit, sync != 0, out);
}

int hm_encrypt_device_seeded_into(hm_context *ctx, const uint8_t *d_values, size_t n, uint32_t L, uint64_t seed, uint64_t first_unit,
                                  hm_batch *out) {
    if (!ctx || !out || (!d_values && n)) return HM_ERR_INVALID_ARGUMENT;
    if (!ctx->has_pk) return HM_ERR_PUBLIC_KEY_UNSET;
    if (out->n != n || out->L != L || L % 8 != 0) return HM_ERR_INVALID_ARGUMENT;
    for (uint32_t k = 0; k < L; ++k)
        if (out->w[k] != ctx->wf) return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    out->degb.assign(L, ctx->fresh_deg);
    if (encrypt_fuses_masks(ctx, (uint64_t)n * L)) return encrypt_exec(ctx, d_values, n, L, nullptr, out, seed, first_unit);
    PoolGuard dm(ctx);
    CK(pool_alloc(ctx, &dm.p, std::max<size_t>(n * L * ((ctx->tau + 7u) / 8u), 16)));
    int rc = masks_generate_device_at(ctx, n * L, seed, first_unit, static_cast<uint8_t *>(dm.p));
    if (rc == HM_OK) rc = encrypt_exec(ctx, d_values, n, L, static_cast<const uint8_t *>(dm.p), out);
    return rc;
}

int hm_decrypt_device(hm_context *ctx, const hm_batch *b, uint8_t *d_values_out) {
    if (!ctx || !b || (!d_values_out && b->n)) return HM_ERR_INVALID_ARGUMENT;
    if (!ctx->has_sk) return HM_ERR_SECRET_KEY_UNSET;       // context.rs:480-488
    if (b->L % 8 != 0) return HM_ERR_INVALID_LENGTH;        // cipher.rs:218
    USE_DEV(ctx);
    if (b->n == 0) return HM_OK;
    const uint64_t units = (uint64_t)b->n * b->L;
    bool uniform = true;
    for (uint32_t k = 1; k < b->L; ++k) uniform = uniform && (b->w[k] == b->w[0]);
    const uint32_t w = b->w[0];
    if (uniform && w <= 32 && ((uintptr_t)d_values_out % 4) == 0) {
        int rc = ensure_decrypt_vector(ctx, (size_t)w * 64);
        if (rc != HM_OK) return rc;
        const size_t smem = ((size_t)((w + 1) & ~1u) + 2 * (size_t)hmk::DEC_THREADS * w) * 8;
        CK(cudaFuncSetAttribute(hmk::decrypt_uniform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / (smem + 1024)));
        const int grid = grid_for(ctx, units, hmk::DEC_THREADS, per_sm);
        hmk::decrypt_uniform_kernel<<<grid, hmk::DEC_THREADS, smem, ctx->stream>>>(b->d, ctx->d_v, d_values_out, units, w);
        LAUNCHED("decrypt_uniform_kernel");
    } else if (!getenv("HM_DECRYPT_NO_TMA") && (b->value_words * 8) % 16 == 0 && b->value_words * 8 * 2 + 4096 <= ctx->smem_optin - 2048 &&
               b->value_words >= 256 && ((uintptr_t)b->d % 16) == 0) {
        // whole values streamed through shared memory by TMA (adder / multiplier results)
        uint32_t wmax = 0;
        for (uint32_t k = 0; k < b->L; ++k) wmax = std::max(wmax, b->w[k]);
        int rc = ensure_decrypt_vector(ctx, (size_t)wmax * 64);
        if (rc != HM_OK) return rc;
        const size_t vbytes = b->value_words * 8, extra = (size_t)wmax * 8 + 16;
        const size_t budget = ctx->smem_optin - 2048;
        static const int decv = getenv("HM_DECV") ? atoi(getenv("HM_DECV")) : 1;
#define DECV_LAUNCH(ST, TH, PER_SM)                                                                                   \
    {                                                                                                                 \
        auto kern = hmk::decrypt_value_tma_kernel<ST, TH>;                                                            \
        const size_t smem = vbytes * ST + extra;                                                                      \
        const int grid = (int)std::min<uint64_t>(b->n, (uint64_t)ctx->sm_count * PER_SM);                             \
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                        \
        kern<<<grid, TH, smem, ctx->stream>>>(b->d, ctx->d_v, wmax, d_values_out, b->n, make_layout(b));              \
    }
        if (decv == 1 && 4 * (vbytes * 4 + extra + 1024) <= budget) DECV_LAUNCH(4, 256, 4)
        else if (decv == 1 && 2 * (vbytes * 2 + extra + 1024) <= budget) DECV_LAUNCH(2, 256, 2)
        else if (decv == 2 && 4 * (vbytes * 1 + extra + 1024) <= budget) DECV_LAUNCH(1, 128, 4)
        else if (vbytes * 4 + extra <= budget) DECV_LAUNCH(4, 512, 1)
        else if (vbytes * 3 + extra <= budget) DECV_LAUNCH(3, 512, 1)
        else DECV_LAUNCH(2, 512, 1)
#undef DECV_LAUNCH
        LAUNCHED("decrypt_value_tma_kernel");
    } else {
        int rc = ensure_vv(ctx, b);
        if (rc != HM_OK) return rc;
        const uint64_t blocks = (units + 7) / 8;
        if (blocks > 0x7fffffffull) return HM_ERR_UNSUPPORTED;
        hmk::decrypt_slots_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(b->d, ctx->d_vv, d_values_out, units,
                                                                            make_layout(b));
        LAUNCHED("decrypt_slots_kernel");
    }
    return HM_OK;
}

static int decrypt_host_impl(hm_context *ctx, const hm_batch *b, uint8_t *values_out, bool sync) {
    if (!ctx || !b || (!values_out && b->n)) return HM_ERR_INVALID_ARGUMENT;
    if (!ctx->has_sk) return HM_ERR_SECRET_KEY_UNSET;
    if (b->L % 8 != 0) return HM_ERR_INVALID_LENGTH;
    USE_DEV(ctx);
    const size_t bytes = b->n * (b->L / 8);
    PoolGuard dout(ctx);
    CK(pool_alloc(ctx, &dout.p, std::max<size_t>(bytes, 16)));
    int rc = hm_decrypt_device(ctx, b, static_cast<uint8_t *>(dout.p));
    if (rc == HM_OK && bytes) {
        cudaError_t e = cudaMemcpyAsync(values_out, dout.p, bytes, cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) rc = fail_cuda(ctx, e, "download plaintext");
    }
    if (sync) {
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (rc == HM_OK && e != cudaSuccess) rc = fail_cuda(ctx, e, "decrypt sync");
    }
    return rc;
}
int hm_decrypt(hm_context *ctx, const hm_batch *b, uint8_t *values_out) { return decrypt_host_impl(ctx, b, values_out, true); }
int hm_decrypt_async(hm_context *ctx, const hm_batch *b, uint8_t *values_out) { return decrypt_host_impl(ctx, b, values_out, false); }

// ---- homomorphic operations -------------------------------------------------------------------
int hm_op_min_d_over_delta(int op) { // reference src/impls/numbers.rs:27-50
    switch (op) {
        case HM_OP_AND: return 2;
        case HM_OP_OR: return 2;
        case HM_OP_XOR: return 1;
        case HM_OP_NOT: return 1;
        case HM_OP_ADD: return 21;
        case HM_OP_MUL: return 64;
        default: return HM_ERR_INVALID_ARGUMENT;
    }
}

static int validate_operation(const hm_context *ctx, int op) { // context.rs:310-323
    const int req = hm_op_min_d_over_delta(op);
    if (req < 0) return req;
    if ((uint32_t)ctx->d < (uint32_t)req * (uint32_t)ctx->delta) return HM_ERR_OPERATION_REQUIREMENT;
    return HM_OK;
}

// Degree bounds of the result slots.  Shared by the planner and hm_result_slot_words.
static int result_bounds(int op, uint32_t L, const uint64_t *da, const uint64_t *db, std::vector<uint64_t> &out) {
    out.assign(L, 0);
    switch (op) {
        case HM_OP_XOR:
            for (uint32_t k = 0; k < L; ++k) out[k] = std::max(da[k], db[k]);
            return HM_OK;
        case HM_OP_AND:
        case HM_OP_OR:
            for (uint32_t k = 0; k < L; ++k) out[k] = da[k] + db[k];
            return HM_OK;
        case HM_OP_ADD: { // common.rs:37-56 with carry_0 = 0
            uint64_t dc = 0;
            bool czero = true;
            for (uint32_t k = 0; k < L; ++k) {
                const uint64_t dpk = std::max(da[k], db[k]);
                out[k] = czero ? dpk : std::max(dpk, dc);
                if (k + 1 >= L) break;
                const uint64_t dg = da[k] + db[k];
                if (czero) {
                    dc = dg; // carry_1 = a_0 * b_0
                    czero = false;
                } else {
                    const uint64_t dcpp = dpk + dc;
                    dc = dg + dcpp; // max(dcpp, dg + dcpp)
                }
            }
            return HM_OK;
        }
        case HM_OP_MUL: { // common.rs:66-105
            std::vector<uint64_t> carries;
            std::vector<uint8_t> czero_flags;
            size_t offset = 0;
            for (uint32_t i = 0; i < L; ++i) {
                const size_t cur = (size_t)i * (i + 1) / 2;
                uint64_t dr = 0;
                bool rz = true;
                for (uint32_t j = 0; j <= i; ++j) {
                    const uint64_t dpp = da[j] + db[i - j];
                    if (i + 1 < L) {
                        carries.push_back(rz ? 0 : dpp + dr);
                        czero_flags.push_back(rz ? 1 : 0);
                    }
                    dr = rz ? dpp : std::max(dr, dpp);
                    rz = false;
                }
                for (size_t j = 0; j < cur; ++j) {
                    const uint64_t dcj = carries[offset + j];
                    const bool cz = czero_flags[offset + j];
                    if (i + 1 < L) {
                        carries.push_back(cz ? 0 : dr + dcj);
                        czero_flags.push_back(cz ? 1 : 0);
                    }
                    if (!cz) dr = std::max(dr, dcj);
                }
                offset += cur;
                out[i] = dr;
                if (dr > ((uint64_t)1 << 31)) return HM_ERR_UNSUPPORTED; // infeasible growth
            }
            return HM_OK;
        }
        default: return HM_ERR_INVALID_ARGUMENT;
    }
}

int hm_result_slot_bounds(int op, uint32_t L, const uint64_t *a_bounds, const uint64_t *b_bounds, uint64_t *out_bounds) {
    if (!a_bounds || !b_bounds || !out_bounds || L == 0 || L > (uint32_t)hmk::MAX_SLOTS) return HM_ERR_INVALID_ARGUMENT;
    for (uint32_t k = 0; k < L; ++k)
        if (a_bounds[k] > MAX_DEGREE_BOUND || b_bounds[k] > MAX_DEGREE_BOUND) return HM_ERR_INVALID_ARGUMENT;
    std::vector<uint64_t> o;
    int rc = result_bounds(op, L, a_bounds, b_bounds, o);
    if (rc != HM_OK) return rc;
    for (uint32_t k = 0; k < L; ++k) out_bounds[k] = o[k];
    return HM_OK;
}

int hm_result_slot_words(const hm_context *ctx, int op, uint32_t L, const uint32_t *a_words, const uint32_t *b_words,
                         uint32_t *out_words) {
    if (!ctx || !a_words || !b_words || !out_words || L == 0 || L > (uint32_t)hmk::MAX_SLOTS) return HM_ERR_INVALID_ARGUMENT;
    std::vector<uint64_t> da(L), db(L), o;
    for (uint32_t k = 0; k < L; ++k) { // a width says nothing about the degree inside: widest case, like hm_batch_upload
        if (!a_words[k] || !b_words[k]) return HM_ERR_INVALID_ARGUMENT;
        da[k] = (uint64_t)a_words[k] * 64 - 1;
        db[k] = (uint64_t)b_words[k] * 64 - 1;
    }
    int rc = result_bounds(op, L, da.data(), db.data(), o);
    if (rc != HM_OK) return rc;
    for (uint32_t k = 0; k < L; ++k) out_words[k] = (uint32_t)(o[k] / 64 + 1);
    return HM_OK;
}

// generic ripple-carry adder from views (any parameters / any input widths), reference order
static int add_generic_seq(hm_context *ctx, const hm_batch *a, const hm_batch *b, hm_batch *o) {
    const uint32_t L = a->L;
    const size_t n = a->n;
    // arena per value: p, carry(2 buffers), cpp, cpp1, g, t — all at the final (largest) width
    uint64_t maxdeg = 0;
    for (uint32_t k = 0; k < L; ++k) maxdeg = std::max(maxdeg, o->degb[k]);
    uint64_t dgmax = 0;
    for (uint32_t k = 0; k < L; ++k) dgmax = std::max(dgmax, a->degb[k] + b->degb[k]);
    const uint32_t wmax = (uint32_t)((maxdeg + dgmax) / 64 + 2);
    const uint32_t NBUF = 7;
    PoolGuard arena_guard(ctx); // released in stream order on every exit path
    const size_t arena_words = (size_t)NBUF * wmax;
    CK(pool_alloc(ctx, &arena_guard.p, std::max<size_t>(n * arena_words * 8, 16)));
    uint64_t *arena = static_cast<uint64_t *>(arena_guard.p);
    CK(cudaMemsetAsync(arena, 0, n * arena_words * 8, ctx->stream));
    auto buf = [&](uint32_t i, uint64_t degb) {
        View v;
        v.base = arena;
        v.stride = arena_words;
        v.off = i * wmax;
        v.w = (uint32_t)(degb / 64 + 1);
        v.deg = degb;
        return v;
    };
    int rc = HM_OK;
    uint64_t dc = 0;
    bool czero = true;
    uint32_t cur = 1; // carry buffers are 1 and 2
    for (uint32_t k = 0; k < L && rc == HM_OK; ++k) {
        const uint64_t dpk = std::max(a->degb[k], b->degb[k]);
        View p = buf(0, dpk);
        rc = launch_xor_views(ctx, p, slot_view(a, k), slot_view(b, k), n); // cb1.xor(cb2)
        if (rc != HM_OK) break;
        View sk = slot_view(o, k);
        rc = launch_xor_views(ctx, sk, p, czero ? null_view() : buf(cur, dc), n); // .xor(&carry)
        if (rc != HM_OK || k + 1 >= L) break;
        const uint64_t dg = a->degb[k] + b->degb[k];
        View g = buf(5, dg);
        std::vector<MulOp> ops(1);
        ops[0] = MulOp{slot_view(a, k), slot_view(b, k), g}; // cb1.and(cb2)
        rc = launch_mul_ops(ctx, ops, n);
        if (rc != HM_OK) break;
        if (czero) { // c_p1_p2 = p * 0 = 0, carry = 0 + g * (0 + 1) = g
            rc = launch_xor_views(ctx, buf(cur, dg), g, null_view(), n);
            dc = dg;
            czero = false;
            continue;
        }
        const uint64_t dcpp = dpk + dc;
        View cpp = buf(3, dcpp);
        ops[0] = MulOp{p, buf(cur, dc), cpp}; // c_p1_p2 = p.and(carry)
        rc = launch_mul_ops(ctx, ops, n);
        if (rc != HM_OK) break;
        View cpp1 = buf(4, dcpp);
        rc = launch_xor_views(ctx, cpp1, cpp, null_view(), n); // copy, then + 1
        if (rc != HM_OK) break;
        {
            Layout one;
            one.L = 1;
            one.value_words = (uint32_t)arena_words;
            one.off[0] = cpp1.off;
            one.off[1] = cpp1.off + cpp1.w;
            const int grid = grid_for(ctx, n, 256, 16);
            hmk::not_kernel<<<grid, 256, 0, ctx->stream>>>(arena, one, n); // c_p1_p2.xor(&one_bit)
            rc = post_launch(ctx, "not_kernel");
            if (rc != HM_OK) break;
        }
        const uint64_t dt = dg + dcpp;
        View t = buf(6, dt);
        ops[0] = MulOp{g, cpp1, t};
        rc = launch_mul_ops(ctx, ops, n);
        if (rc != HM_OK) break;
        const uint32_t nxt = 3 - cur;
        rc = launch_xor_views(ctx, buf(nxt, dt), cpp, t, n); // carry = c_p1_p2.xor(...)
        cur = nxt;
        dc = dt;
    }
    return rc;
}

// one launch for a list of slot XORs (o = a ^ b per entry)
static int launch_xor_ops(hm_context *ctx, const std::vector<MulOp> &ops, size_t n) {
    if (ops.empty() || n == 0) return HM_OK;
    if (ops.size() > 65535) return HM_ERR_UNSUPPORTED;
    size_t slot = 0;
    int rc = ensure_ops(ctx, ops.size(), &slot);
    if (rc != HM_OK) return rc;
    memcpy(ctx->h_ops + slot, ops.data(), ops.size() * sizeof(MulOp));
    CK(cudaMemcpyAsync(ctx->d_ops + slot, ctx->h_ops + slot, ops.size() * sizeof(MulOp), cudaMemcpyHostToDevice, ctx->stream));
    uint32_t wmax = 1;
    for (const MulOp &op : ops) wmax = std::max(wmax, op.o.w);
    const int per_op = std::max(1, grid_for(ctx, (uint64_t)n * wmax, 256, 16) / (int)std::min<size_t>(ops.size(), 16));
    const dim3 grid((unsigned)per_op, (unsigned)ops.size());
    hmk::xor_ops_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_ops + slot, n);
    return post_launch(ctx, "xor_ops_kernel");
}

// Regrouped plan of the same ripple-carry circuit for any degree bounds (the fused kernels cover D in {128, 256}).
// The reference computes carry' = cpp + g (cpp + 1) with cpp = p carry (common.rs:44-53), two long products per bit;
// expanded over GF(2) that is carry' = (p + g p) carry + g = m carry + g, the form the fused kernels use: the same
// polynomial, one long product per bit.  All p_k, g_k, m_k are independent of the chain and are produced by three batched
// launches; the chain itself is one product + one XOR per bit, written straight into the output slots (s_k = c_k + p_k
// is completed by one batched XOR at the end).
static int add_generic(hm_context *ctx, const hm_batch *a, const hm_batch *b, hm_batch *o) {
    if (g_adder_generic_seq) return add_generic_seq(ctx, a, b, o);
    const uint32_t L = a->L;
    const size_t n = a->n;
    std::vector<uint64_t> dp(L), dg(L), dm(L);
    std::vector<uint64_t> offP(L), offG(L), offM(L);
    uint64_t cursor = 0;
    for (uint32_t k = 0; k < L; ++k) {
        dp[k] = std::max(a->degb[k], b->degb[k]);
        dg[k] = a->degb[k] + b->degb[k];
        dm[k] = dg[k] + dp[k];
        offP[k] = cursor; cursor += dp[k] / 64 + 1;
        offG[k] = cursor; cursor += dg[k] / 64 + 1;
        offM[k] = cursor; cursor += dm[k] / 64 + 1;
    }
    const size_t arena_words = cursor;
    if (arena_words >> 32) return HM_ERR_UNSUPPORTED;
    PoolGuard arena_guard(ctx); // released in stream order on every exit path
    CK(pool_alloc(ctx, &arena_guard.p, std::max<size_t>(n * arena_words * 8, 16)));
    uint64_t *arena = static_cast<uint64_t *>(arena_guard.p);
    // products may be accumulated with atomics (thread-per-chunk kernel): clear their destinations once
    CK(cudaMemsetAsync(arena, 0, std::max<size_t>(n * arena_words * 8, 16), ctx->stream));
    CK(cudaMemsetAsync(o->d, 0, std::max<size_t>(n * o->value_words * 8, 16), ctx->stream));
    auto aview = [&](uint64_t off, uint64_t degb) {
        View v;
        v.base = arena;
        v.stride = arena_words;
        v.off = (uint32_t)off;
        v.w = (uint32_t)(degb / 64 + 1);
        v.deg = degb;
        return v;
    };
    int rc = HM_OK;
    std::vector<MulOp> ops;
    for (uint32_t k = 0; k < L; ++k) ops.push_back(MulOp{slot_view(a, k), slot_view(b, k), aview(offP[k], dp[k])}); // p = a + b
    rc = launch_xor_ops(ctx, ops, n);
    if (rc == HM_OK && L >= 2) {
        ops.clear();
        for (uint32_t k = 0; k + 1 < L; ++k) ops.push_back(MulOp{slot_view(a, k), slot_view(b, k), aview(offG[k], dg[k])}); // g = a b
        rc = launch_mul_ops(ctx, ops, n, true);
    }
    if (rc == HM_OK && L >= 3) {
        ops.clear();
        for (uint32_t k = 1; k + 1 < L; ++k) ops.push_back(MulOp{aview(offG[k], dg[k]), aview(offP[k], dp[k]), aview(offM[k], dm[k])}); // g p
        rc = launch_mul_ops(ctx, ops, n, true);
        if (rc == HM_OK) {
            ops.clear();
            for (uint32_t k = 1; k + 1 < L; ++k) ops.push_back(MulOp{aview(offM[k], dm[k]), aview(offP[k], dp[k]), aview(offM[k], dm[k])}); // m = g p + p
            rc = launch_xor_ops(ctx, ops, n);
        }
    }
    // chain: slot k+1 <- c_{k+1} = m_k c_k + g_k   (c_1 = g_0)
    uint64_t dc = 0;
    std::vector<MulOp> one(1);
    for (uint32_t k = 0; k + 1 < L && rc == HM_OK; ++k) {
        View dst = slot_view(o, k + 1);
        if (k == 0) {
            dc = dg[0];
        } else {
            View ck = slot_view(o, k);
            ck.w = (uint32_t)(dc / 64 + 1);
            ck.deg = dc;
            dc = dm[k] + dc;
            View ok = dst;
            ok.w = (uint32_t)std::min<uint64_t>(dst.w, dc / 64 + 1);
            ok.deg = dc;
            one[0] = MulOp{aview(offM[k], dm[k]), ck, ok};
            rc = launch_mul_ops(ctx, one, n, true);
            if (rc != HM_OK) break;
        }
        if (dc > o->degb[k + 1]) { // cannot happen: result_bounds follows the same recurrence
            rc = HM_ERR_UNSUPPORTED;
            break;
        }
        View gk = aview(offG[k], dg[k]);
        View og = dst;
        og.w = std::min(dst.w, gk.w);
        rc = launch_xor_views(ctx, og, og, gk, n); // + g_k over g_k's width
    }
    if (rc == HM_OK) { // s_k = c_k + p_k
        ops.clear();
        for (uint32_t k = 0; k < L; ++k) {
            View pk = aview(offP[k], dp[k]);
            View ok = slot_view(o, k);
            ok.w = std::min(ok.w, pk.w);
            ops.push_back(MulOp{ok, pk, ok});
        }
        rc = launch_xor_ops(ctx, ops, n);
    }
    return rc;
}

// generic unsigned multiplier circuit, reference common.rs:66-105
static int mul_generic_seq(hm_context *ctx, const hm_batch *a, const hm_batch *b, hm_batch *o) {
    const uint32_t L = a->L;
    const size_t n = a->n;
    struct Obj {
        uint32_t off, w_alloc;
        uint64_t degb;
        bool zero;
    };
    // plan: pp[j][k] for j+k < L, carries in push order; results live in the output batch
    std::vector<std::vector<Obj>> pp(L, std::vector<Obj>(L));
    uint64_t cursor = 0;
    auto alloc = [&](uint64_t degb) {
        Obj ob;
        ob.off = (uint32_t)cursor;
        ob.w_alloc = (uint32_t)(degb / 64 + 1);
        ob.degb = degb;
        ob.zero = false;
        cursor += ob.w_alloc;
        return ob;
    };
    for (uint32_t j = 0; j < L; ++j)
        for (uint32_t k = 0; j + k < L; ++k) pp[j][k] = alloc(a->degb[j] + b->degb[k]);
    // dry run for carry shapes
    struct Step {
        int kind; // 0 mul(o, x, y)   1 xor(res_i ^= x)
        Obj o, x, y;
        int res_i;
        bool x_is_res, y_is_res;
    };
    std::vector<Obj> carries;
    std::vector<Step> steps;
    std::vector<uint64_t> dres(L, 0);
    std::vector<bool> rzero(L, true);
    size_t offset = 0;
    for (uint32_t i = 0; i < L; ++i) {
        const size_t curlen = (size_t)i * (i + 1) / 2;
        auto res_obj = [&](uint32_t idx) {
            Obj r;
            r.off = o->off[idx];
            r.w_alloc = o->w[idx];
            r.degb = dres[idx];
            r.zero = rzero[idx];
            return r;
        };
        for (uint32_t j = 0; j <= i; ++j) {
            const Obj &ppo = pp[j][i - j];
            if (i + 1 < L) {
                Obj c;
                if (rzero[i]) {
                    c = Obj{0, 0, 0, true};
                } else {
                    c = alloc(ppo.degb + dres[i]);
                    steps.push_back(Step{0, c, ppo, res_obj(i), (int)i, false, true});
                }
                carries.push_back(c);
            }
            steps.push_back(Step{1, Obj{}, ppo, Obj{}, (int)i, false, false});
            dres[i] = rzero[i] ? ppo.degb : std::max(dres[i], ppo.degb);
            rzero[i] = false;
        }
        for (size_t j = 0; j < curlen; ++j) {
            const Obj cj = carries[offset + j];
            if (i + 1 < L) {
                Obj c;
                if (cj.zero) {
                    c = Obj{0, 0, 0, true};
                } else {
                    c = alloc(dres[i] + cj.degb);
                    steps.push_back(Step{0, c, res_obj(i), cj, (int)i, true, false});
                }
                carries.push_back(c);
            }
            if (!cj.zero) {
                steps.push_back(Step{1, Obj{}, cj, Obj{}, (int)i, false, false});
                dres[i] = std::max(dres[i], cj.degb);
            }
        }
        offset += curlen;
    }
    const size_t arena_words = cursor;
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    if ((double)n * arena_words * 8.0 > 0.9 * (double)free_b) return HM_ERR_UNSUPPORTED;
    PoolGuard arena_guard(ctx); // released in stream order on every exit path
    CK(pool_alloc(ctx, &arena_guard.p, std::max<size_t>(n * arena_words * 8, 16)));
    uint64_t *arena = static_cast<uint64_t *>(arena_guard.p);
    CK(cudaMemsetAsync(o->d, 0, n * o->value_words * 8, ctx->stream));
    auto aview = [&](const Obj &ob) {
        View v;
        v.base = arena;
        v.stride = arena_words;
        v.off = ob.off;
        v.w = (uint32_t)(ob.degb / 64 + 1);
        v.deg = ob.degb;
        return v;
    };
    auto rview = [&](const Obj &ob) {
        View v;
        v.base = o->d;
        v.stride = o->value_words;
        v.off = ob.off;
        v.w = (uint32_t)std::min<uint64_t>(ob.w_alloc, ob.degb / 64 + 1);
        v.deg = ob.degb;
        return v;
    };
    int rc = HM_OK;
    { // all partial products in one launch (independent)
        std::vector<MulOp> ops;
        for (uint32_t j = 0; j < L; ++j)
            for (uint32_t k = 0; j + k < L; ++k) ops.push_back(MulOp{slot_view(a, j), slot_view(b, k), aview(pp[j][k])});
        rc = launch_mul_ops(ctx, ops, n);
    }
    std::vector<MulOp> one(1);
    for (size_t s = 0; s < steps.size() && rc == HM_OK; ++s) {
        const Step &st = steps[s];
        if (st.kind == 0) {
            one[0] = MulOp{st.x_is_res ? rview(st.x) : aview(st.x), st.y_is_res ? rview(st.y) : aview(st.y), aview(st.o)};
            rc = launch_mul_ops(ctx, one, n);
        } else {
            View r;
            r.base = o->d;
            r.stride = o->value_words;
            r.off = o->off[st.res_i];
            r.w = o->w[st.res_i];
            r.deg = o->degb[st.res_i];
            View x = aview(st.x);
            // r ^= x over x's width only (words above are untouched)
            View rr = r;
            rr.w = std::min(r.w, x.w);
            rc = launch_xor_views(ctx, rr, rr, x, n);
        }
    }
    return rc;
}

// Column-batched plan of the same circuit.  In column i the reference XORs the items x_1..x_m (the column's partial
// products, then the carries of column i-1, common.rs:78-101) into result[i] one at a time and, before each XOR, pushes
// the carry x_t * result[i]; result[i] at that moment is the prefix P_{t-1} = x_1 ^ ... ^ x_{t-1}.  So the same carries
// (the same polynomials, bit for bit) are c_t = x_t * P_{t-1}: one prefix pass writes P_2..P_{m-1} and result[i] = P_m,
// and the column's m-1 products are independent and go out as one batch of launches.  ~30 launches per u8 multiply
// instead of ~180, and each product launch has enough (value, chunk) threads to fill the GPU.
// One launch for the whole circuit (kernels_mul.cu) when both operands are fresh u8 batches at D = 256:
// a warp per value, partial products, prefixes and carries in shared memory.  Returns HM_ERR_UNSUPPORTED when it does not apply.
static int mul_fused(hm_context *ctx, const hm_batch *a, const hm_batch *b, hm_batch *o) {
    if (a->L != hmk::K7_L || ctx->fresh_deg != hmk::K7_D || !ctx->has_pk) return HM_ERR_UNSUPPORTED;
    for (uint32_t k = 0; k < a->L; ++k)
        if (a->degb[k] != hmk::K7_D || b->degb[k] != hmk::K7_D || a->w[k] != hmk::K7_D / 64 + 1 || b->w[k] != hmk::K7_D / 64 + 1) return HM_ERR_UNSUPPORTED;
    hmk::K7Host host;
    static const int k7_acc = getenv("HM_K7_ACC") ? atoi(getenv("HM_K7_ACC")) : 8;
    if (!hmk::k7_build_plan(make_layout(o), o->degb.data(), (uint32_t)k7_acc, &host)) return HM_ERR_UNSUPPORTED;
    static const int k7_warps = getenv("HM_K7_WARPS") ? atoi(getenv("HM_K7_WARPS")) : 8;
    if (hmk::k7_smem_bytes(host.plan, k7_warps) > ctx->smem_optin) return HM_ERR_UNSUPPORTED;
    if (!ctx->k7_ready || memcmp(&ctx->k7_plan, &host.plan, sizeof(hmk::K7Plan)) != 0) {
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->k7_ready = false;
        if (ctx->d_k7_items) cudaFree(ctx->d_k7_items);
        if (ctx->d_k7_prods) cudaFree(ctx->d_k7_prods);
        if (ctx->d_k7_units) cudaFree(ctx->d_k7_units);
        ctx->d_k7_items = nullptr;
        ctx->d_k7_prods = nullptr;
        ctx->d_k7_units = nullptr;
        CK(cudaMalloc(&ctx->d_k7_items, host.items.size() * sizeof(hmk::K7Item)));
        CK(cudaMalloc(&ctx->d_k7_prods, host.prods.size() * sizeof(hmk::K7Prod)));
        CK(cudaMalloc(&ctx->d_k7_units, host.units.size() * sizeof(hmk::K7Unit)));
        CK(cudaMemcpy(ctx->d_k7_items, host.items.data(), host.items.size() * sizeof(hmk::K7Item), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ctx->d_k7_prods, host.prods.data(), host.prods.size() * sizeof(hmk::K7Prod), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ctx->d_k7_units, host.units.data(), host.units.size() * sizeof(hmk::K7Unit), cudaMemcpyHostToDevice));
        memcpy(&ctx->k7_plan, &host.plan, sizeof(hmk::K7Plan));
        ctx->k7_ready = true;
    }
    CK(hmk::launch_mul_circuit_fused(k7_warps, a->d, b->d, o->d, a->n, o->value_words, ctx->k7_plan, ctx->d_k7_items, ctx->d_k7_prods, ctx->d_k7_units,
                                     ctx->sm_count, ctx->stream));
    return post_launch(ctx, "mul_circuit_fused_kernel");
}

static int mul_generic(hm_context *ctx, const hm_batch *a, const hm_batch *b, hm_batch *o) {
    if (g_mul_circuit_seq) return mul_generic_seq(ctx, a, b, o);
    if (g_mul_circuit_fused) {
        const int rcf = mul_fused(ctx, a, b, o);
        if (rcf != HM_ERR_UNSUPPORTED) return rcf;
    }
    const uint32_t L = a->L;
    const size_t n = a->n;
    struct Obj {
        uint64_t off; // u64 words from the start of a value's arena
        uint64_t degb;
    };
    uint64_t cursor = 0;
    auto alloc = [&](uint64_t degb) {
        Obj ob{cursor, degb};
        cursor += degb / 64 + 1;
        return ob;
    };
    std::vector<std::vector<Obj>> pp(L, std::vector<Obj>(L));
    for (uint32_t j = 0; j < L; ++j)
        for (uint32_t k = 0; j + k < L; ++k) pp[j][k] = alloc(a->degb[j] + b->degb[k]);
    const uint64_t pp_words = cursor;
    struct Column {
        std::vector<Obj> items, prefix, carry; // prefix[t] = P_{t+1} for 1 <= t+1 < m-1 (arena); carry[t] = x_{t+1} * P_t
    };
    std::vector<Column> cols(L);
    std::vector<Obj> incoming; // non-zero carries of the previous column, in push order
    for (uint32_t i = 0; i < L; ++i) {
        Column &c = cols[i];
        for (uint32_t j = 0; j <= i; ++j) c.items.push_back(pp[j][i - j]);
        c.items.insert(c.items.end(), incoming.begin(), incoming.end());
        incoming.clear();
        uint64_t dres = c.items[0].degb;
        for (size_t t = 1; t < c.items.size(); ++t) {
            if (i + 1 < L) {
                // P_t covers x_1..x_t; P_1 is x_1 itself, later ones need storage (the last prefix is the result slot)
                if (t >= 2) c.prefix.push_back(alloc(dres));
                c.carry.push_back(alloc(c.items[t].degb + dres));
                incoming.push_back(c.carry.back());
            }
            dres = std::max(dres, c.items[t].degb);
        }
        // L = 8 has at most 36 items per column, L = 16 at most 136; the bounds cannot differ (both follow result_bounds)
        if (c.items.size() > hmk::PREFIX_MAX_ITEMS || dres != o->degb[i]) return mul_generic_seq(ctx, a, b, o);
    }
    const size_t arena_words = cursor;
    if (arena_words >> 32) return HM_ERR_UNSUPPORTED;
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    if ((double)n * arena_words * 8.0 > 0.9 * (double)free_b) return HM_ERR_UNSUPPORTED;
    PoolGuard arena_guard(ctx); // released in stream order on every exit path
    CK(pool_alloc(ctx, &arena_guard.p, std::max<size_t>(n * arena_words * 8, 16)));
    uint64_t *arena = static_cast<uint64_t *>(arena_guard.p);
    // carries are accumulated with atomics by the thread-per-chunk kernel: clear the arena once instead of per product
    CK(cudaMemsetAsync(arena, 0, std::max<size_t>(n * arena_words * 8, 16), ctx->stream));
    (void)pp_words;
    auto aview = [&](const Obj &ob) {
        View v;
        v.base = arena;
        v.stride = arena_words;
        v.off = (uint32_t)ob.off;
        v.w = (uint32_t)(ob.degb / 64 + 1);
        v.deg = ob.degb;
        return v;
    };
    int rc = HM_OK;
    { // all partial products in one batch (independent)
        std::vector<MulOp> ops;
        for (uint32_t j = 0; j < L; ++j)
            for (uint32_t k = 0; j + k < L; ++k) ops.push_back(MulOp{slot_view(a, j), slot_view(b, k), aview(pp[j][k])});
        rc = launch_mul_ops(ctx, ops, n, true);
    }
    std::vector<MulOp> pre, ops, small;
    for (uint32_t i = 0; i < L && rc == HM_OK; ++i) {
        const Column &c = cols[i];
        const size_t m = c.items.size();
        pre.clear();
        ops.clear();
        for (size_t t = 0; t < m; ++t) {
            View dst = null_view();
            if (t + 1 == m) dst = slot_view(o, i);
            else if (t >= 1 && !c.prefix.empty()) dst = aview(c.prefix[t - 1]);
            pre.push_back(MulOp{aview(c.items[t]), null_view(), dst});
        }
        size_t slot = 0;
        rc = ensure_ops(ctx, pre.size(), &slot);
        if (rc != HM_OK) break;
        memcpy(ctx->h_ops + slot, pre.data(), pre.size() * sizeof(MulOp));
        CK(cudaMemcpyAsync(ctx->d_ops + slot, ctx->h_ops + slot, pre.size() * sizeof(MulOp), cudaMemcpyHostToDevice, ctx->stream));
        const uint32_t width = o->w[i];
        hmk::prefix_xor_kernel<<<grid_for(ctx, (uint64_t)n * width, 256, 16), 256, 0, ctx->stream>>>(ctx->d_ops + slot, (uint32_t)pre.size(),
                                                                                                   width, n);
        LAUNCHED("prefix_xor_kernel");
        if (i + 1 < L) {
            // the column's products are independent: the big ones (thread-per-chunk kernel) go on the context's stream,
            // the small register-resident ones on the side stream so that they fill the big launch's tail
            small.clear();
            for (size_t t = 1; t < m; ++t) {
                const MulOp op{aview(c.items[t]), t == 1 ? aview(c.items[0]) : aview(c.prefix[t - 2]), aview(c.carry[t - 1])};
                (mul_shape_class(op) ? small : ops).push_back(op);
            }
            const bool fork = !small.empty() && !ops.empty();
            if (fork) {
                CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
                CK(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
            }
            rc = launch_mul_ops(ctx, ops, n, true);
            if (rc == HM_OK && !small.empty()) {
                cudaStream_t main_stream = ctx->stream;
                if (fork) ctx->stream = ctx->side_stream;
                rc = launch_mul_ops(ctx, small, n, true);
                ctx->stream = main_stream;
                if (fork) {
                    CK(cudaEventRecord(ctx->ev_join, ctx->side_stream));
                    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
                }
            }
        }
    }
    return rc;
}

// The fused ripple-carry kernel is instantiated for D = d+d' in {128, 256} (WD = D/32 = 4, 8; its multiplier mask must fit
// 32 bits, i.e. 3 WD + 1 <= 32) and needs both operands fresh (every slot with degree bound D).  Anything else goes
// through the generic view-based circuit.
static int adder_fast_path_wd(const hm_context *ctx, const hm_batch *a, const hm_batch *b, size_t *smem_per_warp) {
    if (!ctx->has_pk) return 0;
    const uint64_t D = ctx->fresh_deg;
    if (D != 128 && D != 256) return 0;
    for (uint32_t k = 0; k < a->L; ++k)
        if (a->degb[k] != D || b->degb[k] != D) return 0;
    if (a->L < 2) return 0;
    const int wd = (int)(D / 32);
    const size_t words = wd == 4 ? hmk::AdderCfg<4>::warp_words(a->L) : hmk::AdderCfg<8>::warp_words(a->L);
    *smem_per_warp = words * 4;
    return words * 4 <= ctx->smem_optin ? wd : 0;
}

// the dynamically scheduled thread-per-value chain (kernels_adder.cu): scheduler state, phase plan, launch
static int launch_chain_adder(hm_context *ctx, int wd, int variant, int ctas_per_sm, const hm_batch *a, const hm_batch *b, hm_batch *o) {
    const size_t n = a->n;
    const size_t words = hmk::adder_chain_sched_words(n);
    if (ctx->sched_words < words) {
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_sched) cudaFree(ctx->d_sched);
        ctx->d_sched = nullptr;
        ctx->sched_words = 0;
        CK(cudaMalloc(&ctx->d_sched, words * 4));
        ctx->sched_words = words;
    }
    CK(cudaMemsetAsync(ctx->d_sched, 0, words * 4, ctx->stream));
    hmk::AdderSched sc;
    sc.counter = ctx->d_sched;
    sc.done = ctx->d_sched + 1;
    sc.ngroups = (uint32_t)((n + 31) / 32);
    // Work units per value.  Splitting a value into phases pays when the batch is several waves of resident warps with a
    // ragged last wave (2^18 values = 3.46 waves: 4.44 -> 4.51 M adds/s with 8 phases); on one or two exact waves the
    // hand-over between warps only costs (75 776 values: 17.7 ms with 1 phase, 21.4 ms with 8).
    const int phases = (int)g_adder_phases;
    const uint64_t resident_warps = (uint64_t)ctx->sm_count * ctas_per_sm * 4;
    // The D = 1024 chain (2 CTAs per SM) measured: one exact wave 265 k adds/s with 1 phase, 242 k with 8; 1.73 waves 230 k / 274 k.
    const bool many_waves = wd == 32 ? (uint64_t)sc.ngroups * 10 >= resident_warps * 11 : (uint64_t)sc.ngroups * 2 >= resident_warps * 5;
    const int auto_phases = many_waves ? 8 : 1;
    hmk::adder_chain_plan(a->L, wd, (uint32_t)(phases > 0 ? phases : auto_phases), &sc);
    CK(hmk::launch_adder_chain(wd, variant, a->d, b->d, o->d, n, a->L, make_layout(o), sc, ctx->sm_count, ctx->stream));
    return post_launch(ctx, wd == 32 ? "adder_chain_wide_kernel" : "adder_chain_kernel");
}

// D = 1024 (config B): fresh operands on both sides and at least one value per resident thread of the thread-per-value chain
// (2 CTAs of 128 threads per SM); smaller batches use the regrouped generic plan, which fills the GPU with (value, chunk)
// threads (measured, u32 adds/s: 4 096 values 173 k generic / 49 k chain, 16 384: 220 k / 195 k, 37 888: 260 k / 265 k,
// 65 536: 258 k / 274 k)
static bool adder_wide_path(const hm_context *ctx, const hm_batch *a, const hm_batch *b) {
    if (!ctx->has_pk || ctx->fresh_deg != 1024 || a->L < 2 || g_adder_wide_min < 0) return false;
    for (uint32_t k = 0; k < a->L; ++k)
        if (a->degb[k] != 1024 || b->degb[k] != 1024) return false;
    return a->n >= (size_t)(g_adder_wide_min > 0 ? g_adder_wide_min : (long)ctx->sm_count * 256);
}

// runs `op` into the already allocated result batch o (layout = result_bounds of the operands)
static int apply2_exec(hm_context *ctx, int op, const hm_batch *a, const hm_batch *b, hm_batch *o, bool force_generic) {
    const size_t n = a->n;
    if (n == 0) return HM_OK;
    int rc = HM_OK;
    switch (op) {
        case HM_OP_XOR: rc = xor_batches(ctx, a, b, o); break;
        case HM_OP_AND:
        case HM_OP_OR: {
            std::vector<MulOp> ops;
            bool flat = same_layout(a, b);
            for (uint32_t k = 0; k < a->L && flat; ++k)
                flat = a->w[k] == a->w[0] && a->degb[k] == a->degb[0] && b->degb[k] == b->degb[0] && o->w[k] == o->w[0];
            if (flat && (uint64_t)n * a->L < ((uint64_t)1 << 40)) {
                // every slot has the same shape: one product per (value, slot) over the flattened batch, so that
                // neighbouring threads touch neighbouring memory
                View va = slot_view(a, 0), vb = slot_view(b, 0), vo = slot_view(o, 0);
                va.stride = va.w;
                vb.stride = vb.w;
                vo.stride = vo.w;
                ops.push_back(MulOp{va, vb, vo});
                int fused = op == HM_OP_OR;
                rc = launch_mul_ops(ctx, ops, n * a->L, false, &fused);
                if (rc == HM_OK && op == HM_OP_OR && !fused) {
                    rc = launch_xor_views(ctx, vo, vo, va, n * a->L);
                    if (rc == HM_OK) rc = launch_xor_views(ctx, vo, vo, vb, n * a->L);
                }
                break;
            }
            for (uint32_t k = 0; k < a->L; ++k) ops.push_back(MulOp{slot_view(a, k), slot_view(b, k), slot_view(o, k)});
            int fused = op == HM_OP_OR;
            rc = launch_mul_ops(ctx, ops, n, false, &fused);
            if (rc == HM_OK && op == HM_OP_OR && !fused) { // a + b + a*b, cipher.rs:76-83
                for (uint32_t k = 0; k < a->L && rc == HM_OK; ++k) {
                    View ok = slot_view(o, k);
                    rc = launch_xor_views(ctx, ok, ok, slot_view(a, k), n);
                    if (rc == HM_OK) rc = launch_xor_views(ctx, ok, ok, slot_view(b, k), n);
                }
            }
            break;
        }
        case HM_OP_ADD: {
            size_t per_warp = 0;
            if (!force_generic && adder_wide_path(ctx, a, b)) {
                rc = launch_chain_adder(ctx, 32, 2, 2, a, b, o);
                break;
            }
            const int wd = force_generic ? 0 : adder_fast_path_wd(ctx, a, b, &per_warp);
            if (wd) {
                int warps = (int)std::min<size_t>(4, ctx->smem_optin / per_warp);
                if (warps < 1) warps = 1;
                const size_t smem = per_warp * warps;
                // HM_ADDER_MODE: 0 / 1 = warp-per-value comb kernel (pair / single uniform branches), m >= 3 = thread-per-value
                // Karatsuba kernel with m-1 CTAs per SM (default: 4 CTAs of 128 threads, 128 registers)
                static const int mode = getenv("HM_ADDER_MODE") ? atoi(getenv("HM_ADDER_MODE")) : 5;
                // the thread-per-value kernel needs tens of thousands of values to fill the GPU (one value per thread);
                // smaller batches (and the chunks of the host pipeline) use the warp-per-value kernel
                // HM_ADDER_CHAIN: 0 = the round-1 thread kernels below; otherwise the dynamically scheduled chain of kernels_adder.cu,
                // value = 10 * (window in shared memory) + CTAs per SM.  HM_ADDER_PHASES = work units per value (default 4).
                const int chain = (int)g_adder_chain;
                const bool big = n >= (g_adder_thread_min >= 0 ? (size_t)g_adder_thread_min : (size_t)ctx->sm_count * 96);
                if (mode >= 3 && chain && big) {
                    rc = launch_chain_adder(ctx, wd, chain, chain % 10, a, b, o);
                    break;
                }
                if (mode >= 3 && wd == 8 && big) { // round-1 thread-per-value kernels; mode - 1 = CTAs per SM (cross-over measured: tools/adder_crossover.py)
                    const int per_sm = (mode - 1 == 3) ? 3 : 4; // 128-thread CTAs per SM (168 or 128 registers per thread)
                    const int blocks = (int)std::min<uint64_t>((n + 127) / 128, (uint64_t)ctx->sm_count * per_sm);
                    static const int use_smem = getenv("HM_ADDER_SMEM") ? atoi(getenv("HM_ADDER_SMEM")) : 1; // 0 = first thread kernel (scratch in global memory)
                    if (use_smem && (per_sm == 3 || per_sm == 4)) { // working set in shared memory, chunks prefetched by cp.async
                        auto sk = per_sm == 3 ? hmk::adder_thread_smem_kernel<3> : hmk::adder_thread_smem_kernel<4>;
                        CK(cudaFuncSetAttribute(sk, cudaFuncAttributeMaxDynamicSharedMemorySize, hmk::ADS_SMEM_BYTES));
                        CK(cudaFuncSetAttribute(sk, cudaFuncAttributePreferredSharedMemoryCarveout, use_smem > 1 ? use_smem : 70));
                        sk<<<blocks, 128, hmk::ADS_SMEM_BYTES, ctx->stream>>>(a->d, b->d, o->d, n, a->L, make_layout(o));
                        rc = post_launch(ctx, "adder_thread_smem_kernel");
                        break;
                    }
                    PoolGuard scratch_guard(ctx);
                    CK(pool_alloc(ctx, &scratch_guard.p, (size_t)blocks * 128 * hmk::ADT_THREAD_WORDS * 4));
                    uint32_t *scratch = static_cast<uint32_t *>(scratch_guard.p);
                    auto tk = hmk::adder_thread_kernel<4>; // the first thread kernel (scratch in global memory), HM_ADDER_SMEM=0
                    tk<<<blocks, 128, 0, ctx->stream>>>(a->d, b->d, o->d, n, a->L, make_layout(o), scratch);
                    rc = post_launch(ctx, "adder_thread_kernel");
                    break;
                }
                auto kern = wd == 4 ? hmk::adder_fused_kernel<4, 1, 0>
                                    : (mode == 0 ? hmk::adder_fused_kernel<8, 0, 0> : hmk::adder_fused_kernel<8, 1, 0>);
                CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                const unsigned grid = (unsigned)((n + warps - 1) / warps);
                kern<<<grid, warps * 32, smem, ctx->stream>>>(a->d, b->d, o->d, n, a->L, make_layout(o));
                rc = post_launch(ctx, "adder_fused_kernel");
            } else {
                rc = add_generic(ctx, a, b, o);
            }
            break;
        }
        case HM_OP_MUL: rc = mul_generic(ctx, a, b, o); break;
        default: rc = HM_ERR_INVALID_ARGUMENT;
    }
    return rc;
}

static int apply2_impl(hm_context *ctx, int op, const hm_batch *a, const hm_batch *b, hm_batch **out,
                       bool force_generic = false) {
    if (!ctx || !a || !b || !out) return HM_ERR_INVALID_ARGUMENT;
    if (a->n != b->n || a->L != b->L) return HM_ERR_INVALID_ARGUMENT;
    if (op == HM_OP_NOT) return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    std::vector<uint64_t> bounds;
    int rc = result_bounds(op, a->L, a->degb.data(), b->degb.data(), bounds);
    if (rc != HM_OK) return rc;
    int nb_status = HM_OK;
    hm_batch *o = new_batch(ctx, a->n, a->L, bounds.data(), &nb_status);
    if (!o) return nb_status;
    rc = alloc_batch(ctx, o);
    if (rc != HM_OK) {
        discard_batch(o);
        return rc;
    }
    rc = apply2_exec(ctx, op, a, b, o, force_generic);
    if (rc != HM_OK) {
        hm_batch_free(ctx, o);
        return rc;
    }
    *out = o;
    return HM_OK;
}

int hm_apply2_unchecked(hm_context *ctx, int op, const hm_batch *a, const hm_batch *b, hm_batch **out) {
    return apply2_impl(ctx, op, a, b, out);
}

int hm_apply2_into(hm_context *ctx, int op, const hm_batch *a, const hm_batch *b, hm_batch *out) {
    if (!ctx || !a || !b || !out) return HM_ERR_INVALID_ARGUMENT;
    int rc = validate_operation(ctx, op);
    if (rc != HM_OK) return rc;
    if (a->n != b->n || a->L != b->L || out->n != a->n || out->L != a->L || op == HM_OP_NOT) return HM_ERR_INVALID_ARGUMENT;
    std::vector<uint64_t> bounds;
    rc = result_bounds(op, a->L, a->degb.data(), b->degb.data(), bounds);
    if (rc != HM_OK) return rc;
    for (uint32_t k = 0; k < a->L; ++k)
        if (out->w[k] != (uint32_t)(bounds[k] / 64 + 1)) return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    out->degb = bounds;
    return apply2_exec(ctx, op, a, b, out, false);
}

int hm_apply2(hm_context *ctx, int op, const hm_batch *a, const hm_batch *b, hm_batch **out) {
    if (!ctx) return HM_ERR_INVALID_ARGUMENT;
    int rc = validate_operation(ctx, op);
    if (rc != HM_OK) return rc;
    return apply2_impl(ctx, op, a, b, out);
}

// ---- struct fields: Ciphered<T>::split_at / new_from_raw / extend_from_slice, examples/simple_struct.rs:32-58 ----------
int hm_batch_slice(hm_context *ctx, const hm_batch *src, uint32_t first_bit, uint32_t n_bits, hm_batch **out) {
    if (!ctx || !src || !out || n_bits == 0) return HM_ERR_INVALID_ARGUMENT;
    if ((uint64_t)first_bit + n_bits > src->L) return HM_ERR_INVALID_LENGTH; // split_at past the end panics in the reference
    USE_DEV(ctx);
    int nb_status = HM_OK;
    hm_batch *o = new_batch(ctx, src->n, n_bits, src->degb.data() + first_bit, &nb_status);
    if (!o) return nb_status;
    int rc = alloc_batch(ctx, o);
    if (rc != HM_OK) {
        discard_batch(o);
        return rc;
    }
    std::vector<MulOp> ops;
    for (uint32_t k = 0; k < n_bits; ++k) ops.push_back(MulOp{slot_view(src, first_bit + k), null_view(), slot_view(o, k)});
    rc = launch_xor_ops(ctx, ops, src->n);
    if (rc != HM_OK) {
        hm_batch_free(ctx, o);
        return rc;
    }
    *out = o;
    return HM_OK;
}

int hm_batch_concat(hm_context *ctx, const hm_batch *const *parts, size_t count, hm_batch **out) {
    if (!ctx || !parts || !out || count == 0) return HM_ERR_INVALID_ARGUMENT;
    std::vector<uint64_t> degb;
    for (size_t i = 0; i < count; ++i) {
        if (!parts[i] || parts[i]->n != parts[0]->n) return HM_ERR_INVALID_ARGUMENT;
        degb.insert(degb.end(), parts[i]->degb.begin(), parts[i]->degb.end());
    }
    if (degb.size() > (size_t)hmk::MAX_SLOTS) return HM_ERR_UNSUPPORTED;
    USE_DEV(ctx);
    int nb_status = HM_OK;
    hm_batch *o = new_batch(ctx, parts[0]->n, (uint32_t)degb.size(), degb.data(), &nb_status);
    if (!o) return nb_status;
    int rc = alloc_batch(ctx, o);
    if (rc != HM_OK) {
        discard_batch(o);
        return rc;
    }
    std::vector<MulOp> ops;
    uint32_t k = 0;
    for (size_t i = 0; i < count; ++i)
        for (uint32_t j = 0; j < parts[i]->L; ++j, ++k) ops.push_back(MulOp{slot_view(parts[i], j), null_view(), slot_view(o, k)});
    rc = launch_xor_ops(ctx, ops, o->n);
    if (rc != HM_OK) {
        hm_batch_free(ctx, o);
        return rc;
    }
    *out = o;
    return HM_OK;
}

int hm_apply2_fields(hm_context *ctx, int op, const hm_batch *a, const hm_batch *b, const uint32_t *field_bits, size_t n_fields,
                     hm_batch **out) {
    if (!ctx || !a || !b || !field_bits || !out || n_fields == 0) return HM_ERR_INVALID_ARGUMENT;
    if (a->n != b->n || a->L != b->L) return HM_ERR_INVALID_ARGUMENT;
    uint64_t total = 0;
    for (size_t f = 0; f < n_fields; ++f) {
        if (field_bits[f] == 0) return HM_ERR_INVALID_ARGUMENT;
        total += field_bits[f];
    }
    if (total != a->L) return HM_ERR_INVALID_LENGTH;
    int rc = validate_operation(ctx, op);
    if (rc != HM_OK) return rc;
    std::vector<hm_batch *> res(n_fields, nullptr);
    uint32_t first = 0;
    for (size_t f = 0; f < n_fields && rc == HM_OK; ++f) {
        hm_batch *fa = nullptr, *fb = nullptr;
        rc = hm_batch_slice(ctx, a, first, field_bits[f], &fa);
        if (rc == HM_OK) rc = hm_batch_slice(ctx, b, first, field_bits[f], &fb);
        if (rc == HM_OK) rc = apply2_impl(ctx, op, fa, fb, &res[f]);
        if (fa) hm_batch_free(ctx, fa);
        if (fb) hm_batch_free(ctx, fb);
        first += field_bits[f];
    }
    if (rc == HM_OK) rc = hm_batch_concat(ctx, res.data(), n_fields, out);
    for (hm_batch *r : res)
        if (r) hm_batch_free(ctx, r);
    return rc;
}

// forces the generic (view based, reference order) circuit even when the fused kernel applies; used to
// cross-check the fused adder on the GPU.
int hm_apply2_generic(hm_context *ctx, int op, const hm_batch *a, const hm_batch *b, hm_batch **out) {
    return apply2_impl(ctx, op, a, b, out, true);
}

int hm_apply1(hm_context *ctx, int op, hm_batch *a) {
    if (!ctx || !a) return HM_ERR_INVALID_ARGUMENT;
    if (op != HM_OP_NOT) return HM_ERR_INVALID_ARGUMENT;
    int rc = validate_operation(ctx, op);
    if (rc != HM_OK) return rc;
    USE_DEV(ctx);
    if (a->n == 0) return HM_OK;
    const int grid = grid_for(ctx, (uint64_t)a->n * a->L, 256, 16);
    hmk::not_kernel<<<grid, 256, 0, ctx->stream>>>(a->d, make_layout(a), a->n);
    LAUNCHED("not_kernel");
    return HM_OK;
}

// Host-buffer pipeline shared by hm_apply2_host / hm_apply2_host_bounded: chunks of values, upload of chunk c+1 and
// download of chunk c-1 overlap the kernels of chunk c on three streams; two stages of device buffers, allocated once.
namespace {
struct HostPipe { // owns the streams, events and stage batches; released on every exit path
    hm_context *ctx;
    cudaStream_t s_up = nullptr, s_down = nullptr;
    struct Stage {
        hm_batch *a = nullptr, *b = nullptr, *o = nullptr;
        cudaEvent_t up = nullptr, done = nullptr, down = nullptr;
        bool used = false;
    } st[2];
    explicit HostPipe(hm_context *c) : ctx(c) {}
    ~HostPipe() {
        if (s_up) cudaStreamSynchronize(s_up);
        cudaStreamSynchronize(ctx->stream);
        if (s_down) cudaStreamSynchronize(s_down);
        for (Stage &s : st) {
            if (s.a) hm_batch_free(ctx, s.a);
            if (s.b) hm_batch_free(ctx, s.b);
            if (s.o) hm_batch_free(ctx, s.o);
            if (s.up) cudaEventDestroy(s.up);
            if (s.done) cudaEventDestroy(s.done);
            if (s.down) cudaEventDestroy(s.down);
        }
        if (s_up) cudaStreamDestroy(s_up);
        if (s_down) cudaStreamDestroy(s_down);
    }
};
} // namespace

static int apply2_host_impl(hm_context *ctx, int op, size_t n, uint32_t L, const uint64_t *da, const uint64_t *a_host, const uint64_t *db,
                            const uint64_t *b_host, uint64_t *out_host) {
    std::vector<uint64_t> bo;
    int rc = result_bounds(op, L, da, db, bo);
    if (rc != HM_OK) return rc;
    const uint64_t vwa = checked_value_words(L, da), vwb = checked_value_words(L, db), vwo = checked_value_words(L, bo.data());
    if (!vwa || !vwb || !vwo) return HM_ERR_INVALID_ARGUMENT;
    if (n == 0) return HM_OK;
    const size_t per_value_bytes = (size_t)(vwa + vwb + vwo) * 8;
    const size_t chunk_bytes = (size_t)(g_host_chunk_mb > 0 ? g_host_chunk_mb : 96) << 20;
    size_t chunk = std::max<size_t>(1, std::min<size_t>(n, chunk_bytes / per_value_bytes));
    if (chunk > 1024) chunk &= ~(size_t)1023;
    cudaStream_t user = ctx->stream;
    HostPipe hp(ctx);
    CK(cudaStreamCreateWithFlags(&hp.s_up, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&hp.s_down, cudaStreamNonBlocking));
    const int nstages = (n > chunk) ? 2 : 1;
    for (int i = 0; i < nstages; ++i) {
        HostPipe::Stage &s = hp.st[i];
        CK(cudaEventCreateWithFlags(&s.up, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s.down, cudaEventDisableTiming));
        int st = HM_OK;
        if (!(s.a = new_batch(ctx, chunk, L, da, &st))) return st;
        if (!(s.b = new_batch(ctx, chunk, L, db, &st))) return st;
        if (!(s.o = new_batch(ctx, chunk, L, bo.data(), &st))) return st;
        if ((rc = alloc_batch(ctx, s.a)) != HM_OK || (rc = alloc_batch(ctx, s.b)) != HM_OK || (rc = alloc_batch(ctx, s.o)) != HM_OK) return rc;
    }
    {   // the stage buffers were allocated in `user` stream order: the copy streams must not touch them earlier
        cudaEvent_t ready = nullptr;
        CK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
        cudaError_t e = cudaEventRecord(ready, user);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(hp.s_up, ready, 0);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(hp.s_down, ready, 0);
        cudaEventDestroy(ready);
        CK(e);
    }
    size_t idx = 0;
    for (size_t first = 0; first < n; first += chunk, ++idx) {
        HostPipe::Stage &s = hp.st[idx % nstages];
        const size_t cnt = std::min(chunk, n - first);
        if (s.used) {
            CK(cudaStreamWaitEvent(hp.s_up, s.done, 0)); // inputs of this stage were consumed
            CK(cudaStreamWaitEvent(user, s.down, 0));    // its previous result has left the device
        }
        s.a->n = s.b->n = s.o->n = cnt;
        CK(cudaMemcpyAsync(s.a->d, a_host + first * vwa, cnt * vwa * 8, cudaMemcpyHostToDevice, hp.s_up));
        CK(cudaMemcpyAsync(s.b->d, b_host + first * vwb, cnt * vwb * 8, cudaMemcpyHostToDevice, hp.s_up));
        CK(cudaEventRecord(s.up, hp.s_up));
        CK(cudaStreamWaitEvent(user, s.up, 0));
        rc = apply2_exec(ctx, op, s.a, s.b, s.o, false);
        if (rc != HM_OK) return rc;
        CK(cudaEventRecord(s.done, user));
        CK(cudaStreamWaitEvent(hp.s_down, s.done, 0));
        CK(cudaMemcpyAsync(out_host + first * vwo, s.o->d, cnt * vwo * 8, cudaMemcpyDeviceToHost, hp.s_down));
        CK(cudaEventRecord(s.down, hp.s_down));
        s.used = true;
    }
    CK(cudaStreamSynchronize(hp.s_up));
    CK(cudaStreamSynchronize(user));
    CK(cudaStreamSynchronize(hp.s_down));
    return HM_OK;
}

int hm_apply2_host_bounded(hm_context *ctx, int op, size_t n, uint32_t L, const uint64_t *a_bounds, const uint64_t *a_host,
                           const uint64_t *b_bounds, const uint64_t *b_host, uint64_t *out_host) {
    if (!ctx || !a_bounds || !b_bounds || (!a_host && n) || (!b_host && n) || (!out_host && n)) return HM_ERR_INVALID_ARGUMENT;
    int rc = validate_operation(ctx, op);
    if (rc != HM_OK) return rc;
    if (L == 0 || L > (uint32_t)hmk::MAX_SLOTS || op == HM_OP_NOT) return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    return apply2_host_impl(ctx, op, n, L, a_bounds, a_host, b_bounds, b_host, out_host);
}

int hm_apply2_host(hm_context *ctx, int op, size_t n, uint32_t L, const uint32_t *a_words, const uint64_t *a_host,
                   const uint32_t *b_words, const uint64_t *b_host, uint64_t *out_host) {
    if (!ctx || !a_words || !b_words || (!a_host && n) || (!b_host && n) || (!out_host && n)) return HM_ERR_INVALID_ARGUMENT;
    int rc = validate_operation(ctx, op);
    if (rc != HM_OK) return rc;
    if (L == 0 || L > (uint32_t)hmk::MAX_SLOTS || op == HM_OP_NOT) return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    // widths alone say nothing about the degrees inside: every slot is taken at its widest (degree <= 64 w - 1)
    std::vector<uint64_t> da(L), db(L);
    for (uint32_t k = 0; k < L; ++k) {
        if (!a_words[k] || !b_words[k]) return HM_ERR_INVALID_ARGUMENT;
        da[k] = (uint64_t)a_words[k] * 64 - 1;
        db[k] = (uint64_t)b_words[k] * 64 - 1;
    }
    return apply2_host_impl(ctx, op, n, L, da.data(), a_host, db.data(), b_host, out_host);
}

// ---- raw polynomial batches ----------------------------------------------------------------------
int hm_poly_add(hm_context *ctx, const hm_batch *a, const hm_batch *b, hm_batch **out) {
    return apply2_impl(ctx, HM_OP_XOR, a, b, out);
}
int hm_poly_mul(hm_context *ctx, const hm_batch *a, const hm_batch *b, hm_batch **out) {
    return apply2_impl(ctx, HM_OP_AND, a, b, out);
}

int hm_poly_rem(hm_context *ctx, const hm_batch *a, hm_batch **out) {
    if (!ctx || !a || !out) return HM_ERR_INVALID_ARGUMENT;
    if (!ctx->has_sk) return HM_ERR_SECRET_KEY_UNSET;
    USE_DEV(ctx);
    std::vector<uint64_t> bounds(a->L, ctx->ds - 1);
    int nb_status = HM_OK;
    hm_batch *o = new_batch(ctx, a->n, a->L, bounds.data(), &nb_status);
    if (!o) return nb_status;
    int rc = alloc_batch(ctx, o);
    for (uint32_t k = 0; k < a->L && rc == HM_OK; ++k) rc = launch_rem(ctx, slot_view(a, k), slot_view(o, k), a->n);
    if (rc != HM_OK) {
        hm_batch_free(ctx, o);
        return rc;
    }
    *out = o;
    return HM_OK;
}

static bool mulrem_fresh_ok(const hm_context *ctx, const hm_batch *a, const hm_batch *b) {
    // fused fast path: both operands fresh-shaped with (D, d) = (256, 128) [config A] or (1024, 512) [config B]
    const bool cfg_a = ctx->fresh_deg == 256 && ctx->ds == 128, cfg_b = ctx->fresh_deg == 1024 && ctx->ds == 512;
    bool fresh = ctx->has_pk && (cfg_a || cfg_b) && same_layout(a, b);
    for (uint32_t k = 0; k < a->L && fresh; ++k) fresh = a->degb[k] == ctx->fresh_deg && b->degb[k] == ctx->fresh_deg;
    return fresh;
}

static int mulrem_fresh_exec(hm_context *ctx, const hm_batch *a, const hm_batch *b, hm_batch *o) {
    const uint64_t pairs = (uint64_t)a->n * a->L;
    if (!pairs) return HM_OK;
    if (ctx->fresh_deg == 1024) {
        // 1 = rolled 32-word product then fold, 2 = operands reduced first (16-word product), 3 = reduce-first with the
        // conflict-free rotated fold table
        static const int mode_b = getenv("HM_MULREM_B_MODE") ? atoi(getenv("HM_MULREM_B_MODE")) : 3;
        CK(hmk::launch_mulrem_fresh_b32(a->d, b->d, o->d, pairs, ctx->d_remT, ctx->sm_count, ctx->stream, mode_b < 1 ? 0 : mode_b - 1));
        ctx->launches++;
        return HM_OK;
    }
    constexpr int WD = 8, WS = 4;
    static const int mode = getenv("HM_MULREM_MODE") ? atoi(getenv("HM_MULREM_MODE")) : 4;
    if (mode == 0) { // shift/mask schoolbook on the ALU pipe
        constexpr int TH = 128;
        const size_t smem = (size_t)4 * 256 * WS * 4 + (size_t)2 * 2 * TH * (WD / 2 + 1) * 8;
        auto kern = hmk::mulrem_fresh_kernel<WD, WS, 0, TH, 1>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid_for(ctx, pairs, TH, 6), TH, smem, ctx->stream>>>(a->d, b->d, o->d, pairs, ctx->d_remT);
    } else if (mode == 1) { // Karatsuba on the multiplier, small CTAs, plain tables
        constexpr int TH = 128;
        const size_t smem = (size_t)4 * 256 * WS * 4 + (size_t)2 * 2 * TH * (WD / 2 + 1) * 8;
        auto kern = hmk::mulrem_fresh_kernel<WD, WS, 2, TH, 1>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid_for(ctx, pairs, TH, 4), TH, smem, ctx->stream>>>(a->d, b->d, o->d, pairs, ctx->d_remT);
    } else if (mode == 4 || mode == 5) { // reduce-first with the rotated conflict-free fold table (32 KB): 1024- or 512-thread CTAs
        const int TH = mode == 4 ? 1024 : 512;
        const size_t smem = (size_t)256 * 4 * 2 * 4 * 4 + (size_t)2 * 2 * TH * (WD / 2 + 1) * 8;
        if (mode == 4) {
            CK(cudaFuncSetAttribute(hmk::mulrem_fresh_a_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            hmk::mulrem_fresh_a_kernel<1024><<<grid_for(ctx, pairs, TH, 1), TH, smem, ctx->stream>>>(a->d, b->d, o->d, pairs, ctx->d_remT);
        } else {
            CK(cudaFuncSetAttribute(hmk::mulrem_fresh_a_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            hmk::mulrem_fresh_a_kernel<512><<<grid_for(ctx, pairs, TH, 1), TH, smem, ctx->stream>>>(a->d, b->d, o->d, pairs, ctx->d_remT);
        }
    } else if (mode == 3) { // operands reduced mod S first, then a WS-word product (same remainder, a third of the leaf products)
        constexpr int TH = 512, REP = 8;
        const size_t smem = (size_t)4 * 256 * REP * WS * 4 + (size_t)2 * 2 * TH * (WD / 2 + 1) * 8;
        auto kern = hmk::mulrem_fresh_kernel<WD, WS, 3, TH, REP>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid_for(ctx, pairs, TH, 1), TH, smem, ctx->stream>>>(a->d, b->d, o->d, pairs, ctx->d_remT);
    } else { // Karatsuba on the multiplier, one 512-thread CTA per SM, 8-way replicated (conflict-free) fold tables
        constexpr int TH = 512, REP = 8;
        const size_t smem = (size_t)4 * 256 * REP * WS * 4 + (size_t)2 * 2 * TH * (WD / 2 + 1) * 8;
        auto kern = hmk::mulrem_fresh_kernel<WD, WS, 2, TH, REP>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid_for(ctx, pairs, TH, 1), TH, smem, ctx->stream>>>(a->d, b->d, o->d, pairs, ctx->d_remT);
    }
    return post_launch(ctx, "mulrem_fresh_kernel");
}

int hm_poly_mulrem(hm_context *ctx, const hm_batch *a, const hm_batch *b, hm_batch **out) {
    if (!ctx || !a || !b || !out) return HM_ERR_INVALID_ARGUMENT;
    if (!ctx->has_sk) return HM_ERR_SECRET_KEY_UNSET;
    if (a->n != b->n || a->L != b->L) return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    if (mulrem_fresh_ok(ctx, a, b)) {
        std::vector<uint64_t> bounds(a->L, ctx->ds - 1);
        int nb_status = HM_OK;
        hm_batch *o = new_batch(ctx, a->n, a->L, bounds.data(), &nb_status);
        if (!o) return nb_status;
        int rc = alloc_batch(ctx, o);
        if (rc == HM_OK) rc = mulrem_fresh_exec(ctx, a, b, o);
        if (rc != HM_OK) {
            hm_batch_free(ctx, o);
            return rc;
        }
        *out = o;
        return HM_OK;
    }
    // any other shapes: rem(mul(a, b)) = rem(mul(rem(a), rem(b))) — the remainder is unique — so the product is never wider
    // than 2 d bits however long the operands are (a u32 sum times another is 23 552 x 23 552 bits otherwise)
    hm_batch *ra = nullptr, *rb = nullptr, *prod = nullptr;
    int rc = hm_poly_rem(ctx, a, &ra);
    if (rc == HM_OK) rc = hm_poly_rem(ctx, b, &rb);
    if (rc == HM_OK) rc = hm_poly_mul(ctx, ra, rb, &prod);
    if (rc == HM_OK) rc = hm_poly_rem(ctx, prod, out);
    if (ra) hm_batch_free(ctx, ra);
    if (rb) hm_batch_free(ctx, rb);
    if (prod) hm_batch_free(ctx, prod);
    return rc;
}

int hm_poly_mulrem_into(hm_context *ctx, const hm_batch *a, const hm_batch *b, hm_batch *out) {
    if (!ctx || !a || !b || !out) return HM_ERR_INVALID_ARGUMENT;
    if (!ctx->has_sk) return HM_ERR_SECRET_KEY_UNSET;
    if (a->n != b->n || a->L != b->L || out->n != a->n || out->L != a->L) return HM_ERR_INVALID_ARGUMENT;
    for (uint32_t k = 0; k < a->L; ++k)
        if (out->w[k] != (uint32_t)((ctx->ds - 1) / 64 + 1)) return HM_ERR_INVALID_ARGUMENT;
    if (!mulrem_fresh_ok(ctx, a, b)) return HM_ERR_UNSUPPORTED; // only the fused kernel writes in place
    USE_DEV(ctx);
    return mulrem_fresh_exec(ctx, a, b, out);
}

// ---- measured integer-logic peak ------------------------------------------------------------------
int hm_measure_alu_peak(hm_context *ctx, double *lop3_lane_ops_per_s, double *sm_clock_mhz) {
    if (!ctx || !lop3_lane_ops_per_s) return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    const int blocks = ctx->sm_count * 8, threads = 256;
    uint32_t *sink = nullptr;
    unsigned long long *clk = nullptr;
    CK(cudaMalloc(&sink, (size_t)blocks * threads * 4));
    CK(cudaMalloc(&clk, (size_t)blocks * 8));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 1 << 16;
    hmk::lop3_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(sink, clk, iters, 0x9e3779b9u);
    double best = 0, mhz = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0, ctx->stream);
        hmk::lop3_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(sink, clk, iters, 0x9e3779b9u + rep);
        cudaEventRecord(e1, ctx->stream);
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        std::vector<unsigned long long> h(blocks);
        CK(cudaMemcpy(h.data(), clk, (size_t)blocks * 8, cudaMemcpyDeviceToHost));
        double avg = 0;
        for (int i = 0; i < blocks; ++i) avg += (double)h[i];
        avg /= blocks;
        const double ops = (double)blocks * threads * iters * hmk::PEAK_ILP;
        const double rate = ops / (ms * 1e-3);
        if (rate > best) {
            best = rate;
            mhz = avg / (ms * 1e3);
        }
    }
    ctx->launches += 4;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    cudaFree(clk);
    *lop3_lane_ops_per_s = best;
    if (sm_clock_mhz) *sm_clock_mhz = mhz;
    return HM_OK;
}

int hm_measure_kara8_peak(hm_context *ctx, double *products_per_s) {
    if (!ctx || !products_per_s) return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    const int blocks = ctx->sm_count * 4, threads = 128, iters = 1500;
    uint32_t *sink = nullptr;
    CK(cudaMalloc(&sink, (size_t)blocks * threads * 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    hmk::kara8_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(sink, iters, 0x9e3779b9u);
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0, ctx->stream);
        hmk::kara8_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(sink, iters, 0x9e3779b9u + rep);
        cudaEventRecord(e1, ctx->stream);
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        best = std::max(best, (double)blocks * threads * iters / (ms * 1e-3));
    }
    ctx->launches += 4;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *products_per_s = best;
    return HM_OK;
}

int hm_measure_pipe_peaks(hm_context *ctx, double min_ms, double *out12) {
    if (!ctx || !out12 || !(min_ms > 0)) return HM_ERR_INVALID_ARGUMENT;
    USE_DEV(ctx);
    hmk::PipeProbe p[3];
    CK(cudaStreamSynchronize(ctx->stream));
    CK(hmk::measure_pipe_peaks(ctx->sm_count, ctx->stream, min_ms, p));
    for (int i = 0; i < 3; ++i) {
        out12[4 * i + 0] = p[i].warp_instr_per_s;
        out12[4 * i + 1] = p[i].ms;
        out12[4 * i + 2] = (double)p[i].warps_per_sm;
        out12[4 * i + 3] = p[i].cycle_counter_mhz;
    }
    ctx->launches += 12;
    return HM_OK;
}

// ---- host-side helpers ---------------------------------------------------------------------------
uint32_t hm_fresh_slot_words(const hm_context *ctx) {
    if (!ctx) return 0;
    if (ctx->has_pk) return ctx->wf;
    return (uint32_t)(((size_t)ctx->d + ctx->dp) / 64 + 1);
}

int hm_decrypt_vector(const hm_context *ctx, size_t nbits, uint64_t *v_out) {
    if (!ctx || !v_out) return HM_ERR_INVALID_ARGUMENT;
    if (!ctx->has_sk) return HM_ERR_SECRET_KEY_UNSET;
    gf2::words v = gf2::decrypt_vector(ctx->S, ctx->ds, nbits);
    for (size_t i = 0; i < (nbits + 63) / 64; ++i) v_out[i] = v[i];
    if (nbits % 64) v_out[nbits / 64] &= (((uint64_t)1 << (nbits % 64)) - 1);
    return HM_OK;
}

size_t hm_poly_degree(const uint64_t *words, size_t n_words) {
    if (!words || n_words == 0) return 0;
    return gf2::degree(words, n_words);
}

} // extern "C"
