// hmgroup.cu — one logical batch over several GPUs of a box, driven from one host thread (include/hmgpu.h, "Device groups").
//
// Ciphertexts are independent (reference src/cipher.rs:180-185, :227-237; the circuits of common.rs touch only their two
// operands), so a batch of n values is cut into contiguous index ranges, range r on device r, with the keys and their derived
// tables replicated: there is NO exchange step and no collective.  Every per-device call only enqueues work on that device's
// stream; the group calls fan the work out over all devices first and synchronise afterwards, so the devices run concurrently.
// Everything here is written against the public C ABI of one context — a Rust or C++ host could do the same by hand.
#include <cuda_runtime.h>

#include <new>
#include <vector>

#include "../../include/hmgpu.h"

struct hm_group {
    std::vector<hm_context *> ctx;
    uint16_t tau = 0;
};

struct hm_group_batch {
    hm_group *g = nullptr;
    size_t n = 0;
    uint32_t L = 0;
    std::vector<hm_batch *> part; // part[r] holds values [first(r), first(r) + count(r))
};

extern "C" {

// Contiguous split of [0, n) over `world` ranks: sizes differ by at most one, rank order == index order.
int hm_shard_range(size_t n, int rank, int world, size_t *first, size_t *count) {
    if (world <= 0 || rank < 0 || rank >= world || !first || !count) return HM_ERR_INVALID_ARGUMENT;
    const size_t base = n / (size_t)world, extra = n % (size_t)world, r = (size_t)rank;
    *first = r * base + (r < extra ? r : extra);
    *count = base + (r < extra ? 1 : 0);
    return HM_OK;
}

int hm_group_create(uint16_t d, uint16_t dp, uint16_t delta, uint16_t tau, const int *device_ids, int n_dev, hm_group **out) {
    if (!out || !device_ids || n_dev <= 0) return HM_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    hm_group *g = new (std::nothrow) hm_group;
    if (!g) return HM_ERR_OUT_OF_MEMORY;
    g->tau = tau;
    for (int i = 0; i < n_dev; ++i) {
        hm_context *c = nullptr;
        const int rc = hm_context_create(d, dp, delta, tau, device_ids[i], &c);
        if (rc != HM_OK) {
            for (hm_context *p : g->ctx) hm_context_destroy(p);
            delete g;
            return rc;
        }
        g->ctx.push_back(c);
    }
    *out = g;
    return HM_OK;
}

void hm_group_destroy(hm_group *g) {
    if (!g) return;
    for (hm_context *c : g->ctx) hm_context_destroy(c);
    delete g;
}

int hm_group_size(const hm_group *g) { return g ? (int)g->ctx.size() : 0; }
hm_context *hm_group_context(const hm_group *g, int i) { return (g && i >= 0 && i < (int)g->ctx.size()) ? g->ctx[i] : nullptr; }

int hm_group_set_secret_key(hm_group *g, const uint8_t *bytes, size_t len) {
    if (!g) return HM_ERR_INVALID_ARGUMENT;
    for (hm_context *c : g->ctx) {
        const int rc = hm_set_secret_key(c, bytes, len);
        if (rc != HM_OK) return rc;
    }
    return HM_OK;
}

int hm_group_set_public_key(hm_group *g, const uint8_t *const *polys, const size_t *lens, size_t n_polys) {
    if (!g) return HM_ERR_INVALID_ARGUMENT;
    for (hm_context *c : g->ctx) {
        const int rc = hm_set_public_key(c, polys, lens, n_polys);
        if (rc != HM_OK) return rc;
    }
    return HM_OK;
}

int hm_group_synchronize(hm_group *g) {
    if (!g) return HM_ERR_INVALID_ARGUMENT;
    int rc = HM_OK;
    for (hm_context *c : g->ctx) {
        const int r = hm_context_synchronize(c);
        if (rc == HM_OK) rc = r;
    }
    return rc;
}

size_t hm_group_batch_len(const hm_group_batch *b) { return b ? b->n : 0; }
uint32_t hm_group_batch_bits(const hm_group_batch *b) { return b ? b->L : 0; }
hm_batch *hm_group_batch_part(const hm_group_batch *b, int i) { return (b && i >= 0 && i < (int)b->part.size()) ? b->part[i] : nullptr; }

void hm_group_batch_free(hm_group_batch *b) {
    if (!b) return;
    for (hm_batch *p : b->part)
        if (p) hm_batch_free(nullptr, p);
    delete b;
}

static hm_group_batch *new_group_batch(hm_group *g, size_t n, uint32_t L) {
    hm_group_batch *b = new (std::nothrow) hm_group_batch;
    if (!b) return nullptr;
    b->g = g;
    b->n = n;
    b->L = L;
    b->part.assign(g->ctx.size(), nullptr);
    return b;
}

// masks == NULL: device-side Philox masks from `seed`; shard r continues the stream at its first bit-ciphertext, so the group
// produces exactly the ciphertexts one device would for the whole batch.
static int group_encrypt(hm_group *g, const uint8_t *values, size_t n, uint32_t L, const uint8_t *masks, uint64_t seed, hm_group_batch **out) {
    if (!g || !out || (!values && n)) return HM_ERR_INVALID_ARGUMENT;
    if (L == 0 || L % 8 != 0) return HM_ERR_INVALID_ARGUMENT;
    hm_group_batch *b = new_group_batch(g, n, L);
    if (!b) return HM_ERR_OUT_OF_MEMORY;
    const int world = (int)g->ctx.size();
    const size_t vb = L / 8, mb = (size_t)L * ((g->tau + 7u) / 8u);
    int rc = HM_OK;
    for (int r = 0; r < world && rc == HM_OK; ++r) { // enqueue on every device ...
        size_t first = 0, cnt = 0;
        hm_shard_range(n, r, world, &first, &cnt);
        if (masks) rc = hm_encrypt_async(g->ctx[r], values + first * vb, cnt, L, masks + first * mb, &b->part[r]);
        else rc = hm_encrypt_seeded_at(g->ctx[r], values + first * vb, cnt, L, seed, (uint64_t)first * L, 0, &b->part[r]);
    }
    const int src = hm_group_synchronize(g); // ... then wait: the host buffers may be reused after return
    if (rc == HM_OK) rc = src;
    if (rc != HM_OK) {
        hm_group_batch_free(b);
        return rc;
    }
    *out = b;
    return HM_OK;
}

int hm_group_encrypt(hm_group *g, const uint8_t *values, size_t n, uint32_t L, const uint8_t *masks, hm_group_batch **out) {
    if (!masks && n) return HM_ERR_INVALID_ARGUMENT;
    return group_encrypt(g, values, n, L, masks ? masks : reinterpret_cast<const uint8_t *>(""), 0, out);
}
int hm_group_encrypt_seeded(hm_group *g, const uint8_t *values, size_t n, uint32_t L, uint64_t seed, hm_group_batch **out) {
    return group_encrypt(g, values, n, L, nullptr, seed, out);
}

int hm_group_apply2(hm_group *g, int op, const hm_group_batch *a, const hm_group_batch *b, hm_group_batch **out) {
    if (!g || !a || !b || !out || a->g != g || b->g != g) return HM_ERR_INVALID_ARGUMENT;
    if (a->n != b->n || a->L != b->L) return HM_ERR_INVALID_ARGUMENT;
    hm_group_batch *o = new_group_batch(g, a->n, a->L);
    if (!o) return HM_ERR_OUT_OF_MEMORY;
    int rc = HM_OK;
    for (size_t r = 0; r < g->ctx.size() && rc == HM_OK; ++r) rc = hm_apply2(g->ctx[r], op, a->part[r], b->part[r], &o->part[r]); // asynchronous
    if (rc != HM_OK) {
        hm_group_synchronize(g);
        hm_group_batch_free(o);
        return rc;
    }
    *out = o;
    return HM_OK;
}

int hm_group_apply2_into(hm_group *g, int op, const hm_group_batch *a, const hm_group_batch *b, hm_group_batch *out) {
    if (!g || !a || !b || !out || a->g != g || b->g != g || out->g != g) return HM_ERR_INVALID_ARGUMENT;
    if (a->n != b->n || a->L != b->L || out->n != a->n || out->L != a->L) return HM_ERR_INVALID_ARGUMENT;
    int rc = HM_OK;
    for (size_t r = 0; r < g->ctx.size() && rc == HM_OK; ++r) rc = hm_apply2_into(g->ctx[r], op, a->part[r], b->part[r], out->part[r]);
    return rc;
}

int hm_group_apply1(hm_group *g, int op, hm_group_batch *a) {
    if (!g || !a || a->g != g) return HM_ERR_INVALID_ARGUMENT;
    int rc = HM_OK;
    for (size_t r = 0; r < g->ctx.size() && rc == HM_OK; ++r) rc = hm_apply1(g->ctx[r], op, a->part[r]);
    return rc;
}

int hm_group_decrypt(hm_group *g, const hm_group_batch *b, uint8_t *values_out) {
    if (!g || !b || b->g != g || (!values_out && b->n)) return HM_ERR_INVALID_ARGUMENT;
    const int world = (int)g->ctx.size();
    const size_t vb = b->L / 8;
    int rc = HM_OK;
    for (int r = 0; r < world && rc == HM_OK; ++r) { // kernels + D2H copies enqueued on every device, plaintexts land in index order
        size_t first = 0, cnt = 0;
        hm_shard_range(b->n, r, world, &first, &cnt);
        rc = hm_decrypt_async(g->ctx[r], b->part[r], values_out + first * vb);
    }
    const int src = hm_group_synchronize(g);
    return rc == HM_OK ? src : rc;
}

// Gathers the shards into one host buffer in the padded layout (n * value_words u64 words), index order.
int hm_group_batch_download(hm_group *g, const hm_group_batch *b, uint64_t *host) {
    if (!g || !b || b->g != g || (!host && b->n)) return HM_ERR_INVALID_ARGUMENT;
    const int world = (int)g->ctx.size();
    int rc = HM_OK;
    for (int r = 0; r < world && rc == HM_OK; ++r) {
        size_t first = 0, cnt = 0;
        hm_shard_range(b->n, r, world, &first, &cnt);
        rc = hm_batch_download(g->ctx[r], b->part[r], host + first * hm_batch_value_words(b->part[r]));
    }
    return rc;
}

} // extern "C"
