/*
 * hm_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE ONLY; see hm_oracle.h).
 *
 * Restates, function by function, the algorithms of mathisbot/homomorph-rust:
 *   src/polynomial.rs, src/cipher.rs, src/context.rs (keygen only),
 *   src/impls/numbers/common.rs.
 * The restatement keeps the reference's cost profile on purpose (bit-serial
 * multiply, long-division remainder, one heap allocation per operation,
 * degree rescans) so that it can double as the same-box CPU baseline.
 */
#include "hm_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#define WBITS 64u /* BITS_PER_COEFF — src/polynomial.rs:9 (usize on a 64-bit target) */

typedef struct {
    uint64_t *c;   /* coefficients, LSB-first: X^i is bit i%64 of word i/64 (src/polynomial.rs:144,172) */
    size_t len;    /* allocated words (>= 1) */
    size_t degree; /* tracked degree; the null polynomial has degree 0 (src/polynomial.rs:132-137) */
} poly;

struct orc_vec {
    poly *p;
    size_t n;
};

/* ------------------------------------------------------------------ helpers */

static uint64_t *walloc(size_t n) {
    uint64_t *w = (uint64_t *)calloc(n ? n : 1, sizeof(uint64_t));
    if (!w) abort();
    return w;
}

static inline unsigned lz64(uint64_t x) { return x ? (unsigned)__builtin_clzll(x) : 64u; }
static inline unsigned tz64(uint64_t x) { return x ? (unsigned)__builtin_ctzll(x) : 64u; }

static void poly_free(poly *p) {
    free(p->c);
    p->c = NULL;
    p->len = 0;
    p->degree = 0;
}

/* src/polynomial.rs:35-42 — index of the highest set bit, 0 when all words are zero */
static size_t compute_degree(const uint64_t *c, size_t len) {
    for (size_t i = len; i-- > 0;) {
        if (c[i] != 0) return WBITS - 1 - lz64(c[i]) + WBITS * i;
    }
    return 0;
}

/* src/polynomial.rs:53-63 — takes ownership of `words` */
static poly poly_new_owned(uint64_t *words, size_t len) {
    poly r = {words, len, compute_degree(words, len)};
    return r;
}

/* src/polynomial.rs:132-137 */
static poly poly_null(void) {
    poly r = {walloc(1), 1, 0};
    return r;
}

/* src/polynomial.rs:142-150 */
static poly poly_monomial(size_t degree) {
    size_t len = degree / WBITS + 1;
    poly r = {walloc(len), len, degree};
    r.c[degree / WBITS] = (uint64_t)1 << (degree % WBITS);
    return r;
}

/* src/polynomial.rs:73-96 — getrandom::fill replaced by caller bytes (native = little endian words) */
static poly poly_random(size_t degree, const uint8_t *rnd) {
    size_t n = degree / WBITS + 1;
    poly r = {walloc(n), n, degree};
    for (size_t i = 0; i < n; i++) {
        uint64_t w = 0;
        for (unsigned b = 0; b < 8; b++) w |= (uint64_t)rnd[8 * i + b] << (8 * b);
        r.c[i] = w;
    }
    r.c[n - 1] &= ((uint64_t)1 << (degree % WBITS)) - 1;
    r.c[n - 1] |= (uint64_t)1 << (degree % WBITS);
    return r;
}

/* src/polynomial.rs:108-122 — little-endian words, last chunk zero-padded */
static poly poly_from_bytes(const uint8_t *bytes, size_t nbytes) {
    size_t n = (nbytes + 7) / 8;
    uint64_t *w = walloc(n);
    for (size_t i = 0; i < nbytes; i++) w[i / 8] |= (uint64_t)bytes[i] << (8 * (i % 8));
    return poly_new_owned(w, n);
}

/* src/polynomial.rs:404-414 — keeps only degree/64+1 words */
static poly poly_clone(const poly *a) {
    size_t n = a->degree / WBITS + 1;
    poly r = {walloc(n), n, a->degree};
    memcpy(r.c, a->c, n * sizeof(uint64_t));
    return r;
}

/* src/polynomial.rs:416-426 */
static int poly_eq(const poly *a, const poly *b) {
    if (a->degree != b->degree) return 0;
    size_t n = a->degree / WBITS + 1;
    return memcmp(a->c, b->c, n * sizeof(uint64_t)) == 0;
}

/* src/polynomial.rs:168-181 */
static int poly_evaluate(const poly *a, int x) {
    if (!x) return (int)(a->c[0] & 1);
    unsigned ones = 0;
    for (size_t i = 0; i < a->len; i++) ones += (unsigned)__builtin_popcountll(a->c[i]);
    return (int)(ones % 2);
}

/* src/polynomial.rs:190-213 */
static poly poly_add(const poly *a, const poly *b) {
    size_t max_deg = a->degree > b->degree ? a->degree : b->degree;
    size_t n = max_deg / WBITS + 1;
    uint64_t *r = walloc(n);
    for (size_t i = 0; i < n; i++) {
        uint64_t x = i < a->len ? a->c[i] : 0;
        uint64_t y = i < b->len ? b->c[i] : 0;
        r[i] = x ^ y;
    }
    if (a->degree == b->degree) return poly_new_owned(r, n);
    poly out = {r, n, max_deg};
    return out;
}

/* src/polynomial.rs:216-235 */
static void poly_add_assign(poly *a, const poly *b) {
    size_t lhs_len = a->degree / WBITS + 1;
    size_t rhs_len = b->degree / WBITS + 1;
    if (rhs_len > lhs_len) {
        uint64_t *c = walloc(rhs_len);
        memcpy(c, a->c, lhs_len * sizeof(uint64_t));
        free(a->c);
        a->c = c;
        a->len = rhs_len;
    }
    size_t n = a->len < b->len ? a->len : b->len; /* zip stops at the shorter buffer */
    for (size_t i = 0; i < n; i++) a->c[i] ^= b->c[i];
    a->degree = compute_degree(a->c, a->len); /* full rescan, as in the reference (:234) */
}

/* src/polynomial.rs:238-243 */
static void poly_add_bool_assign(poly *a, int x) {
    if (x) {
        a->c[0] ^= 1;
        a->degree = compute_degree(a->c, a->len);
    }
}

/* src/polynomial.rs:252-310 — bit-serial schoolbook carry-less product */
static poly poly_mul(const poly *a, const poly *b) {
    if ((a->degree == 0 && (a->c[0] & 1) == 0) || (b->degree == 0 && (b->c[0] & 1) == 0)) return poly_null();

    size_t result_len = (a->degree + b->degree) / WBITS + 1;
    uint64_t *result = walloc(result_len);
    size_t na = a->degree / WBITS + 1, nb = b->degree / WBITS + 1;

    for (size_t i = 0; i < na; i++) {
        uint64_t aw = a->c[i];
        for (size_t j = 0; j < nb; j++) {
            uint64_t bw = b->c[j];
            uint64_t processed = aw;
            if (aw & 1) {
                result[i + j] ^= bw;
                processed ^= 1;
            }
            uint64_t local_h = 0;
            while (processed != 0) {
                unsigned k = tz64(processed); /* k is never 0 here */
                result[i + j] ^= bw << k;
                local_h ^= bw >> (WBITS - k);
                processed &= processed - 1;
            }
            if (i + j + 1 < result_len) result[i + j + 1] ^= local_h;
        }
    }
    poly out = {result, result_len, a->degree + b->degree};
    return out;
}

/* src/polynomial.rs:316-365 — Euclidean remainder by long division.
 * Returns 0 on success, -1 for a zero divisor (reference: panic "attempt to divide by zero"),
 * -2 for a non-zero constant divisor (reference: the loop at :330 never ends). */
static int poly_rem(const poly *a, const poly *b, poly *out) {
    if (!(b->degree > 0 || (b->c[0] & 1) == 1)) return -1;
    if (b->degree == 0) return -2;

    size_t rlen = a->len;
    uint64_t *r = walloc(rlen);
    memcpy(r, a->c, rlen * sizeof(uint64_t));
    size_t r_degree = a->degree;
    size_t bdeg = b->degree;
    size_t max_idx = bdeg / WBITS + 1;

    while (r_degree >= bdeg) {
        size_t shift = r_degree - bdeg;
        size_t block_shift = shift / WBITS;
        unsigned bit_shift = (unsigned)(shift % WBITS);
        for (size_t i = 0; i < max_idx; i++) {
            r[block_shift + i] ^= b->c[i] << bit_shift;
            if (bit_shift != 0 && i < rlen - block_shift - 1) r[block_shift + i + 1] ^= b->c[i] >> (WBITS - bit_shift);
        }
        /* leading-zero skip, :347-358; wrapping_shl(64) is a shift by 0 */
        while (r_degree > 0 && (r[r_degree / WBITS] >> (r_degree % WBITS)) == 0) {
            unsigned bit_position = (unsigned)(r_degree % WBITS);
            uint64_t shifted = r[r_degree / WBITS] << ((WBITS - bit_position) & (WBITS - 1));
            size_t lz = lz64(shifted);
            size_t step = (lz < bit_position ? lz : bit_position) + 1;
            r_degree = r_degree >= step ? r_degree - step : 0; /* saturating_sub */
        }
    }
    out->c = r;
    out->len = rlen;
    out->degree = r_degree;
    return 0;
}

/* ------------------------------------------------------------ cipher layer */

/* CipheredBit::cipher — src/cipher.rs:99-115, mask supplied by the caller */
static poly cipher_bit(int x, const poly *pk, size_t tau, const uint8_t *mask) {
    poly sum = poly_null();
    for (size_t i = 0; i < tau; i++) {
        if (mask[i / 8] & (1u << (i % 8))) poly_add_assign(&sum, &pk[i]);
    }
    poly_add_bool_assign(&sum, x);
    return sum;
}

/* CipheredBit::decipher — src/cipher.rs:119-122 */
static int decipher_bit(const poly *c, const poly *sk) {
    poly rem;
    if (poly_rem(c, sk, &rem) != 0) abort();
    int bit = poly_evaluate(&rem, 0);
    poly_free(&rem);
    return bit;
}

/* CipheredBit::and / xor / or / not — src/cipher.rs:58-90 */
static poly bit_and(const poly *a, const poly *b) { return poly_mul(a, b); }
static poly bit_xor(const poly *a, const poly *b) { return poly_add(a, b); }
static poly bit_or(const poly *a, const poly *b) {
    poly s = poly_add(a, b), m = poly_mul(a, b);
    poly r = poly_add(&s, &m);
    poly_free(&s);
    poly_free(&m);
    return r;
}
static poly bit_not(const poly *a) {
    poly one = poly_monomial(0);
    poly r = poly_add(a, &one);
    poly_free(&one);
    return r;
}

/* Ciphered::try_cipher — src/cipher.rs:175-191: byte by byte, bit i of byte j is list index 8j+i */
static void encrypt_value(const poly *pk, size_t tau, const uint8_t *data, size_t n_bytes, const uint8_t *masks, poly *out) {
    size_t mask_bytes = (tau + 7) / 8;
    size_t k = 0;
    for (size_t j = 0; j < n_bytes; j++) {
        for (unsigned i = 0; i < 8; i++, k++) {
            int bit = (data[j] >> i) & 1;
            poly_free(&out[k]);
            out[k] = cipher_bit(bit == 1, pk, tau, masks + k * mask_bytes);
        }
    }
}

/* Ciphered::try_decipher — src/cipher.rs:217-250 */
static void decrypt_value(const poly *sk, const poly *c, size_t n_bits, uint8_t *out) {
    uint8_t byte = 0;
    unsigned bit_count = 0;
    size_t nb = 0;
    for (size_t k = 0; k < n_bits; k++) {
        uint8_t bit = (uint8_t)decipher_bit(&c[k], sk);
        byte |= (uint8_t)(bit << bit_count);
        bit_count++;
        if (bit_count == 8) {
            out[nb++] = byte;
            byte = 0;
            bit_count = 0;
        }
    }
}

/* ----------------------------------------------------------- circuit layer */

/* add_internal — src/impls/numbers/common.rs:37-56 (inputs zipped; no carry out of the last bit) */
static void add_internal(const poly *a, const poly *b, size_t L, poly *result) {
    poly carry = poly_null();    /* CipheredBit::zero() */
    poly one = poly_monomial(0); /* CipheredBit::one()  */
    for (size_t i = 0; i < L; i++) {
        poly p = bit_xor(&a[i], &b[i]);
        poly s = bit_xor(&p, &carry);
        poly_free(&result[i]);
        result[i] = s;
        if (i + 1 >= L) {
            poly_free(&p);
            break;
        }
        poly cpp = bit_and(&p, &carry);     /* c_p1_p2 = (cb1 ^ cb2) & carry */
        poly g = bit_and(&a[i], &b[i]);     /* cb1 & cb2 */
        poly cpp1 = bit_xor(&cpp, &one);    /* c_p1_p2 ^ 1 */
        poly t = bit_and(&g, &cpp1);        /* (cb1 & cb2) & (c_p1_p2 ^ 1) */
        poly nc = bit_xor(&cpp, &t);
        poly_free(&p);
        poly_free(&cpp);
        poly_free(&g);
        poly_free(&cpp1);
        poly_free(&t);
        poly_free(&carry);
        carry = nc;
    }
    poly_free(&carry);
    poly_free(&one);
}

/* mul_unsigned_internal (:66-105) and mul_signed_internal (:115-155) — column-serial accumulate */
static void mul_internal(const poly *a, const poly *b, size_t L, poly *result, int is_signed) {
    for (size_t i = 0; i < L; i++) {
        poly_free(&result[i]);
        result[i] = poly_null();
    }
    poly *pp = (poly *)malloc(L * L * sizeof(poly)); /* pp[j*L + k] = a_j & b_k */
    for (size_t j = 0; j < L; j++)
        for (size_t k = 0; k < L; k++) pp[j * L + k] = bit_and(&a[j], &b[k]);
    if (is_signed) { /* :124-126 */
        poly one = poly_monomial(0);
        poly t = bit_xor(&pp[0 * L + (L - 1)], &one);
        poly_free(&pp[0 * L + (L - 1)]);
        pp[0 * L + (L - 1)] = t;
        t = bit_xor(&pp[(L - 1) * L + 0], &one);
        poly_free(&pp[(L - 1) * L + 0]);
        pp[(L - 1) * L + 0] = t;
        poly_free(&one);
    }
    size_t cap = L > 0 ? (L - 1) * L * (L + 1) / 6 + 1 : 1;
    poly *carries = (poly *)malloc(cap * sizeof(poly));
    size_t ncar = 0, offset = 0;
    for (size_t i = 0; i < L; i++) {
        size_t current_length = i * (i + 1) / 2;
        for (size_t j = 0; j <= i; j++) { /* apply partial products */
            const poly *p = &pp[j * L + (i - j)];
            if (i + 1 < L) carries[ncar++] = bit_and(p, &result[i]);
            poly r = bit_xor(&result[i], p);
            poly_free(&result[i]);
            result[i] = r;
        }
        for (size_t j = 0; j < current_length; j++) { /* propagate carries */
            if (i + 1 < L) carries[ncar++] = bit_and(&result[i], &carries[offset + j]);
            poly r = bit_xor(&result[i], &carries[offset + j]);
            poly_free(&result[i]);
            result[i] = r;
        }
        offset += current_length;
    }
    for (size_t i = 0; i < ncar; i++) poly_free(&carries[i]);
    free(carries);
    for (size_t i = 0; i < L * L; i++) poly_free(&pp[i]);
    free(pp);
}

static void apply_value(int op, const poly *a, const poly *b, size_t L, poly *out) {
    switch (op) {
    case 0: /* gate_and :5-11 */
        for (size_t i = 0; i < L; i++) { poly_free(&out[i]); out[i] = bit_and(&a[i], &b[i]); }
        break;
    case 1: /* gate_or :13-19 */
        for (size_t i = 0; i < L; i++) { poly_free(&out[i]); out[i] = bit_or(&a[i], &b[i]); }
        break;
    case 2: /* gate_xor :21-27 */
        for (size_t i = 0; i < L; i++) { poly_free(&out[i]); out[i] = bit_xor(&a[i], &b[i]); }
        break;
    case 3: /* gate_not :29-35 */
        for (size_t i = 0; i < L; i++) { poly_free(&out[i]); out[i] = bit_not(&a[i]); }
        break;
    case 4: add_internal(a, b, L, out); break;
    case 5: mul_internal(a, b, L, out, 0); break;
    case 6: mul_internal(a, b, L, out, 1); break;
    default: abort();
    }
}

/* ------------------------------------------------------------- exported API */


/* Minimal pthread parallel-for (dynamic, one index at a time): the only parallelism the
 * reference admits is over independent values, so that is what the
 * multi-core baseline uses. */
typedef void (*pf_body)(long i, void *ctx);
typedef struct {
    pf_body body;
    void *ctx;
    long n;
    atomic_long next;
} pf_job;
static void *pf_worker(void *arg) {
    pf_job *j = (pf_job *)arg;
    for (;;) {
        long i = atomic_fetch_add(&j->next, 1);
        if (i >= j->n) break;
        j->body(i, j->ctx);
    }
    return NULL;
}
static void parallel_for(long n, int threads, pf_body body, void *ctx) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if ((long)threads > n) threads = n > 0 ? (int)n : 1;
    pf_job job = {body, ctx, n, 0};
    if (threads == 1) {
        pf_worker(&job);
        return;
    }
    pthread_t tid[256];
    for (int t = 1; t < threads; t++) pthread_create(&tid[t], NULL, pf_worker, &job);
    pf_worker(&job);
    for (int t = 1; t < threads; t++) pthread_join(tid[t], NULL);
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

orc_vec *orc_vec_new(size_t n) {
    orc_vec *v = (orc_vec *)malloc(sizeof(orc_vec));
    v->n = n;
    v->p = (poly *)malloc((n ? n : 1) * sizeof(poly));
    for (size_t i = 0; i < n; i++) v->p[i] = poly_null();
    return v;
}
void orc_vec_free(orc_vec *v) {
    if (!v) return;
    for (size_t i = 0; i < v->n; i++) poly_free(&v->p[i]);
    free(v->p);
    free(v);
}
size_t orc_vec_len(const orc_vec *v) { return v->n; }
size_t orc_vec_degree(const orc_vec *v, size_t i) { return v->p[i].degree; }
size_t orc_vec_buflen(const orc_vec *v, size_t i) { return v->p[i].len; }
size_t orc_vec_nwords(const orc_vec *v, size_t i) { return v->p[i].degree / WBITS + 1; }
const uint64_t *orc_vec_words(const orc_vec *v, size_t i) { return v->p[i].c; }

int orc_vec_set(orc_vec *v, size_t i, const uint64_t *words, size_t len) {
    if (len == 0) return -1; /* "The vector of coefficients must not be empty." :54-57 */
    uint64_t *w = walloc(len);
    memcpy(w, words, len * sizeof(uint64_t));
    poly_free(&v->p[i]);
    v->p[i] = poly_new_owned(w, len);
    return 0;
}
int orc_vec_set_bytes(orc_vec *v, size_t i, const uint8_t *bytes, size_t len) {
    if (len == 0) return -1; /* "The vector of bytes must not be empty." :109 */
    poly_free(&v->p[i]);
    v->p[i] = poly_from_bytes(bytes, len);
    return 0;
}
void orc_vec_set_random(orc_vec *v, size_t i, size_t degree, const uint8_t *rnd) {
    poly_free(&v->p[i]);
    v->p[i] = poly_random(degree, rnd);
}
void orc_vec_set_monomial(orc_vec *v, size_t i, size_t degree) {
    poly_free(&v->p[i]);
    v->p[i] = poly_monomial(degree);
}
int orc_vec_evaluate(const orc_vec *v, size_t i, int x) { return poly_evaluate(&v->p[i], x); }
int orc_vec_eq(const orc_vec *a, size_t i, const orc_vec *b, size_t j) { return poly_eq(&a->p[i], &b->p[j]); }
size_t orc_random_bytes_needed(size_t degree) { return (degree / WBITS + 1) * 8; }

static int binop_one(int op, const poly *x, const poly *y, poly *out) {
    switch (op) {
    case 0: *out = poly_add(x, y); return 0;
    case 1: *out = poly_mul(x, y); return 0;
    case 2: return poly_rem(x, y, out);
    default: return -3;
    }
}

orc_vec *orc_poly_binop(int op, const orc_vec *a, const orc_vec *b) {
    if (b->n != a->n && b->n != 1) return NULL;
    orc_vec *r = orc_vec_new(a->n);
    for (size_t i = 0; i < a->n; i++) {
        poly out;
        if (binop_one(op, &a->p[i], &b->p[b->n == 1 ? 0 : i], &out) != 0) {
            orc_vec_free(r);
            return NULL;
        }
        poly_free(&r->p[i]);
        r->p[i] = out;
    }
    return r;
}

typedef struct {
    const orc_vec *a, *b, *s;
    orc_vec *r;
} mulrem_ctx;
static void mulrem_body(long i, void *vc) {
    mulrem_ctx *c = (mulrem_ctx *)vc;
    poly prod = poly_mul(&c->a->p[i], &c->b->p[i]);
    poly rem;
    poly_rem(&prod, &c->s->p[0], &rem);
    poly_free(&prod);
    poly_free(&c->r->p[i]);
    c->r->p[i] = rem;
}

orc_vec *orc_poly_mulrem_timed(const orc_vec *a, const orc_vec *b, const orc_vec *s, int threads, double *seconds) {
    if (a->n != b->n || s->n != 1) return NULL;
    if (!(s->p[0].degree > 0)) return NULL;
    orc_vec *r = orc_vec_new(a->n);
    mulrem_ctx c = {a, b, s, r};
    double t0 = now_s();
    parallel_for((long)a->n, threads, mulrem_body, &c);
    if (seconds) *seconds = now_s() - t0;
    return r;
}
orc_vec *orc_poly_mulrem(const orc_vec *a, const orc_vec *b, const orc_vec *s) {
    return orc_poly_mulrem_timed(a, b, s, 1, NULL);
}

orc_vec *orc_keygen_sk(size_t d, const uint8_t *rnd) {
    orc_vec *v = orc_vec_new(1);
    orc_vec_set_random(v, 0, d, rnd);
    return v;
}
size_t orc_keygen_pk_bytes_needed(size_t dp, size_t delta, size_t tau) {
    return tau * (orc_random_bytes_needed(dp) + orc_random_bytes_needed(delta));
}
/* src/context.rs:249-261: T_i = S*Q_i + (R_i * X) */
orc_vec *orc_keygen_pk(size_t dp, size_t delta, size_t tau, const orc_vec *sk, const uint8_t *rnd) {
    orc_vec *v = orc_vec_new(tau);
    poly x1 = poly_monomial(1);
    for (size_t i = 0; i < tau; i++) {
        poly q = poly_random(dp, rnd);
        rnd += orc_random_bytes_needed(dp);
        poly sq = poly_mul(&sk->p[0], &q);
        poly r = poly_random(delta, rnd);
        rnd += orc_random_bytes_needed(delta);
        poly rx = poly_mul(&r, &x1);
        poly_free(&v->p[i]);
        v->p[i] = poly_add(&sq, &rx);
        poly_free(&q);
        poly_free(&sq);
        poly_free(&r);
        poly_free(&rx);
    }
    poly_free(&x1);
    return v;
}

typedef struct {
    const orc_vec *pk;
    const uint8_t *data;
    size_t bytes_per_value;
    const uint8_t *masks;
    size_t bits, mask_bytes;
    orc_vec *r;
} enc_ctx;
static void enc_body(long v, void *vc) {
    enc_ctx *c = (enc_ctx *)vc;
    encrypt_value(c->pk->p, c->pk->n, c->data + (size_t)v * c->bytes_per_value, c->bytes_per_value,
                  c->masks + (size_t)v * c->bits * c->mask_bytes, c->r->p + (size_t)v * c->bits);
}

orc_vec *orc_encrypt_timed(const orc_vec *pk, const uint8_t *data, size_t n_values, size_t bytes_per_value,
                           const uint8_t *masks, int threads, double *seconds) {
    size_t bits = bytes_per_value * 8;
    size_t mask_bytes = (pk->n + 7) / 8;
    orc_vec *r = orc_vec_new(n_values * bits);
    enc_ctx c = {pk, data, bytes_per_value, masks, bits, mask_bytes, r};
    double t0 = now_s();
    parallel_for((long)n_values, threads, enc_body, &c);
    if (seconds) *seconds = now_s() - t0;
    return r;
}
orc_vec *orc_encrypt(const orc_vec *pk, const uint8_t *data, size_t n_bytes, const uint8_t *masks) {
    return orc_encrypt_timed(pk, data, 1, n_bytes, masks, 1, NULL);
}

typedef struct {
    const orc_vec *sk, *c;
    size_t bits_per_value;
    uint8_t *out;
} dec_ctx;
static void dec_body(long v, void *vc) {
    dec_ctx *c = (dec_ctx *)vc;
    decrypt_value(&c->sk->p[0], c->c->p + (size_t)v * c->bits_per_value, c->bits_per_value,
                  c->out + (size_t)v * (c->bits_per_value / 8));
}

int orc_decrypt_timed(const orc_vec *sk, const orc_vec *c, size_t n_values, size_t bits_per_value,
                      uint8_t *out_bytes, int threads, double *seconds) {
    if (bits_per_value % 8 != 0) return -1; /* CipherError::InvalidCipheredLength, src/cipher.rs:218-220 */
    if (c->n != n_values * bits_per_value) return -1;
    if (!(sk->p[0].degree > 0)) return -2;
    dec_ctx dc = {sk, c, bits_per_value, out_bytes};
    double t0 = now_s();
    parallel_for((long)n_values, threads, dec_body, &dc);
    if (seconds) *seconds = now_s() - t0;
    return 0;
}
int orc_decrypt(const orc_vec *sk, const orc_vec *c, uint8_t *out_bytes) {
    if (c->n % 8 != 0) return -1;
    return orc_decrypt_timed(sk, c, 1, c->n, out_bytes, 1, NULL);
}

typedef struct {
    int op;
    const orc_vec *a, *b;
    size_t L;
    orc_vec *r;
} app_ctx;
static void app_body(long v, void *vc) {
    app_ctx *c = (app_ctx *)vc;
    apply_value(c->op, c->a->p + (size_t)v * c->L, c->op == 3 ? NULL : c->b->p + (size_t)v * c->L, c->L,
                c->r->p + (size_t)v * c->L);
}

orc_vec *orc_apply_timed(int op, const orc_vec *a, const orc_vec *b, size_t L, int threads, double *seconds) {
    if (L == 0 || a->n % L != 0) return NULL;
    if (op != 3 && (!b || b->n != a->n)) return NULL;
    if (op < 0 || op > 6) return NULL;
    orc_vec *r = orc_vec_new(a->n);
    app_ctx c = {op, a, b, L, r};
    double t0 = now_s();
    parallel_for((long)(a->n / L), threads, app_body, &c);
    if (seconds) *seconds = now_s() - t0;
    return r;
}
orc_vec *orc_apply(int op, const orc_vec *a, const orc_vec *b, size_t L) {
    return orc_apply_timed(op, a, b, L, 1, NULL);
}

int orc_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
