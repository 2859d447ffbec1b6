/*
 * hm_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * A plain-C restatement of the GF(2)[X] ciphertext arithmetic of
 * mathisbot/homomorph-rust (crate `homomorph` v1.1.0).  It exists to CHECK the
 * CUDA engine and to serve as the timed CPU baseline; it is never linked,
 * imported or called by the product (libhmgpu.so / homomorph_rust_b200).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may use it.
 *
 * Parity status: the reference cannot be compiled here (no Rust toolchain).
 *   - polynomial layer (degree/add/mul/rem/evaluate/eq/bytes): PINNED by every
 *     known-answer vector in src/polynomial.rs:439-612 (tests/test_oracle_kat.py);
 *   - cipher / circuit layer: the reference holds no ciphertext-level vectors
 *     (SURVEY.md §8c); pinned by the reference's plaintext-level expectations
 *     (src/impls/numbers/uint.rs:109-293, src/cipher.rs:276-304) and by an
 *     independent big-integer model (oracle/pymodel.py).
 *
 * Every function cites the reference file:line it follows.
 */
#ifndef HM_ORACLE_H
#define HM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* A growable list of polynomials; the unit the Python side talks in. */
typedef struct orc_vec orc_vec;

orc_vec *orc_vec_new(size_t n);                 /* n null polynomials            */
void     orc_vec_free(orc_vec *v);
size_t   orc_vec_len(const orc_vec *v);
size_t   orc_vec_degree(const orc_vec *v, size_t i);
size_t   orc_vec_buflen(const orc_vec *v, size_t i);   /* allocated words        */
size_t   orc_vec_nwords(const orc_vec *v, size_t i);   /* degree/64+1 (canonical) */
const uint64_t *orc_vec_words(const orc_vec *v, size_t i);
/* Polynomial::new — src/polynomial.rs:53-63; returns -1 on empty input (the reference panics). */
int      orc_vec_set(orc_vec *v, size_t i, const uint64_t *words, size_t len);
/* Polynomial::from_bytes — src/polynomial.rs:108-122 */
int      orc_vec_set_bytes(orc_vec *v, size_t i, const uint8_t *bytes, size_t len);
/* Polynomial::random with caller-supplied bytes — src/polynomial.rs:73-96 */
void     orc_vec_set_random(orc_vec *v, size_t i, size_t degree, const uint8_t *rnd);
/* Polynomial::monomial / null — src/polynomial.rs:132-150 */
void     orc_vec_set_monomial(orc_vec *v, size_t i, size_t degree);
/* Polynomial::evaluate — src/polynomial.rs:168-181 */
int      orc_vec_evaluate(const orc_vec *v, size_t i, int x);
/* PartialEq — src/polynomial.rs:416-426 */
int      orc_vec_eq(const orc_vec *a, size_t i, const orc_vec *b, size_t j);
/* Bytes needed by orc_vec_set_random for a given degree. */
size_t   orc_random_bytes_needed(size_t degree);

/* Elementwise polynomial ops over lists; `b` of length 1 broadcasts.
 * op: 0 = add (src/polynomial.rs:190-213), 1 = mul (:252-310), 2 = rem (:316-365).
 * Returns NULL when rem is asked with a zero divisor (reference panics) or a
 * constant divisor (reference never terminates). */
orc_vec *orc_poly_binop(int op, const orc_vec *a, const orc_vec *b);
/* mul followed by rem by s[0] — the `mul+rem` unit of BASELINE.json. */
orc_vec *orc_poly_mulrem(const orc_vec *a, const orc_vec *b, const orc_vec *s);

/* Key generation with caller-supplied randomness.
 * SecretKey::random — src/context.rs:160-162.
 * PublicKey::random — src/context.rs:249-261: per i, Q = random(dp) then
 * R = random(delta), bytes consumed in that order. */
orc_vec *orc_keygen_sk(size_t d, const uint8_t *rnd);
orc_vec *orc_keygen_pk(size_t dp, size_t delta, size_t tau, const orc_vec *sk, const uint8_t *rnd);
size_t   orc_keygen_pk_bytes_needed(size_t dp, size_t delta, size_t tau);

/* Ciphered::<T>::try_cipher — src/cipher.rs:175-191, with CipheredBit::cipher
 * (:99-115) taking its subset mask from `masks` instead of getrandom (:92-97):
 * ceil(tau/8) bytes per bit, consumed in bit order.  `data` = n_bytes of the
 * bincode (fixint, LE) encoding, i.e. the integers' little-endian bytes. */
orc_vec *orc_encrypt(const orc_vec *pk, const uint8_t *data, size_t n_bytes, const uint8_t *masks);
/* Ciphered::<T>::try_decipher — src/cipher.rs:217-250; returns -1 if len % 8 != 0. */
int      orc_decrypt(const orc_vec *sk, const orc_vec *c, uint8_t *out_bytes);

/* Gates and circuits of src/impls/numbers/common.rs applied value by value:
 * a and b hold n_values * L bit-ciphertexts (LSB first).
 * op: 0 and (:5-11), 1 or (:13-19), 2 xor (:21-27), 3 not (:29-35, b ignored),
 *     4 add_internal (:37-56), 5 mul_unsigned_internal (:66-105),
 *     6 mul_signed_internal (:115-155). */
orc_vec *orc_apply(int op, const orc_vec *a, const orc_vec *b, size_t L);

/* Multi-threaded, internally timed variants for the CPU baseline.  Work is
 * split by value (encrypt/decrypt/apply) or by element (binop) over `threads`
 * pthreads; *seconds receives the wall time of the parallel region. */
orc_vec *orc_apply_timed(int op, const orc_vec *a, const orc_vec *b, size_t L, int threads, double *seconds);
orc_vec *orc_encrypt_timed(const orc_vec *pk, const uint8_t *data, size_t n_values, size_t bytes_per_value,
                           const uint8_t *masks, int threads, double *seconds);
int      orc_decrypt_timed(const orc_vec *sk, const orc_vec *c, size_t n_values, size_t bits_per_value,
                           uint8_t *out_bytes, int threads, double *seconds);
orc_vec *orc_poly_mulrem_timed(const orc_vec *a, const orc_vec *b, const orc_vec *s, int threads, double *seconds);
int      orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
