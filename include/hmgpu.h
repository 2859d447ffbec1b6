/*
 * hmgpu.h — C ABI of libhmgpu.so, the B200-native batched ciphertext engine for the
 * GF(2)[X] hot path of mathisbot/homomorph-rust (crate `homomorph` v1.1.0).
 *
 * The reference has NO FFI/plugin interface; its only seams are
 * Rust-level.  Each entry point below names the reference interface it replaces
 * (file:line into the reference tree) so that a Rust shim can bind it 1:1 — the
 * `extern "C"` block a maintainer would add is in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; opaque handles; no exceptions/panics cross the ABI;
 *   - every function returns an hm_status (0 = ok, negative = error mirroring the
 *     reference's error enums / panics); hm_status_string() names it;
 *   - a *batch* is n values x L bit-ciphertexts ("slots", LSB first — src/cipher.rs:180-185)
 *     resident in HBM.  Slot k of every value has the same fixed width w[k] in 64-bit
 *     words; words are the reference's own coefficient words: LSB-first, coefficient of
 *     X^i is bit i%64 of word i/64 (src/polynomial.rs:144,172), zero padded above the
 *     degree.  The canonical polynomial the reference would hold is (degree = highest set
 *     bit, words[0 ..= degree/64]) — src/polynomial.rs:404-426 ignore anything above;
 *   - host-side layout of a batch ("padded layout"): value-major, slot-minor, word-minor:
 *         host[(v * value_words) + slot_offset[k] + j],  value_words = sum_k w[k];
 *   - work is enqueued on the context's CUDA stream; calls that return data to host
 *     memory synchronise that stream before returning;
 *   - a context may be used by one host thread at a time; contexts are independent.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails
 *     with HM_ERR_CUDA.
 */
#ifndef HMGPU_H
#define HMGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hm_context hm_context; /* Context            — src/context.rs:301-305 */
typedef struct hm_batch hm_batch;     /* Vec<Ciphered<T>>   — src/cipher.rs:126-130, device resident */
typedef struct hm_group hm_group;             /* one Context replicated on several GPUs of a box            */
typedef struct hm_group_batch hm_group_batch; /* one logical batch, contiguous index shards on the devices  */

typedef enum hm_status {
    HM_OK = 0,
    HM_ERR_INVALID_PARAMETERS = -1,  /* Parameters::new asserts            — src/context.rs:87-94   */
    HM_ERR_PUBLIC_KEY_UNSET = -2,    /* ContextCryptoError::PublicKeyUnset — src/context.rs:41-52   */
    HM_ERR_SECRET_KEY_UNSET = -3,    /* ContextCryptoError::SecretKeyUnset — src/context.rs:41-52   */
    HM_ERR_OPERATION_REQUIREMENT = -4, /* OperationError::InvalidParameters — src/operations.rs:11-18 */
    HM_ERR_INVALID_LENGTH = -5,      /* CipherError::InvalidCipheredLength — src/cipher.rs:17-24    */
    HM_ERR_CUDA = -6,                /* device missing / CUDA runtime failure                       */
    HM_ERR_UNSUPPORTED = -7,         /* shape outside what the kernels are built for                */
    HM_ERR_INVALID_ARGUMENT = -8,    /* NULL pointer, mismatched batches, empty polynomial          */
    HM_ERR_DIVIDE_BY_ZERO = -9,      /* Polynomial::rem panic              — src/polynomial.rs:319-322 */
    HM_ERR_OUT_OF_MEMORY = -10       /* host or device allocation failed (the reference aborts on OOM)   */
} hm_status;

/* Operation selectors = the marker types of src/impls/numbers.rs:7-25. */
typedef enum hm_op {
    HM_OP_AND = 0, /* HomomorphicAndGate         MIN_D_OVER_DELTA  2 — src/impls/numbers.rs:27-29 */
    HM_OP_OR = 1,  /* HomomorphicOrGate                            2 — :31-33 */
    HM_OP_XOR = 2, /* HomomorphicXorGate                           1 — :35-37 */
    HM_OP_NOT = 3, /* HomomorphicNotGate                           1 — :39-41 */
    HM_OP_ADD = 4, /* HomomorphicAddition                         21 — :43-45 */
    HM_OP_MUL = 5  /* HomomorphicMultiplication                   64 — :47-50 */
} hm_op;

const char *hm_status_string(int status);
/* Last CUDA error text recorded on this context (empty string if none). */
const char *hm_last_error(const hm_context *ctx);
/* Number of CUDA devices visible; 0 when there is no driver/GPU (never an error). */
int hm_device_count(void);

/* ---- Context / Parameters ------------------------------------------------------------
 * Parameters::new + Context::new — src/context.rs:87-94, :341-347.  Rejects d, dp, delta or
 * tau == 0 and delta >= d with HM_ERR_INVALID_PARAMETERS (the reference panics).
 * `device` is the CUDA ordinal the context (keys, tables, stream, batches) lives on. */
int hm_context_create(uint16_t d, uint16_t dp, uint16_t delta, uint16_t tau, int device, hm_context **out);
/* Drop: zeroises the secret key, its derived tables and staging (src/context.rs:199-206). */
void hm_context_destroy(hm_context *ctx);
int hm_context_parameters(const hm_context *ctx, uint16_t *d, uint16_t *dp, uint16_t *delta, uint16_t *tau);
/* Use a caller-owned CUDA stream (cudaStream_t) for all subsequent work; NULL = own stream. */
int hm_context_set_stream(hm_context *ctx, void *cuda_stream);
void *hm_context_stream(const hm_context *ctx);
int hm_context_synchronize(hm_context *ctx);
/* Counter of kernels this library has launched on this context (bench.py "gpu_launches"). */
uint64_t hm_context_kernel_launches(const hm_context *ctx);
/* CUDA ordinal the context lives on. */
int hm_context_device(const hm_context *ctx);
/* Process-wide kernel-selection knobs (tests / experiments).  "adder_thread_min": smallest batch for which the
 * u32-class adder uses the thread-per-value Karatsuba kernel instead of the warp-per-value kernel (-1 = default).
 * "mul_thread_min": smallest (values x 24-word chunks) for which a general product uses the thread-per-chunk Karatsuba
 * kernel (-1 = default).  "mul_circuit_sequential": 1 = launch the multiplier circuit's carry products one at a time
 * instead of one batch per column; "adder_generic_sequential": 1 = the generic adder evaluates common.rs:44-53 literally
 * (two long products per bit) instead of the regrouped one-product-per-bit form; "mul_thread_chunk": 24 | 32 words per
 * chunk of the thread-per-chunk product kernel; "adder_chain": 0 = the round-1 thread-per-value adder kernels, otherwise
 * the dynamically scheduled chain of kernels_adder.cu as 10 * (window in shared memory) + CTAs per SM (4 = default, 3, 12, 13);
 * "adder_phases": work units per value of that chain (1..8; 0 = default: 8 for batches of at least 2.5 waves of resident warps, else 1); "host_chunk_mb": bytes per stage of the host-buffer pipeline
 * of hm_apply2_host (MiB, default 96); "pool_max_mb": batches up to this size come from the context's stream-ordered pool,
 * larger ones from plain cudaMalloc (MiB, default 16384).
 * All of them give the same polynomials: they are the A/B arms of tests. */
int hm_set_tuning(const char *key, long value);

/* Context::set_secret_key(SecretKey::from_bytes(bytes)) — src/context.rs:153-155, :568-571.
 * `bytes` is SecretKey::to_bytes(): little-endian u64 words (src/polynomial.rs:99-122).
 * Clears the public key, like the reference.  Like the reference, the degree of S is taken from the bytes (it is not
 * compared with d); the zero polynomial is refused with HM_ERR_DIVIDE_BY_ZERO and S = 1 with HM_ERR_INVALID_ARGUMENT. */
int hm_set_secret_key(hm_context *ctx, const uint8_t *bytes, size_t len);
/* Context::set_public_key(PublicKey::from_bytes(..)) — src/context.rs:239-245, :592-595.
 * polys[i]/lens[i] = PublicKey::to_bytes()[i]; n_polys must equal tau. */
int hm_set_public_key(hm_context *ctx, const uint8_t *const *polys, const size_t *lens, size_t n_polys);
/* Context::generate_secret_key + generate_public_key (src/context.rs:421-454) from a 64-bit seed instead of getrandom:
 * S = Polynomial::random(d) on the host, T_i = S * Q_i + X * R_i (src/context.rs:249-261) for all i on the device (one
 * product launch with S broadcast + one XOR launch).  The random bytes are a documented Philox4x32-10 stream —
 * hm_key_stream_host(seed, stream, ...) returns it: stream 0 feeds S, stream 1 feeds Q_0, R_0, Q_1, R_1, ... in the order
 * and byte counts the reference draws them (Polynomial::random: (degree/64+1)*8 bytes each) — so the reference fed with the
 * same bytes (a getrandom shim) holds the same keys.  REPRODUCIBILITY / TESTING ONLY: Philox is not a cryptographic
 * generator and 64 bits of seed are not a key; production keys come from hm_set_secret_key / hm_set_public_key. */
int hm_generate_keys_seeded(hm_context *ctx, uint64_t seed);
int hm_key_stream_host(uint64_t seed, uint32_t stream, size_t nbytes, uint8_t *out);
/* SecretKey::to_bytes() / PublicKey::to_bytes()[i] of the keys the context holds (src/context.rs:192-194, :291-297).
 * *len receives the byte count; out == NULL only queries it. */
int hm_secret_key_bytes(const hm_context *ctx, uint8_t *out, size_t capacity, size_t *len);
int hm_public_key_bytes(const hm_context *ctx, size_t i, uint8_t *out, size_t capacity, size_t *len);
int hm_has_secret_key(const hm_context *ctx);
int hm_has_public_key(const hm_context *ctx);

/* ---- Batches -------------------------------------------------------------------------- */
size_t hm_batch_len(const hm_batch *b);          /* n values                                   */
uint32_t hm_batch_bits(const hm_batch *b);       /* L bit-ciphertexts per value (Ciphered::len) */
size_t hm_batch_value_words(const hm_batch *b);  /* sum of slot widths, in u64 words           */
/* Copies the L slot widths (u64 words) into widths_out[0..L). */
int hm_batch_slot_words(const hm_batch *b, uint32_t *widths_out);
void *hm_batch_device_ptr(const hm_batch *b);    /* the padded layout, in HBM                   */
/* Releases a batch.  The owning context is recorded in the batch, so `ctx` may be NULL.  Batches still alive when their
 * context is destroyed lose their device memory at that point (they become orphans); freeing an orphan afterwards only
 * releases the handle, so either destruction order is safe. */
void hm_batch_free(hm_context *ctx, hm_batch *b);
/* Host <-> HBM in the padded layout described above (n * value_words u64 words).
 * upload = what a shim does with Vec<CipheredBit> built by Ciphered::new_from_raw
 * (src/cipher.rs:151-156); download = reading the polynomials back. */
int hm_batch_upload(hm_context *ctx, size_t n, uint32_t L, const uint32_t *slot_words, const uint64_t *host,
                    hm_batch **out);
int hm_batch_download(hm_context *ctx, const hm_batch *b, uint64_t *host);
/* Values [first, first + count) only (count * value_words words): spot checks and streaming readers of large results. */
int hm_batch_download_range(hm_context *ctx, const hm_batch *b, size_t first, size_t count, uint64_t *host);
/* Same as hm_batch_upload, but the caller states a degree bound per slot (>= the true degree of every
 * polynomial in that slot); slot k is degree_bounds[k]/64+1 words wide.  Tight bounds keep the results of
 * multiplications as narrow as the reference's (out len = (da+db)/64+1, src/polynomial.rs:264). */
int hm_batch_upload_bounded(hm_context *ctx, size_t n, uint32_t L, const uint64_t *degree_bounds,
                            const uint64_t *host, hm_batch **out);
/* Copies the L per-slot degree bounds the engine tracks for this batch. */
int hm_batch_slot_degree_bounds(const hm_batch *b, uint64_t *bounds_out);
/* ---- Wire format (engine-defined: the reference has no ciphertext serialisation, src/cipher.rs:30) ----
 * little endian: "HMB1" | u16 d, dp, delta, tau | u32 L | u64 n | L x u64 degree bound |
 *                n * value_words x u64 coefficient words (padded layout above).
 * deserialize rejects a buffer written under other parameters with HM_ERR_INVALID_PARAMETERS. */
size_t hm_batch_serialized_size(const hm_batch *b);
int hm_batch_serialize(hm_context *ctx, const hm_batch *b, uint8_t *out, size_t capacity);
int hm_batch_deserialize(hm_context *ctx, const uint8_t *in, size_t len, hm_batch **out);
/* Header check of a wire buffer without a context or a GPU: every field is validated against `len` with overflow-free
 * arithmetic (degree bounds <= 2^31, sum of slot widths < 2^32, body length == n * value_words * 8 exactly) — the buffer
 * may come from another party.  Outputs (each may be NULL): params_out = {d, dp, delta, tau}. */
int hm_batch_wire_inspect(const uint8_t *in, size_t len, uint16_t params_out[4], uint32_t *L_out, uint64_t *n_out,
                          uint64_t *value_words_out);
/* Canonical export for offline comparison with a Rust build of the reference: per polynomial (value-major,
 * slot-minor) `u64 degree` followed by degree/64+1 words — what Polynomial holds (src/polynomial.rs:22-26).
 * out may be NULL to only count; *written = number of u64. */
int hm_batch_download_canonical(hm_context *ctx, const hm_batch *b, uint64_t *out, size_t capacity_words,
                                size_t *written);
/* The inverse: n * L canonical polynomials in the same order and format (e.g. exported by a Rust build of the reference)
 * become a device batch.  Slot k is as wide as the largest degree seen in slot k, or as degree_bounds[k] when
 * degree_bounds != NULL (each must cover its slot).  A stated degree that is not the highest set bit of its words is
 * refused (HM_ERR_INVALID_ARGUMENT); a truncated or over-long buffer gives HM_ERR_INVALID_LENGTH. */
int hm_batch_upload_canonical(hm_context *ctx, size_t n, uint32_t L, const uint64_t *in, size_t in_words,
                              const uint64_t *degree_bounds, hm_batch **out);
/* Page-locked host memory for the host-buffer entry points (hm_encrypt, hm_decrypt, hm_apply2_host,
 * upload/download): pageable memory works too, but is staged by the driver. NULL on failure. */
void *hm_host_alloc(size_t bytes);
void hm_host_free(void *p);
/* Device-resident duplicate (Ciphered::clone). */
int hm_batch_clone(hm_context *ctx, const hm_batch *b, hm_batch **out);

/* Fields of a user struct (examples/simple_struct.rs:32-58).  hm_batch_slice: slots [first_bit, first_bit + n_bits) of every
 * value as a new batch — `Ciphered::split_at` + `Ciphered::new_from_raw` on one field (first_bit + n_bits > L ->
 * HM_ERR_INVALID_LENGTH; the reference's split_at panics).  hm_batch_concat: the slots of `count` batches of equal n, in
 * order — the `extend_from_slice` merge of the per-field results (more than 128 slots -> HM_ERR_UNSUPPORTED).
 * hm_apply2_fields: the whole pattern — field f covers field_bits[f] consecutive slots (sum must be L), `op` is applied to
 * each field of a and b as its own integer and the results are concatenated; requirement check as hm_apply2. */
int hm_batch_slice(hm_context *ctx, const hm_batch *src, uint32_t first_bit, uint32_t n_bits, hm_batch **out);
int hm_batch_concat(hm_context *ctx, const hm_batch *const *parts, size_t count, hm_batch **out);
int hm_apply2_fields(hm_context *ctx, int op, const hm_batch *a, const hm_batch *b, const uint32_t *field_bits, size_t n_fields,
                     hm_batch **out);

/* ---- Encrypt / decrypt ----------------------------------------------------------------
 * Context::encrypt -> Ciphered::try_cipher -> CipheredBit::cipher for n integers at once —
 * src/context.rs:463-471, src/cipher.rs:175-191, :99-115.
 *   values : n * (L/8) bytes, each integer little endian (bincode fixint LE, src/cipher.rs:6-13,176)
 *   L      : bits per value, multiple of 8 (u8 = 8, u32 = 32, ...)
 *   masks  : the subset U for every bit, replacing getrandom (src/cipher.rs:92-97):
 *            n * L * ceil(tau/8) bytes, value-major then bit-minor; polynomial T_i is included
 *            iff masks[..][i/8] & (1 << (i%8))  (src/cipher.rs:106).
 * `values`/`masks` are HOST pointers; the *_device variant takes device pointers. */
int hm_encrypt(hm_context *ctx, const uint8_t *values, size_t n, uint32_t L, const uint8_t *masks, hm_batch **out);
int hm_encrypt_device(hm_context *ctx, const uint8_t *d_values, size_t n, uint32_t L, const uint8_t *d_masks,
                      hm_batch **out);
/* Same, writing into an existing batch of n values x L fresh-width slots (no allocation). */
int hm_encrypt_device_into(hm_context *ctx, const uint8_t *d_values, size_t n, uint32_t L, const uint8_t *d_masks,
                           hm_batch *out);
/* Seeded subset masks (Philox4x32-10, Random123) — REPRODUCIBILITY / TESTING ONLY: the masks are the only randomness behind
 * ciphertext indistinguishability, the reference draws them from the OS (getrandom, src/cipher.rs:92-97) and Philox keyed
 * by a 64-bit seed is not a cryptographic generator.  Production callers pass masks from a CSPRNG to hm_encrypt.
 * Stream layout: bytes [16 b, 16 b + 16) of the mask of bit-ciphertext u are
 * Philox(counter = (u_lo, u_hi, b, 0), key = (seed_lo, seed_hi)), words little endian, truncated to ceil(tau/8) bytes.
 * The *_host variant computes the same stream on the CPU (no GPU needed), so a seeded encryption can be reproduced
 * bit for bit by the reference/oracle fed with these masks.  hm_encrypt_seeded = hm_encrypt with device-generated
 * masks: only the plaintext crosses PCIe. */
int hm_masks_generate_host(uint16_t tau, size_t units, uint64_t seed, uint8_t *masks_out);
int hm_masks_generate_device(hm_context *ctx, size_t units, uint64_t seed, uint8_t *d_masks_out);
int hm_encrypt_seeded(hm_context *ctx, const uint8_t *values, size_t n, uint32_t L, uint64_t seed, hm_batch **out);
/* The same for a shard of a larger logical batch: bit-ciphertext u of this call takes position first_unit + u of the mask
 * stream (first_unit = first value of the shard x L), so shards encrypted separately equal the batch encrypted at once.
 * sync == 0 returns once the work is enqueued (see the *_async calls). */
int hm_encrypt_seeded_at(hm_context *ctx, const uint8_t *values, size_t n, uint32_t L, uint64_t seed, uint64_t first_unit, int sync,
                         hm_batch **out);
/* Seeded encryption of plaintexts already in HBM into an existing batch of n x L fresh-width slots (no allocation, no copy):
 * the masks of stream positions [first_unit, first_unit + n L) are drawn inside the encrypt kernel where the parameter set
 * has a fused kernel (d + d' = 256, tau = 128), else written to a temporary buffer first.  Same ciphertexts as
 * hm_encrypt_device_into fed with hm_masks_generate_host(tau, ...). */
int hm_encrypt_device_seeded_into(hm_context *ctx, const uint8_t *d_values, size_t n, uint32_t L, uint64_t seed, uint64_t first_unit,
                                  hm_batch *out);
/* Asynchronous forms of hm_encrypt / hm_decrypt: they return as soon as the copies and kernels are enqueued on the
 * context's stream.  The host buffers must stay valid (and, for decrypt, unread) until hm_context_synchronize; pinned
 * memory (hm_host_alloc) keeps the copies asynchronous.  They let one host thread drive several devices at once. */
int hm_encrypt_async(hm_context *ctx, const uint8_t *values, size_t n, uint32_t L, const uint8_t *masks, hm_batch **out);
int hm_decrypt_async(hm_context *ctx, const hm_batch *b, uint8_t *values_out);
/* Context::decrypt -> Ciphered::try_decipher -> CipheredBit::decipher — src/context.rs:480-488,
 * src/cipher.rs:217-250, :119-122.  Writes n * (L/8) bytes.  L % 8 != 0 -> HM_ERR_INVALID_LENGTH. */
int hm_decrypt(hm_context *ctx, const hm_batch *b, uint8_t *values_out);
int hm_decrypt_device(hm_context *ctx, const hm_batch *b, uint8_t *d_values_out);

/* ---- Homomorphic operations -------------------------------------------------------------
 * Context::apply2::<O, T>(&a, &b) — src/context.rs:515-527 — for every value of the batch:
 * validate_operation (d >= MIN_D_OVER_DELTA * delta, src/context.rs:310-323, else
 * HM_ERR_OPERATION_REQUIREMENT) then the circuit of src/impls/numbers/common.rs
 * (gate_and :5-11, gate_or :13-19, gate_xor :21-27, add_internal :37-56,
 * mul_unsigned_internal :66-105).  a and b must have the same n and L. */
int hm_apply2(hm_context *ctx, int op, const hm_batch *a, const hm_batch *b, hm_batch **out);
/* Context::apply1::<HomomorphicNotGate, T>(&mut a) — src/context.rs:496-507, common.rs:29-35. In place. */
int hm_apply1(hm_context *ctx, int op, hm_batch *a);
/* The reference's `unsafe { O::apply(..) }` (src/operations.rs:81,140): same, without the
 * parameter check. */
int hm_apply2_unchecked(hm_context *ctx, int op, const hm_batch *a, const hm_batch *b, hm_batch **out);
/* Like hm_apply2 but writes into an existing batch `out` whose slot widths already equal
 * hm_result_slot_words() of the operands (reuses the allocation; no cudaMalloc on the hot path). */
int hm_apply2_into(hm_context *ctx, int op, const hm_batch *a, const hm_batch *b, hm_batch *out);
/* Same operation through the generic, slot-by-slot kernels in the reference's own evaluation order
 * (no fused circuit kernel).  Results are identical; used to cross-check the fused adder. */
int hm_apply2_generic(hm_context *ctx, int op, const hm_batch *a, const hm_batch *b, hm_batch **out);
/* MIN_D_OVER_DELTA of an op, or a negative status for an unknown op. */
int hm_op_min_d_over_delta(int op);
/* Degree bounds of the result slots of `op` on operands with the given per-slot degree bounds (the recurrences of
 * common.rs:37-105 on degrees; slot k is bounds[k]/64+1 words wide).  No context or GPU needed. */
int hm_result_slot_bounds(int op, uint32_t L, const uint64_t *a_bounds, const uint64_t *b_bounds, uint64_t *out_bounds);
/* Same from slot widths alone: every operand slot is taken at its widest (degree 64 w - 1), because a width says
 * nothing about the degrees inside.  Use hm_result_slot_bounds with bounds[k] = d + dp for fresh ciphertexts. */
int hm_result_slot_words(const hm_context *ctx, int op, uint32_t L, const uint32_t *a_words, const uint32_t *b_words,
                         uint32_t *out_words);
/* End-to-end convenience = upload a, upload b, apply2, download, with the copies overlapped with compute in chunks of
 * values (three streams, two stages).  a/b/out are HOST buffers in the padded layout; the result layout is
 * hm_result_slot_bounds of the operand bounds.  The *_bounded form takes per-slot degree bounds like
 * hm_batch_upload_bounded — state d + dp for fresh ciphertexts and the fused circuit kernels apply; the plain form
 * takes widths and assumes the widest degrees (always safe, generic kernels).  Both check every CUDA call and release
 * their streams, events and stage buffers on every exit path. */
int hm_apply2_host_bounded(hm_context *ctx, int op, size_t n, uint32_t L, const uint64_t *a_bounds, const uint64_t *a_host,
                           const uint64_t *b_bounds, const uint64_t *b_host, uint64_t *out_host);
int hm_apply2_host(hm_context *ctx, int op, size_t n, uint32_t L, const uint32_t *a_words, const uint64_t *a_host,
                   const uint32_t *b_words, const uint64_t *b_host, uint64_t *out_host);

/* ---- Device groups: one logical batch over several GPUs ------------------------------------------------------
 * Ciphertexts are independent (src/cipher.rs:180-185, :227-237; the circuits of common.rs touch only their two operands),
 * so a batch shards by value index with NO exchange step: hm_group_create replicates the parameters on `n_dev` devices
 * (device_ids may repeat a device: two contexts on one GPU — useful for tests), hm_group_set_*_key replicate the keys and
 * their derived tables, hm_group_encrypt / _apply2 / _apply1 / _decrypt split [0, n) into contiguous ranges (hm_shard_range:
 * sizes differ by at most one, rank order = index order), enqueue every device's share from the calling thread and only then
 * synchronise, and hm_group_decrypt gathers the plaintexts in index order.  Results are word for word those of one device.
 * The per-device contexts and batch parts are reachable (hm_group_context, hm_group_batch_part) for anything not wrapped. */
int hm_shard_range(size_t n, int rank, int world, size_t *first, size_t *count);
int hm_group_create(uint16_t d, uint16_t dp, uint16_t delta, uint16_t tau, const int *device_ids, int n_dev, hm_group **out);
void hm_group_destroy(hm_group *g);
int hm_group_size(const hm_group *g);
hm_context *hm_group_context(const hm_group *g, int i);
int hm_group_set_secret_key(hm_group *g, const uint8_t *bytes, size_t len);
int hm_group_set_public_key(hm_group *g, const uint8_t *const *polys, const size_t *lens, size_t n_polys);
int hm_group_synchronize(hm_group *g);
int hm_group_encrypt(hm_group *g, const uint8_t *values, size_t n, uint32_t L, const uint8_t *masks, hm_group_batch **out);
int hm_group_encrypt_seeded(hm_group *g, const uint8_t *values, size_t n, uint32_t L, uint64_t seed, hm_group_batch **out);
int hm_group_apply2(hm_group *g, int op, const hm_group_batch *a, const hm_group_batch *b, hm_group_batch **out);
int hm_group_apply2_into(hm_group *g, int op, const hm_group_batch *a, const hm_group_batch *b, hm_group_batch *out);
int hm_group_apply1(hm_group *g, int op, hm_group_batch *a);
int hm_group_decrypt(hm_group *g, const hm_group_batch *b, uint8_t *values_out);
int hm_group_batch_download(hm_group *g, const hm_group_batch *b, uint64_t *host);
size_t hm_group_batch_len(const hm_group_batch *b);
uint32_t hm_group_batch_bits(const hm_group_batch *b);
hm_batch *hm_group_batch_part(const hm_group_batch *b, int i);
void hm_group_batch_free(hm_group_batch *b);

/* ---- Raw polynomial batches (L = 1) -------------------------------------------------------
 * Polynomial::add / mul / rem over batches — src/polynomial.rs:190-213, :252-310, :316-365.
 * rem is by the context's secret key S (the only divisor on the hot path, src/cipher.rs:120);
 * mulrem is the fused `(a*b) mod S` unit of BASELINE.json's second metric: the output is the remainder the reference's
 * rem(mul(a, b)) gives, bit for bit; internally both operands are reduced mod S first (the remainder is unique), so the
 * product is never wider than 2 d bits. */
int hm_poly_add(hm_context *ctx, const hm_batch *a, const hm_batch *b, hm_batch **out);
int hm_poly_mul(hm_context *ctx, const hm_batch *a, const hm_batch *b, hm_batch **out);
int hm_poly_rem(hm_context *ctx, const hm_batch *a, hm_batch **out);
int hm_poly_mulrem(hm_context *ctx, const hm_batch *a, const hm_batch *b, hm_batch **out);
/* Fused mul+rem into an existing result batch (fresh operands at a tuned parameter set only; else
 * HM_ERR_UNSUPPORTED). */
int hm_poly_mulrem_into(hm_context *ctx, const hm_batch *a, const hm_batch *b, hm_batch *out);

/* ---- Measurement support ---------------------------------------------------------------------
 * Runs a LOP3 issue-rate probe on the context's device and reports 32-bit LOP3 lane-operations per
 * second (x32 = bit-MACs/s: one `r ^= a & m` is 32 AND-XOR pairs) and the SM clock seen meanwhile.
 * This is the measured denominator of the integer-logic roofline of the multiply kernels. */
int hm_measure_alu_peak(hm_context *ctx, double *lop3_lane_ops_per_s, double *sm_clock_mhz);
/* Rate of the 8x8-word (256 x 256 bit) Karatsuba carry-less product in isolation, products per second on the whole
 * device: the measured ceiling of the kernels that are chains of such products (thread-per-value adder: 2 881 per
 * u32 add; fused mul+rem at d=d'=128: 1 per pair, plus the fold). */
int hm_measure_kara8_peak(hm_context *ctx, double *products_per_s);

/* Hardware pipe peaks for the integer rooflines, measured on the context's device.  Three probes, each >= min_ms long
 * (use >= 100 so that the SM clock has ramped; fastest of three launches), 8 independent chains per thread, every CTA resident:
 *   0: IMAD.WIDE.U32 with both operands in per-thread registers (the 32x32->64 products of the carry-less leaf),
 *   1: LOP3 on three registers,   2: the kernels' own mix, 1 IMAD.WIDE : 2 LOP3.
 * out12[4 i ..] = { warp-instructions per second (whole device, CUDA-event time), duration in ms, resident warps per SM,
 * clock64() ticks per microsecond (diagnostic only: on B200 that counter runs below the SM clock, sample nvidia-smi) }. */
int hm_measure_pipe_peaks(hm_context *ctx, double min_ms, double *out12);

/* ---- Host-side helpers (no GPU needed) ---------------------------------------------------- */
/* Worst-case slot widths of a freshly encrypted value: (d+dp)/64+1 words per slot. */
uint32_t hm_fresh_slot_words(const hm_context *ctx);
/* v[k] = (X^k mod S)(0), k < nbits, as LSB-first u64 words: decryption is parity(C AND v),
 * identical to rem + evaluate(false) (src/cipher.rs:119-122).  Needs the secret key. */
int hm_decrypt_vector(const hm_context *ctx, size_t nbits, uint64_t *v_out);
/* Canonical form of one slot: degree (highest set bit, 0 for the null polynomial,
 * src/polynomial.rs:35-42) of a zero-padded word run. */
size_t hm_poly_degree(const uint64_t *words, size_t n_words);

#ifdef __cplusplus
}
#endif
#endif /* HMGPU_H */
