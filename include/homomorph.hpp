// homomorph.hpp — C++17 host-side mirror of the reference crate's public surface for the GF(2)[X] hot path, over
// the C ABI of libhmgpu.so (include/hmgpu.h).  Header only; link with -lhmgpu.
//
// The reference (mathisbot/homomorph-rust, crate `homomorph` v1.1.0) is Rust and cannot be built in this image, so
// this is what a user of the reference would switch to: same names, argument meaning and error behaviour —
//   Parameters::new                       src/context.rs:87-94      (panics -> std::invalid_argument)
//   SecretKey / PublicKey                 src/context.rs:122-298    (same byte formats, src/polynomial.rs:99-122)
//   Context::{new, generate_*_key, get_*/set_*_key, encrypt, decrypt, apply1, apply2}   src/context.rs:301-596
//   ContextCryptoError, CipherError, OperationError   src/context.rs:41-52, src/cipher.rs:17-24, src/operations.rs:11-18
//   Homomorphic{And,Or,Xor,Not}Gate, HomomorphicAddition, HomomorphicMultiplication + MIN_D_OVER_DELTA
//                                         src/impls/numbers.rs:7-50
// with one deliberate difference: everything is batched.  Ciphered<T> holds n values of T (n x 8*sizeof(T)
// bit-ciphertexts resident in HBM) and the operations act on whole batches.  All polynomial arithmetic on
// ciphertexts runs on the GPU; the only host arithmetic is key generation (tau+1 small polynomials once per context).
#pragma once
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <optional>
#include <random>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "hmgpu.h"

namespace homomorph {

// ---- errors ---------------------------------------------------------------------------------------------
struct ContextCryptoError : std::runtime_error { // src/context.rs:41-52
    enum Kind { PublicKeyUnset, SecretKeyUnset } kind;
    explicit ContextCryptoError(Kind k) : std::runtime_error(k == PublicKeyUnset ? "PublicKeyUnset" : "SecretKeyUnset"), kind(k) {}
};
struct CipherError : std::runtime_error { // src/cipher.rs:17-24
    explicit CipherError(const std::string &m) : std::runtime_error(m) {}
};
struct OperationError : std::runtime_error { // src/operations.rs:11-18 (InvalidParameters)
    uint16_t required_min_d_over_delta, actual_d, actual_delta;
    OperationError(uint16_t req, uint16_t d, uint16_t delta)
        : std::runtime_error("invalid parameters: d/delta must be at least " + std::to_string(req)),
          required_min_d_over_delta(req), actual_d(d), actual_delta(delta) {}
};
struct EngineError : std::runtime_error { // CUDA failure / unsupported shape / bad argument from libhmgpu.so
    int status;
    EngineError(int s, const std::string &detail)
        : std::runtime_error(std::string(hm_status_string(s)) + (detail.empty() ? "" : ": " + detail)), status(s) {}
};

// ---- operation markers (src/impls/numbers.rs:7-50) ----------------------------------------------------------
struct HomomorphicAndGate { static constexpr int code = HM_OP_AND; static constexpr uint16_t MIN_D_OVER_DELTA = 2; };
struct HomomorphicOrGate { static constexpr int code = HM_OP_OR; static constexpr uint16_t MIN_D_OVER_DELTA = 2; };
struct HomomorphicXorGate { static constexpr int code = HM_OP_XOR; static constexpr uint16_t MIN_D_OVER_DELTA = 1; };
struct HomomorphicNotGate { static constexpr int code = HM_OP_NOT; static constexpr uint16_t MIN_D_OVER_DELTA = 1; };
struct HomomorphicAddition { static constexpr int code = HM_OP_ADD; static constexpr uint16_t MIN_D_OVER_DELTA = 21; };
struct HomomorphicMultiplication { static constexpr int code = HM_OP_MUL; static constexpr uint16_t MIN_D_OVER_DELTA = 64; };

// ---- randomness: the reference calls getrandom::fill (src/polynomial.rs:87, src/cipher.rs:95); here a byte source ---
using ByteSource = std::function<void(uint8_t *, size_t)>;
inline ByteSource os_random() {
    return [](uint8_t *p, size_t n) {
        std::random_device rd;
        for (size_t i = 0; i < n; ++i) p[i] = (uint8_t)rd();
    };
}
inline ByteSource seeded_random(uint64_t seed) { // reproducible runs (std::mt19937_64, one byte per draw)
    auto eng = std::make_shared<std::mt19937_64>(seed);
    return [eng](uint8_t *p, size_t n) {
        for (size_t i = 0; i < n; ++i) p[i] = (uint8_t)((*eng)() >> 24);
    };
}

namespace detail {
using words = std::vector<uint64_t>; // coefficient of X^i = bit i%64 of word i/64 (src/polynomial.rs:144,172)
inline size_t degree(const words &w) {
    for (size_t i = w.size(); i-- > 0;)
        if (w[i]) return 64 * i + 63 - (size_t)__builtin_clzll(w[i]);
    return 0;
}
// Polynomial::random — src/polynomial.rs:73-96: fill the words, clear above the degree, force the leading coefficient
inline words random_poly(size_t deg, const ByteSource &rnd) {
    const size_t n = deg / 64 + 1;
    std::vector<uint8_t> bytes(8 * n);
    rnd(bytes.data(), bytes.size());
    words w(n);
    for (size_t i = 0; i < n; ++i) {
        uint64_t x = 0;
        for (int b = 0; b < 8; ++b) x |= (uint64_t)bytes[8 * i + b] << (8 * b);
        w[i] = x;
    }
    w[n - 1] &= ((uint64_t)1 << (deg % 64)) - 1;
    w[n - 1] |= (uint64_t)1 << (deg % 64);
    return w;
}
// key generation only: tau products of a (d+1)-bit by a (d'+1)-bit polynomial (src/context.rs:249-261)
inline words clmul(const words &a, const words &b) {
    words r(a.size() + b.size(), 0);
    for (size_t i = 0; i < a.size(); ++i)
        for (int bit = 0; bit < 64; ++bit)
            if ((a[i] >> bit) & 1)
                for (size_t j = 0; j < b.size(); ++j) {
                    r[i + j] ^= b[j] << bit;
                    if (bit) r[i + j + 1] ^= b[j] >> (64 - bit);
                }
    r.resize(degree(r) / 64 + 1);
    return r;
}
inline std::vector<uint8_t> to_bytes(const words &w) { // Polynomial::to_bytes — src/polynomial.rs:99-105
    std::vector<uint8_t> out(8 * w.size());
    for (size_t i = 0; i < w.size(); ++i)
        for (int b = 0; b < 8; ++b) out[8 * i + b] = (uint8_t)(w[i] >> (8 * b));
    return out;
}
inline words from_bytes(const std::vector<uint8_t> &bytes) { // Polynomial::from_bytes — src/polynomial.rs:108-122
    if (bytes.empty()) throw std::invalid_argument("The vector of bytes must not be empty.");
    words w((bytes.size() + 7) / 8, 0);
    for (size_t i = 0; i < bytes.size(); ++i) w[i / 8] |= (uint64_t)bytes[i] << (8 * (i % 8));
    return w;
}
} // namespace detail

// ---- Parameters (src/context.rs:33-119) ---------------------------------------------------------------------
class Parameters {
    uint16_t d_, dp_, delta_, tau_;

  public:
    Parameters(uint16_t d, uint16_t dp, uint16_t delta, uint16_t tau) : d_(d), dp_(dp), delta_(delta), tau_(tau) {
        if (d == 0 || dp == 0 || delta == 0 || tau == 0) throw std::invalid_argument("Parameters must be strictly positive");
        if (!(delta < d)) throw std::invalid_argument("Delta must be strictly less than d");
    }
    static Parameters create(uint16_t d, uint16_t dp, uint16_t delta, uint16_t tau) { return Parameters(d, dp, delta, tau); } // `new`
    uint16_t d() const { return d_; }
    uint16_t dp() const { return dp_; }
    uint16_t delta() const { return delta_; }
    uint16_t tau() const { return tau_; }
};

// ---- keys (src/context.rs:122-298) ----------------------------------------------------------------------------
class SecretKey {
    detail::words p_;
    friend class PublicKey;

  public:
    explicit SecretKey(detail::words p) : p_(std::move(p)) {}
    ~SecretKey() { // zeroize on drop, src/context.rs:199-206
        volatile uint64_t *q = p_.data();
        for (size_t i = 0; i < p_.size(); ++i) q[i] = 0;
    }
    SecretKey(const SecretKey &) = default;
    SecretKey &operator=(const SecretKey &) = default;
    static SecretKey from_bytes(const std::vector<uint8_t> &b) { return SecretKey(detail::from_bytes(b)); }
    static SecretKey random(uint16_t d, const ByteSource &rnd) { return SecretKey(detail::random_poly(d, rnd)); } // :160-162
    std::vector<uint8_t> to_bytes() const { return detail::to_bytes(p_); }
};

class PublicKey {
    std::vector<detail::words> t_;

  public:
    explicit PublicKey(std::vector<detail::words> t) : t_(std::move(t)) {}
    static PublicKey from_bytes(const std::vector<std::vector<uint8_t>> &rows) {
        std::vector<detail::words> t;
        for (const auto &r : rows) t.push_back(detail::from_bytes(r));
        return PublicKey(std::move(t));
    }
    // T_i = S*Q_i + X*R_i; per i the bytes of Q are drawn before those of R — src/context.rs:249-261
    static PublicKey random(uint16_t dp, uint16_t delta, uint16_t tau, const SecretKey &sk, const ByteSource &rnd) {
        std::vector<detail::words> t;
        for (uint16_t i = 0; i < tau; ++i) {
            detail::words q = detail::random_poly(dp, rnd);
            detail::words sq = detail::clmul(sk.p_, q);
            detail::words r = detail::random_poly(delta, rnd);
            detail::words xr(r.size() + 1, 0); // R * X
            for (size_t k = 0; k < r.size(); ++k) {
                xr[k] |= r[k] << 1;
                xr[k + 1] |= r[k] >> 63;
            }
            if (xr.size() > sq.size()) sq.resize(xr.size(), 0);
            for (size_t k = 0; k < xr.size(); ++k) sq[k] ^= xr[k];
            sq.resize(detail::degree(sq) / 64 + 1);
            t.push_back(std::move(sq));
        }
        return PublicKey(std::move(t));
    }
    std::vector<std::vector<uint8_t>> to_bytes() const {
        std::vector<std::vector<uint8_t>> out;
        for (const auto &p : t_) out.push_back(detail::to_bytes(p));
        return out;
    }
    size_t size() const { return t_.size(); }
};

class Context;

// ---- Ciphered<T>: a batch of the reference's Ciphered<T> (src/cipher.rs:126-259) ---------------------------------
template <class T> class Ciphered {
    // integers, or packed structs of integers (bincode writes the fields in order, each fixint little endian: a struct
    // without padding has the same bytes) — examples/simple_struct.rs:12-17
    static_assert(std::is_integral<T>::value || (std::is_trivially_copyable<T>::value && std::has_unique_object_representations<T>::value),
                  "bincode fixint little endian is reproduced for integers and padding-free structs of integers only");
    hm_context *ctx_ = nullptr;
    hm_batch *b_ = nullptr;
    friend class Context;
    Ciphered(hm_context *c, hm_batch *b) : ctx_(c), b_(b) {}

  public:
    Ciphered() = default;
    Ciphered(const Ciphered &) = delete;
    Ciphered &operator=(const Ciphered &) = delete;
    Ciphered(Ciphered &&o) noexcept : ctx_(o.ctx_), b_(o.b_) { o.b_ = nullptr; }
    Ciphered &operator=(Ciphered &&o) noexcept {
        if (this != &o) {
            reset();
            ctx_ = o.ctx_;
            b_ = o.b_;
            o.b_ = nullptr;
        }
        return *this;
    }
    ~Ciphered() { reset(); }
    void reset() {
        if (b_) hm_batch_free(ctx_, b_);
        b_ = nullptr;
    }
    size_t size() const { return hm_batch_len(b_); }       // number of values
    uint32_t bits() const { return hm_batch_bits(b_); }    // Ciphered::len() of each value == T::BITS (cipher.rs:286)
    size_t value_words() const { return hm_batch_value_words(b_); }
    const hm_batch *raw() const { return b_; }
    // the coefficient words of every polynomial, padded layout of include/hmgpu.h
    std::vector<uint64_t> to_host() const {
        std::vector<uint64_t> out(size() * value_words());
        int rc = hm_batch_download(ctx_, b_, out.data());
        if (rc != HM_OK) throw EngineError(rc, hm_last_error(ctx_));
        return out;
    }
};

// ---- Context (src/context.rs:301-596) ---------------------------------------------------------------------------
class Context {
    Parameters params_;
    hm_context *h_ = nullptr;
    std::optional<SecretKey> sk_;
    std::optional<PublicKey> pk_;

    void check(int rc) const {
        if (rc == HM_OK) return;
        if (rc == HM_ERR_PUBLIC_KEY_UNSET) throw ContextCryptoError(ContextCryptoError::PublicKeyUnset);
        if (rc == HM_ERR_SECRET_KEY_UNSET) throw ContextCryptoError(ContextCryptoError::SecretKeyUnset);
        if (rc == HM_ERR_INVALID_LENGTH) throw CipherError("InvalidCipheredLength");
        if (rc == HM_ERR_DIVIDE_BY_ZERO) throw std::domain_error("attempt to divide by zero");
        throw EngineError(rc, h_ ? hm_last_error(h_) : "");
    }
    template <class O> void validate() const { // src/context.rs:310-323
        if ((uint32_t)params_.d() < (uint32_t)O::MIN_D_OVER_DELTA * params_.delta())
            throw OperationError(O::MIN_D_OVER_DELTA, params_.d(), params_.delta());
    }

  public:
    explicit Context(const Parameters &p, int device = 0) : params_(p) {
        int rc = hm_context_create(p.d(), p.dp(), p.delta(), p.tau(), device, &h_);
        if (rc != HM_OK) throw EngineError(rc, rc == HM_ERR_CUDA ? "no usable CUDA device (the engine has no CPU fallback)" : "");
    }
    ~Context() { hm_context_destroy(h_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;

    const Parameters &parameters() const { return params_; }
    void generate_secret_key(const ByteSource &rnd = os_random()) { set_secret_key(SecretKey::random(params_.d(), rnd)); } // :421-425
    void generate_public_key(const ByteSource &rnd = os_random()) {                                                        // :440-454
        if (!sk_) throw ContextCryptoError(ContextCryptoError::SecretKeyUnset);
        set_public_key(PublicKey::random(params_.dp(), params_.delta(), params_.tau(), *sk_, rnd));
    }
    const std::optional<SecretKey> &get_secret_key() const { return sk_; }
    const std::optional<PublicKey> &get_public_key() const { return pk_; }
    void set_secret_key(const SecretKey &sk) { // also clears the public key, src/context.rs:568-571
        const auto bytes = sk.to_bytes();
        check(hm_set_secret_key(h_, bytes.data(), bytes.size()));
        sk_ = sk;
        pk_.reset();
    }
    void set_public_key(const PublicKey &pk) { // src/context.rs:592-595
        const auto rows = pk.to_bytes();
        std::vector<const uint8_t *> ptrs;
        std::vector<size_t> lens;
        for (const auto &r : rows) {
            ptrs.push_back(r.data());
            lens.push_back(r.size());
        }
        check(hm_set_public_key(h_, ptrs.data(), lens.data(), rows.size()));
        pk_ = pk;
    }
    size_t mask_bytes() const { return ((size_t)params_.tau() + 7) / 8; }

    // Context::encrypt (src/context.rs:463-471) for a whole vector.  `masks`: n * bits * ceil(tau/8) bytes, the subset U
    // of every bit (value-major, bit-minor), replacing CipheredBit::part's getrandom call (src/cipher.rs:92-97).
    template <class T> Ciphered<T> encrypt(const std::vector<T> &values, const std::vector<uint8_t> &masks) const {
        if (!pk_) throw ContextCryptoError(ContextCryptoError::PublicKeyUnset);
        const uint32_t L = 8 * sizeof(T);
        if (masks.size() != values.size() * L * mask_bytes()) throw std::invalid_argument("masks must hold n * bits * ceil(tau/8) bytes");
        hm_batch *b = nullptr;
        check(hm_encrypt(h_, reinterpret_cast<const uint8_t *>(values.data()), values.size(), L, masks.data(), &b)); // LE host
        return Ciphered<T>(h_, b);
    }
    // same with masks drawn on the device from a seed (Philox4x32-10; hm_masks_generate_host gives the same stream)
    template <class T> Ciphered<T> encrypt(const std::vector<T> &values, uint64_t seed) const {
        if (!pk_) throw ContextCryptoError(ContextCryptoError::PublicKeyUnset);
        hm_batch *b = nullptr;
        check(hm_encrypt_seeded(h_, reinterpret_cast<const uint8_t *>(values.data()), values.size(), 8 * sizeof(T), seed, &b));
        return Ciphered<T>(h_, b);
    }
    // Context::decrypt (src/context.rs:480-488)
    template <class T> std::vector<T> decrypt(const Ciphered<T> &c) const {
        if (!sk_) throw ContextCryptoError(ContextCryptoError::SecretKeyUnset);
        std::vector<T> out(c.size());
        check(hm_decrypt(h_, c.raw(), reinterpret_cast<uint8_t *>(out.data())));
        return out;
    }
    // Context::apply2 (src/context.rs:515-527)
    template <class O, class T> Ciphered<T> apply2(const Ciphered<T> &a, const Ciphered<T> &b) const {
        validate<O>();
        hm_batch *o = nullptr;
        check(hm_apply2(h_, O::code, a.raw(), b.raw(), &o));
        return Ciphered<T>(h_, o);
    }
    // User structs (examples/simple_struct.rs:32-58).  slice: the bits [first_bit, first_bit + 8 sizeof(U)) of every value as
    // a Ciphered<U> (split_at + new_from_raw on one field); concat: the extend_from_slice merge into a Ciphered<T>;
    // apply2_fields: O on every field of a and b, field f being field_bits[f] consecutive bits.
    template <class U, class T> Ciphered<U> slice(const Ciphered<T> &c, uint32_t first_bit) const {
        hm_batch *o = nullptr;
        check(hm_batch_slice(h_, c.raw(), first_bit, 8 * sizeof(U), &o));
        return Ciphered<U>(h_, o);
    }
    template <class T> Ciphered<T> concat(const std::vector<const hm_batch *> &parts) const {
        hm_batch *o = nullptr;
        check(hm_batch_concat(h_, parts.data(), parts.size(), &o));
        if (hm_batch_bits(o) != 8 * sizeof(T)) {
            hm_batch_free(h_, o);
            throw CipherError("InvalidCipheredLength");
        }
        return Ciphered<T>(h_, o);
    }
    template <class O, class T> Ciphered<T> apply2_fields(const Ciphered<T> &a, const Ciphered<T> &b, const std::vector<uint32_t> &field_bits) const {
        validate<O>();
        hm_batch *o = nullptr;
        check(hm_apply2_fields(h_, O::code, a.raw(), b.raw(), field_bits.data(), field_bits.size(), &o));
        return Ciphered<T>(h_, o);
    }
    // Context::apply1 (src/context.rs:496-507): in place
    template <class O, class T> void apply1(Ciphered<T> &a) const {
        validate<O>();
        check(hm_apply1(h_, O::code, const_cast<hm_batch *>(a.raw())));
    }
    hm_context *raw() const { return h_; }
};

} // namespace homomorph
