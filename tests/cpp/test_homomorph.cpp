// The reference's own tests, restated against the C++ mirror (include/homomorph.hpp) on batches.
//   --cpu : what needs no GPU (src/context.rs:602-635 parameter panics, key byte round trips; no-CPU-fallback check)
//   --gpu : src/context.rs:638-677, src/cipher.rs:276-304, src/impls/numbers/uint.rs:109-293
#include <cstdio>
#include <cstring>
#include <iostream>

#include "homomorph.hpp"

using namespace homomorph;

static int failures = 0;
#define CHECK(cond)                                                                  \
    do {                                                                             \
        if (!(cond)) {                                                               \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);              \
            ++failures;                                                              \
        }                                                                            \
    } while (0)
template <class E, class F> static bool throws(F f) {
    try {
        f();
    } catch (const E &) {
        return true;
    } catch (...) {
    }
    return false;
}
static std::vector<uint8_t> masks(const Context &c, size_t n, size_t bits, uint64_t seed) {
    std::vector<uint8_t> m(n * bits * c.mask_bytes());
    seeded_random(seed)(m.data(), m.size());
    return m;
}

static void cpu_tests() {
    // context.rs:602-613 — Parameters::new panics
    CHECK(throws<std::invalid_argument>([] { Parameters(0, 1, 1, 1); }));
    CHECK(throws<std::invalid_argument>([] { Parameters(8, 0, 1, 1); }));
    CHECK(throws<std::invalid_argument>([] { Parameters(8, 1, 0, 1); }));
    CHECK(throws<std::invalid_argument>([] { Parameters(8, 1, 1, 0); }));
    CHECK(throws<std::invalid_argument>([] { Parameters(8, 4, 8, 4); }));
    Parameters p(6, 3, 2, 5);
    CHECK(p.d() == 6 && p.dp() == 3 && p.delta() == 2 && p.tau() == 5);
    // context.rs:616-635 — key byte round trips; random keys have the exact degree (polynomial.rs:490-496)
    auto rnd = seeded_random(42);
    SecretKey sk = SecretKey::random(128, rnd);
    CHECK(sk.to_bytes().size() == 24 && sk.to_bytes()[16] == 1);
    CHECK(SecretKey::from_bytes(sk.to_bytes()).to_bytes() == sk.to_bytes());
    PublicKey pk = PublicKey::random(128, 1, 16, sk, rnd);
    CHECK(pk.size() == 16);
    for (const auto &row : pk.to_bytes()) CHECK(row.size() == 40 && row[32] == 1); // exact degree d+d' = 256
    CHECK(PublicKey::from_bytes(pk.to_bytes()).to_bytes() == pk.to_bytes());
    CHECK(throws<std::invalid_argument>([] { SecretKey::from_bytes({}); }));
    CHECK(HomomorphicAddition::MIN_D_OVER_DELTA == 21 && HomomorphicMultiplication::MIN_D_OVER_DELTA == 64);
    CHECK(hm_op_min_d_over_delta(HomomorphicAddition::code) == 21);
    if (hm_device_count() == 0) // no CPU fallback: a Context cannot exist without a CUDA device
        CHECK(throws<EngineError>([] { Context c(Parameters(64, 32, 8, 32)); }));
}

static void gpu_tests() {
    {   // context.rs:638-677
        Context ctx(Parameters(64, 32, 8, 32));
        CHECK(!ctx.get_secret_key() && !ctx.get_public_key());
        CHECK(throws<ContextCryptoError>([&] { ctx.generate_public_key(seeded_random(1)); })); // SecretKeyUnset
        ctx.generate_secret_key(seeded_random(1));
        ctx.generate_public_key(seeded_random(2));
        CHECK(ctx.get_secret_key() && ctx.get_public_key());
        ctx.set_secret_key(*ctx.get_secret_key());
        CHECK(!ctx.get_public_key()); // set_secret_key clears the public key
        CHECK(throws<ContextCryptoError>([&] { ctx.encrypt(std::vector<uint8_t>{1}, masks(ctx, 1, 8, 3)); }));
        ctx.generate_public_key(seeded_random(2));
        // cipher.rs:276-304 — round trips, len == T::BITS
        std::vector<uint8_t> a8 = {0b10101010, 0, 255};
        auto c8 = ctx.encrypt(a8, masks(ctx, a8.size(), 8, 4));
        CHECK(c8.bits() == 8 && c8.size() == 3);
        CHECK(ctx.decrypt(c8) == a8);
        std::vector<uint64_t> a64 = {0x0123456789ABCDEFull, 42};
        auto c64 = ctx.encrypt(a64, masks(ctx, a64.size(), 64, 5));
        CHECK(c64.bits() == 64 && ctx.decrypt(c64) == a64);
        auto cs = ctx.encrypt(a64, (uint64_t)77); // device-side masks
        CHECK(ctx.decrypt(cs) == a64);
    }
    {   // uint.rs:109-174 — gates (delta = 1 so that products of fresh bits decrypt correctly)
        Context ctx(Parameters(32, 8, 1, 8));
        ctx.generate_secret_key(seeded_random(11));
        ctx.generate_public_key(seeded_random(12));
        std::vector<uint8_t> a = {0b1010}, b = {0b1100};
        auto ca = ctx.encrypt(a, masks(ctx, 1, 8, 6)), cb = ctx.encrypt(b, masks(ctx, 1, 8, 7));
        CHECK((ctx.decrypt(ctx.apply2<HomomorphicAndGate>(ca, cb))[0] == 0b1000));
        CHECK((ctx.decrypt(ctx.apply2<HomomorphicOrGate>(ca, cb))[0] == 0b1110));
        CHECK((ctx.decrypt(ctx.apply2<HomomorphicXorGate>(ca, cb))[0] == 0b0110));
        ctx.apply1<HomomorphicNotGate>(ca);
        CHECK(ctx.decrypt(ca)[0] == 0b11110101);
    }
    {   // uint.rs:176-208 — addition at (64,16,1,16)
        Context ctx(Parameters(64, 16, 1, 16));
        ctx.generate_secret_key(seeded_random(21));
        ctx.generate_public_key(seeded_random(22));
        std::vector<uint8_t> a = {22, 255}, b = {20, 240};
        auto s = ctx.apply2<HomomorphicAddition>(ctx.encrypt(a, masks(ctx, 2, 8, 8)), ctx.encrypt(b, masks(ctx, 2, 8, 9)));
        auto d = ctx.decrypt(s);
        CHECK(d[0] == 42 && d[1] == 239); // wrapping overflow
        std::vector<uint16_t> x = {12345, 999}, y = {4321, 1};
        auto s16 = ctx.apply2<HomomorphicAddition>(ctx.encrypt(x, masks(ctx, 2, 16, 10)), ctx.encrypt(y, masks(ctx, 2, 16, 11)));
        auto d16 = ctx.decrypt(s16);
        CHECK(d16[0] == 16666 && d16[1] == 1000);
    }
    {   // fused u32 adder at the benchmark parameters (benches/u32.rs:26-50, README.md:65-69)
        Context ctx(Parameters(128, 128, 1, 128));
        ctx.generate_secret_key(seeded_random(31));
        ctx.generate_public_key(seeded_random(32));
        std::vector<uint32_t> a = {0xDEADBEEFu, 22, 0xFFFFFFFFu}, b = {0x12345678u, 20, 1};
        auto s = ctx.apply2<HomomorphicAddition>(ctx.encrypt(a, masks(ctx, 3, 32, 12)), ctx.encrypt(b, masks(ctx, 3, 32, 13)));
        CHECK(s.value_words() == 5864);
        auto d = ctx.decrypt(s);
        CHECK(d[0] == 0xDEADBEEFu + 0x12345678u && d[1] == 42 && d[2] == 0);
    }
    {   // uint.rs:254-293 — multiplication on u8 at (128,64,1,64)
        Context ctx(Parameters(128, 64, 1, 64));
        ctx.generate_secret_key(seeded_random(41));
        ctx.generate_public_key(seeded_random(42));
        std::vector<uint8_t> a = {6, 0, 255}, b = {7, 151, 240};
        auto p = ctx.apply2<HomomorphicMultiplication>(ctx.encrypt(a, masks(ctx, 3, 8, 14)), ctx.encrypt(b, masks(ctx, 3, 8, 15)));
        auto d = ctx.decrypt(p);
        CHECK(d[0] == 42 && d[1] == 0 && d[2] == 16);
    }
    {   // examples/simple_struct.rs — Vec3 { x, y, z: u16 } added field by field at (64,32,1,32): {1,2,3} + {4,5,6} = {5,7,9}
        struct Vec3 {
            uint16_t x, y, z;
            bool operator==(const Vec3 &o) const { return x == o.x && y == o.y && z == o.z; }
        };
        Context ctx(Parameters(64, 32, 1, 32));
        ctx.generate_secret_key(seeded_random(61));
        ctx.generate_public_key(seeded_random(62));
        std::vector<Vec3> a = {{1, 2, 3}, {65535, 1000, 7}}, b = {{4, 5, 6}, {1, 2345, 65530}};
        auto ca = ctx.encrypt(a, masks(ctx, 2, 48, 17)), cb = ctx.encrypt(b, masks(ctx, 2, 48, 18));
        CHECK(ca.bits() == 48);
        auto c = ctx.apply2_fields<HomomorphicAddition>(ca, cb, {16, 16, 16});
        auto d = ctx.decrypt(c);
        CHECK((d[0] == Vec3{5, 7, 9}) && (d[1] == Vec3{0, 3345, 1}));
        // the same by hand: split_at / new_from_raw / extend_from_slice
        auto ay = ctx.slice<uint16_t>(ca, 16), by = ctx.slice<uint16_t>(cb, 16);
        CHECK(ctx.decrypt(ay) == (std::vector<uint16_t>{2, 1000}));
        auto x = ctx.apply2<HomomorphicAddition>(ctx.slice<uint16_t>(ca, 0), ctx.slice<uint16_t>(cb, 0));
        auto y = ctx.apply2<HomomorphicAddition>(ay, by);
        auto z = ctx.apply2<HomomorphicAddition>(ctx.slice<uint16_t>(ca, 32), ctx.slice<uint16_t>(cb, 32));
        auto m = ctx.concat<Vec3>({x.raw(), y.raw(), z.raw()});
        CHECK(m.to_host() == c.to_host());
        CHECK(throws<CipherError>([&] { ctx.slice<uint16_t>(ca, 40); }));
    }
    {   // context.rs:310-323 — requirement check: d/delta = 16 < 21
        Context ctx(Parameters(64, 16, 4, 16));
        ctx.generate_secret_key(seeded_random(51));
        ctx.generate_public_key(seeded_random(52));
        auto c = ctx.encrypt(std::vector<uint8_t>{1}, masks(ctx, 1, 8, 16));
        bool ok = false;
        try {
            ctx.apply2<HomomorphicAddition>(c, c);
        } catch (const OperationError &e) {
            ok = e.required_min_d_over_delta == 21 && e.actual_d == 64 && e.actual_delta == 4;
        }
        CHECK(ok);
        CHECK(throws<ContextCryptoError>([&] { // decrypt without the secret key (fresh context)
            Context c2(Parameters(64, 16, 4, 16));
            c2.decrypt(c);
        }));
    }
}

int main(int argc, char **argv) {
    const bool gpu = argc > 1 && std::strcmp(argv[1], "--gpu") == 0;
    cpu_tests();
    if (gpu) {
        if (hm_device_count() == 0) {
            std::printf("FAIL no CUDA device\n");
            return 2;
        }
        gpu_tests();
    }
    std::printf("%s: %d failure(s)\n", gpu ? "gpu" : "cpu", failures);
    return failures ? 1 : 0;
}
