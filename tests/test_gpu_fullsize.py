"""BASELINE.json's full-size configurations through size-independent properties (plus oracle spot checks):
  config 2: 2^20 u32 encrypt -> decrypt round trip;      config 3: u32 add on 2^18 encrypted pairs;
  config 4: u8 multiply on 2^14 pairs (SURVEY.md finding 4: L=8 is the feasible width);
  config 5: d=d'=512, tau=256, delta=8 mul+rem sweep 2^10 .. 2^22.
"""
import numpy as np
import pytest

from helpers import engine_context, expected_padded, keys, oracle_encrypt

pytestmark = pytest.mark.gpu

CONFIG_A = (128, 128, 1, 128)
CONFIG_B = (512, 512, 8, 256)


@pytest.fixture(scope="module")
def hm():
    import homomorph_rust_b200 as h

    assert h.lib().hm_device_count() > 0
    return h


def rbytes(seed, n):
    return np.frombuffer(np.random.default_rng(seed).bytes(n), dtype=np.uint8)


def test_config2_roundtrip_and_linearity(oracle, hm):
    sk, pk, skb, pkb = keys(oracle, *CONFIG_A, 101)
    ctx = engine_context(hm, *CONFIG_A, skb, pkb)
    n = 1 << 20
    rng = np.random.default_rng(1)
    a = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    ma = rbytes(2, n * 32 * 16)
    ca = ctx.encrypt(a, ma)
    assert len(ca) == n and ca.bits == 32 and list(ca.slot_words()) == [5] * 32
    np.testing.assert_array_equal(ctx.decrypt(ca), a)  # encode -> decode round trip, src/cipher.rs:276-304
    # oracle spot check on a strided sample of values
    idx = np.arange(0, n, n // 64)
    want = oracle_encrypt(oracle, pk, a[idx], ma.reshape(n, 512)[idx].reshape(-1))
    host = ca.to_host()
    np.testing.assert_array_equal(host[idx], expected_padded(want, idx.size, [5] * 32))
    # linearity of the subset-XOR: enc(x, U) + enc(y, V) == enc(x ^ y, U ^ V) on a 2^18 slice
    m = 1 << 18
    b = rng.integers(0, 2**32, size=m, dtype=np.uint32)
    mb = rbytes(3, m * 32 * 16)
    c1 = ctx.encrypt(a[:m], ma[: m * 512])
    c2 = ctx.encrypt(b, mb)
    c3 = ctx.encrypt(a[:m] ^ b, ma[: m * 512] ^ mb)
    x = ctx.apply2(hm.HomomorphicXorGate, c1, c2)
    np.testing.assert_array_equal(x.to_host(), c3.to_host())
    np.testing.assert_array_equal(ctx.decrypt(x), a[:m] ^ b)


def test_config3_u32_add_full(oracle, hm):
    sk, pk, skb, pkb = keys(oracle, *CONFIG_A, 102)
    ctx = engine_context(hm, *CONFIG_A, skb, pkb)
    n = 1 << 18
    rng = np.random.default_rng(5)
    a = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    b = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    a[:4] = [0, 0xFFFFFFFF, 0xFFFFFFFF, 22]
    b[:4] = [0, 1, 0xFFFFFFFF, 20]
    ma, mb = rbytes(6, n * 512), rbytes(7, n * 512)
    ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
    s = ctx.apply2(hm.HomomorphicAddition, ca, cb)
    assert int(s.slot_words().sum()) == 5864  # 46 912 B per u32 sum
    dec = ctx.decrypt(s)
    # delta = 1: the decrypted sum is the plaintext sum for (practically) every value
    assert np.mean(dec == a + b) > 0.9999
    assert list(dec[:4]) == [0, 0, 0xFFFFFFFE, 42]
    # what the reference would decrypt, on a sample: oracle add + oracle decrypt
    idx = np.arange(0, n, n // 16)
    oa = oracle_encrypt(oracle, pk, a[idx], ma.reshape(n, 512)[idx].reshape(-1))
    ob = oracle_encrypt(oracle, pk, b[idx], mb.reshape(n, 512)[idx].reshape(-1))
    want, _ = oracle.apply(oracle.OP_ADD, oa, ob, 32, threads=oracle.max_threads())
    od, _ = oracle.decrypt(sk, want, 32)
    np.testing.assert_array_equal(dec[idx].view(np.uint8), od)
    # commutativity of the circuit's result polynomials is NOT expected (carry chain is symmetric in a, b, so it is):
    s2 = ctx.apply2(hm.HomomorphicAddition, cb, ca)
    np.testing.assert_array_equal(ctx.decrypt(s2), dec)
    del s2
    # fused kernel == generic reference-order path on a slice, bit for bit
    m = 2048
    sa, sb = ctx.encrypt(a[:m], ma[: m * 512]), ctx.encrypt(b[:m], mb[: m * 512])
    f = ctx.apply2(hm.HomomorphicAddition, sa, sb).to_host()
    g = ctx.apply2(hm.HomomorphicAddition, sa, sb, generic=True).to_host()
    np.testing.assert_array_equal(f, g)
    np.testing.assert_array_equal(f[:16], expected_padded(
        oracle.apply(oracle.OP_ADD, oracle_encrypt(oracle, pk, a[:16], ma[: 16 * 512]), oracle_encrypt(oracle, pk, b[:16], mb[: 16 * 512]), 32,
                     threads=oracle.max_threads())[0], 16, s.slot_words()))


def test_config4_u8_mul_batch(oracle, hm):
    sk, pk, skb, pkb = keys(oracle, *CONFIG_A, 103)
    ctx = engine_context(hm, *CONFIG_A, skb, pkb)
    n = 1 << 14
    rng = np.random.default_rng(9)
    a = rng.integers(0, 256, size=n, dtype=np.uint8)
    b = rng.integers(0, 256, size=n, dtype=np.uint8)
    ma, mb = rbytes(10, n * 8 * 16), rbytes(11, n * 8 * 16)
    ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
    p = ctx.apply2(hm.HomomorphicMultiplication, ca, cb)
    assert int(p.slot_words().sum()) == 528
    dec = ctx.decrypt(p)
    assert np.mean(dec == a * b) > 0.999
    idx = np.arange(0, n, n // 8)
    oa = oracle_encrypt(oracle, pk, a[idx], ma.reshape(n, 128)[idx].reshape(-1))
    ob = oracle_encrypt(oracle, pk, b[idx], mb.reshape(n, 128)[idx].reshape(-1))
    want, _ = oracle.apply(oracle.OP_MUL, oa, ob, 8, threads=oracle.max_threads())
    np.testing.assert_array_equal(p.to_host()[idx], expected_padded(want, idx.size, p.slot_words()))
    od, _ = oracle.decrypt(sk, want, 8)
    np.testing.assert_array_equal(dec[idx], od)
    # u32 multiplication is infeasible by construction (SURVEY.md finding 4) and must be refused, not attempted
    c32 = ctx.encrypt(np.array([3], dtype=np.uint32), rbytes(1, 512))
    with pytest.raises(hm.EngineError):
        ctx.apply2(hm.HomomorphicMultiplication, c32, c32)


@pytest.mark.parametrize("log2n", [10, 16, 22])
def test_config5_stress_mulrem_sweep(oracle, hm, log2n):
    """d=d'=512, tau=256, delta=8: carry-less mul + rem over 2^10 .. 2^22 fresh ciphertext pairs."""
    sk, pk, skb, pkb = keys(oracle, *CONFIG_B, 104)
    ctx = engine_context(hm, *CONFIG_B, skb, pkb)
    n = 1 << log2n  # pairs; each u8 value carries 8 pairs
    nv = n // 8
    rng = np.random.default_rng(log2n)
    va, vb, vc = (rng.integers(0, 256, size=nv, dtype=np.uint8) for _ in range(3))
    ma, mb, mc = rbytes(1, nv * 8 * 32), rbytes(2, nv * 8 * 32), rbytes(3, nv * 8 * 32)
    ca, cb, cc = ctx.encrypt(va, ma), ctx.encrypt(vb, mb), ctx.encrypt(vc, mc)
    assert list(ca.slot_words()) == [17] * 8
    r_ab = ctx.poly_mulrem(ca, cb)
    assert list(r_ab.slot_words()) == [8] * 8  # degree < 512
    h_ab = r_ab.to_host()
    # oracle on a sample
    k = min(nv, 16)
    oa = oracle_encrypt(oracle, pk, va[:k], ma[: k * 256])
    ob = oracle_encrypt(oracle, pk, vb[:k], mb[: k * 256])
    want, _ = oracle.poly_mulrem(oa, ob, sk, threads=oracle.max_threads())
    np.testing.assert_array_equal(h_ab[:k], expected_padded(want, k, [8] * 8))
    # linearity over the whole batch: (a + c) * b mod S == a*b mod S + c*b mod S
    r_cb = ctx.poly_mulrem(cc, cb)
    r_sum = ctx.poly_mulrem(ctx.poly_add(ca, cc), cb)
    np.testing.assert_array_equal(r_sum.to_host(), h_ab ^ r_cb.to_host())
    # and commutativity
    np.testing.assert_array_equal(ctx.poly_mulrem(cb, ca).to_host(), h_ab)


def test_config_b_u32_add_one_wave(oracle, hm):
    """Config B (d = d' = 512, tau = 256, delta = 8) u32 add on 37 888 + 40 pairs — above the dispatcher's switch to the fused
    thread-per-value chain (adder_chain_wide_kernel), so this is the default path at that size: ONE launch; every slot of a
    sample of values equals the oracle's add_internal (common.rs:37-56) word for word, all values decrypt to a + b, and the
    result equals the regrouped generic plan's on a slice."""
    sk, pk, skb, pkb = keys(oracle, *CONFIG_B, 105)
    ctx = engine_context(hm, *CONFIG_B, skb, pkb)
    n = 148 * 256 + 40
    rng = np.random.default_rng(11)
    a = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    b = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    a[:3] = [0, 0xFFFFFFFF, 22]
    b[:3] = [0, 1, 20]
    ma, mb = rbytes(12, n * 32 * 32), rbytes(13, n * 32 * 32)
    ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
    l0 = ctx.kernel_launches()
    s = ctx.apply2(hm.HomomorphicAddition, ca, cb)
    assert ctx.kernel_launches() - l0 == 1
    dec = ctx.decrypt(s)
    np.testing.assert_array_equal(dec, a + b)  # d / delta = 64 >= 21: exact (operations.rs MIN_D_OVER_DELTA for u32 add)
    idx = np.array([0, 1, 2, 31, 32, n // 2, n - 41, n - 1])
    per = 32 * 32
    oa = oracle_encrypt(oracle, pk, a[idx], ma.reshape(n, per)[idx].reshape(-1))
    ob = oracle_encrypt(oracle, pk, b[idx], mb.reshape(n, per)[idx].reshape(-1))
    want, _ = oracle.apply(oracle.OP_ADD, oa, ob, 32, threads=oracle.max_threads())
    got = np.concatenate([s.rows_to_host(int(i), 1) for i in idx])
    np.testing.assert_array_equal(got, expected_padded(want, len(idx), s.slot_words()))
    m = 64
    sa, sb = ctx.encrypt(a[:m], ma[: m * per]), ctx.encrypt(b[:m], mb[: m * per])
    g = ctx.apply2(hm.HomomorphicAddition, sa, sb)  # 64 values: the regrouped generic plan
    np.testing.assert_array_equal(g.to_host(), s.rows_to_host(0, m))
