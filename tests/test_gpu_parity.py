"""GPU parity: the CUDA engine (through the C ABI / Python mirror) against the CPU oracle, bit for bit.

Reference behaviours covered: CipheredBit::cipher/decipher (src/cipher.rs:99-122), try_cipher /
try_decipher ordering (:175-250), the gates and circuits of src/impls/numbers/common.rs:5-105,
Polynomial::add/mul/rem (src/polynomial.rs:190-365).
"""
import numpy as np
import pytest

from helpers import engine_context, expected_padded, keys, oracle_encrypt, philox_masks, random_polys

pytestmark = pytest.mark.gpu

CONFIG_A = (128, 128, 1, 128)
CONFIG_B = (512, 512, 8, 256)


@pytest.fixture(scope="module")
def hm():
    import homomorph_rust_b200 as h

    assert h.lib().hm_device_count() > 0, "no CUDA device: the GPU tests must not silently pass"
    return h


def setup(oracle, hm, params, seed):
    sk, pk, skb, pkb = keys(oracle, *params, seed)
    return sk, pk, engine_context(hm, *params, skb, pkb)


def masks_for(rng, n, L, tau):
    return rng.integers(0, 256, size=n * L * ((tau + 7) // 8), dtype=np.uint8)


@pytest.mark.parametrize(
    "params,dtype,n",
    [
        (CONFIG_A, np.uint32, 700),  # 22 400 bit-cts: full tiles + ragged tail of the table kernel
        (CONFIG_A, np.uint8, 37),
        (CONFIG_B, np.uint32, 130),
        ((64, 32, 8, 32), np.uint8, 50),   # src/cipher.rs:277
        ((64, 32, 8, 32), np.uint64, 9),
        ((6, 3, 2, 5), np.uint8, 20),      # doctest parameters, tau not a multiple of 8
        ((32, 16, 16, 16), np.uint16, 11),
    ],
)
def test_encrypt_decrypt(oracle, hm, params, dtype, n):
    rng = np.random.default_rng(hash((params, n)) % 2**32)
    sk, pk, ctx = setup(oracle, hm, params, 5)
    L = np.dtype(dtype).itemsize * 8
    values = rng.integers(0, np.iinfo(dtype).max, size=n, dtype=dtype, endpoint=True)
    masks = masks_for(rng, n, L, params[3])
    ct = ctx.encrypt(values, masks)
    assert ct.bits == L and len(ct) == n  # len == T::BITS, src/cipher.rs:286,292
    want = oracle_encrypt(oracle, pk, values, masks)
    got = ct.to_host()
    np.testing.assert_array_equal(got, expected_padded(want, n, ct.slot_words()))
    dec = ctx.decrypt(ct)
    odec, _ = oracle.decrypt(sk, want, L)
    np.testing.assert_array_equal(dec.view(np.uint8), odec)
    # δ small enough for a fresh ciphertext to decrypt to its plaintext in these parameter sets
    if params[2] * 2 <= params[0]:
        np.testing.assert_array_equal(dec, values)


def test_encrypt_empty_and_errors(oracle, hm):
    sk, pk, skb, pkb = keys(oracle, *CONFIG_A, 1)
    ctx = hm.Context(hm.Parameters(*CONFIG_A))
    with pytest.raises(hm.PublicKeyUnset):
        ctx.encrypt(np.zeros(1, dtype=np.uint8))
    ctx.set_public_key(hm.PublicKey.from_bytes(pkb))
    ct = ctx.encrypt(np.zeros(0, dtype=np.uint32))
    assert len(ct) == 0 and ct.bits == 32
    one = ctx.encrypt(np.array([7], dtype=np.uint8), rng=np.random.default_rng(0))
    with pytest.raises(hm.SecretKeyUnset):
        ctx.decrypt(one)
    ctx.set_secret_key(hm.SecretKey.from_bytes(skb))
    assert ctx.get_public_key() is None  # set_secret_key clears the public key, src/context.rs:657-667
    with pytest.raises(hm.PublicKeyUnset):
        ctx.encrypt(np.zeros(1, dtype=np.uint8))
    assert int(ctx.decrypt(one)[0]) == 7
    assert ctx.decrypt(ct).size == 0


@pytest.mark.parametrize("params", [(32, 16, 16, 16), (32, 8, 1, 8), CONFIG_A])
def test_gates(oracle, hm, params):
    rng = np.random.default_rng(12)
    sk, pk, ctx = setup(oracle, hm, params, 9)
    n, L = 33, 8
    a = rng.integers(0, 256, size=n, dtype=np.uint8)
    b = rng.integers(0, 256, size=n, dtype=np.uint8)
    a[0], b[0] = 0b1010, 0b1100  # uint.rs:109-174
    ma, mb = masks_for(rng, n, L, params[3]), masks_for(rng, n, L, params[3])
    ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
    oa, ob = oracle_encrypt(oracle, pk, a, ma), oracle_encrypt(oracle, pk, b, mb)
    for op, oop in [(hm.HomomorphicAndGate, oracle.OP_AND), (hm.HomomorphicOrGate, oracle.OP_OR), (hm.HomomorphicXorGate, oracle.OP_XOR)]:
        if params[0] < op.MIN_D_OVER_DELTA * params[2]:
            with pytest.raises(hm.OperationError):
                ctx.apply2(op, ca, cb)
            continue
        r = ctx.apply2(op, ca, cb)
        want, _ = oracle.apply(oop, oa, ob, L)
        np.testing.assert_array_equal(r.to_host(), expected_padded(want, n, r.slot_words()))
        od, _ = oracle.decrypt(sk, want, L)
        np.testing.assert_array_equal(ctx.decrypt(r).view(np.uint8), od)
    c = ca.clone()
    ctx.apply1(hm.HomomorphicNotGate, c)
    want, _ = oracle.apply(oracle.OP_NOT, oa, None, L)
    np.testing.assert_array_equal(c.to_host(), expected_padded(want, n, c.slot_words()))
    if params[2] == 1:
        assert int(ctx.decrypt(c)[0]) == 0b1111_0101  # uint.rs:165-168


@pytest.mark.parametrize("dtype,n", [(np.uint32, 40), (np.uint8, 70), (np.uint16, 9)])
def test_add_fused_config_a(oracle, hm, dtype, n):
    """HomomorphicAddition at d=d'=128 (fused ripple-carry kernel) — common.rs:37-56."""
    rng = np.random.default_rng(n)
    sk, pk, ctx = setup(oracle, hm, CONFIG_A, 21)
    L = np.dtype(dtype).itemsize * 8
    a = rng.integers(0, np.iinfo(dtype).max, size=n, dtype=dtype, endpoint=True)
    b = rng.integers(0, np.iinfo(dtype).max, size=n, dtype=dtype, endpoint=True)
    a[:3] = [22, np.iinfo(dtype).max, 0]
    b[:3] = [20, 240, 0]
    ma, mb = masks_for(rng, n, L, 128), masks_for(rng, n, L, 128)
    ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
    r = ctx.apply2(hm.HomomorphicAddition, ca, cb)
    oa, ob = oracle_encrypt(oracle, pk, a, ma), oracle_encrypt(oracle, pk, b, mb)
    want, _ = oracle.apply(oracle.OP_ADD, oa, ob, L, threads=oracle.max_threads())
    widths = r.slot_words()
    D = 256
    assert list(widths) == [D // 64 + 1] + [((3 * k - 1) * D) // 64 + 1 for k in range(1, L)]
    got = r.to_host()
    np.testing.assert_array_equal(got, expected_padded(want, n, widths))
    # the generic slot-by-slot path in the reference's own evaluation order gives the same polynomials
    g = ctx.apply2(hm.HomomorphicAddition, ca, cb, generic=True)
    np.testing.assert_array_equal(g.to_host(), got)
    # decrypt after add
    dec = ctx.decrypt(r)
    od, _ = oracle.decrypt(sk, want, L)
    np.testing.assert_array_equal(dec.view(np.uint8), od)
    assert int(dec[0]) == 42  # uint.rs:186-190


@pytest.mark.parametrize("params,dtype", [((64, 16, 1, 16), np.uint8), ((64, 16, 1, 16), np.uint16), ((32, 8, 1, 8), np.uint8), (CONFIG_B, np.uint8)])
def test_add_generic(oracle, hm, params, dtype):
    """uint.rs:176-208 at (64,16,1,16): 22+20=42, 255+240=239 (wrap), random pairs."""
    rng = np.random.default_rng(4)
    sk, pk, ctx = setup(oracle, hm, params, 33)
    n, L = 6, np.dtype(dtype).itemsize * 8
    a = rng.integers(0, np.iinfo(dtype).max // 2, size=n, dtype=dtype)
    b = rng.integers(0, np.iinfo(dtype).max // 2, size=n, dtype=dtype)
    a[:2], b[:2] = [22, 255], [20, 240]
    ma, mb = masks_for(rng, n, L, params[3]), masks_for(rng, n, L, params[3])
    ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
    r = ctx.apply2(hm.HomomorphicAddition, ca, cb)
    oa, ob = oracle_encrypt(oracle, pk, a, ma), oracle_encrypt(oracle, pk, b, mb)
    want, _ = oracle.apply(oracle.OP_ADD, oa, ob, L)
    np.testing.assert_array_equal(r.to_host(), expected_padded(want, n, r.slot_words()))
    od, _ = oracle.decrypt(sk, want, L)
    np.testing.assert_array_equal(ctx.decrypt(r).view(np.uint8), od)
    if params == (64, 16, 1, 16) and dtype == np.uint8:
        assert int(ctx.decrypt(r)[0]) == 42


@pytest.mark.parametrize("params", [(128, 64, 1, 64), CONFIG_A])
def test_mul_u8(oracle, hm, params):
    """HomomorphicMultiplication on u8 — common.rs:66-105; uint.rs:255-293 at (128,64,1,64)."""
    rng = np.random.default_rng(8)
    sk, pk, ctx = setup(oracle, hm, params, 44)
    a = np.array([6, 0, 255, 11, 3], dtype=np.uint8)
    b = np.array([7, 151, 240, 19, 200], dtype=np.uint8)
    n, L = a.size, 8
    ma, mb = masks_for(rng, n, L, params[3]), masks_for(rng, n, L, params[3])
    ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
    r = ctx.apply2(hm.HomomorphicMultiplication, ca, cb)
    oa, ob = oracle_encrypt(oracle, pk, a, ma), oracle_encrypt(oracle, pk, b, mb)
    want, _ = oracle.apply(oracle.OP_MUL, oa, ob, L, threads=oracle.max_threads())
    np.testing.assert_array_equal(r.to_host(), expected_padded(want, n, r.slot_words()))
    dec = ctx.decrypt(r)
    od, _ = oracle.decrypt(sk, want, L)
    np.testing.assert_array_equal(dec.view(np.uint8), od)
    assert list(dec[:3]) == [42, 0, 16]  # uint.rs:264-292


def test_mul_requirement(oracle, hm):
    sk, pk, ctx = setup(oracle, hm, (64, 16, 4, 16), 3)
    c = ctx.encrypt(np.array([1], dtype=np.uint8), rng=np.random.default_rng(1))
    with pytest.raises(hm.OperationError) as e:  # src/context.rs:310-323
        ctx.apply2(hm.HomomorphicMultiplication, c, c)
    assert e.value.required_min_d_over_delta == 64 and e.value.actual_d == 64 and e.value.actual_delta == 4


@pytest.mark.parametrize("wa,wb", [(1, 1), (5, 5), (2, 9), (17, 17), (40, 3), (9, 361)])
def test_poly_mul_rem(oracle, hm, wa, wb):
    """Polynomial::mul / rem on raw batches (src/polynomial.rs:252-365), any widths."""
    rng = np.random.default_rng(wa * 1000 + wb)
    sk, pk, ctx = setup(oracle, hm, CONFIG_A, 2)
    n = 45
    A, B = random_polys(rng, n, wa), random_polys(rng, n, wb)
    A[0] = 0  # null polynomial (src/polynomial.rs:257-261)
    B[1] = 0
    A[2, :] = 0
    A[2, 0] = 1  # the constant 1
    ba, bb = ctx.upload(A, [wa]), ctx.upload(B, [wb])
    prod = ctx.poly_mul(ba, bb)
    want = oracle.poly_binop(oracle.POLY_MUL, oracle.PolyVec.from_padded(A), oracle.PolyVec.from_padded(B))
    np.testing.assert_array_equal(prod.to_host(), expected_padded(want, n, prod.slot_words()))
    rem = ctx.poly_rem(prod)
    wantr = oracle.poly_binop(oracle.POLY_REM, want, sk)
    np.testing.assert_array_equal(rem.to_host(), expected_padded(wantr, n, rem.slot_words()))
    assert list(rem.slot_words()) == [2]  # degree < 128
    s = ctx.poly_add(ba, bb)
    wants = oracle.poly_binop(oracle.POLY_ADD, oracle.PolyVec.from_padded(A), oracle.PolyVec.from_padded(B))
    np.testing.assert_array_equal(s.to_host(), expected_padded(wants, n, s.slot_words()))


@pytest.mark.parametrize("params,n", [(CONFIG_A, 1000), (CONFIG_A, 128 * 7), (CONFIG_B, 200), (CONFIG_B, 512 * 3 + 40), ((64, 32, 8, 32), 77), ((6, 3, 2, 5), 30)])
def test_mulrem_fresh(oracle, hm, params, n):
    """The BASELINE `mul+rem` unit on pairs of fresh ciphertexts (fused kernel at config A)."""
    rng = np.random.default_rng(n)
    sk, pk, ctx = setup(oracle, hm, params, 6)
    va = rng.integers(0, 256, size=(n + 7) // 8, dtype=np.uint8)
    vb = rng.integers(0, 256, size=(n + 7) // 8, dtype=np.uint8)
    ma, mb = masks_for(rng, va.size, 8, params[3]), masks_for(rng, vb.size, 8, params[3])
    ca, cb = ctx.encrypt(va, ma), ctx.encrypt(vb, mb)
    r = ctx.poly_mulrem(ca, cb)
    oa, ob = oracle_encrypt(oracle, pk, va, ma), oracle_encrypt(oracle, pk, vb, mb)
    want, _ = oracle.poly_mulrem(oa, ob, sk, threads=oracle.max_threads())
    np.testing.assert_array_equal(r.to_host(), expected_padded(want, va.size, r.slot_words()))
    # separate mul then rem agrees with the fused kernel
    two = ctx.poly_rem(ctx.poly_mul(ca, cb))
    np.testing.assert_array_equal(two.to_host(), r.to_host())


@pytest.mark.parametrize("dtype,n", [(np.uint64, 3), (np.int32, 17), (np.int8, 40), (np.int64, 2)])
def test_add_other_widths_and_signed(oracle, hm, dtype, n):
    """u64 (uint.rs:210-230, ignored there as 'long test') and the signed types: addition of iN is the same circuit
    on the same bits (int.rs:71-73), so -20 + 22 == 2 (int.rs:187-191)."""
    rng = np.random.default_rng(77)
    sk, pk, ctx = setup(oracle, hm, CONFIG_A, 55)
    info = np.iinfo(dtype)
    L = np.dtype(dtype).itemsize * 8
    a = rng.integers(info.min // 2, info.max // 2, size=n, dtype=dtype)
    b = rng.integers(info.min // 2, info.max // 2, size=n, dtype=dtype)
    if info.min < 0:
        a[0], b[0] = 22, -20
    ma, mb = masks_for(rng, n, L, 128), masks_for(rng, n, L, 128)
    ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
    r = ctx.apply2(hm.HomomorphicAddition, ca, cb)
    u = np.dtype(f"u{np.dtype(dtype).itemsize}")
    oa, ob = oracle_encrypt(oracle, pk, a.view(u), ma), oracle_encrypt(oracle, pk, b.view(u), mb)
    want, _ = oracle.apply(oracle.OP_ADD, oa, ob, L, threads=oracle.max_threads())
    np.testing.assert_array_equal(r.to_host(), expected_padded(want, n, r.slot_words()))
    dec = ctx.decrypt(r)
    od, _ = oracle.decrypt(sk, want, L)
    np.testing.assert_array_equal(dec.view(np.uint8), od)
    assert dec.dtype == np.dtype(dtype)
    if info.min < 0:
        assert int(dec[0]) == 2


def test_mul_signed_i8(oracle, hm):
    """mul_signed_internal (common.rs:115-155; int.rs:248-268 at (512,64,1,64)): 6 * -7 = -42, 0 * 10 = 0.  The two
    extra XORs with `one` land in the last column and cancel, so the unsigned circuit gives the same polynomials."""
    rng = np.random.default_rng(5)
    params = (512, 64, 1, 64)
    sk, pk, ctx = setup(oracle, hm, params, 66)
    a = np.array([6, 0, -3, 11], dtype=np.int8)
    b = np.array([-7, 10, -5, -11], dtype=np.int8)
    n, L = a.size, 8
    ma, mb = masks_for(rng, n, L, 64), masks_for(rng, n, L, 64)
    ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
    r = ctx.apply2(hm.HomomorphicMultiplication, ca, cb)
    oa, ob = oracle_encrypt(oracle, pk, a.view(np.uint8), ma), oracle_encrypt(oracle, pk, b.view(np.uint8), mb)
    want, _ = oracle.apply(oracle.OP_MUL_SIGNED, oa, ob, L, threads=oracle.max_threads())
    np.testing.assert_array_equal(r.to_host(), expected_padded(want, n, r.slot_words()))
    dec = ctx.decrypt(r)
    assert list(dec[:2]) == [-42, 0]
    np.testing.assert_array_equal(dec, a * b)


def test_ragged_and_mismatched_batches(oracle, hm):
    """Edge cases: operands of different widths (xor of a fresh batch with a product), mismatched n / L rejected."""
    rng = np.random.default_rng(3)
    sk, pk, ctx = setup(oracle, hm, CONFIG_A, 4)
    n, L = 9, 8
    a = rng.integers(0, 256, size=n, dtype=np.uint8)
    b = rng.integers(0, 256, size=n, dtype=np.uint8)
    ma, mb = masks_for(rng, n, L, 128), masks_for(rng, n, L, 128)
    ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
    oa, ob = oracle_encrypt(oracle, pk, a, ma), oracle_encrypt(oracle, pk, b, mb)
    prod = ctx.apply2(hm.HomomorphicAndGate, ca, cb)          # 9-word slots
    mixed = ctx.apply2(hm.HomomorphicXorGate, prod, ca)       # 9-word xor 5-word
    oprod, _ = oracle.apply(oracle.OP_AND, oa, ob, L)
    omixed, _ = oracle.apply(oracle.OP_XOR, oprod, oa, L)
    np.testing.assert_array_equal(mixed.to_host(), expected_padded(omixed, n, mixed.slot_words()))
    # (a & b) & a : 9-word x 5-word products
    p3 = ctx.apply2(hm.HomomorphicAndGate, prod, ca)
    op3, _ = oracle.apply(oracle.OP_AND, oprod, oa, L)
    np.testing.assert_array_equal(p3.to_host(), expected_padded(op3, n, p3.slot_words()))
    # add on non-fresh operands takes the generic path and still matches
    s = ctx.apply2(hm.HomomorphicAddition, prod, cb)
    os_, _ = oracle.apply(oracle.OP_ADD, oprod, ob, L)
    np.testing.assert_array_equal(s.to_host(), expected_padded(os_, n, s.slot_words()))
    other = ctx.encrypt(a[:4], ma[: 4 * L * 16])
    with pytest.raises(hm.EngineError):
        ctx.apply2(hm.HomomorphicXorGate, ca, other)
    wide = ctx.encrypt(a.astype(np.uint16), masks_for(rng, n, 16, 128))
    with pytest.raises(hm.EngineError):
        ctx.apply2(hm.HomomorphicXorGate, ca, wide)


def test_wire_format_and_canonical_export(oracle, hm):
    rng = np.random.default_rng(8)
    sk, pk, ctx = setup(oracle, hm, CONFIG_A, 12)
    n = 5
    a = rng.integers(0, 2**16, size=n, dtype=np.uint16)
    b = rng.integers(0, 2**16, size=n, dtype=np.uint16)
    ma, mb = masks_for(rng, n, 16, 128), masks_for(rng, n, 16, 128)
    s = ctx.apply2(hm.HomomorphicAddition, ctx.encrypt(a, ma), ctx.encrypt(b, mb))
    blob = s.to_bytes()
    assert blob[:4] == b"HMB1" and len(blob) == 24 + 16 * 8 + n * s.value_words * 8
    back = hm.Ciphered.from_bytes(ctx, blob)
    np.testing.assert_array_equal(back.to_host(), s.to_host())
    np.testing.assert_array_equal(back.slot_degree_bounds(), s.slot_degree_bounds())
    np.testing.assert_array_equal(ctx.decrypt(back, np.uint16), a + b)
    other = hm.Context(hm.Parameters(128, 128, 2, 128))
    with pytest.raises(ValueError):
        hm.Ciphered.from_bytes(other, blob)
    with pytest.raises((hm.EngineError, hm.CipherError)):
        hm.Ciphered.from_bytes(ctx, blob[:-8])
    # canonical (degree, words) pairs == the oracle's polynomials
    want, _ = oracle.apply(oracle.OP_ADD, oracle_encrypt(oracle, pk, a, ma), oracle_encrypt(oracle, pk, b, mb), 16)
    can = s.canonical()
    assert len(can) == n * 16
    for i, (deg, words) in enumerate(can):
        assert deg == want.degree(i)
        np.testing.assert_array_equal(words, want.words(i))


def test_apply2_host_pipeline(oracle, hm):
    """hm_apply2_host: host ciphertexts in, host result out, chunked over three streams — equals the device-resident
    path and the oracle (several chunks: 18 000 u32 adds = 0.9 GB of result)."""
    import ctypes as C

    from homomorph_rust_b200 import _native as N

    rng = np.random.default_rng(21)
    sk, pk, ctx = setup(oracle, hm, CONFIG_A, 31)
    lib = hm.lib()
    n, L = 18000, 32
    a = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    b = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    ma = np.frombuffer(rng.bytes(n * L * 16), dtype=np.uint8)
    mb = np.frombuffer(rng.bytes(n * L * 16), dtype=np.uint8)
    ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
    ha, hb = ca.to_host(), cb.to_host()
    dev = ctx.apply2(hm.HomomorphicAddition, ca, cb)
    wo = dev.slot_words().astype(np.uint32)
    wa = np.full(L, 5, dtype=np.uint32)
    fresh = np.full(L, 256, dtype=np.uint64)  # degree bound of a fresh ciphertext: d + dp
    u32p, u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    out = np.zeros((n, int(wo.sum())), dtype=np.uint64)
    rc = lib.hm_apply2_host_bounded(ctx._h, N.HM_OP_ADD, n, L, fresh.ctypes.data_as(u64p), ha.ctypes.data, fresh.ctypes.data_as(u64p), hb.ctypes.data, out.ctypes.data)
    assert rc == 0
    check = np.concatenate([np.arange(0, 64), np.arange(n - 64, n), rng.integers(0, n, 256)])
    dh = dev.to_host()
    np.testing.assert_array_equal(out[check], dh[check])
    assert np.array_equal(out, dh)
    k = 8
    want, _ = oracle.apply(oracle.OP_ADD, oracle_encrypt(oracle, pk, a[:k], ma[: k * L * 16]), oracle_encrypt(oracle, pk, b[:k], mb[: k * L * 16]), L,
                           threads=oracle.max_threads())
    np.testing.assert_array_equal(out[:k], expected_padded(want, k, wo))
    # expected result bounds are also what hm_result_slot_bounds reports
    rb = np.zeros(L, dtype=np.uint64)
    assert lib.hm_result_slot_bounds(N.HM_OP_ADD, L, fresh.ctypes.data_as(u64p), fresh.ctypes.data_as(u64p), rb.ctypes.data_as(u64p)) == 0
    np.testing.assert_array_equal(rb // 64 + 1, wo)
    np.testing.assert_array_equal(rb, dev.slot_degree_bounds())
    # XOR through the widths form of the call (memory-bound op, one chunk)
    outx = np.zeros((n, 160), dtype=np.uint64)
    assert lib.hm_apply2_host(ctx._h, N.HM_OP_XOR, n, L, wa.ctypes.data_as(u32p), ha.ctypes.data, wa.ctypes.data_as(u32p), hb.ctypes.data, outx.ctypes.data) == 0
    np.testing.assert_array_equal(outx, ha ^ hb)
    # The widths form assumes the widest degrees a slot can hold (64 w - 1), so its result layout is wider than the fresh
    # one; the polynomials are the same.  A few values through the generic kernels:
    m = 6
    rw = np.zeros(L, dtype=np.uint32)
    assert lib.hm_result_slot_words(ctx._h, N.HM_OP_ADD, L, wa.ctypes.data_as(u32p), wa.ctypes.data_as(u32p), rw.ctypes.data_as(u32p)) == 0
    assert rw[0] == 5 and rw[1] == 10 and np.all(rw >= wo)
    outw = np.zeros((m, int(rw.sum())), dtype=np.uint64)
    assert lib.hm_apply2_host(ctx._h, N.HM_OP_ADD, m, L, wa.ctypes.data_as(u32p), ha.ctypes.data, wa.ctypes.data_as(u32p), hb.ctypes.data, outw.ctypes.data) == 0
    np.testing.assert_array_equal(outw, expected_padded(want, k, rw)[:m])
    # requirement check happens before any work (src/context.rs:310-323)
    ctx2 = hm.Context(hm.Parameters(64, 16, 8, 16))
    assert lib.hm_apply2_host(ctx2._h, N.HM_OP_ADD, 1, 8, wa.ctypes.data_as(u32p), ha.ctypes.data, wa.ctypes.data_as(u32p), hb.ctypes.data, out.ctypes.data) == N.HM_ERR_OPERATION_REQUIREMENT


@pytest.mark.parametrize("params,dtype,n", [(CONFIG_A, np.uint32, 300), (CONFIG_B, np.uint8, 40), ((6, 3, 2, 5), np.uint8, 11), ((64, 32, 8, 200), np.uint16, 7)])
def test_encrypt_seeded_device_masks(oracle, hm, params, dtype, n):
    """Masks generated on the device (Philox4x32-10) == the documented host stream, and the resulting ciphertexts ==
    the oracle fed with that stream (so a seeded run is reproducible by the reference with a getrandom shim)."""
    rng = np.random.default_rng(n)
    sk, pk, ctx = setup(oracle, hm, params, 77)
    L = np.dtype(dtype).itemsize * 8
    values = rng.integers(0, np.iinfo(dtype).max, size=n, dtype=dtype, endpoint=True)
    seed = 0xC0FFEE_0000_0001 + n
    ct = ctx.encrypt(values, seed=seed)
    masks = ctx.seeded_masks(n * L, seed)
    want = oracle_encrypt(oracle, pk, values, masks)
    np.testing.assert_array_equal(ct.to_host(), expected_padded(want, n, ct.slot_words()))
    # same seed, explicit host masks: identical batch
    np.testing.assert_array_equal(ctx.encrypt(values, masks).to_host(), ct.to_host())
    assert not np.array_equal(ctx.encrypt(values, seed=seed + 1).to_host(), ct.to_host())
    if params[2] * 2 <= params[0]:
        np.testing.assert_array_equal(ctx.decrypt(ct), values)


@pytest.mark.parametrize("params,n,first_unit,odd_key", [
    (CONFIG_A, 257, 0, False), (CONFIG_A, 1, 5, False), (CONFIG_A, 1031, (1 << 33) + 7, True), (CONFIG_A, 40, 96, True),
    (CONFIG_B, 131, 0, False), (CONFIG_B, 1, 3, True), (CONFIG_B, 517, (1 << 32) + 11, True)])
def test_encrypt_fused_philox(oracle, hm, params, n, first_unit, odd_key):
    """encrypt_tab4_kernel / encrypt_tab4b_kernel with the Philox masks drawn inside the kernel (hm_encrypt_device_seeded_into):
    equal, word for word, to the oracle's cipher (cipher.rs:99-115) fed with the documented host stream from the same stream
    position, and to the host-mask path of the same kernel.  odd_key: a public key whose polynomials do not all reach degree
    d + d', so the computed leading coefficient (parity of mask AND topmask) is exercised with a non-trivial topmask."""
    import torch

    rng = np.random.default_rng(n)
    sk, pk, ctx = setup(oracle, hm, params, 91)
    tau, wfull = params[3], (params[0] + params[1]) // 64
    if odd_key:
        polys = [pk.words(i).copy() for i in range(len(pk))]
        for i in range(0, len(polys), 3):  # every third polynomial loses its leading coefficient and a few more
            polys[i] = polys[i][:wfull].copy()
            polys[i][wfull - 1] &= np.uint64((1 << (40 + i % 20)) - 1)
        pk = oracle.PolyVec.from_words(polys)
        ctx.set_public_key(hm.PublicKey.from_bytes([p.astype("<u8").tobytes() for p in polys]))
    L = 32
    mbytes = tau // 8
    lib = hm.lib()
    values = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    seed = 0xABCD_0000_1234 + n
    out = ctx.encrypt(np.zeros(n, dtype=np.uint32), np.zeros(n * L * mbytes, dtype=np.uint8))
    dv = torch.from_numpy(values.view(np.uint8).copy()).cuda()
    torch.cuda.synchronize()
    l0 = ctx.kernel_launches()
    assert lib.hm_encrypt_device_seeded_into(ctx._h, dv.data_ptr(), n, L, seed, first_unit, out._h) == 0
    ctx.synchronize()
    assert ctx.kernel_launches() - l0 == 1  # no mask_fill_kernel launch
    masks = philox_masks(n * L, seed, first_unit, mbytes)  # the documented stream, written independently in numpy
    if first_unit == 0:
        np.testing.assert_array_equal(masks, ctx.seeded_masks(n * L, seed))
    want = oracle_encrypt(oracle, pk, values, masks)
    np.testing.assert_array_equal(out.to_host(), expected_padded(want, n, out.slot_words()))
    np.testing.assert_array_equal(ctx.encrypt(values, masks).to_host(), out.to_host())


@pytest.mark.parametrize("force_thread", [0, 24, 32])
def test_poly_mul_random_shapes(oracle, hm, force_thread):
    """Randomised widths and degree bounds through every multiply kernel class (thread-per-product Karatsuba for the
    'k*256 + 1 bit' shapes; for the rest warp-cooperative tiles + scalar tail, or — forced here for small batches — the
    thread-per-chunk Karatsuba kernels with atomic accumulation, 24- and 32-word chunks) against Polynomial::mul of the
    oracle."""
    rng = np.random.default_rng(2026)
    sk, pk, ctx = setup(oracle, hm, CONFIG_A, 2)
    hm.lib().hm_set_tuning(b"mul_thread_min", 0 if force_thread else 1 << 50)
    hm.lib().hm_set_tuning(b"mul_thread_chunk", force_thread or 32)
    shapes = [(256, 256), (512, 512), (512, 1024), (1024, 512), (1024, 1024), (256, 512), (2048, 1024)]
    for _ in range(25):
        shapes.append((int(rng.integers(0, 3000)), int(rng.integers(0, 3000))))
    shapes += [(0, 0), (0, 700), (63, 64), (64, 63), (4095, 1), (23552, 768), (768, 23552), (14336, 8192)]
    for da, db in shapes:
        n = 7
        wa, wb = da // 64 + 1, db // 64 + 1
        A = random_polys(rng, n, wa, (da % 64) + 1)
        B = random_polys(rng, n, wb, (db % 64) + 1)
        A[:, -1] |= np.uint64(1) << np.uint64(da % 64)  # exact degree da for most rows
        A[0] = 0
        B[1] = 0
        ba = ctx.upload(A, [wa], [da])
        bb = ctx.upload(B, [wb], [db])
        prod = ctx.poly_mul(ba, bb)
        assert list(prod.slot_words()) == [(da + db) // 64 + 1]  # out len, src/polynomial.rs:264
        want = oracle.poly_binop(oracle.POLY_MUL, oracle.PolyVec.from_padded(A), oracle.PolyVec.from_padded(B))
        np.testing.assert_array_equal(prod.to_host(), expected_padded(want, n, prod.slot_words()), err_msg=f"da={da} db={db}")
        if da + db >= 128:
            rem = ctx.poly_rem(prod)
            wantr = oracle.poly_binop(oracle.POLY_REM, want, sk)
            np.testing.assert_array_equal(rem.to_host(), expected_padded(wantr, n, rem.slot_words()), err_msg=f"rem da={da} db={db}")
    hm.lib().hm_set_tuning(b"mul_thread_min", -1)
    hm.lib().hm_set_tuning(b"mul_thread_chunk", 32)


@pytest.mark.parametrize("params,dtype,n,force_thread", [((256, 256, 1, 64), np.uint8, 9, 0), ((256, 256, 1, 64), np.uint16, 5, 32), (CONFIG_B, np.uint8, 4, 32), ((64, 16, 1, 16), np.uint16, 6, 24), (CONFIG_A, np.uint32, 3, 32)])
def test_add_generic_plans_agree(oracle, hm, params, dtype, n, force_thread):
    """The regrouped generic adder (batched p, g, m = p + g p, then one product per bit) against the literal evaluation
    of common.rs:44-53 (two long products per bit) and against the oracle, with the warp-cooperative and (forced) the
    thread-per-chunk product kernels."""
    rng = np.random.default_rng(n * 7 + params[0])
    sk, pk, ctx = setup(oracle, hm, params, 17)
    lib = hm.lib()
    L = np.dtype(dtype).itemsize * 8
    a = rng.integers(0, np.iinfo(dtype).max, size=n, dtype=dtype, endpoint=True)
    b = rng.integers(0, np.iinfo(dtype).max, size=n, dtype=dtype, endpoint=True)
    a[:2], b[:2] = [0, np.iinfo(dtype).max], [0, 1]
    ma, mb = masks_for(rng, n, L, params[3]), masks_for(rng, n, L, params[3])
    ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
    try:
        lib.hm_set_tuning(b"mul_thread_min", 0 if force_thread else 1 << 50)
        lib.hm_set_tuning(b"mul_thread_chunk", force_thread or 32)
        assert lib.hm_set_tuning(b"adder_generic_sequential", 1) == 0
        seq = ctx.apply2(hm.HomomorphicAddition, ca, cb, generic=True).to_host()
        assert lib.hm_set_tuning(b"adder_generic_sequential", 0) == 0
        r = ctx.apply2(hm.HomomorphicAddition, ca, cb, generic=True)
    finally:
        lib.hm_set_tuning(b"mul_thread_min", -1)
        lib.hm_set_tuning(b"mul_thread_chunk", 32)
        lib.hm_set_tuning(b"adder_generic_sequential", 0)
    np.testing.assert_array_equal(r.to_host(), seq)
    oa, ob = oracle_encrypt(oracle, pk, a, ma), oracle_encrypt(oracle, pk, b, mb)
    want, _ = oracle.apply(oracle.OP_ADD, oa, ob, L, threads=oracle.max_threads())
    np.testing.assert_array_equal(r.to_host(), expected_padded(want, n, r.slot_words()))
    np.testing.assert_array_equal(ctx.decrypt(r), a + b)


@pytest.mark.parametrize("force_thread", [0, 24, 32])
def test_mul_circuit_plans_agree(oracle, hm, force_thread):
    """The column-batched multiplier circuit (prefix XORs + one batch of carry products per column) against the
    one-product-at-a-time plan and against the oracle's mul_unsigned_internal (common.rs:66-105), bit for bit, with the
    warp-cooperative and (forced) the thread-per-chunk Karatsuba product kernels."""
    rng = np.random.default_rng(31)
    sk, pk, ctx = setup(oracle, hm, CONFIG_A, 12)
    lib = hm.lib()
    n, L = 37, 8
    a = rng.integers(0, 256, size=n, dtype=np.uint8)
    b = rng.integers(0, 256, size=n, dtype=np.uint8)
    a[:3] = [0, 255, 1]
    b[:3] = [77, 255, 0]
    ma, mb = masks_for(rng, n, L, CONFIG_A[3]), masks_for(rng, n, L, CONFIG_A[3])
    ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
    try:
        lib.hm_set_tuning(b"mul_thread_min", 0 if force_thread else 1 << 50)
        lib.hm_set_tuning(b"mul_thread_chunk", force_thread or 32)
        assert lib.hm_set_tuning(b"mul_circuit_sequential", 1) == 0
        seq = ctx.apply2(hm.HomomorphicMultiplication, ca, cb).to_host()
        assert lib.hm_set_tuning(b"mul_circuit_sequential", 0) == 0
        assert lib.hm_set_tuning(b"mul_circuit_fused", 0) == 0
        l0 = ctx.kernel_launches()
        r = ctx.apply2(hm.HomomorphicMultiplication, ca, cb)
        assert ctx.kernel_launches() - l0 > 1  # the column-batched plan
        assert lib.hm_set_tuning(b"mul_circuit_fused", 1) == 0
        l0 = ctx.kernel_launches()
        fused = ctx.apply2(hm.HomomorphicMultiplication, ca, cb)
        assert ctx.kernel_launches() - l0 == 1  # the fused column multiplier (kernels_mul.cu)
    finally:
        lib.hm_set_tuning(b"mul_thread_min", -1)
        lib.hm_set_tuning(b"mul_thread_chunk", 32)
        lib.hm_set_tuning(b"mul_circuit_sequential", 0)
        lib.hm_set_tuning(b"mul_circuit_fused", 1)
    np.testing.assert_array_equal(r.to_host(), seq)
    np.testing.assert_array_equal(fused.to_host(), seq)
    oa, ob = oracle_encrypt(oracle, pk, a, ma), oracle_encrypt(oracle, pk, b, mb)
    want, _ = oracle.apply(oracle.OP_MUL, oa, ob, L, threads=oracle.max_threads())
    np.testing.assert_array_equal(r.to_host(), expected_padded(want, n, r.slot_words()))
    np.testing.assert_array_equal(ctx.decrypt(r), a * b)


@pytest.mark.parametrize("dtype,n", [(np.uint8, 1), (np.uint8, 9), (np.int8, 300), (np.uint8, 1300)])
def test_mul_fused_column_multiplier(oracle, hm, dtype, n):
    """The one-launch fused column multiplier (a warp per value, partial products, prefixes and carries in
    shared memory) on ragged batch sizes — fewer values than warps in a CTA, more than one CTA per SM's worth — against the
    oracle's mul_unsigned_internal / mul_signed_internal (common.rs:66-155), every slot word for word."""
    rng = np.random.default_rng(n)
    sk, pk, ctx = setup(oracle, hm, CONFIG_A, 57)
    L = 8
    info = np.iinfo(dtype)
    a = rng.integers(info.min, info.max, size=n, dtype=dtype, endpoint=True)
    b = rng.integers(info.min, info.max, size=n, dtype=dtype, endpoint=True)
    a[:1] = [info.max]
    b[:1] = [info.max]
    ma, mb = masks_for(rng, n, L, 128), masks_for(rng, n, L, 128)
    ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
    l0 = ctx.kernel_launches()
    r = ctx.apply2(hm.HomomorphicMultiplication, ca, cb)
    assert ctx.kernel_launches() - l0 == 1
    k = min(n, 48)
    oa, ob = oracle_encrypt(oracle, pk, a[:k].view(np.uint8), ma[: k * L * 16]), oracle_encrypt(oracle, pk, b[:k].view(np.uint8), mb[: k * L * 16])
    want, _ = oracle.apply(oracle.OP_MUL_SIGNED if info.min < 0 else oracle.OP_MUL, oa, ob, L, threads=oracle.max_threads())
    np.testing.assert_array_equal(r.to_host()[:k], expected_padded(want, k, r.slot_words()))
    with np.errstate(over="ignore"):
        np.testing.assert_array_equal(ctx.decrypt(r), (a * b).astype(dtype))
    # a second batch through the same context reuses the uploaded plan
    r2 = ctx.apply2(hm.HomomorphicMultiplication, cb, ca)
    np.testing.assert_array_equal(ctx.decrypt(r2), ctx.decrypt(r))


@pytest.mark.parametrize("params", [(64, 32, 1, 32), CONFIG_A])
def test_struct_fields(oracle, hm, params):
    """examples/simple_struct.rs: Vec3 { x, y, z: u16 } added field by field (split_at / new_from_raw / extend_from_slice,
    :32-58; main :62-74 expects {1,2,3} + {4,5,6} = {5,7,9} at (64,32,1,32)) — hm_batch_slice / hm_batch_concat /
    hm_apply2_fields against the oracle's adder run on every field separately."""
    rng = np.random.default_rng(11)
    sk, pk, ctx = setup(oracle, hm, params, 23)
    vec3 = np.dtype([("x", "<u2"), ("y", "<u2"), ("z", "<u2")])
    n, L = 9, 48
    a = np.zeros(n, dtype=vec3)
    b = np.zeros(n, dtype=vec3)
    for f in vec3.names:
        a[f] = rng.integers(0, 1 << 16, size=n)
        b[f] = rng.integers(0, 1 << 16, size=n)
    a[0], b[0] = (1, 2, 3), (4, 5, 6)
    ma = masks_for(rng, n, L, params[3]).reshape(n, L, -1)
    mb = masks_for(rng, n, L, params[3]).reshape(n, L, -1)
    ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
    assert ca.bits == 48
    r = ctx.apply2_fields(hm.HomomorphicAddition, ca, cb, [16, 16, 16])
    widths = list(r.slot_words())
    assert len(widths) == 48
    exp = []
    for i, f in enumerate(vec3.names):
        oa = oracle_encrypt(oracle, pk, np.ascontiguousarray(a[f]), np.ascontiguousarray(ma[:, 16 * i:16 * i + 16]).reshape(-1))
        ob = oracle_encrypt(oracle, pk, np.ascontiguousarray(b[f]), np.ascontiguousarray(mb[:, 16 * i:16 * i + 16]).reshape(-1))
        want, _ = oracle.apply(oracle.OP_ADD, oa, ob, 16, threads=oracle.max_threads())
        exp.append(expected_padded(want, n, widths[16 * i:16 * i + 16]))
    np.testing.assert_array_equal(r.to_host().reshape(n, -1), np.concatenate(exp, axis=1))
    dec = ctx.decrypt(r, dtype=vec3)
    assert tuple(dec[0]) == (5, 7, 9)
    for f in vec3.names:
        np.testing.assert_array_equal(dec[f], a[f] + b[f])
    # the pieces on their own: slice + concat is the identity, a field is an integer batch, bounds are checked
    parts = [ca.slice(0, 16, np.uint16), ca.slice(16, 32)]
    np.testing.assert_array_equal(ctx.concat(parts).to_host(), ca.to_host())
    np.testing.assert_array_equal(ctx.decrypt(parts[0]), a["x"])
    y = ctx.apply2(hm.HomomorphicAddition, ca.slice(16, 16, np.uint16), cb.slice(16, 16, np.uint16))
    np.testing.assert_array_equal(ctx.decrypt(y), a["y"] + b["y"])
    with pytest.raises(hm.InvalidCipheredLength):
        ca.slice(40, 16)
    with pytest.raises(hm.InvalidCipheredLength):
        ctx.apply2_fields(hm.HomomorphicAddition, ca, cb, [16, 16])


def test_add_u128(oracle, hm):
    """u128 (impl list of src/impls/numbers/uint.rs:98-105): 128 bit-ciphertexts per value, the same ripple-carry circuit;
    slot 127 has degree bound 380 * 256.  Fused kernel and regrouped generic plan against the oracle."""
    rng = np.random.default_rng(128)
    sk, pk, ctx = setup(oracle, hm, CONFIG_A, 28)
    u128 = np.dtype([("lo", "<u8"), ("hi", "<u8")])
    n, L = 3, 128
    a = np.zeros(n, dtype=u128)
    b = np.zeros(n, dtype=u128)
    for f in u128.names:
        a[f] = rng.integers(0, 1 << 64, size=n, dtype=np.uint64)
        b[f] = rng.integers(0, 1 << 64, size=n, dtype=np.uint64)
    a[0], b[0] = (2**64 - 1, 2**64 - 1), (1, 0)  # carry through all 128 bits -> 0
    ma, mb = masks_for(rng, n, L, 128), masks_for(rng, n, L, 128)
    ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
    assert ca.bits == 128
    r = ctx.apply2(hm.HomomorphicAddition, ca, cb)
    widths = list(r.slot_words())
    assert widths[0] == 5 and widths[127] == (380 * 256) // 64 + 1
    raw = lambda x: np.frombuffer(x.tobytes(), dtype=np.uint8)
    oa, _ = oracle.encrypt(pk, raw(a), 16, ma, threads=oracle.max_threads())
    ob, _ = oracle.encrypt(pk, raw(b), 16, mb, threads=oracle.max_threads())
    want, _ = oracle.apply(oracle.OP_ADD, oa, ob, L, threads=oracle.max_threads())
    got = r.to_host()
    np.testing.assert_array_equal(got, expected_padded(want, n, widths))
    np.testing.assert_array_equal(ctx.apply2(hm.HomomorphicAddition, ca, cb, generic=True).to_host(), got)
    try:  # and the thread-per-value kernel (normally for batches of >= 256 values per SM)
        hm.lib().hm_set_tuning(b"adder_thread_min", 0)
        np.testing.assert_array_equal(ctx.apply2(hm.HomomorphicAddition, ca, cb).to_host(), got)
    finally:
        hm.lib().hm_set_tuning(b"adder_thread_min", -1)
    dec = ctx.decrypt(r, dtype=u128)
    for i in range(n):
        x = (int(a["hi"][i]) << 64 | int(a["lo"][i])) + (int(b["hi"][i]) << 64 | int(b["lo"][i]))
        assert (int(dec["hi"][i]) << 64 | int(dec["lo"][i])) == x % (1 << 128)
    assert tuple(dec[0]) == (0, 0)


def test_mulrem_on_sums(oracle, hm):
    """mul+rem of non-fresh operands (two u8 sums: slot k has degree bound (3k-1) D): the generic path reduces both operands
    first — rem(mul(a, b)) = rem(mul(rem(a), rem(b))) — and must still equal the oracle's literal mul then rem."""
    rng = np.random.default_rng(61)
    sk, pk, ctx = setup(oracle, hm, CONFIG_A, 9)
    n, L = 11, 8
    vals = [rng.integers(0, 256, size=n, dtype=np.uint8) for _ in range(4)]
    ms = [masks_for(rng, n, L, CONFIG_A[3]) for _ in range(4)]
    cs = [ctx.encrypt(v, m) for v, m in zip(vals, ms)]
    os_ = [oracle_encrypt(oracle, pk, v, m) for v, m in zip(vals, ms)]
    s1 = ctx.apply2(hm.HomomorphicAddition, cs[0], cs[1])
    s2 = ctx.apply2(hm.HomomorphicAddition, cs[2], cs[3])
    w1, _ = oracle.apply(oracle.OP_ADD, os_[0], os_[1], L, threads=oracle.max_threads())
    w2, _ = oracle.apply(oracle.OP_ADD, os_[2], os_[3], L, threads=oracle.max_threads())
    r = ctx.poly_mulrem(s1, s2)
    want, _ = oracle.poly_mulrem(w1, w2, sk, threads=oracle.max_threads())
    np.testing.assert_array_equal(r.to_host(), expected_padded(want, n, r.slot_words()))
    np.testing.assert_array_equal(ctx.poly_rem(ctx.poly_mul(s1, s2)).to_host(), r.to_host())


def test_empty_batches(oracle, hm):
    """n = 0 everywhere (the reference's Vec-based API accepts empty inputs)."""
    sk, pk, ctx = setup(oracle, hm, CONFIG_A, 2)
    e = ctx.encrypt(np.zeros(0, dtype=np.uint8))
    assert len(e) == 0
    for op in (hm.HomomorphicXorGate, hm.HomomorphicAndGate, hm.HomomorphicOrGate, hm.HomomorphicAddition, hm.HomomorphicMultiplication):
        r = ctx.apply2(op, e, e)
        assert len(r) == 0 and r.bits == 8
        assert ctx.decrypt(r).size == 0
    ctx.apply1(hm.HomomorphicNotGate, e)
    assert len(ctx.poly_mulrem(e, e)) == 0
    assert hm.Ciphered.from_bytes(ctx, e.to_bytes()).bits == 8


@pytest.mark.parametrize("params,dtype,n", [((64, 64, 1, 32), np.uint8, 20), ((64, 64, 1, 32), np.uint32, 5), ((256, 256, 1, 64), np.uint8, 9), ((256, 256, 1, 64), np.uint16, 4)])
def test_add_fused_other_degrees(oracle, hm, params, dtype, n):
    """The fused ripple-carry kernel at D = d+d' = 128 (instantiation WD = 4) and the generic circuit at D = 512, against
    the oracle and against each other."""
    rng = np.random.default_rng(n + params[0])
    sk, pk, ctx = setup(oracle, hm, params, 91)
    L = np.dtype(dtype).itemsize * 8
    a = rng.integers(0, np.iinfo(dtype).max, size=n, dtype=dtype, endpoint=True)
    b = rng.integers(0, np.iinfo(dtype).max, size=n, dtype=dtype, endpoint=True)
    ma, mb = masks_for(rng, n, L, params[3]), masks_for(rng, n, L, params[3])
    ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
    r = ctx.apply2(hm.HomomorphicAddition, ca, cb)
    oa, ob = oracle_encrypt(oracle, pk, a, ma), oracle_encrypt(oracle, pk, b, mb)
    want, _ = oracle.apply(oracle.OP_ADD, oa, ob, L, threads=oracle.max_threads())
    D = params[0] + params[1]
    assert list(r.slot_words()) == [D // 64 + 1] + [((3 * k - 1) * D) // 64 + 1 for k in range(1, L)]
    got = r.to_host()
    np.testing.assert_array_equal(got, expected_padded(want, n, r.slot_words()))
    np.testing.assert_array_equal(ctx.apply2(hm.HomomorphicAddition, ca, cb, generic=True).to_host(), got)
    od, _ = oracle.decrypt(sk, want, L)
    np.testing.assert_array_equal(ctx.decrypt(r).view(np.uint8), od)


@pytest.mark.parametrize("dtype,n", [(np.uint32, 300), (np.uint8, 130), (np.uint64, 5)])
def test_add_thread_kernel_small_batches(oracle, hm, dtype, n):
    """The thread-per-value Karatsuba adder (default only for >= 256 values per SM) forced on small batches: bit-exact
    against the oracle and against the warp-per-value kernel, ragged thread counts included."""
    rng = np.random.default_rng(n)
    sk, pk, ctx = setup(oracle, hm, CONFIG_A, 23)
    L = np.dtype(dtype).itemsize * 8
    a = rng.integers(0, np.iinfo(dtype).max, size=n, dtype=dtype, endpoint=True)
    b = rng.integers(0, np.iinfo(dtype).max, size=n, dtype=dtype, endpoint=True)
    ma, mb = masks_for(rng, n, L, 128), masks_for(rng, n, L, 128)
    ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
    lib = hm.lib()
    try:
        assert lib.hm_set_tuning(b"adder_thread_min", 0) == 0
        r_thread = ctx.apply2(hm.HomomorphicAddition, ca, cb).to_host()
        assert lib.hm_set_tuning(b"adder_thread_min", 1 << 40) == 0
        r_warp = ctx.apply2(hm.HomomorphicAddition, ca, cb)
    finally:
        lib.hm_set_tuning(b"adder_thread_min", -1)
    np.testing.assert_array_equal(r_thread, r_warp.to_host())
    k = min(n, 24)
    want, _ = oracle.apply(oracle.OP_ADD, oracle_encrypt(oracle, pk, a[:k], ma[: k * L * 16]), oracle_encrypt(oracle, pk, b[:k], mb[: k * L * 16]), L,
                           threads=oracle.max_threads())
    np.testing.assert_array_equal(r_thread[:k], expected_padded(want, k, r_warp.slot_words()))


@pytest.mark.parametrize("params,dtype,n,chain,phases", [
    (CONFIG_A, np.uint32, 300, 4, 4), (CONFIG_A, np.uint32, 97, 4, 1), (CONFIG_A, np.uint32, 70, 3, 8), (CONFIG_A, np.uint32, 300, 13, 4),
    (CONFIG_A, np.uint8, 130, 12, 3), (CONFIG_A, np.uint64, 5, 4, 8), (CONFIG_A, np.uint16, 33, 13, 2), (CONFIG_A, np.uint32, 64, 0, 4),
    ((64, 64, 1, 32), np.uint32, 200, 4, 4), ((64, 64, 1, 32), np.uint8, 45, 13, 2), ((64, 64, 1, 32), np.uint64, 9, 4, 8)])
def test_add_chain_kernel_variants(oracle, hm, params, dtype, n, chain, phases):
    """The dynamically scheduled thread-per-value chain (kernels_adder.cu) in every variant the dispatcher can pick — window
    in registers or shared memory, 2/3/4 CTAs per SM, 1..8 work units per value, D = 256 and D = 128 — forced on small ragged
    batches: bit-exact against the oracle and against the warp-per-value kernel (common.rs:37-56)."""
    rng = np.random.default_rng(n * 31 + chain + phases)
    sk, pk, ctx = setup(oracle, hm, params, 29)
    tau = params[3]
    L = np.dtype(dtype).itemsize * 8
    a = rng.integers(0, np.iinfo(dtype).max, size=n, dtype=dtype, endpoint=True)
    b = rng.integers(0, np.iinfo(dtype).max, size=n, dtype=dtype, endpoint=True)
    ma, mb = masks_for(rng, n, L, tau), masks_for(rng, n, L, tau)
    ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
    lib = hm.lib()
    try:
        assert lib.hm_set_tuning(b"adder_thread_min", 0) == 0
        assert lib.hm_set_tuning(b"adder_chain", chain) == 0
        assert lib.hm_set_tuning(b"adder_phases", phases) == 0
        l0 = ctx.kernel_launches()
        r_chain = ctx.apply2(hm.HomomorphicAddition, ca, cb)
        assert ctx.kernel_launches() - l0 == 1  # one fused launch
        got = r_chain.to_host()
        assert lib.hm_set_tuning(b"adder_thread_min", 1 << 40) == 0
        r_warp = ctx.apply2(hm.HomomorphicAddition, ca, cb)
    finally:
        lib.hm_set_tuning(b"adder_thread_min", -1)
        lib.hm_set_tuning(b"adder_chain", 4)
        lib.hm_set_tuning(b"adder_phases", 0)
    np.testing.assert_array_equal(got, r_warp.to_host())
    k = min(n, 40)
    mbytes = (tau + 7) // 8
    want, _ = oracle.apply(oracle.OP_ADD, oracle_encrypt(oracle, pk, a[:k], ma[: k * L * mbytes]), oracle_encrypt(oracle, pk, b[:k], mb[: k * L * mbytes]), L,
                           threads=oracle.max_threads())
    np.testing.assert_array_equal(got[:k], expected_padded(want, k, r_chain.slot_words()))
    od, _ = oracle.decrypt(sk, want, L)
    np.testing.assert_array_equal(ctx.decrypt(r_chain)[:k].view(np.uint8), od)


@pytest.mark.parametrize("dtype,n,phases", [(np.uint8, 70, 1), (np.uint32, 37, 8), (np.uint16, 33, 3), (np.uint64, 3, 2)])
def test_add_chain_kernel_wide(oracle, hm, dtype, n, phases):
    """Config B (d = d' = 512, D = 1024): the thread-per-value chain with four sub-multipliers (adder_chain_wide_kernel) forced
    on small ragged batches: one launch, bit-exact against the regrouped generic plan and against the oracle
    (common.rs:37-56)."""
    rng = np.random.default_rng(n * 17 + phases)
    sk, pk, ctx = setup(oracle, hm, CONFIG_B, 31)
    tau = CONFIG_B[3]
    L = np.dtype(dtype).itemsize * 8
    a = rng.integers(0, np.iinfo(dtype).max, size=n, dtype=dtype, endpoint=True)
    b = rng.integers(0, np.iinfo(dtype).max, size=n, dtype=dtype, endpoint=True)
    ma, mb = masks_for(rng, n, L, tau), masks_for(rng, n, L, tau)
    ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
    lib = hm.lib()
    try:
        assert lib.hm_set_tuning(b"adder_wide_min", 1) == 0
        assert lib.hm_set_tuning(b"adder_phases", phases) == 0
        l0 = ctx.kernel_launches()
        r_wide = ctx.apply2(hm.HomomorphicAddition, ca, cb)
        assert ctx.kernel_launches() - l0 == 1  # one fused launch
        got = r_wide.to_host()
        assert lib.hm_set_tuning(b"adder_wide_min", -1) == 0
        l0 = ctx.kernel_launches()
        r_gen = ctx.apply2(hm.HomomorphicAddition, ca, cb)
        assert ctx.kernel_launches() - l0 > 1
    finally:
        lib.hm_set_tuning(b"adder_wide_min", 0)
        lib.hm_set_tuning(b"adder_phases", 0)
    np.testing.assert_array_equal(got, r_gen.to_host())
    k = min(n, 6 if L > 16 else 24)
    mbytes = (tau + 7) // 8
    want, _ = oracle.apply(oracle.OP_ADD, oracle_encrypt(oracle, pk, a[:k], ma[: k * L * mbytes]), oracle_encrypt(oracle, pk, b[:k], mb[: k * L * mbytes]), L,
                           threads=oracle.max_threads())
    np.testing.assert_array_equal(got[:k], expected_padded(want, k, r_wide.slot_words()))
    od, _ = oracle.decrypt(sk, want, L)
    np.testing.assert_array_equal(ctx.decrypt(r_wide)[:k].view(np.uint8), od)


def test_into_and_device_entry_points(oracle, hm):
    """The entry points bench.py times — hm_encrypt_device_into, hm_apply2_into, hm_poly_mulrem_into, hm_decrypt_device —
    compared word for word with the oracle (they share the kernels of the allocating calls, but are separate ABI paths)."""
    import ctypes as C

    import torch

    from homomorph_rust_b200 import _native as N

    rng = np.random.default_rng(77)
    sk, pk, ctx = setup(oracle, hm, CONFIG_A, 41)
    lib = hm.lib()
    n, L = 257, 32  # ragged against every tile size
    a = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    b = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    ma, mb = masks_for(rng, n, L, 128), masks_for(rng, n, L, 128)
    oa, ob = oracle_encrypt(oracle, pk, a, ma), oracle_encrypt(oracle, pk, b, mb)
    # hm_encrypt_device_into: values + masks already in HBM, into existing batches
    ca = ctx.encrypt(np.zeros(n, dtype=np.uint32), np.zeros(n * L * 16, dtype=np.uint8))
    cb = ca.clone()
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x).view(np.uint8).copy()).cuda()
    for batch, v, m in ((ca, a, ma), (cb, b, mb)):
        dv, dm = dev(v), dev(m)
        torch.cuda.synchronize()
        assert lib.hm_encrypt_device_into(ctx._h, dv.data_ptr(), n, L, dm.data_ptr(), batch._h) == 0
        ctx.synchronize()
    fresh_w = [5] * L
    np.testing.assert_array_equal(ca.to_host(), expected_padded(oa, n, fresh_w))
    np.testing.assert_array_equal(cb.to_host(), expected_padded(ob, n, fresh_w))
    assert lib.hm_encrypt_device_into(ctx._h, dv.data_ptr(), n + 1, L, dm.data_ptr(), cb._h) == N.HM_ERR_INVALID_ARGUMENT
    # hm_apply2_into for ADD (dispatches to the warp kernel at this size), AND, XOR
    for op, oop in ((hm.HomomorphicAddition, oracle.OP_ADD), (hm.HomomorphicAndGate, oracle.OP_AND), (hm.HomomorphicXorGate, oracle.OP_XOR)):
        out = ctx.apply2(op, ca, cb)
        want, _ = oracle.apply(oop, oa, ob, L, threads=oracle.max_threads())
        exp = expected_padded(want, n, out.slot_words())
        # overwrite the result with garbage, then recompute in place
        junk = ctx.upload(np.full_like(exp, 0xDEADBEEFCAFEF00D), [int(w) for w in out.slot_words()], [int(x) for x in out.slot_degree_bounds()])
        assert lib.hm_apply2_into(ctx._h, op.code, ca._h, cb._h, junk._h) == 0
        np.testing.assert_array_equal(junk.to_host(), exp)
        # a result batch of the wrong shape is refused
        if op is not hm.HomomorphicXorGate:
            assert lib.hm_apply2_into(ctx._h, op.code, ca._h, cb._h, ca._h) == N.HM_ERR_INVALID_ARGUMENT
        if op is hm.HomomorphicAddition:
            # hm_decrypt_device on the ragged result and on a fresh batch
            dout = torch.zeros(n * 4, dtype=torch.uint8, device="cuda")
            assert lib.hm_decrypt_device(ctx._h, junk._h, dout.data_ptr()) == 0
            ctx.synchronize()
            od, _ = oracle.decrypt(sk, want, L)
            np.testing.assert_array_equal(dout.cpu().numpy(), od)
            np.testing.assert_array_equal(dout.cpu().numpy().view(np.uint32), a + b)
        junk.free()
        out.free()
    ca2, cb2 = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
    dout = torch.zeros(n * 4, dtype=torch.uint8, device="cuda")
    assert lib.hm_decrypt_device(ctx._h, ca2._h, dout.data_ptr()) == 0
    ctx.synchronize()
    np.testing.assert_array_equal(dout.cpu().numpy().view(np.uint32), a)
    # hm_poly_mulrem_into: n * L fresh pairs -> d-bit remainders
    mr = ctx.poly_mulrem(ca2, cb2)
    want, _ = oracle.poly_mulrem(oa, ob, sk, threads=oracle.max_threads())
    exp = expected_padded(want, n, mr.slot_words())
    junk = ctx.upload(np.full_like(exp, 0x5555AAAA5555AAAA), [int(w) for w in mr.slot_words()], [int(x) for x in mr.slot_degree_bounds()])
    assert lib.hm_poly_mulrem_into(ctx._h, ca2._h, cb2._h, junk._h) == 0
    np.testing.assert_array_equal(junk.to_host(), exp)
    np.testing.assert_array_equal(mr.to_host(), exp)
    assert lib.hm_poly_mulrem_into(ctx._h, ca2._h, cb2._h, ca2._h) == N.HM_ERR_INVALID_ARGUMENT


def test_upload_canonical_round_trip_and_orphans(oracle, hm):
    """hm_batch_upload_canonical is the inverse of hm_batch_download_canonical (ragged degrees, null polynomials), rejects
    inconsistent input, and batches outliving their context can still be freed (they are orphans)."""
    import ctypes as C

    from homomorph_rust_b200 import _native as N

    rng = np.random.default_rng(5)
    sk, pk, ctx = setup(oracle, hm, CONFIG_A, 43)
    lib = hm.lib()
    n, L = 9, 8
    a = rng.integers(0, 256, size=n, dtype=np.uint8)
    b = rng.integers(0, 256, size=n, dtype=np.uint8)
    ca, cb = ctx.encrypt(a, masks_for(rng, n, L, 128)), ctx.encrypt(b, masks_for(rng, n, L, 128))
    s = ctx.apply2(hm.HomomorphicAddition, ca, cb)
    can = s.canonical()
    back = ctx.upload_canonical(can, n, L, degree_bounds=[int(x) for x in s.slot_degree_bounds()])
    np.testing.assert_array_equal(back.to_host(), s.to_host())
    np.testing.assert_array_equal(ctx.decrypt(back, np.uint8), a + b)
    # without bounds the slots are as wide as the largest degree seen; the polynomials are the same
    tight = ctx.upload_canonical(can, n, L)
    assert all(int(w) <= int(w0) for w, w0 in zip(tight.slot_words(), s.slot_words()))
    assert [(d, list(w)) for d, w in tight.canonical()] == [(d, list(w)) for d, w in can]
    # a null polynomial is (degree 0, [0]) — src/polynomial.rs:132-137
    z = ctx.upload_canonical([(0, [0])] * 8, 1, 8)
    assert not z.to_host().any() and list(z.slot_words()) == [1] * 8
    # stated degree must be the highest set bit; buffer must be exactly consumed; bounds must cover
    u64p = C.POINTER(C.c_uint64)
    out = C.c_void_p()
    bad = np.array([5, 0b1000], dtype=np.uint64)  # degree 5 stated, true degree 3
    assert lib.hm_batch_upload_canonical(ctx._h, 1, 1, bad.ctypes.data, 2, None, C.byref(out)) == N.HM_ERR_INVALID_ARGUMENT
    good = np.array([3, 0b1000], dtype=np.uint64)
    assert lib.hm_batch_upload_canonical(ctx._h, 1, 1, good.ctypes.data, 1, None, C.byref(out)) == N.HM_ERR_INVALID_LENGTH
    long = np.array([3, 0b1000, 0], dtype=np.uint64)
    assert lib.hm_batch_upload_canonical(ctx._h, 1, 1, long.ctypes.data, 3, None, C.byref(out)) == N.HM_ERR_INVALID_LENGTH
    small = np.array([2], dtype=np.uint64)
    assert lib.hm_batch_upload_canonical(ctx._h, 1, 1, good.ctypes.data, 2, small.ctypes.data_as(u64p), C.byref(out)) == N.HM_ERR_INVALID_ARGUMENT
    assert lib.hm_batch_upload_canonical(ctx._h, 1, 1, good.ctypes.data, 2, None, C.byref(out)) == 0
    lib.hm_batch_free(None, out)  # NULL context: the batch knows its owner
    # orphans: destroy the context first, free the batches afterwards
    ctx.close()
    for batch in (ca, cb, s, back, tight, z):
        batch.free()


@pytest.mark.parametrize("n_dev", [2, 3])
def test_device_group_equals_one_device(oracle, hm, n_dev):
    """hm_group_*: one logical batch sharded by value index over several contexts (distinct GPUs when the box has them,
    else several contexts on GPU 0) gives word for word what one device gives — host masks and seeded masks, ADD / AND /
    XOR / NOT, decrypt gathered in index order; ragged split (n not divisible by the group size)."""
    import ctypes as C

    from homomorph_rust_b200 import _native as N
    from homomorph_rust_b200.api import ContextGroup

    lib = hm.lib()
    have = lib.hm_device_count()
    devices = [i % have for i in range(n_dev)]
    rng = np.random.default_rng(100 + n_dev)
    sk, pk, skb, pkb = keys(oracle, *CONFIG_A, 47)
    ctx = engine_context(hm, *CONFIG_A, skb, pkb)
    grp = ContextGroup(hm.Parameters(*CONFIG_A), devices)
    assert len(grp) == n_dev
    grp.set_secret_key(hm.SecretKey.from_bytes(skb))
    grp.set_public_key(hm.PublicKey.from_bytes(pkb))
    n, L = 101, 32
    a = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    b = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    ma, mb = masks_for(rng, n, L, 128), masks_for(rng, n, L, 128)
    ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, seed=99)
    ga, gb = grp.encrypt(a, ma), grp.encrypt(b, seed=99)
    sizes = [ga.part_len(i) for i in range(n_dev)]
    assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    np.testing.assert_array_equal(ga.to_host(), ca.to_host())
    np.testing.assert_array_equal(gb.to_host(), cb.to_host())  # shard r continues the Philox stream where shard r-1 stopped
    for op in (hm.HomomorphicAddition, hm.HomomorphicAndGate, hm.HomomorphicXorGate, hm.HomomorphicOrGate):
        one = ctx.apply2(op, ca, cb)
        many = grp.apply2(op, ga, gb)
        np.testing.assert_array_equal(many.to_host(), one.to_host())
        np.testing.assert_array_equal(grp.decrypt(many), ctx.decrypt(one))
        if op is hm.HomomorphicAddition:
            np.testing.assert_array_equal(grp.decrypt(many), a + b)
            want, _ = oracle.apply(oracle.OP_ADD, oracle_encrypt(oracle, pk, a[:6], ma[: 6 * L * 16]),
                                   oracle_encrypt(oracle, pk, b[:6], ctx.seeded_masks(n * L, 99)[: 6 * L * 16]), L, threads=oracle.max_threads())
            np.testing.assert_array_equal(many.to_host()[:6], expected_padded(want, 6, many.slot_words()))
        many.free(); one.free()
    grp.apply1(hm.HomomorphicNotGate, ga)
    ctx.apply1(hm.HomomorphicNotGate, ca)
    np.testing.assert_array_equal(ga.to_host(), ca.to_host())
    np.testing.assert_array_equal(grp.decrypt(ga), ~a)
    # tiny batches: fewer values than devices leaves empty shards
    tiny = grp.encrypt(a[:1], ma[: L * 16])
    assert [tiny.part_len(i) for i in range(n_dev)] == [1] + [0] * (n_dev - 1)
    np.testing.assert_array_equal(grp.decrypt(tiny), a[:1])
    empty = grp.encrypt(a[:0], ma[:0])
    assert grp.decrypt(empty).size == 0
    # errors surface like on one context
    g2 = ContextGroup(hm.Parameters(*CONFIG_A), devices[:1])
    with pytest.raises(hm.PublicKeyUnset):
        g2.encrypt(a, ma)
    g2.close()
    for x in (ga, gb, tiny, empty):
        x.free()
    grp.close()


@pytest.mark.parametrize("params", [CONFIG_A, CONFIG_B, (64, 32, 8, 32), (6, 3, 2, 5), (200, 130, 63, 17)])
def test_generate_keys_on_device(oracle, hm, params):
    """hm_generate_keys_seeded: T_i = S * Q_i + X * R_i computed on the GPU equals the oracle's keygen (src/context.rs:249-261,
    src/polynomial.rs:73-96) fed with the same documented byte stream; the keys then encrypt / decrypt correctly."""
    import ctypes as C

    d, dp, delta, tau = params
    lib = hm.lib()
    seed = 0xC0FFEE + d
    ctx = hm.Context(hm.Parameters(*params))
    ctx.generate_keys_seeded(seed)
    nb = lambda deg: (deg // 64 + 1) * 8
    rs = np.zeros(nb(d), dtype=np.uint8)
    assert lib.hm_key_stream_host(seed, 0, rs.size, rs.ctypes.data) == 0
    rp = np.zeros(tau * (nb(dp) + nb(delta)), dtype=np.uint8)
    assert lib.hm_key_stream_host(seed, 1, rp.size, rp.ctypes.data) == 0
    u8p = C.POINTER(C.c_uint8)
    L = oracle.lib()
    osk = oracle.PolyVec(L.orc_keygen_sk(d, rs.ctypes.data_as(u8p)))
    opk = oracle.PolyVec(L.orc_keygen_pk(dp, delta, tau, osk._h, rp.ctypes.data_as(u8p)))
    assert ctx.get_secret_key().to_bytes() == osk.words(0).astype("<u8").tobytes()
    got = ctx.get_public_key().to_bytes()
    assert len(got) == tau
    for i in range(tau):
        assert got[i] == opk.words(i).astype("<u8").tobytes(), f"T_{i} differs"
        assert opk.degree(i) == d + dp  # exact degree: both leading coefficients are forced
    rng = np.random.default_rng(3)
    v = rng.integers(0, 256, size=50, dtype=np.uint8)
    m = masks_for(rng, 50, 8, tau)
    c = ctx.encrypt(v, m)
    np.testing.assert_array_equal(c.to_host(), expected_padded(oracle_encrypt(oracle, opk, v, m), 50, c.slot_words()))
    np.testing.assert_array_equal(ctx.decrypt(c), v)
    # a different seed gives different keys; the stream is deterministic
    ctx2 = hm.Context(hm.Parameters(*params))
    ctx2.generate_keys_seeded(seed)
    assert ctx2.get_public_key().to_bytes() == got
    ctx2.generate_keys_seeded(seed + 1)
    assert ctx2.get_public_key().to_bytes() != got


def test_multiword_kat_on_the_engine(oracle, hm):
    """The committed 3- and 5-word known-answer vectors (tests/golden/kat_multiword.json, minted by the big-integer model and
    re-derived by numpy bit vectors in tests/test_oracle_model.py) through hm_poly_mul and hm_poly_rem."""
    import json
    import os

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat_multiword.json")) as f:
        cases = json.load(f)["cases"]
    ctx = hm.Context(hm.Parameters(128, 128, 1, 128))
    for c in cases:
        a = np.array([[int(x, 16) for x in c["a"]]], dtype=np.uint64)
        b = np.array([[int(x, 16) for x in c["b"]]], dtype=np.uint64)
        want = np.array([int(x, 16) for x in c["out"]], dtype=np.uint64)
        if c["kind"] == "mul":
            ca, cb = ctx.upload(a, [a.shape[1]]), ctx.upload(b, [b.shape[1]])
            got = ctx.poly_mul(ca, cb).to_host()[0]
        else:
            s = int("".join(reversed(c["b"])), 16)
            if s.bit_length() - 1 < 1:
                continue
            ctx.set_secret_key(hm.SecretKey.from_bytes(b[0].astype("<u8").tobytes()))
            got = ctx.poly_rem(ctx.upload(a, [a.shape[1]])).to_host()[0]
        assert np.array_equal(got[: want.size], want) and not got[want.size :].any(), c


@pytest.mark.timeout(900, method="thread")
def test_mul_u16_column_plan(oracle, hm):
    """HomomorphicMultiplication at L = 16 (936 products, 680 carries per value) through the column-batched
    plan — common.rs:66-105 — against the oracle at a small D (the oracle's bit-serial products make D = 256 minutes per
    value), and at config A the batched plan against the one-product-per-launch plan."""
    rng = np.random.default_rng(16)
    params = (64, 16, 1, 16)  # d >= 64 delta (numbers.rs:47-50); D = 80 is not a multiple of 64: generic kernels
    sk, pk, ctx = setup(oracle, hm, params, 45)
    a = np.array([6, 65535, 300, 12345], dtype=np.uint16)
    b = np.array([7, 65535, 200, 54321], dtype=np.uint16)
    n, L = a.size, 16
    ma, mb = masks_for(rng, n, L, params[3]), masks_for(rng, n, L, params[3])
    ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
    l0 = ctx.kernel_launches()
    r = ctx.apply2(hm.HomomorphicMultiplication, ca, cb)
    launches = ctx.kernel_launches() - l0
    assert launches < 120, launches  # one prefix pass + at most a few product launches per column, not one per product
    oa, ob = oracle_encrypt(oracle, pk, a, ma), oracle_encrypt(oracle, pk, b, mb)
    want, _ = oracle.apply(oracle.OP_MUL, oa, ob, L, threads=oracle.max_threads())
    np.testing.assert_array_equal(r.to_host(), expected_padded(want, n, r.slot_words()))
    od, _ = oracle.decrypt(sk, want, L)
    np.testing.assert_array_equal(ctx.decrypt(r).view(np.uint8), od)
    # config A: result degrees per column follow SURVEY.md A.3, and the two plans agree word for word
    sk2, pk2, ctx2 = setup(oracle, hm, CONFIG_A, 46)
    a2 = rng.integers(0, 65536, size=2, dtype=np.uint16)
    b2 = rng.integers(0, 65536, size=2, dtype=np.uint16)
    c2a, c2b = ctx2.encrypt(a2, seed=5), ctx2.encrypt(b2, seed=6)
    p = ctx2.apply2(hm.HomomorphicMultiplication, c2a, c2b)
    assert [int(x) // 256 for x in p.slot_degree_bounds()] == [2, 2, 4, 6, 10, 18, 32, 56, 98, 174, 314, 572, 1044, 1902, 3460, 6302]
    lib = hm.lib()
    try:
        assert lib.hm_set_tuning(b"mul_circuit_sequential", 1) == 0
        q = ctx2.apply2(hm.HomomorphicMultiplication, c2a, c2b)
    finally:
        lib.hm_set_tuning(b"mul_circuit_sequential", 0)
    assert np.array_equal(p.to_host(), q.to_host())
