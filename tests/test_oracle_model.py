"""Cross-checks the C oracle against the independent big-integer model (oracle/pymodel.py)
on seeded random inputs: polynomial ops of mixed widths, keygen, encrypt, decrypt, the adder
and multiplier circuits, and the identities the CUDA kernels rely on."""
import numpy as np
import pytest

from oracle import pymodel as pm


def rand_poly_words(rng, max_words):
    n = int(rng.integers(1, max_words + 1))
    w = rng.integers(0, 1 << 63, n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, n, dtype=np.uint64)
    if rng.random() < 0.3:  # ragged: zero high words, degree not at the buffer end
        w[int(rng.integers(0, n)):] = 0
    if rng.random() < 0.3:
        w[-1] >>= np.uint64(int(rng.integers(0, 64)))
    return [int(x) for x in w]


def test_poly_ops_match_model(oracle):
    rng = np.random.default_rng(20261018)
    A = [rand_poly_words(rng, 7) for _ in range(400)]
    B = [rand_poly_words(rng, 5) for _ in range(400)]
    va, vb = oracle.PolyVec.from_words(A), oracle.PolyVec.from_words(B)
    add = oracle.poly_binop(oracle.POLY_ADD, va, vb)
    mul = oracle.poly_binop(oracle.POLY_MUL, va, vb)
    for i, (a, b) in enumerate(zip(A, B)):
        ia, ib = pm.from_words(a), pm.from_words(b)
        assert pm.from_words(add.words(i)) == ia ^ ib and add.degree(i) == pm.degree(ia ^ ib)
        prod = pm.clmul(ia, ib)
        assert pm.from_words(mul.words(i)) == prod and mul.degree(i) == pm.degree(prod)
        # words past the tracked degree are zero (needed for the engine's padded slots)
        assert pm.from_words(mul.buffer(i)) == prod


def test_rem_matches_model(oracle):
    rng = np.random.default_rng(7)
    for _ in range(300):
        a = rand_poly_words(rng, 9)
        while True:
            s = rand_poly_words(rng, 3)
            if pm.from_words(s) > 1:
                break
        r = oracle.poly_binop(oracle.POLY_REM, oracle.PolyVec.from_words([a]), oracle.PolyVec.from_words([s]))
        want = pm.polymod(pm.from_words(a), pm.from_words(s))
        assert pm.from_words(r.buffer(0)) == want
        assert r.degree(0) == pm.degree(want)


@pytest.mark.parametrize("params", [(6, 3, 2, 5), (64, 32, 8, 32), (128, 128, 1, 128), (32, 16, 16, 16)])
def test_keygen_encrypt_decrypt_match_model(oracle, params):
    d, dp, delta, tau = params
    rng = np.random.default_rng(1234)
    rng2 = np.random.default_rng(1234)
    sk, pk = oracle.keygen(d, dp, delta, tau, rng)
    # same byte stream through the model
    s = pm.random_poly(d, rng2.integers(0, 256, (d // 64 + 1) * 8, dtype=np.uint8).tobytes())
    nbytes = tau * ((dp // 64 + 1) * 8 + (delta // 64 + 1) * 8)
    T = pm.keygen_pk(s, dp, delta, tau, rng2.integers(0, 256, nbytes, dtype=np.uint8).tobytes())
    assert pm.from_words(sk.words(0)) == s and sk.degree(0) == d
    assert len(pk) == tau
    for i in range(tau):
        assert pm.from_words(pk.words(i)) == T[i]
        assert pk.degree(i) == d + dp  # exact degree D (src/context.rs:252-256)
    data = rng.integers(0, 256, 6, dtype=np.uint8)
    mb = (tau + 7) // 8
    masks = rng.integers(0, 256, 6 * 8 * mb, dtype=np.uint8)
    ct, _ = oracle.encrypt(pk, data, 6, masks)
    want = pm.encrypt_bytes(data.tobytes(), T, masks.tobytes())
    assert len(ct) == 48  # ciphertext length = 8 * bytes (src/cipher.rs:286,292)
    for k in range(48):
        assert pm.from_words(ct.words(k)) == want[k] and ct.degree(k) == pm.degree(want[k])
    out, _ = oracle.decrypt(sk, ct, 48)
    assert out.tobytes() == pm.decrypt_bytes(want, s)
    if delta * 2 <= d:  # round trip (src/cipher.rs:276-304)
        assert out.tobytes() == data.tobytes()


def test_circuits_match_model(oracle):
    rng = np.random.default_rng(99)
    d, dp, delta, tau = 64, 16, 1, 16
    sk, pk = oracle.keygen(d, dp, delta, tau, rng)
    s = pm.from_words(sk.words(0))
    mb = (tau + 7) // 8
    vals = rng.integers(0, 256, 4, dtype=np.uint8)  # two u16 values... as 2 x 2 bytes
    a, _ = oracle.encrypt(pk, vals[:2], 2, rng.integers(0, 256, 16 * mb, dtype=np.uint8))
    b, _ = oracle.encrypt(pk, vals[2:], 2, rng.integers(0, 256, 16 * mb, dtype=np.uint8))
    ia = [pm.from_words(a.words(i)) for i in range(16)]
    ib = [pm.from_words(b.words(i)) for i in range(16)]
    add, _ = oracle.apply(oracle.OP_ADD, a, b, 16)
    want = pm.add_circuit(ia, ib)
    for i in range(16):
        assert pm.from_words(add.words(i)) == want[i] and add.degree(i) == pm.degree(want[i])
    # u8 multiplier on the low bytes
    a8 = oracle.PolyVec.from_words([a.words(i) for i in range(8)])
    b8 = oracle.PolyVec.from_words([b.words(i) for i in range(8)])
    mul, _ = oracle.apply(oracle.OP_MUL, a8, b8, 8)
    wantm = pm.mul_circuit(ia[:8], ib[:8])
    for i in range(8):
        assert pm.from_words(mul.words(i)) == wantm[i] and mul.degree(i) == pm.degree(wantm[i])
    # signed multiplier equals the unsigned one (the two XOR-with-one cancel)
    smul, _ = oracle.apply(oracle.OP_MUL_SIGNED, a8, b8, 8)
    for i in range(8):
        assert smul.eq(i, mul, i)
    assert s > 1


def test_identities_used_by_kernels(oracle):
    """Decrypt-as-linear-functional, adder restructuring and Barrett exactness."""
    rng = np.random.default_rng(5)
    d = 128
    s = pm.random_poly(d, rng.integers(0, 256, 24, dtype=np.uint8).tobytes())
    nbits = 3000
    v = pm.decrypt_vector(s, nbits)
    for _ in range(50):
        c = int.from_bytes(rng.integers(0, 256, nbits // 8, dtype=np.uint8).tobytes(), "little")
        assert bin(c & v).count("1") % 2 == pm.decrypt_bit(c, s)
    for _ in range(50):
        p = int.from_bytes(rng.integers(0, 256, 33, dtype=np.uint8).tobytes(), "little")
        g = int.from_bytes(rng.integers(0, 256, 65, dtype=np.uint8).tobytes(), "little")
        c = int.from_bytes(rng.integers(0, 256, 200, dtype=np.uint8).tobytes(), "little")
        cpp = pm.clmul(p, c)
        ref = cpp ^ pm.clmul(g, cpp ^ 1)  # src/impls/numbers/common.rs:51-52
        m = p ^ pm.clmul(g, p)
        assert ref == pm.clmul(m, c) ^ g
    # (A mod S)(B mod S) mod S == (A B) mod S : pre-reduction used by the fused mul+rem kernel
    for _ in range(50):
        a = int.from_bytes(rng.integers(0, 256, 33, dtype=np.uint8).tobytes(), "little")
        b = int.from_bytes(rng.integers(0, 256, 33, dtype=np.uint8).tobytes(), "little")
        assert pm.polymod(pm.clmul(pm.polymod(a, s), pm.polymod(b, s)), s) == pm.polymod(pm.clmul(a, b), s)


# ---- multi-word known-answer vectors (tests/golden/kat_multiword.json) -------------------------------------------------
# The reference's own KATs stop at two words (src/polynomial.rs:538-582).  The committed vectors were minted by the big-integer
# model; here they are re-derived by a THIRD implementation that shares nothing with it — numpy bit vectors, products by
# convolution of 0/1 arrays, remainders by schoolbook long division on arrays — and the C oracle is held to them.
def _kat_cases():
    import json
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat_multiword.json")
    with open(path) as f:
        return json.load(f)["cases"]


def _bits(words_hex):
    w = np.array([int(x, 16) for x in words_hex], dtype=np.uint64)
    return np.unpackbits(w.view(np.uint8), bitorder="little").astype(np.int64)  # bit i = coefficient of X^i


def _trim(bits):
    nz = np.nonzero(bits)[0]
    return bits[: nz[-1] + 1] if nz.size else bits[:1] * 0


def bitvec_mul(a_bits, b_bits):
    return np.convolve(a_bits, b_bits) % 2  # coefficient k = sum_{i+j=k} a_i b_j over Z, reduced mod 2


def bitvec_rem(a_bits, s_bits):
    a, s = _trim(a_bits).copy(), _trim(s_bits)
    ds = s.size - 1
    for top in range(a.size - 1, ds - 1, -1):  # cancel leading coefficients one by one (src/polynomial.rs:330-359)
        if a[top]:
            a[top - ds : top + 1] ^= s
    return a[:ds] if ds else a[:1] * 0


def _words_of(bits, nwords):
    out = np.zeros(nwords * 64, dtype=np.uint8)
    out[: min(bits.size, out.size)] = bits[: out.size]
    assert not bits[out.size :].any()
    return [int(x) for x in np.packbits(out, bitorder="little").view(np.uint64)]


def test_multiword_kat_by_bitvectors_and_oracle(oracle):
    cases = _kat_cases()
    assert sum(c["kind"] == "mul" for c in cases) >= 30 and sum(c["kind"] == "rem" for c in cases) >= 20
    assert any(len(c["a"]) == 3 for c in cases) and any(len(c["a"]) == 5 for c in cases) and any(len(c["b"]) == 5 for c in cases)
    for c in cases:
        a, b, want = _bits(c["a"]), _bits(c["b"]), [int(x, 16) for x in c["out"]]
        got = bitvec_mul(a, b) if c["kind"] == "mul" else bitvec_rem(a, b)
        assert _words_of(got, len(want)) == want, c
        nz = np.nonzero(got)[0]
        assert (int(nz[-1]) if nz.size else 0) == c["degree"]
        # big-integer model (the generator) and the C oracle
        ia, ib = pm.from_words([int(x, 16) for x in c["a"]]), pm.from_words([int(x, 16) for x in c["b"]])
        assert pm.to_words(pm.clmul(ia, ib) if c["kind"] == "mul" else pm.polymod(ia, ib)) == want
        va = oracle.PolyVec.from_words([[int(x, 16) for x in c["a"]]])
        vb = oracle.PolyVec.from_words([[int(x, 16) for x in c["b"]]])
        r = oracle.poly_binop(oracle.POLY_MUL if c["kind"] == "mul" else oracle.POLY_REM, va, vb)
        assert [int(x) for x in r.words(0)] == want and r.degree(0) == c["degree"], c
        assert pm.from_words(r.buffer(0)) == pm.from_words(want)  # nothing above the tracked degree
