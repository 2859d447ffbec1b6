"""Pins the CPU oracle against every known-answer vector the reference holds for the hot path.

All vectors are transcribed from the reference's in-file unit tests
(/root/reference/src/polynomial.rs:428-613 and src/context.rs:615-635); the line each
one comes from is cited next to it.  The reference has no ciphertext-level vectors
 — those are covered by tests/test_oracle_model.py and the plaintext-level
expectations in tests/test_oracle_circuits.py.
"""
import numpy as np
import pytest

U64_MAX = (1 << 64) - 1
BITS = 64


def vec(oracle, *polys):
    return oracle.PolyVec.from_words(polys)


def test_new_empty_rejected(oracle):
    # src/polynomial.rs:432-437  #[should_panic = "The vector of coefficients must not be empty."]
    v = oracle.PolyVec.zeros(1)
    with pytest.raises(ValueError, match="must not be empty"):
        v.set_words(0, [])


def test_compute_degree(oracle):
    # src/polynomial.rs:439-449
    v = vec(oracle, [0b10010], [0b10010, 0b1], [0b10010, 0b0])
    assert v.degree(0) == 4
    assert v.degree(1) == BITS
    assert v.degree(2) == 4


def test_eq(oracle):
    # src/polynomial.rs:451-472
    long = [0b1001, 0b1000_0011_0101_1010, 0b0, 0b1, 0b0]
    v = vec(oracle, [0b1001], [0b1001], long, long, [0b1001, 0b0], [0b1000], [0b1000, 0b10, 0b0], [0b1000, 0b0, 0b0])
    assert v.eq(0, v, 1)
    assert v.eq(2, v, 3)
    assert v.eq(0, v, 4)  # trailing zero words are not part of the value
    assert not v.eq(0, v, 5)
    assert not v.eq(6, v, 7)


def test_monomial(oracle):
    # src/polynomial.rs:474-487
    v = oracle.PolyVec.zeros(3)
    for i, d in enumerate([5, BITS - 1, BITS]):
        v.set_monomial(i, d)
        assert v.degree(i) == d
        w = v.words(i)
        assert int(w[d // 64]) == 1 << (d % 64) and int(np.sum(w != 0)) == 1


def test_random_has_exact_degree(oracle):
    # src/polynomial.rs:489-496 (randomness injected instead of getrandom)
    rng = np.random.default_rng(1)
    v = oracle.PolyVec.zeros(2)
    for i, d in enumerate([5, BITS]):
        v.set_random(i, d, rng.integers(0, 256, 16, dtype=np.uint8).tobytes())
        assert v.degree(i) == d
        # compute_degree over the stored words agrees with the tracked degree
        w = vec(oracle, list(map(int, v.buffer(i))))
        assert w.degree(0) == d
    # all-ones fill: bits above the degree are cleared, the degree bit is forced (:89-90)
    v.set_random(0, 5, b"\xff" * 8)
    assert list(map(int, v.words(0))) == [0b111111]
    v.set_random(0, 5, b"\x00" * 8)
    assert list(map(int, v.words(0))) == [0b100000]


def test_evaluate(oracle):
    # src/polynomial.rs:511-520
    v = vec(oracle, [0b1001], [0b1111_00010, 0b1001])
    assert not v.evaluate(0, True)
    assert v.evaluate(0, False)
    assert v.evaluate(1, True)
    assert not v.evaluate(1, False)


def test_add(oracle):
    # src/polynomial.rs:522-535 — the reference compares the whole coefficient buffer
    r = oracle.poly_binop(oracle.POLY_ADD, vec(oracle, [0b1001]), vec(oracle, [0b0011]))
    assert list(map(int, r.buffer(0))) == [0b1010]
    r = oracle.poly_binop(oracle.POLY_ADD, vec(oracle, [0b1001, 0b1]), vec(oracle, [0b0101, 0b1]))
    assert list(map(int, r.buffer(0))) == [0b1100, 0b0]
    assert r.degree(0) == 3


def test_mul(oracle):
    # src/polynomial.rs:537-561
    cases = [
        ([0b1001], [0b11], [0b11011]),
        ([0b111], [0b11], [0b1001]),
        ([U64_MAX], [0b11], [0b1, 0b1]),  # carry across the word boundary
        ([0], [0b11], [0]),  # null polynomial
    ]
    for a, b, want in cases:
        r = oracle.poly_binop(oracle.POLY_MUL, vec(oracle, a), vec(oracle, b))
        assert list(map(int, r.buffer(0))) == want
    r = oracle.poly_binop(oracle.POLY_MUL, vec(oracle, [0]), vec(oracle, [0b11]))
    assert r.degree(0) == 0


def test_rem(oracle):
    # src/polynomial.rs:563-582
    cases = [
        ([0b1001], [0b11], [0]),
        ([0b1], [0b10], [1]),
        ([0b10_1010_1101], [0b11011], [0b1010]),
    ]
    for a, b, want in cases:
        d = vec(oracle, b)
        r = oracle.poly_binop(oracle.POLY_REM, vec(oracle, a), d)
        assert r.degree(0) < d.degree(0)
        assert list(map(int, r.buffer(0))) == want


def test_rem_zero_divisor(oracle):
    # src/polynomial.rs:584-590  #[should_panic = "attempt to divide by zero"]
    with pytest.raises(ValueError):
        oracle.poly_binop(oracle.POLY_REM, vec(oracle, [0b1001]), vec(oracle, [0]))


def test_byte_conversion(oracle):
    # src/polynomial.rs:606-612
    p = vec(oracle, [0b1001, 0b1000_0011_0101_1010, 0b0, 0b1, 0b0])
    q = oracle.PolyVec.zeros(1)
    q.set_bytes(0, p.to_bytes(0))
    assert p.eq(0, q, 0)
    # little-endian words, 8 bytes each (:99-105)
    assert p.to_bytes(0)[:16] == (0b1001).to_bytes(8, "little") + (0b1000_0011_0101_1010).to_bytes(8, "little")


def test_key_bytes_roundtrip(oracle):
    # src/context.rs:615-635 — from_bytes zero-pads a short tail (src/polynomial.rs:110-116)
    for raw in ([5, 14, 8], [4, 7, 5], [1, 2, 3], [5, 4, 6]):
        a = oracle.PolyVec.zeros(1)
        a.set_bytes(0, bytes(raw))
        b = oracle.PolyVec.zeros(1)
        b.set_bytes(0, a.to_bytes(0))
        assert a.eq(0, b, 0)
        assert int(a.words(0)[0]) == int.from_bytes(bytes(raw), "little")
