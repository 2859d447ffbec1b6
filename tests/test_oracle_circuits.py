"""Plaintext-level expectations of the reference's own end-to-end tests, run through the oracle
with seeded keys and masks (the reference draws fresh random keys on every run and asserts on the
decrypted plaintext only — src/impls/numbers/uint.rs:109-293, src/cipher.rs:276-304)."""
import numpy as np
import pytest


def enc(oracle, pk, value, nbytes, rng):
    data = np.frombuffer(int(value).to_bytes(nbytes, "little"), dtype=np.uint8)
    mb = (len(pk) + 7) // 8
    ct, _ = oracle.encrypt(pk, data, nbytes, rng.integers(0, 256, nbytes * 8 * mb, dtype=np.uint8))
    return ct


def dec(oracle, sk, ct):
    out, _ = oracle.decrypt(sk, ct, len(ct))
    return int.from_bytes(out.tobytes(), "little")


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_gates(oracle, seed):
    rng = np.random.default_rng(seed)
    # uint.rs:109-137 at (32,8,8,8); uint.rs:140-174 at (32,16,16,16).  With delta = d/4 or d/2 a
    # product of two fresh bits is only *probably* right, so the AND/OR cases use delta=1 here.
    sk, pk = oracle.keygen(32, 8, 1, 8, rng)
    a, b = enc(oracle, pk, 0b1010, 1, rng), enc(oracle, pk, 0b1100, 1, rng)
    assert dec(oracle, sk, oracle.apply(oracle.OP_AND, a, b, 8)[0]) == 0b1000  # uint.rs:117-121
    assert dec(oracle, sk, oracle.apply(oracle.OP_OR, a, b, 8)[0]) == 0b1110  # uint.rs:133-137
    sk, pk = oracle.keygen(32, 16, 16, 16, rng)
    a, b = enc(oracle, pk, 0b1010, 1, rng), enc(oracle, pk, 0b1100, 1, rng)
    assert dec(oracle, sk, oracle.apply(oracle.OP_XOR, a, b, 8)[0]) == 0b0110  # uint.rs:149-153
    assert dec(oracle, sk, oracle.apply(oracle.OP_NOT, a, None, 8)[0]) == 0b1111_0101  # uint.rs:165-168
    assert dec(oracle, sk, oracle.apply(oracle.OP_NOT, b, None, 8)[0]) == 0b1111_0011  # uint.rs:170-173


@pytest.mark.parametrize("seed", [0, 1])
def test_addition(oracle, seed):
    # uint.rs:176-208 at (64,16,1,16)
    rng = np.random.default_rng(100 + seed)
    sk, pk = oracle.keygen(64, 16, 1, 16, rng)
    a, b = enc(oracle, pk, 22, 1, rng), enc(oracle, pk, 20, 1, rng)
    assert dec(oracle, sk, oracle.apply(oracle.OP_ADD, a, b, 8)[0]) == 42
    x, y = int(rng.integers(0, 1 << 15)), int(rng.integers(0, 1 << 15))
    a, b = enc(oracle, pk, x, 2, rng), enc(oracle, pk, y, 2, rng)
    assert dec(oracle, sk, oracle.apply(oracle.OP_ADD, a, b, 16)[0]) == x + y
    a, b = enc(oracle, pk, 255, 1, rng), enc(oracle, pk, 240, 1, rng)
    assert dec(oracle, sk, oracle.apply(oracle.OP_ADD, a, b, 8)[0]) == 239  # wrapping overflow


def test_multiplication_u8(oracle):
    # uint.rs:254-293 at (128,64,1,64)
    rng = np.random.default_rng(77)
    sk, pk = oracle.keygen(128, 64, 1, 64, rng)
    for x, y, want in [(6, 7, 42), (0, 151, 0), (int(rng.integers(0, 13)), int(rng.integers(0, 20)), None), (255, 240, 16)]:
        a, b = enc(oracle, pk, x, 1, rng), enc(oracle, pk, y, 1, rng)
        got = dec(oracle, sk, oracle.apply(oracle.OP_MUL, a, b, 8)[0])
        assert got == ((x * y) & 0xFF if want is None else want)


def test_roundtrip_and_length(oracle):
    # src/cipher.rs:276-304 at (64,32,8,32): u8 and usize round trip, len == 8 * bytes
    rng = np.random.default_rng(3)
    sk, pk = oracle.keygen(64, 32, 8, 32, rng)
    for value, nbytes in [(0b1010_1010, 1), (0x0123_4567_89AB_CDEF, 8)]:
        ct = enc(oracle, pk, value, nbytes, rng)
        assert len(ct) == 8 * nbytes
        assert dec(oracle, sk, ct) == value


def test_decrypt_invalid_length(oracle):
    # src/cipher.rs:218-220 CipherError::InvalidCipheredLength
    rng = np.random.default_rng(3)
    sk, pk = oracle.keygen(64, 32, 8, 32, rng)
    ct = enc(oracle, pk, 1, 1, rng)
    seven = oracle.PolyVec.from_words([ct.words(i) for i in range(7)])
    with pytest.raises(ValueError, match="InvalidCipheredLength"):
        oracle.decrypt(sk, seven, 7)


def test_adder_shapes(oracle):
    """Worst-case widths the engine allocates: deg s_k <= (3k-1)D for k>=2."""
    rng = np.random.default_rng(11)
    sk, pk = oracle.keygen(128, 128, 1, 128, rng)
    a, b = enc(oracle, pk, 0xDEADBEEF, 4, rng), enc(oracle, pk, 0x12345678, 4, rng)
    s, _ = oracle.apply(oracle.OP_ADD, a, b, 32)
    D = 256
    bound = [D, 2 * D] + [(3 * k - 1) * D for k in range(2, 32)]
    assert all(s.degree(k) <= bound[k] for k in range(32))
    assert dec(oracle, sk, s) == (0xDEADBEEF + 0x12345678) & 0xFFFFFFFF
