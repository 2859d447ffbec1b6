"""Independent big-integer model of the scheme (TEST INFRASTRUCTURE ONLY).

A second, deliberately different statement of the math in
README.md:97-133 of the reference: a polynomial over Z/2Z is a Python int whose
bit i is the coefficient of X^i.  It shares no code with oracle/hm_oracle.c and is
used to cross-check it (tests/test_oracle_model.py) — pure-Python loops, small
cases only.
"""
from __future__ import annotations

from typing import List, Sequence


def clmul(a: int, b: int) -> int:
    """Carry-less product (what `Polynomial::mul`, src/polynomial.rs:252-310, computes)."""
    if a.bit_length() > b.bit_length():
        a, b = b, a
    r = 0
    while a:
        low = a & -a
        r ^= b << (low.bit_length() - 1)
        a ^= low
    return r


def polymod(a: int, s: int) -> int:
    """Euclidean remainder (`Polynomial::rem`, src/polynomial.rs:316-365)."""
    if s == 0:
        raise ZeroDivisionError("attempt to divide by zero")
    ds = s.bit_length() - 1
    while a and a.bit_length() - 1 >= ds:
        a ^= s << (a.bit_length() - 1 - ds)
    return a


def degree(a: int) -> int:
    """Tracked degree; the null polynomial reports 0 (src/polynomial.rs:132-137)."""
    return max(a.bit_length() - 1, 0)


def to_words(a: int) -> List[int]:
    n = degree(a) // 64 + 1
    return [(a >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


def from_words(words: Sequence[int]) -> int:
    r = 0
    for i, w in enumerate(words):
        r |= int(w) << (64 * i)
    return r


def random_poly(deg: int, rnd: bytes) -> int:
    """`Polynomial::random` (src/polynomial.rs:73-96) fed with caller bytes."""
    n = deg // 64 + 1
    v = int.from_bytes(rnd[: 8 * n], "little")
    v &= (1 << deg) - 1
    return v | (1 << deg)


def keygen_pk(s: int, dp: int, delta: int, tau: int, rnd: bytes) -> List[int]:
    """T_i = S*Q_i + X*R_i (src/context.rs:249-261); Q bytes then R bytes per i."""
    out, off = [], 0
    nq, nr = (dp // 64 + 1) * 8, (delta // 64 + 1) * 8
    for _ in range(tau):
        q = random_poly(dp, rnd[off : off + nq])
        off += nq
        r = random_poly(delta, rnd[off : off + nr])
        off += nr
        out.append(clmul(s, q) ^ (r << 1))
    return out


def encrypt_bit(x: int, pk: Sequence[int], mask: bytes) -> int:
    """C = sum_{i in U} T_i + x (src/cipher.rs:99-115); U from the mask bytes (:106)."""
    c = 0
    for i, t in enumerate(pk):
        if mask[i // 8] & (1 << (i % 8)):
            c ^= t
    return c ^ (x & 1)


def decrypt_bit(c: int, s: int) -> int:
    """(C mod S)(0) (src/cipher.rs:119-122)."""
    return polymod(c, s) & 1


def encrypt_bytes(data: bytes, pk: Sequence[int], masks: bytes) -> List[int]:
    """LSB-first per byte (src/cipher.rs:180-185)."""
    mb = (len(pk) + 7) // 8
    out = []
    for j, byte in enumerate(data):
        for i in range(8):
            k = 8 * j + i
            out.append(encrypt_bit((byte >> i) & 1, pk, masks[k * mb : (k + 1) * mb]))
    return out


def decrypt_bytes(cts: Sequence[int], s: int) -> bytes:
    """src/cipher.rs:227-237."""
    assert len(cts) % 8 == 0
    out = bytearray(len(cts) // 8)
    for k, c in enumerate(cts):
        out[k // 8] |= decrypt_bit(c, s) << (k % 8)
    return bytes(out)


def add_circuit(a: Sequence[int], b: Sequence[int]) -> List[int]:
    """Ripple-carry adder of src/impls/numbers/common.rs:37-56."""
    res, carry = [], 0
    for i, (x, y) in enumerate(zip(a, b)):
        p = x ^ y
        res.append(p ^ carry)
        if i + 1 >= len(a):
            break
        cpp = clmul(p, carry)
        carry = cpp ^ clmul(clmul(x, y), cpp ^ 1)
    return res


def mul_circuit(a: Sequence[int], b: Sequence[int]) -> List[int]:
    """Column-serial multiplier of src/impls/numbers/common.rs:66-105."""
    n = len(a)
    res = [0] * n
    pp = [[clmul(x, y) for y in b] for x in a]
    carries: List[int] = []
    offset = 0
    for i in range(n):
        cur = i * (i + 1) // 2
        for j in range(i + 1):
            p = pp[j][i - j]
            if i + 1 < n:
                carries.append(clmul(p, res[i]))
            res[i] ^= p
        for j in range(cur):
            if i + 1 < n:
                carries.append(clmul(res[i], carries[offset + j]))
            res[i] ^= carries[offset + j]
        offset += cur
    return res


def decrypt_vector(s: int, nbits: int) -> int:
    """v with bit k = (X^k mod S)(0): decryption is parity(C AND v)."""
    v, cur = 0, 1
    ds = s.bit_length() - 1
    for k in range(nbits):
        v |= (cur & 1) << k
        cur <<= 1
        if (cur >> ds) & 1:
            cur ^= s
    return v
