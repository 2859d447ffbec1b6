"""Shared helpers for the parity tests: seeded keys/masks for BOTH the oracle and the engine."""
from __future__ import annotations

import numpy as np


def keys(oracle, d, dp, delta, tau, seed):
    """Oracle key pair + the same keys as the byte formats the C ABI takes
    (SecretKey::to_bytes / PublicKey::to_bytes, src/polynomial.rs:99-105)."""
    rng = np.random.default_rng(seed)
    sk, pk = oracle.keygen(d, dp, delta, tau, rng)
    sk_bytes = sk.words(0).astype("<u8").tobytes()
    pk_bytes = [pk.words(i).astype("<u8").tobytes() for i in range(len(pk))]
    return sk, pk, sk_bytes, pk_bytes


def engine_context(hm, d, dp, delta, tau, sk_bytes, pk_bytes, device=0):
    ctx = hm.Context(hm.Parameters(d, dp, delta, tau), device=device)
    ctx.set_secret_key(hm.SecretKey.from_bytes(sk_bytes))
    ctx.set_public_key(hm.PublicKey.from_bytes(pk_bytes))
    return ctx


def oracle_encrypt(oracle, pk, values: np.ndarray, masks: np.ndarray):
    nbytes = values.dtype.itemsize
    data = np.frombuffer(values.astype(values.dtype.newbyteorder("<")).tobytes(), dtype=np.uint8)
    ct, _ = oracle.encrypt(pk, data, nbytes, masks, threads=oracle.max_threads())
    return ct


def expected_padded(polyvec, n, widths) -> np.ndarray:
    """Oracle polynomials (value-major, slot-minor) laid out like an engine batch."""
    L = len(widths)
    off = np.concatenate([[0], np.cumsum(widths)]).astype(int)
    out = np.zeros((n, off[-1]), dtype=np.uint64)
    for v in range(n):
        for k in range(L):
            w = polyvec.words(v * L + k)
            if w.size > widths[k]:
                assert not np.any(w[widths[k]:]), "oracle polynomial exceeds the engine's slot width"
                w = w[: widths[k]]
            out[v, off[k] : off[k] + w.size] = w
    return out


def random_polys(rng, n, nwords, top_bits=64) -> np.ndarray:
    a = rng.integers(0, 1 << 63, size=(n, nwords), dtype=np.uint64) * 2 + rng.integers(0, 2, size=(n, nwords), dtype=np.uint64)
    if top_bits < 64:
        a[:, -1] &= np.uint64((1 << top_bits) - 1)
    return a


def philox_masks(units: int, seed: int, first_unit: int = 0, mask_bytes: int = 16) -> np.ndarray:
    """The documented mask stream (include/hmgpu.h: Philox4x32-10, counter = (u_lo, u_hi, block, 0), key = seed), written
    independently of the engine in numpy: bytes [16 b, 16 b + 16) of bit-ciphertext u, for u in [first_unit, first_unit + units)."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    blocks = (mask_bytes + 15) // 16
    u = np.repeat(np.arange(units, dtype=np.uint64) + np.uint64(first_unit), blocks)
    c0 = (u & np.uint64(0xFFFFFFFF)).astype(np.uint64)
    c1 = (u >> np.uint64(32)).astype(np.uint64)
    c2 = np.tile(np.arange(blocks, dtype=np.uint64), units)
    c3 = np.zeros_like(c0)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    mask32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n1 = p1 & mask32
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        n3 = p0 & mask32
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    words = np.stack([c0, c1, c2, c3], axis=1).astype("<u4")  # (units * blocks, 4)
    by = words.view(np.uint8).reshape(units, blocks * 16)
    return np.ascontiguousarray(by[:, :mask_bytes]).reshape(-1)
