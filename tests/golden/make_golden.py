"""Mints the committed golden vectors from the CPU oracle (seeded).  The reference holds no
ciphertext-level vectors and cannot be run here (Rust, no toolchain), so these pin the
oracle against regressions and give the GPU tests a fixture that does not depend on rebuilding it.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from oracle import hmoracle as orc  # noqa: E402
from helpers import expected_padded, keys, oracle_encrypt  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def adder_widths(D, L):
    return [D // 64 + 1] + [((3 * k - 1) * D) // 64 + 1 for k in range(1, L)]


def make(name, params, dtype, n, seed):
    d, dp, delta, tau = params
    sk, pk, skb, pkb = keys(orc, d, dp, delta, tau, seed)
    rng = np.random.default_rng(seed + 1)
    L = np.dtype(dtype).itemsize * 8
    a = rng.integers(0, np.iinfo(dtype).max, size=n, dtype=dtype, endpoint=True)
    b = rng.integers(0, np.iinfo(dtype).max, size=n, dtype=dtype, endpoint=True)
    mb = (tau + 7) // 8
    ma = rng.integers(0, 256, size=n * L * mb, dtype=np.uint8)
    mbk = rng.integers(0, 256, size=n * L * mb, dtype=np.uint8)
    ca, cb = oracle_encrypt(orc, pk, a, ma), oracle_encrypt(orc, pk, b, mbk)
    wf = (d + dp) // 64 + 1
    ct_a = expected_padded(ca, n, [wf] * L)
    ct_b = expected_padded(cb, n, [wf] * L)
    s, _ = orc.apply(orc.OP_ADD, ca, cb, L, threads=orc.max_threads())
    widths = adder_widths(d + dp, L)
    sum_words = expected_padded(s, n, widths)
    dec_sum, _ = orc.decrypt(sk, s, L)
    x, _ = orc.apply(orc.OP_XOR, ca, cb, L)
    an, _ = orc.apply(orc.OP_AND, ca, cb, L)
    mr, _ = orc.poly_mulrem(ca, cb, sk)
    out = dict(
        params=np.array(params), sk=np.frombuffer(skb, dtype=np.uint8), pk=np.stack([np.frombuffer(p, dtype=np.uint8) for p in pkb]),
        a=a, b=b, masks_a=ma, masks_b=mbk, ct_a=ct_a, ct_b=ct_b,
        sum_widths=np.array(widths), sum_sha256=np.array([sha(sum_words[v]) for v in range(n)]), sum_first=sum_words[0],
        dec_sum=dec_sum, xor_words=expected_padded(x, n, [wf] * L),
        and_words=expected_padded(an, n, [2 * (d + dp) // 64 + 1] * L),
        mulrem_words=expected_padded(mr, n, [(d - 1) // 64 + 1] * L),
    )
    if L == 8:
        m, _ = orc.apply(orc.OP_MUL, ca, cb, L, threads=orc.max_threads())
        mw = [m.max_nwords()] * L
        dm, _ = orc.decrypt(sk, m, L)
        mwords = expected_padded(m, n, mw)
        out.update(mul_sha256=np.array([sha(np.concatenate([m.words(v * L + k) for k in range(L)])) for v in range(n)]), dec_mul=dm)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    make("config_a_u32.npz", (128, 128, 1, 128), np.uint32, 3, 20261018)
    make("config_a_u8.npz", (128, 128, 1, 128), np.uint8, 3, 20261019)
    make("small_u8.npz", (64, 64, 1, 16), np.uint8, 4, 20261020)
