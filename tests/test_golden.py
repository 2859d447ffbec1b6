"""Committed golden vectors (tests/golden/*.npz, minted by tests/golden/make_golden.py from the seeded oracle).
CPU: the oracle still reproduces them.  GPU: the CUDA engine reproduces them through the C ABI."""
import hashlib
import os

import numpy as np
import pytest

from helpers import expected_padded, oracle_encrypt

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = ["config_a_u32.npz", "config_a_u8.npz", "small_u8.npz"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def canonical_concat(row, widths):
    """Concatenation of the canonical words (degree/64+1, src/polynomial.rs:404-426) of every slot of one value."""
    out, off = [], 0
    for w in widths:
        s = row[off : off + w]
        nz = np.nonzero(s)[0]
        out.append(s[: (nz[-1] + 1) if nz.size else 1])
        off += w
    return np.concatenate(out)


@pytest.mark.parametrize("fname", FILES)
def test_oracle_reproduces_golden(oracle, fname):
    g = np.load(os.path.join(GOLDEN, fname))
    d, dp, delta, tau = (int(x) for x in g["params"])
    sk = oracle.PolyVec.zeros(1)
    sk.set_bytes(0, g["sk"].tobytes())
    pk = oracle.PolyVec.zeros(tau)
    for i in range(tau):
        pk.set_bytes(i, g["pk"][i].tobytes())
    a, b = g["a"], g["b"]
    n, L = a.size, a.dtype.itemsize * 8
    ca, cb = oracle_encrypt(oracle, pk, a, g["masks_a"]), oracle_encrypt(oracle, pk, b, g["masks_b"])
    wf = (d + dp) // 64 + 1
    np.testing.assert_array_equal(expected_padded(ca, n, [wf] * L), g["ct_a"])
    np.testing.assert_array_equal(expected_padded(cb, n, [wf] * L), g["ct_b"])
    s, _ = oracle.apply(oracle.OP_ADD, ca, cb, L, threads=oracle.max_threads())
    sw = expected_padded(s, n, list(g["sum_widths"]))
    assert [sha(sw[v]) for v in range(n)] == list(g["sum_sha256"])
    np.testing.assert_array_equal(sw[0], g["sum_first"])
    np.testing.assert_array_equal(oracle.decrypt(sk, s, L)[0], g["dec_sum"])
    if fname.startswith("config_a"):  # decryption of a sum is only probabilistically the plaintext sum
        np.testing.assert_array_equal(g["dec_sum"].view(a.dtype), a + b)
    mr, _ = oracle.poly_mulrem(ca, cb, sk)
    np.testing.assert_array_equal(expected_padded(mr, n, [(d - 1) // 64 + 1] * L), g["mulrem_words"])
    if "dec_mul" in g and fname.startswith("config_a"):
        np.testing.assert_array_equal(g["dec_mul"].view(a.dtype), a * b)


@pytest.mark.gpu
@pytest.mark.parametrize("fname", FILES)
def test_engine_reproduces_golden(fname):
    import homomorph_rust_b200 as hm

    assert hm.lib().hm_device_count() > 0
    g = np.load(os.path.join(GOLDEN, fname))
    d, dp, delta, tau = (int(x) for x in g["params"])
    ctx = hm.Context(hm.Parameters(d, dp, delta, tau))
    ctx.set_secret_key(hm.SecretKey.from_bytes(g["sk"].tobytes()))
    ctx.set_public_key(hm.PublicKey.from_bytes([r.tobytes() for r in g["pk"]]))
    a, b = g["a"], g["b"]
    n = a.size
    ca, cb = ctx.encrypt(a, g["masks_a"]), ctx.encrypt(b, g["masks_b"])
    np.testing.assert_array_equal(ca.to_host(), g["ct_a"])
    np.testing.assert_array_equal(cb.to_host(), g["ct_b"])
    np.testing.assert_array_equal(ctx.apply2(hm.HomomorphicXorGate, ca, cb).to_host(), g["xor_words"])
    np.testing.assert_array_equal(ctx.apply2(hm.HomomorphicAndGate, ca, cb).to_host(), g["and_words"])
    s = ctx.apply2(hm.HomomorphicAddition, ca, cb)
    assert list(s.slot_words()) == list(g["sum_widths"])
    sw = s.to_host()
    assert [sha(sw[v]) for v in range(n)] == list(g["sum_sha256"])
    np.testing.assert_array_equal(sw[0], g["sum_first"])
    np.testing.assert_array_equal(ctx.decrypt(s).view(np.uint8), g["dec_sum"])
    np.testing.assert_array_equal(ctx.poly_mulrem(ca, cb).to_host(), g["mulrem_words"])
    if "mul_sha256" in g:
        m = ctx.apply2(hm.HomomorphicMultiplication, ca, cb)
        mh, mw = m.to_host(), list(m.slot_words())
        assert [sha(canonical_concat(mh[v], mw)) for v in range(n)] == list(g["mul_sha256"])
        np.testing.assert_array_equal(ctx.decrypt(m).view(np.uint8), g["dec_mul"])
