"""Mints the multi-word known-answer vectors for Polynomial::mul / Polynomial::rem (reference src/polynomial.rs:252-365).

The reference's own KATs (src/polynomial.rs:439-612) stop at two words, so the carry paths of `mul` across several words
and the degree scan of `rem` are pinned here instead: operands of 3 and 5 words (plus ragged and sparse shapes) whose
products and remainders are computed by the big-integer model (oracle/pymodel.py) and, independently, by a numpy
bit-vector implementation (tests/test_oracle_model.py::bitvec_*).  The C oracle and the CUDA engine are then held to the
committed file.  Run from the repo root:  python tests/golden/make_kat_multiword.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pymodel as pm  # noqa: E402


def words(x: int):
    return [f"{w:016x}" for w in pm.to_words(x)]


def main():
    rng = np.random.default_rng(0x4B4154)  # "KAT"
    cases = []

    def rnd(nwords, top_bits=64):
        v = int.from_bytes(rng.bytes(8 * nwords), "little")
        if top_bits < 64:
            v &= (1 << (64 * (nwords - 1) + top_bits)) - 1
        return v | (1 << (64 * (nwords - 1) + top_bits - 1))  # exact length

    shapes = [(3, 3), (3, 5), (5, 5), (5, 3), (1, 5), (5, 1), (3, 2), (4, 5)]
    for na, nb in shapes:
        for top in (64, 1, 33):
            a, b = rnd(na, top), rnd(nb, 64 if top == 1 else top)
            cases.append({"kind": "mul", "a": words(a), "b": words(b), "out": words(pm.clmul(a, b)), "degree": pm.degree(pm.clmul(a, b))})
    # sparse / structured operands: all ones, single bits at word boundaries, alternating patterns
    allones3, allones5 = (1 << 192) - 1, (1 << 320) - 1
    for a, b in [(allones3, allones5), (allones5, allones5), (1 << 191, 1 << 319), ((1 << 128) | 1, (1 << 256) | (1 << 64) | 1),
                 (int("aaaaaaaaaaaaaaaa" * 3, 16), int("5555555555555555" * 5, 16)), (0, allones5), (allones3, 0), (1, allones5)]:
        cases.append({"kind": "mul", "a": words(a), "b": words(b), "out": words(pm.clmul(a, b)), "degree": pm.degree(pm.clmul(a, b))})
    # remainders: dividends of 3, 5 and 9 words by divisors of 1..3 words (exact degrees 64, 65, 127, 128, 129, 191)
    for ds in (5, 64, 65, 127, 128, 129, 191):
        for na in (3, 5, 9):
            s = (int.from_bytes(rng.bytes(8 * (ds // 64 + 1)), "little") & ((1 << ds) - 1)) | (1 << ds)
            a = rnd(na, 64)
            r = pm.polymod(a, s)
            cases.append({"kind": "rem", "a": words(a), "b": words(s), "out": words(r), "degree": pm.degree(r)})
    # dividend shorter than / equal to the divisor, and a dividend that is a multiple of the divisor (remainder 0)
    s = (1 << 128) | int.from_bytes(rng.bytes(16), "little")
    for a in (rnd(1), rnd(2), s, pm.clmul(s, rnd(3)), pm.clmul(s, rnd(3)) ^ 1):
        r = pm.polymod(a, s)
        cases.append({"kind": "rem", "a": words(a), "b": words(s), "out": words(r), "degree": pm.degree(r)})
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat_multiword.json")
    with open(out, "w") as f:
        json.dump({"format": "u64 words, LSB-first word order, hex; degree = highest set bit (0 for the null polynomial)", "cases": cases}, f, indent=0)
    print(len(cases), "cases ->", out)


if __name__ == "__main__":
    main()
