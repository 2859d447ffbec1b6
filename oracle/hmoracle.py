"""ctypes front-end of the CPU oracle (oracle/hm_oracle.c).

TEST INFRASTRUCTURE ONLY.  Nothing under ``homomorph_rust_b200`` may import this
module; it is used by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

The oracle restates the reference's algorithms (src/polynomial.rs, src/cipher.rs,
src/impls/numbers/common.rs, keygen of src/context.rs); see hm_oracle.h for the
parity status and the per-function citations.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libhmoracle.so")

OP_AND, OP_OR, OP_XOR, OP_NOT, OP_ADD, OP_MUL, OP_MUL_SIGNED = range(7)
POLY_ADD, POLY_MUL, POLY_REM = range(3)


def _host_signature() -> str:
    """The oracle is compiled with -march=native (mirroring the reference's target-cpu=native), so a
    library built on another machine may not run here: rebuild when the CPU changes."""
    import hashlib

    try:
        with open("/proc/cpuinfo") as f:
            lines = [l for l in f if l.startswith(("model name", "flags"))][:2]
    except OSError:
        lines = []
    return hashlib.sha256("".join(lines).encode()).hexdigest()


def build(force: bool = False) -> str:
    """Compile libhmoracle.so with the committed Makefile (gcc -O3 -march=native -flto)."""
    src = os.path.join(_HERE, "hm_oracle.c")
    stamp = _LIB_PATH + ".host"
    sig = _host_signature()
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "hm_oracle.h"))
    )
    try:
        stale = stale or open(stamp).read().strip() != sig
    except OSError:
        stale = True
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libhmoracle.so"], check=True, capture_output=True)
        with open(stamp, "w") as f:
            f.write(sig)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    try:
        L = C.CDLL(_LIB_PATH)
    except OSError:
        # built with -march=native on another host: rebuild for this one
        build(force=True)
        L = C.CDLL(_LIB_PATH)
    vp, sz, u8p, u64p = C.c_void_p, C.c_size_t, C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)
    dp = C.POINTER(C.c_double)
    sig = {
        "orc_vec_new": (vp, [sz]),
        "orc_vec_free": (None, [vp]),
        "orc_vec_len": (sz, [vp]),
        "orc_vec_degree": (sz, [vp, sz]),
        "orc_vec_buflen": (sz, [vp, sz]),
        "orc_vec_nwords": (sz, [vp, sz]),
        "orc_vec_words": (u64p, [vp, sz]),
        "orc_vec_set": (C.c_int, [vp, sz, u64p, sz]),
        "orc_vec_set_bytes": (C.c_int, [vp, sz, u8p, sz]),
        "orc_vec_set_random": (None, [vp, sz, sz, u8p]),
        "orc_vec_set_monomial": (None, [vp, sz, sz]),
        "orc_vec_evaluate": (C.c_int, [vp, sz, C.c_int]),
        "orc_vec_eq": (C.c_int, [vp, sz, vp, sz]),
        "orc_random_bytes_needed": (sz, [sz]),
        "orc_poly_binop": (vp, [C.c_int, vp, vp]),
        "orc_poly_mulrem": (vp, [vp, vp, vp]),
        "orc_keygen_sk": (vp, [sz, u8p]),
        "orc_keygen_pk": (vp, [sz, sz, sz, vp, u8p]),
        "orc_keygen_pk_bytes_needed": (sz, [sz, sz, sz]),
        "orc_encrypt": (vp, [vp, u8p, sz, u8p]),
        "orc_decrypt": (C.c_int, [vp, vp, u8p]),
        "orc_apply": (vp, [C.c_int, vp, vp, sz]),
        "orc_apply_timed": (vp, [C.c_int, vp, vp, sz, C.c_int, dp]),
        "orc_encrypt_timed": (vp, [vp, u8p, sz, sz, u8p, C.c_int, dp]),
        "orc_decrypt_timed": (C.c_int, [vp, vp, sz, sz, u8p, C.c_int, dp]),
        "orc_poly_mulrem_timed": (vp, [vp, vp, vp, C.c_int, dp]),
        "orc_max_threads": (C.c_int, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _u8(buf) -> Tuple[np.ndarray, "C._Pointer"]:
    arr = np.ascontiguousarray(np.frombuffer(bytes(buf), dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf, dtype=np.uint8)
    if arr.size == 0:
        arr = np.zeros(1, dtype=np.uint8)
    return arr, arr.ctypes.data_as(C.POINTER(C.c_uint8))


class PolyVec:
    """Owned list of oracle polynomials (each = tracked degree + LSB-first u64 words)."""

    def __init__(self, handle: int):
        if not handle:
            raise ValueError("oracle returned NULL (invalid arguments, zero or constant divisor)")
        self._h = C.c_void_p(handle)

    # -- construction ---------------------------------------------------------
    @classmethod
    def zeros(cls, n: int) -> "PolyVec":
        return cls(lib().orc_vec_new(n))

    @classmethod
    def from_words(cls, polys: Iterable[Sequence[int]]) -> "PolyVec":
        polys = [list(p) for p in polys]
        v = cls.zeros(len(polys))
        for i, p in enumerate(polys):
            v.set_words(i, p)
        return v

    @classmethod
    def from_padded(cls, arr: np.ndarray) -> "PolyVec":
        """One polynomial per row of a 2-D uint64 array (zero padded)."""
        arr = np.ascontiguousarray(arr, dtype=np.uint64)
        v = cls.zeros(arr.shape[0])
        L = lib()
        for i in range(arr.shape[0]):
            row = arr[i]
            if L.orc_vec_set(v._h, i, row.ctypes.data_as(C.POINTER(C.c_uint64)), row.size) != 0:
                raise ValueError("empty polynomial")
        return v

    def set_words(self, i: int, words: Sequence[int]) -> None:
        a = np.asarray(list(words), dtype=np.uint64)
        rc = lib().orc_vec_set(self._h, i, a.ctypes.data_as(C.POINTER(C.c_uint64)), a.size)
        if rc != 0:
            raise ValueError("The vector of coefficients must not be empty.")

    def set_bytes(self, i: int, data: bytes) -> None:
        arr, p = _u8(data)
        if lib().orc_vec_set_bytes(self._h, i, p, len(data)) != 0:
            raise ValueError("The vector of bytes must not be empty.")

    def set_random(self, i: int, degree: int, rnd: bytes) -> None:
        need = lib().orc_random_bytes_needed(degree)
        assert len(rnd) >= need
        arr, p = _u8(rnd)
        lib().orc_vec_set_random(self._h, i, degree, p)

    def set_monomial(self, i: int, degree: int) -> None:
        lib().orc_vec_set_monomial(self._h, i, degree)

    # -- inspection -------------------------------------------------------------
    def __len__(self) -> int:
        return lib().orc_vec_len(self._h)

    def degree(self, i: int) -> int:
        return lib().orc_vec_degree(self._h, i)

    def words(self, i: int) -> np.ndarray:
        """Canonical words: degree/64+1 of them (src/polynomial.rs:404-426 ignore the rest)."""
        n = lib().orc_vec_nwords(self._h, i)
        p = lib().orc_vec_words(self._h, i)
        return np.ctypeslib.as_array(p, shape=(n,)).copy()

    def buffer(self, i: int) -> np.ndarray:
        """The whole coefficient buffer, as `Polynomial::coefficients()` would return it."""
        n = lib().orc_vec_buflen(self._h, i)
        p = lib().orc_vec_words(self._h, i)
        return np.ctypeslib.as_array(p, shape=(n,)).copy()

    def to_bytes(self, i: int) -> bytes:
        """Polynomial::to_bytes — src/polynomial.rs:99-105 (whole buffer, little-endian words)."""
        return self.buffer(i).astype("<u8").tobytes()

    def evaluate(self, i: int, x: bool) -> bool:
        return bool(lib().orc_vec_evaluate(self._h, i, int(bool(x))))

    def eq(self, i: int, other: "PolyVec", j: int) -> bool:
        return bool(lib().orc_vec_eq(self._h, i, other._h, j))

    def padded(self, width_words: int, start: int = 0, count: Optional[int] = None) -> np.ndarray:
        """Rows of zero-padded canonical words — the layout the engine's batches use."""
        count = len(self) - start if count is None else count
        out = np.zeros((count, width_words), dtype=np.uint64)
        for r in range(count):
            w = self.words(start + r)
            if w.size > width_words:
                if np.any(w[width_words:]):
                    raise ValueError(f"polynomial {start + r} does not fit in {width_words} words")
                w = w[:width_words]
            out[r, : w.size] = w
        return out

    def max_nwords(self) -> int:
        L = lib()
        return max((L.orc_vec_nwords(self._h, i) for i in range(len(self))), default=1)

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and _lib is not None:
                _lib.orc_vec_free(self._h)
                self._h = None
        except Exception:
            pass


# ---------------------------------------------------------------------------- ops


def poly_binop(op: int, a: PolyVec, b: PolyVec) -> PolyVec:
    return PolyVec(lib().orc_poly_binop(op, a._h, b._h))


def poly_mulrem(a: PolyVec, b: PolyVec, s: PolyVec, threads: int = 1) -> Tuple[PolyVec, float]:
    sec = C.c_double(0.0)
    h = lib().orc_poly_mulrem_timed(a._h, b._h, s._h, threads, C.byref(sec))
    return PolyVec(h), sec.value


def keygen(d: int, dp: int, delta: int, tau: int, rng: np.random.Generator) -> Tuple[PolyVec, PolyVec]:
    """Seeded key pair: S then the tau public polynomials (src/context.rs:160-162, :249-261)."""
    L = lib()
    rnd = rng.integers(0, 256, size=L.orc_random_bytes_needed(d), dtype=np.uint8)
    sk = PolyVec(L.orc_keygen_sk(d, rnd.ctypes.data_as(C.POINTER(C.c_uint8))))
    rnd = rng.integers(0, 256, size=L.orc_keygen_pk_bytes_needed(dp, delta, tau), dtype=np.uint8)
    pk = PolyVec(L.orc_keygen_pk(dp, delta, tau, sk._h, rnd.ctypes.data_as(C.POINTER(C.c_uint8))))
    return sk, pk


def encrypt(pk: PolyVec, data: np.ndarray, bytes_per_value: int, masks: np.ndarray, threads: int = 1) -> Tuple[PolyVec, float]:
    """data: n_values*bytes_per_value LE bytes; masks: per bit ceil(tau/8) bytes, value-major, bit-minor."""
    data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
    masks = np.ascontiguousarray(masks, dtype=np.uint8).reshape(-1)
    n_values = data.size // bytes_per_value
    mask_bytes = (len(pk) + 7) // 8
    assert masks.size == n_values * bytes_per_value * 8 * mask_bytes, "mask stream has the wrong length"
    sec = C.c_double(0.0)
    h = lib().orc_encrypt_timed(
        pk._h, data.ctypes.data_as(C.POINTER(C.c_uint8)), n_values, bytes_per_value,
        masks.ctypes.data_as(C.POINTER(C.c_uint8)), threads, C.byref(sec),
    )
    return PolyVec(h), sec.value


def decrypt(sk: PolyVec, c: PolyVec, bits_per_value: int, threads: int = 1) -> Tuple[np.ndarray, float]:
    n = len(c)
    if bits_per_value % 8 != 0 or n % bits_per_value != 0:
        raise ValueError(f"InvalidCipheredLength {{ len: {n} }}")
    n_values = n // bits_per_value
    out = np.zeros(max(1, n_values * bits_per_value // 8), dtype=np.uint8)
    sec = C.c_double(0.0)
    rc = lib().orc_decrypt_timed(sk._h, c._h, n_values, bits_per_value, out.ctypes.data_as(C.POINTER(C.c_uint8)), threads, C.byref(sec))
    if rc != 0:
        raise ValueError(f"oracle decrypt failed ({rc})")
    return out[: n_values * bits_per_value // 8], sec.value


def apply(op: int, a: PolyVec, b: Optional[PolyVec], L: int, threads: int = 1) -> Tuple[PolyVec, float]:
    sec = C.c_double(0.0)
    h = lib().orc_apply_timed(op, a._h, b._h if b is not None else None, L, threads, C.byref(sec))
    return PolyVec(h), sec.value


def max_threads() -> int:
    return lib().orc_max_threads()
