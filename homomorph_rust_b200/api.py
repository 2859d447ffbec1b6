"""Host-side mirror of the reference's public surface for the GF(2)[X] hot path.

Names, argument meaning and error behaviour follow mathisbot/homomorph-rust (crate
`homomorph` v1.1.0); every class cites the reference item it stands for.  The one
deliberate difference is that everything is *batched*: ``Context.encrypt`` takes an array
of integers and returns a :class:`Ciphered` holding n values resident in HBM, and the
operations of ``src/operations.rs`` act on whole batches.  All polynomial arithmetic is
done by libhmgpu.so on the GPU; this module holds no arithmetic fallback.

Randomness: the reference draws key material and subset masks from ``getrandom``
(src/polynomial.rs:87, src/cipher.rs:95).  Here the caller may pass a seeded
``numpy.random.Generator`` (or explicit masks) so that runs are reproducible and can be
compared bit for bit with the oracle.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _native as N


# ------------------------------------------------------------------------------- errors
class ContextCryptoError(Exception):
    """reference src/context.rs:41-52"""


class PublicKeyUnset(ContextCryptoError):
    pass


class SecretKeyUnset(ContextCryptoError):
    pass


class CipherError(Exception):
    """reference src/cipher.rs:17-24"""


class InvalidCipheredLength(CipherError):
    pass


class OperationError(Exception):
    """reference src/operations.rs:11-18 (InvalidParameters)"""

    def __init__(self, required_min_d_over_delta: int, actual_d: int, actual_delta: int):
        super().__init__(
            f"invalid parameters: d/delta must be at least {required_min_d_over_delta} "
            f"(d = {actual_d}, delta = {actual_delta})"
        )
        self.required_min_d_over_delta = required_min_d_over_delta
        self.actual_d = actual_d
        self.actual_delta = actual_delta


class EngineError(RuntimeError):
    """CUDA failure / unsupported shape / bad argument reported by libhmgpu.so."""

    def __init__(self, status: int, detail: str = ""):
        msg = N.lib().hm_status_string(status).decode()
        super().__init__(f"{msg} [{status}]" + (f": {detail}" if detail else ""))
        self.status = status


# --------------------------------------------------------------------------- operations
class _Op:
    code: int
    MIN_D_OVER_DELTA: int  # reference src/operations.rs:24-27, src/impls/numbers.rs:27-50


class HomomorphicAndGate(_Op):
    code, MIN_D_OVER_DELTA = N.HM_OP_AND, 2


class HomomorphicOrGate(_Op):
    code, MIN_D_OVER_DELTA = N.HM_OP_OR, 2


class HomomorphicXorGate(_Op):
    code, MIN_D_OVER_DELTA = N.HM_OP_XOR, 1


class HomomorphicNotGate(_Op):
    code, MIN_D_OVER_DELTA = N.HM_OP_NOT, 1


class HomomorphicAddition(_Op):
    code, MIN_D_OVER_DELTA = N.HM_OP_ADD, 21


class HomomorphicMultiplication(_Op):
    code, MIN_D_OVER_DELTA = N.HM_OP_MUL, 64


# --------------------------------------------------------------------- host polynomials
# Key material only (tau+1 small polynomials once per context, SURVEY.md §2 "keygen stays on
# the host").  A polynomial is a Python int, bit i = coefficient of X^i.
def _clmul(a: int, b: int) -> int:
    if a.bit_length() > b.bit_length():
        a, b = b, a
    r = 0
    while a:
        low = a & -a
        r ^= b << (low.bit_length() - 1)
        a ^= low
    return r


def _random_poly(degree: int, rnd: bytes) -> int:
    """Polynomial::random (src/polynomial.rs:73-96): fill words, clear above the degree, force the top bit."""
    n = degree // 64 + 1
    v = int.from_bytes(rnd[: 8 * n], "little")
    return (v & ((1 << degree) - 1)) | (1 << degree)


def _poly_bytes(p: int) -> bytes:
    """Polynomial::to_bytes (src/polynomial.rs:99-105): degree/64+1 little-endian u64 words."""
    n = max(p.bit_length() - 1, 0) // 64 + 1
    return p.to_bytes(8 * n, "little")


def _rng_bytes(rng: Optional[np.random.Generator], n: int) -> bytes:
    if rng is None:
        import os

        return os.urandom(n)
    return rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()


# ------------------------------------------------------------------------------- types
class Parameters:
    """reference src/context.rs:33-119"""

    def __init__(self, d: int, dp: int, delta: int, tau: int):
        # the reference panics (src/context.rs:87-94)
        if not (0 < d <= 0xFFFF and 0 < dp <= 0xFFFF and 0 < delta <= 0xFFFF and 0 < tau <= 0xFFFF):
            raise ValueError("Parameters must be strictly positive (and fit in u16)")
        if not delta < d:
            raise ValueError("Delta must be strictly less than d")
        self._d, self._dp, self._delta, self._tau = d, dp, delta, tau

    @classmethod
    def new(cls, d: int, dp: int, delta: int, tau: int) -> "Parameters":
        return cls(d, dp, delta, tau)

    def d(self) -> int:
        return self._d

    def dp(self) -> int:
        return self._dp

    def delta(self) -> int:
        return self._delta

    def tau(self) -> int:
        return self._tau

    def __repr__(self):
        return f"Parameters(d={self._d}, dp={self._dp}, delta={self._delta}, tau={self._tau})"


class SecretKey:
    """reference src/context.rs:122-206"""

    def __init__(self, poly: int):
        self._p = poly

    @classmethod
    def from_bytes(cls, data: bytes) -> "SecretKey":
        if len(data) == 0:
            raise ValueError("The vector of bytes must not be empty.")
        return cls(int.from_bytes(bytes(data), "little"))

    @classmethod
    def random(cls, d: int, rng: Optional[np.random.Generator] = None) -> "SecretKey":
        return cls(_random_poly(d, _rng_bytes(rng, 8 * (d // 64 + 1))))  # src/context.rs:160-162

    def to_bytes(self) -> bytes:
        return _poly_bytes(self._p)

    def zeroize(self) -> None:
        self._p = 0


class PublicKey:
    """reference src/context.rs:209-298"""

    def __init__(self, polys: Sequence[int]):
        self._t = list(polys)

    @classmethod
    def from_bytes(cls, data: Sequence[bytes]) -> "PublicKey":
        return cls([int.from_bytes(bytes(b), "little") for b in data])

    @classmethod
    def random(cls, dp: int, delta: int, tau: int, secret_key: SecretKey, rng: Optional[np.random.Generator] = None) -> "PublicKey":
        """src/context.rs:249-261: T_i = S*Q_i + X*R_i; per i the bytes of Q are drawn before those of R."""
        polys = []
        for _ in range(tau):
            q = _random_poly(dp, _rng_bytes(rng, 8 * (dp // 64 + 1)))
            sq = _clmul(secret_key._p, q)
            r = _random_poly(delta, _rng_bytes(rng, 8 * (delta // 64 + 1)))
            polys.append(sq ^ (r << 1))
        return cls(polys)

    def to_bytes(self) -> List[bytes]:
        return [_poly_bytes(p) for p in self._t]

    def __len__(self):
        return len(self._t)


def _check(ctx_handle, rc: int) -> None:
    if rc == N.HM_OK:
        return
    detail = N.lib().hm_last_error(ctx_handle).decode() if ctx_handle else ""
    if rc == N.HM_ERR_PUBLIC_KEY_UNSET:
        raise PublicKeyUnset("PublicKeyUnset")
    if rc == N.HM_ERR_SECRET_KEY_UNSET:
        raise SecretKeyUnset("SecretKeyUnset")
    if rc == N.HM_ERR_INVALID_LENGTH:
        raise InvalidCipheredLength("InvalidCipheredLength")
    if rc == N.HM_ERR_DIVIDE_BY_ZERO:
        raise ZeroDivisionError("attempt to divide by zero")
    if rc == N.HM_ERR_OUT_OF_MEMORY:
        raise MemoryError(detail or "out of host or device memory")
    raise EngineError(rc, detail)


class Ciphered:
    """A batch of ``Ciphered<T>`` (reference src/cipher.rs:126-259): n values x L bit-ciphertexts in HBM."""

    def __init__(self, ctx: "Context", handle: int):
        self._ctx = ctx
        self._h = C.c_void_p(handle)

    # -- shape ------------------------------------------------------------------------
    def __len__(self) -> int:
        return N.lib().hm_batch_len(self._h)

    @property
    def bits(self) -> int:
        """Ciphered::len() of every value (src/cipher.rs:252-259 via Deref)."""
        return N.lib().hm_batch_bits(self._h)

    @property
    def value_words(self) -> int:
        return N.lib().hm_batch_value_words(self._h)

    def slot_words(self) -> np.ndarray:
        out = np.zeros(self.bits, dtype=np.uint32)
        _check(self._ctx._h, N.lib().hm_batch_slot_words(self._h, out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out

    def slot_degree_bounds(self) -> np.ndarray:
        out = np.zeros(self.bits, dtype=np.uint64)
        _check(self._ctx._h, N.lib().hm_batch_slot_degree_bounds(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out

    def device_ptr(self) -> int:
        return N.lib().hm_batch_device_ptr(self._h) or 0

    # -- data -------------------------------------------------------------------------
    def to_host(self) -> np.ndarray:
        """(n, value_words) uint64: slot k of value v is row v, columns off[k] .. off[k]+w[k]."""
        out = np.zeros((len(self), self.value_words), dtype=np.uint64)
        _check(self._ctx._h, N.lib().hm_batch_download(self._ctx._h, self._h, out.ctypes.data))
        return out

    def rows_to_host(self, first: int, count: int) -> np.ndarray:
        """Values [first, first + count) only: (count, value_words) uint64."""
        out = np.zeros((count, self.value_words), dtype=np.uint64)
        _check(self._ctx._h, N.lib().hm_batch_download_range(self._ctx._h, self._h, first, count, out.ctypes.data))
        return out

    def slot(self, host: np.ndarray, k: int) -> np.ndarray:
        w = self.slot_words()
        off = int(w[:k].sum())
        return host[:, off : off + int(w[k])]

    def to_bytes(self) -> bytes:
        """Engine wire format (include/hmgpu.h); the reference has no ciphertext serialisation."""
        size = N.lib().hm_batch_serialized_size(self._h)
        buf = np.zeros(size, dtype=np.uint8)
        _check(self._ctx._h, N.lib().hm_batch_serialize(self._ctx._h, self._h, buf.ctypes.data, size))
        return buf.tobytes()

    @classmethod
    def from_bytes(cls, ctx: "Context", data: bytes) -> "Ciphered":
        arr = np.frombuffer(data, dtype=np.uint8)
        out = C.c_void_p()
        rc = N.lib().hm_batch_deserialize(ctx._h, arr.ctypes.data, arr.size, C.byref(out))
        if rc == N.HM_ERR_INVALID_PARAMETERS:
            raise ValueError("batch was written under different parameters")
        _check(ctx._h, rc)
        return cls(ctx, out.value)

    def canonical(self):
        """[(degree, words)] per polynomial, value-major then slot-minor: the reference's Polynomial state."""
        cnt = C.c_size_t(0)
        _check(self._ctx._h, N.lib().hm_batch_download_canonical(self._ctx._h, self._h, None, 0, C.byref(cnt)))
        buf = np.zeros(cnt.value, dtype=np.uint64)
        _check(self._ctx._h, N.lib().hm_batch_download_canonical(self._ctx._h, self._h, buf.ctypes.data, buf.size, C.byref(cnt)))
        out, pos = [], 0
        while pos < cnt.value:
            deg = int(buf[pos])
            nw = deg // 64 + 1
            out.append((deg, buf[pos + 1 : pos + 1 + nw].copy()))
            pos += 1 + nw
        return out

    def clone(self) -> "Ciphered":
        out = C.c_void_p()
        _check(self._ctx._h, N.lib().hm_batch_clone(self._ctx._h, self._h, C.byref(out)))
        return Ciphered(self._ctx, out.value)

    def slice(self, first_bit: int, n_bits: int, dtype=None) -> "Ciphered":
        """Slots [first_bit, first_bit + n_bits) of every value as their own batch: `split_at` + `new_from_raw` on one
        field of a user struct (examples/simple_struct.rs:32-45)."""
        out = C.c_void_p()
        _check(self._ctx._h, N.lib().hm_batch_slice(self._ctx._h, self._h, first_bit, n_bits, C.byref(out)))
        r = Ciphered(self._ctx, out.value)
        r.dtype = np.dtype(dtype) if dtype is not None else None
        return r

    def free(self) -> None:
        if self._h is not None and self._h.value:
            N.lib().hm_batch_free(self._ctx._h, self._h)
            self._h = None

    def __del__(self):
        # safe in either order: a batch whose context is already destroyed is an orphan (its device memory went with the
        # context) and freeing it only releases the handle
        try:
            self.free()
        except Exception:
            pass


class Context:
    """reference src/context.rs:301-596 — parameters + optional keys, bound to one CUDA device."""

    def __init__(self, parameters: Parameters, device: int = 0):
        self._params = parameters
        self._sk: Optional[SecretKey] = None
        self._pk: Optional[PublicKey] = None
        h = C.c_void_p()
        rc = N.lib().hm_context_create(parameters.d(), parameters.dp(), parameters.delta(), parameters.tau(), device, C.byref(h))
        self._h = None
        if rc == N.HM_ERR_CUDA:
            raise EngineError(rc, f"no usable CUDA device {device} (the engine has no CPU fallback)")
        _check(None, rc)
        self._h = h

    @classmethod
    def new(cls, parameters: Parameters, device: int = 0) -> "Context":
        return cls(parameters, device)

    def close(self) -> None:
        if self._h is not None:
            N.lib().hm_context_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- keys ---------------------------------------------------------------------------
    def parameters(self) -> Parameters:
        return self._params

    def generate_secret_key(self, rng: Optional[np.random.Generator] = None) -> None:
        """src/context.rs:421-425"""
        self.set_secret_key(SecretKey.random(self._params.d(), rng))

    def generate_public_key(self, rng: Optional[np.random.Generator] = None) -> None:
        """src/context.rs:440-454 — SecretKeyUnset when there is no secret key."""
        if self._sk is None:
            raise SecretKeyUnset("SecretKeyUnset")
        p = self._params
        self.set_public_key(PublicKey.random(p.dp(), p.delta(), p.tau(), self._sk, rng))

    def generate_keys_seeded(self, seed: int) -> None:
        """Both keys from a 64-bit seed, the public polynomials computed on the GPU (hm_generate_keys_seeded).  Reproducible
        by the reference fed with hm_key_stream_host's bytes; for tests and benchmarks, not for production keys."""
        _check(self._h, N.lib().hm_generate_keys_seeded(self._h, seed))
        ln = C.c_size_t()
        _check(self._h, N.lib().hm_secret_key_bytes(self._h, None, 0, C.byref(ln)))
        buf = (C.c_uint8 * ln.value)()
        _check(self._h, N.lib().hm_secret_key_bytes(self._h, buf, ln.value, C.byref(ln)))
        self._sk = SecretKey.from_bytes(bytes(buf))
        rows = []
        for i in range(self._params.tau()):
            _check(self._h, N.lib().hm_public_key_bytes(self._h, i, None, 0, C.byref(ln)))
            b = (C.c_uint8 * ln.value)()
            _check(self._h, N.lib().hm_public_key_bytes(self._h, i, b, ln.value, C.byref(ln)))
            rows.append(bytes(b))
        self._pk = PublicKey.from_bytes(rows)

    def get_secret_key(self) -> Optional[SecretKey]:
        return self._sk

    def get_public_key(self) -> Optional[PublicKey]:
        return self._pk

    def set_secret_key(self, secret_key: SecretKey) -> None:
        """src/context.rs:568-571 — also clears the public key."""
        data = secret_key.to_bytes()
        buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
        _check(self._h, N.lib().hm_set_secret_key(self._h, C.addressof(buf), len(data)))
        self._sk = secret_key
        self._pk = None

    def set_public_key(self, public_key: PublicKey) -> None:
        """src/context.rs:592-595"""
        rows = public_key.to_bytes()
        bufs = [(C.c_uint8 * len(r)).from_buffer_copy(r) for r in rows]
        ptrs = (C.c_void_p * len(rows))(*[C.addressof(b) for b in bufs])
        lens = (C.c_size_t * len(rows))(*[len(r) for r in rows])
        _check(self._h, N.lib().hm_set_public_key(self._h, ptrs, lens, len(rows)))
        self._pk = public_key

    # -- encrypt / decrypt ----------------------------------------------------------------
    def mask_bytes(self) -> int:
        return (self._params.tau() + 7) // 8

    def seeded_masks(self, units: int, seed: int) -> np.ndarray:
        """The Philox4x32-10 mask stream `encrypt(..., seed=)` uses, computed on the host (for parity checks)."""
        out = np.zeros(units * self.mask_bytes(), dtype=np.uint8)
        _check(self._h, N.lib().hm_masks_generate_host(self._params.tau(), units, seed, out.ctypes.data))
        return out

    def encrypt(self, values: np.ndarray, masks: Optional[np.ndarray] = None, rng: Optional[np.random.Generator] = None,
                seed: Optional[int] = None) -> Ciphered:
        """Context::encrypt (src/context.rs:463-471) for an array of unsigned/signed integers.

        ``masks``: (n, L, ceil(tau/8)) uint8, the subset U of every bit (src/cipher.rs:92-97,106);
        drawn from ``rng`` (or the OS) when omitted.
        """
        if self._pk is None:
            raise PublicKeyUnset("PublicKeyUnset")
        values = np.ascontiguousarray(values)
        if values.dtype.kind == "V" and values.dtype.names:  # a struct of integers: fields in order, each little endian
            if any(values.dtype[f].kind not in "ui" or values.dtype[f].byteorder == ">" for f in values.dtype.names) or \
                    values.dtype.itemsize != sum(values.dtype[f].itemsize for f in values.dtype.names):
                raise TypeError("packed structs of little-endian integers only (bincode fixint, src/cipher.rs:6-13)")
            le = values
        elif values.dtype.kind not in "ui":
            raise TypeError("integers only (bincode fixint little-endian, src/cipher.rs:6-13)")
        else:
            le = values.astype(values.dtype.newbyteorder("<"), copy=False)
        n, L = values.size, values.dtype.itemsize * 8
        raw = np.frombuffer(le.tobytes(), dtype=np.uint8)
        if seed is not None:  # masks generated on the device from a counter-based PRNG; nothing but plaintext is uploaded
            out = C.c_void_p()
            _check(self._h, N.lib().hm_encrypt_seeded(self._h, raw.ctypes.data, n, L, seed, C.byref(out)))
            c = Ciphered(self, out.value)
            c.dtype = values.dtype
            return c
        if masks is None:
            masks = np.frombuffer(_rng_bytes(rng, n * L * self.mask_bytes()), dtype=np.uint8)
        masks = np.ascontiguousarray(masks, dtype=np.uint8).reshape(-1)
        if masks.size != n * L * self.mask_bytes():
            raise ValueError("masks must hold n * L * ceil(tau/8) bytes")
        out = C.c_void_p()
        _check(self._h, N.lib().hm_encrypt(self._h, raw.ctypes.data, n, L, masks.ctypes.data, C.byref(out)))
        c = Ciphered(self, out.value)
        c.dtype = values.dtype
        return c

    def decrypt(self, c: Ciphered, dtype=None) -> np.ndarray:
        """Context::decrypt (src/context.rs:480-488)."""
        if self._sk is None:
            raise SecretKeyUnset("SecretKeyUnset")
        L = c.bits
        if L % 8 != 0:
            raise InvalidCipheredLength(f"InvalidCipheredLength {{ len: {L} }}")
        out = np.zeros(len(c) * (L // 8), dtype=np.uint8)
        _check(self._h, N.lib().hm_decrypt(self._h, c._h, out.ctypes.data))
        dtype = dtype or getattr(c, "dtype", None)
        if dtype is not None and np.dtype(dtype).kind == "V":
            return out.view(np.dtype(dtype))
        dtype = dtype or (np.dtype(f"<u{L // 8}") if L // 8 in (1, 2, 4, 8) else np.dtype((np.uint8, (L // 8,))))
        return out.view(np.dtype(dtype).newbyteorder("<")).astype(dtype)

    # -- operations -----------------------------------------------------------------------
    def _validate(self, op) -> None:
        """src/context.rs:310-323"""
        d, delta = self._params.d(), self._params.delta()
        if d < op.MIN_D_OVER_DELTA * delta:
            raise OperationError(op.MIN_D_OVER_DELTA, d, delta)

    def apply1(self, op, a: Ciphered) -> None:
        """Context::apply1 (src/context.rs:496-507): in place."""
        self._validate(op)
        _check(self._h, N.lib().hm_apply1(self._h, op.code, a._h))

    def apply2(self, op, a: Ciphered, b: Ciphered, generic: bool = False) -> Ciphered:
        """Context::apply2 (src/context.rs:515-527)."""
        self._validate(op)
        out = C.c_void_p()
        fn = N.lib().hm_apply2_generic if generic else N.lib().hm_apply2
        _check(self._h, fn(self._h, op.code, a._h, b._h, C.byref(out)))
        r = Ciphered(self, out.value)
        r.dtype = getattr(a, "dtype", None)
        return r

    def concat(self, parts: Sequence[Ciphered], dtype=None) -> Ciphered:
        """The `extend_from_slice` merge of per-field results (examples/simple_struct.rs:52-58)."""
        arr = (C.c_void_p * len(parts))(*[p._h.value for p in parts])
        out = C.c_void_p()
        _check(self._h, N.lib().hm_batch_concat(self._h, arr, len(parts), C.byref(out)))
        r = Ciphered(self, out.value)
        r.dtype = np.dtype(dtype) if dtype is not None else None
        return r

    def apply2_fields(self, op, a: Ciphered, b: Ciphered, field_bits: Sequence[int]) -> Ciphered:
        """An operation on every field of a struct at once: field f is `field_bits[f]` consecutive bits (the Vec3Add of
        examples/simple_struct.rs with field_bits = [16, 16, 16])."""
        self._validate(op)
        fb = (C.c_uint32 * len(field_bits))(*[int(x) for x in field_bits])
        out = C.c_void_p()
        _check(self._h, N.lib().hm_apply2_fields(self._h, op.code, a._h, b._h, fb, len(field_bits), C.byref(out)))
        r = Ciphered(self, out.value)
        r.dtype = getattr(a, "dtype", None)
        return r

    # -- raw polynomial batches (crate-private Polynomial in the reference) -------------------
    def upload(self, host: np.ndarray, slot_words: Sequence[int], degree_bounds: Optional[Sequence[int]] = None) -> Ciphered:
        host = np.ascontiguousarray(host, dtype=np.uint64)
        L = len(slot_words)
        n = host.size // max(int(np.sum(slot_words)), 1)
        out = C.c_void_p()
        if degree_bounds is None:
            sw = np.asarray(slot_words, dtype=np.uint32)
            rc = N.lib().hm_batch_upload(self._h, n, L, sw.ctypes.data_as(C.POINTER(C.c_uint32)), host.ctypes.data, C.byref(out))
        else:
            db = np.asarray(degree_bounds, dtype=np.uint64)
            assert all(int(b) // 64 + 1 == int(w) for b, w in zip(db, slot_words))
            rc = N.lib().hm_batch_upload_bounded(self._h, n, L, db.ctypes.data_as(C.POINTER(C.c_uint64)), host.ctypes.data, C.byref(out))
        _check(self._h, rc)
        return Ciphered(self, out.value)

    def upload_canonical(self, polys, n: int, bits: int, degree_bounds: Optional[Sequence[int]] = None) -> Ciphered:
        """n * bits canonical polynomials [(degree, words)] (value-major, slot-minor) — e.g. exported by a Rust build of the
        reference, or by Ciphered.canonical() — as a device batch: the inverse of Ciphered.canonical()."""
        flat = []
        for deg, words in polys:
            flat.append(np.asarray([deg], dtype=np.uint64))
            flat.append(np.asarray(words, dtype=np.uint64))
        buf = np.ascontiguousarray(np.concatenate(flat) if flat else np.zeros(0, dtype=np.uint64))
        out = C.c_void_p()
        db = None if degree_bounds is None else np.asarray(degree_bounds, dtype=np.uint64)
        rc = N.lib().hm_batch_upload_canonical(self._h, n, bits, buf.ctypes.data, buf.size,
                                               None if db is None else db.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(out))
        _check(self._h, rc)
        return Ciphered(self, out.value)

    def poly_add(self, a: Ciphered, b: Ciphered) -> Ciphered:
        out = C.c_void_p()
        _check(self._h, N.lib().hm_poly_add(self._h, a._h, b._h, C.byref(out)))
        return Ciphered(self, out.value)

    def poly_mul(self, a: Ciphered, b: Ciphered) -> Ciphered:
        out = C.c_void_p()
        _check(self._h, N.lib().hm_poly_mul(self._h, a._h, b._h, C.byref(out)))
        return Ciphered(self, out.value)

    def poly_rem(self, a: Ciphered) -> Ciphered:
        out = C.c_void_p()
        _check(self._h, N.lib().hm_poly_rem(self._h, a._h, C.byref(out)))
        return Ciphered(self, out.value)

    def poly_mulrem(self, a: Ciphered, b: Ciphered) -> Ciphered:
        out = C.c_void_p()
        _check(self._h, N.lib().hm_poly_mulrem(self._h, a._h, b._h, C.byref(out)))
        return Ciphered(self, out.value)

    def synchronize(self) -> None:
        _check(self._h, N.lib().hm_context_synchronize(self._h))

    def kernel_launches(self) -> int:
        return N.lib().hm_context_kernel_launches(self._h)


class GroupCiphered:
    """One logical batch whose values are split by index over the devices of a ContextGroup."""

    def __init__(self, group: "ContextGroup", handle: int, dtype=None):
        self._g = group
        self._h = C.c_void_p(handle)
        self.dtype = dtype

    def __len__(self) -> int:
        return int(N.lib().hm_group_batch_len(self._h))

    @property
    def bits(self) -> int:
        return int(N.lib().hm_group_batch_bits(self._h))

    def part_len(self, i: int) -> int:
        return int(N.lib().hm_batch_len(N.lib().hm_group_batch_part(self._h, i)))

    def slot_words(self) -> np.ndarray:
        p = N.lib().hm_group_batch_part(self._h, 0)
        out = np.zeros(self.bits, dtype=np.uint32)
        _check(None, N.lib().hm_batch_slot_words(p, out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out

    def to_host(self) -> np.ndarray:
        """(n, value_words) u64, index order — the concatenation of the shards."""
        vw = int(self.slot_words().sum())
        out = np.zeros((len(self), vw), dtype=np.uint64)
        _check(None, N.lib().hm_group_batch_download(self._g._h, self._h, out.ctypes.data))
        return out

    def free(self) -> None:
        if self._h is not None and self._h.value:
            N.lib().hm_group_batch_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ContextGroup:
    """A Context replicated over several GPUs of one box (include/hmgpu.h "Device groups"): keys and tables on every
    device, batches split by value index into contiguous ranges, no collective — ciphertexts are independent
    (reference src/cipher.rs:180-185, :227-237)."""

    def __init__(self, parameters: Parameters, devices: Sequence[int]):
        self._params = parameters
        ids = (C.c_int * len(devices))(*[int(x) for x in devices])
        h = C.c_void_p()
        rc = N.lib().hm_group_create(parameters.d(), parameters.dp(), parameters.delta(), parameters.tau(), ids, len(devices), C.byref(h))
        self._h = None
        if rc == N.HM_ERR_CUDA:
            raise EngineError(rc, f"no usable CUDA device among {list(devices)} (the engine has no CPU fallback)")
        _check(None, rc)
        self._h = h
        self._sk = self._pk = None

    def __len__(self) -> int:
        return int(N.lib().hm_group_size(self._h))

    def close(self) -> None:
        if self._h is not None:
            N.lib().hm_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_secret_key(self, secret_key: SecretKey) -> None:
        raw = secret_key.to_bytes()
        buf = (C.c_uint8 * len(raw)).from_buffer_copy(raw)
        _check(None, N.lib().hm_group_set_secret_key(self._h, C.addressof(buf), len(raw)))
        self._sk, self._pk = secret_key, None

    def set_public_key(self, public_key: PublicKey) -> None:
        rows = public_key.to_bytes()
        bufs = [(C.c_uint8 * len(r)).from_buffer_copy(r) for r in rows]
        ptrs = (C.c_void_p * len(rows))(*[C.addressof(b) for b in bufs])
        lens = (C.c_size_t * len(rows))(*[len(r) for r in rows])
        _check(None, N.lib().hm_group_set_public_key(self._h, ptrs, lens, len(rows)))
        self._pk = public_key

    def synchronize(self) -> None:
        _check(None, N.lib().hm_group_synchronize(self._h))

    def encrypt(self, values: np.ndarray, masks: Optional[np.ndarray] = None, seed: Optional[int] = None) -> GroupCiphered:
        if self._pk is None:
            raise PublicKeyUnset("PublicKeyUnset")
        values = np.ascontiguousarray(values)
        if values.dtype.kind not in "ui":
            raise TypeError("integers only (bincode fixint little-endian, src/cipher.rs:6-13)")
        n, L = values.size, values.dtype.itemsize * 8
        raw = np.frombuffer(values.astype(values.dtype.newbyteorder("<"), copy=False).tobytes(), dtype=np.uint8)
        out = C.c_void_p()
        if seed is not None:
            _check(None, N.lib().hm_group_encrypt_seeded(self._h, raw.ctypes.data, n, L, seed, C.byref(out)))
        else:
            if masks is None:
                raise ValueError("masks or seed")
            masks = np.ascontiguousarray(masks, dtype=np.uint8).reshape(-1)
            if masks.size != n * L * ((self._params.tau() + 7) // 8):
                raise ValueError("masks must hold n * L * ceil(tau/8) bytes")
            _check(None, N.lib().hm_group_encrypt(self._h, raw.ctypes.data, n, L, masks.ctypes.data, C.byref(out)))
        return GroupCiphered(self, out.value, values.dtype)

    def apply2(self, op, a: GroupCiphered, b: GroupCiphered) -> GroupCiphered:
        out = C.c_void_p()
        rc = N.lib().hm_group_apply2(self._h, op.code, a._h, b._h, C.byref(out))
        if rc == N.HM_ERR_OPERATION_REQUIREMENT:
            raise OperationError(op.MIN_D_OVER_DELTA, self._params.d(), self._params.delta())
        _check(None, rc)
        return GroupCiphered(self, out.value, a.dtype)

    def apply1(self, op, a: GroupCiphered) -> None:
        _check(None, N.lib().hm_group_apply1(self._h, op.code, a._h))

    def decrypt(self, c: GroupCiphered, dtype=None) -> np.ndarray:
        if self._sk is None:
            raise SecretKeyUnset("SecretKeyUnset")
        L = c.bits
        out = np.zeros(len(c) * (L // 8), dtype=np.uint8)
        rc = N.lib().hm_group_decrypt(self._h, c._h, out.ctypes.data)
        if rc == N.HM_ERR_INVALID_LENGTH:
            raise InvalidCipheredLength(f"InvalidCipheredLength {{ len: {L} }}")
        _check(None, rc)
        dtype = dtype or c.dtype or np.dtype(f"<u{L // 8}")
        return out.view(np.dtype(dtype).newbyteorder("<")).astype(dtype)
