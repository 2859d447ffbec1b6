"""CPU-side checks of the product: the C-ABI library loads and exports every symbol of include/hmgpu.h,
argument validation that needs no GPU, and the host-side mirror of the reference API (keys, parameters)."""
import ctypes as C

import numpy as np
import pytest

import homomorph_rust_b200 as hm
from homomorph_rust_b200 import _native as N


def test_library_exports_every_declared_symbol():
    lib = hm.lib()
    names = N.exported_symbols()
    assert len(names) >= 45
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/hmgpu.h but not exported by libhmgpu.so"


def test_status_strings_and_requirements():
    lib = hm.lib()
    assert lib.hm_status_string(0) == b"ok"
    assert b"divide by zero" in lib.hm_status_string(N.HM_ERR_DIVIDE_BY_ZERO)
    # MIN_D_OVER_DELTA, reference src/impls/numbers.rs:27-50
    assert [lib.hm_op_min_d_over_delta(op) for op in range(6)] == [2, 2, 1, 1, 21, 64]
    assert lib.hm_op_min_d_over_delta(17) == N.HM_ERR_INVALID_ARGUMENT
    for op, want in [(hm.HomomorphicAndGate, 2), (hm.HomomorphicOrGate, 2), (hm.HomomorphicXorGate, 1),
                     (hm.HomomorphicNotGate, 1), (hm.HomomorphicAddition, 21), (hm.HomomorphicMultiplication, 64)]:
        assert op.MIN_D_OVER_DELTA == want == lib.hm_op_min_d_over_delta(op.code)


def test_parameter_validation_needs_no_gpu():
    lib = hm.lib()
    h = C.c_void_p()
    # Parameters::new asserts (src/context.rs:87-94; tests :602-613) are checked before the device is touched
    for bad in [(0, 1, 1, 1), (8, 0, 1, 1), (8, 1, 0, 1), (8, 1, 1, 0), (8, 4, 8, 4), (8, 4, 9, 4)]:
        assert lib.hm_context_create(*bad, 0, C.byref(h)) == N.HM_ERR_INVALID_PARAMETERS
    for bad in [(0, 1, 1, 1), (8, 4, 8, 4)]:
        with pytest.raises(ValueError):
            hm.Parameters(*bad)
    p = hm.Parameters.new(6, 3, 2, 5)  # doctest parameters, src/context.rs:25
    assert (p.d(), p.dp(), p.delta(), p.tau()) == (6, 3, 2, 5)


def test_no_cpu_fallback():
    """Without a CUDA device every compute path fails loudly (HM_ERR_CUDA), it never computes on the host."""
    lib = hm.lib()
    if lib.hm_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert lib.hm_context_create(128, 128, 1, 128, 0, C.byref(h)) == N.HM_ERR_CUDA
    with pytest.raises(hm.EngineError):
        hm.Context(hm.Parameters(128, 128, 1, 128))


def test_null_handles_are_rejected_without_a_gpu():
    """Every entry point checks its handles before touching CUDA: NULL context / batch -> HM_ERR_INVALID_ARGUMENT."""
    lib = hm.lib()
    out = C.c_void_p()
    bad = N.HM_ERR_INVALID_ARGUMENT
    assert lib.hm_batch_slice(None, None, 0, 8, C.byref(out)) == bad
    assert lib.hm_batch_concat(None, None, 0, C.byref(out)) == bad
    assert lib.hm_apply2_fields(None, N.HM_OP_ADD, None, None, None, 0, C.byref(out)) == bad
    assert lib.hm_apply2(None, N.HM_OP_ADD, None, None, C.byref(out)) == bad
    assert lib.hm_apply1(None, N.HM_OP_NOT, None) == bad
    assert lib.hm_decrypt(None, None, None) == bad
    assert lib.hm_set_tuning(b"no_such_knob", 1) == bad
    assert lib.hm_set_tuning(b"mul_thread_chunk", 17) == bad  # 24 or 32 only
    assert lib.hm_set_tuning(b"mul_thread_chunk", 32) == N.HM_OK


def test_poly_degree_helper():
    lib = hm.lib()
    u64p = C.POINTER(C.c_uint64)

    def deg(words):
        a = np.asarray(words, dtype=np.uint64)
        return lib.hm_poly_degree(a.ctypes.data_as(u64p), a.size)

    # src/polynomial.rs:440-449
    assert deg([0b1]) == 0 and deg([0b10]) == 1 and deg([0b1001]) == 3
    assert deg([0, 1]) == 64 and deg([0, 0]) == 0 and deg([1 << 63, 0, 0]) == 63


class _ByteStream:
    """numpy-Generator stand-in that serves a fixed byte string (so two keygens can share one stream)."""

    def __init__(self, data: bytes):
        self.data, self.pos = data, 0

    def integers(self, lo, hi, size, dtype):
        out = np.frombuffer(self.data[self.pos : self.pos + size], dtype=np.uint8)
        self.pos += size
        return out


@pytest.mark.parametrize("params", [(128, 128, 1, 128), (64, 32, 8, 32), (6, 3, 2, 5), (512, 512, 8, 16), (100, 70, 3, 9)])
def test_host_keygen_matches_oracle(oracle, params):
    """SecretKey::random / PublicKey::random (src/context.rs:160-162, :249-261) of the host mirror vs the oracle."""
    d, dp, delta, tau = params
    L = oracle.lib()
    rng = np.random.default_rng(sum(params))
    sk_rnd = rng.integers(0, 256, size=L.orc_random_bytes_needed(d), dtype=np.uint8).tobytes()
    pk_rnd = rng.integers(0, 256, size=L.orc_keygen_pk_bytes_needed(dp, delta, tau), dtype=np.uint8).tobytes()
    u8p = C.POINTER(C.c_uint8)
    a1 = np.frombuffer(sk_rnd, dtype=np.uint8).copy()
    a2 = np.frombuffer(pk_rnd, dtype=np.uint8).copy()
    osk = oracle.PolyVec(L.orc_keygen_sk(d, a1.ctypes.data_as(u8p)))
    opk = oracle.PolyVec(L.orc_keygen_pk(dp, delta, tau, osk._h, a2.ctypes.data_as(u8p)))
    sk = hm.SecretKey.random(d, _ByteStream(sk_rnd))
    pk = hm.PublicKey.random(dp, delta, tau, sk, _ByteStream(pk_rnd))
    assert sk.to_bytes() == osk.words(0).astype("<u8").tobytes()
    assert len(pk) == tau
    for i, row in enumerate(pk.to_bytes()):
        assert row == opk.words(i).astype("<u8").tobytes()
    # key byte round trips, src/context.rs:616-635
    assert hm.SecretKey.from_bytes(sk.to_bytes()).to_bytes() == sk.to_bytes()
    assert hm.PublicKey.from_bytes(pk.to_bytes()).to_bytes() == pk.to_bytes()


def test_philox_mask_stream_host():
    """hm_masks_generate_host: Philox4x32-10 known answers (Random123 kat_vectors) and the documented stream layout."""
    lib = hm.lib()
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85

    def philox(c, k):
        c, k = list(c), list(k)
        for _ in range(10):
            p0, p1 = M0 * c[0], M1 * c[2]
            c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
            k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
        return c

    assert philox([0] * 4, [0] * 2) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    for tau, units, seed in [(128, 5, 0), (256, 3, 0x1234_5678_9ABC_DEF0), (5, 9, 7), (200, 4, 2**64 - 1)]:
        mb = (tau + 7) // 8
        out = np.zeros(units * mb, dtype=np.uint8)
        assert lib.hm_masks_generate_host(tau, units, seed, out.ctypes.data) == 0
        want = bytearray()
        for u in range(units):
            row = bytearray()
            for b in range((mb + 15) // 16):
                for w in philox([u & 0xFFFFFFFF, u >> 32, b, 0], [seed & 0xFFFFFFFF, seed >> 32]):
                    row += int(w).to_bytes(4, "little")
            want += row[:mb]
        assert out.tobytes() == bytes(want)


def _wire(params, L, n, bounds, body_words=None, magic=b"HMB1"):
    import struct

    vw = sum(b // 64 + 1 for b in bounds)
    body = np.zeros(n * vw if body_words is None else body_words, dtype=np.uint64).tobytes()
    return magic + struct.pack("<4HIQ", *params, L, n) + struct.pack(f"<{len(bounds)}Q", *bounds) + body


def test_wire_header_is_validated_without_trusting_it():
    """hm_batch_wire_inspect (the header check of hm_batch_deserialize) on malformed buffers: crafted sizes must not wrap
    (n = 2^61 made n * vw * 8 overflow to 0 and the length check pass)."""
    import struct

    lib = hm.lib()
    P = (128, 128, 1, 128)

    def inspect(buf):
        prm = (C.c_uint16 * 4)()
        L, n, vw = C.c_uint32(), C.c_uint64(), C.c_uint64()
        rc = lib.hm_batch_wire_inspect(buf, len(buf), prm, C.byref(L), C.byref(n), C.byref(vw))
        return rc, tuple(prm), L.value, n.value, vw.value

    good = _wire(P, 2, 3, [256, 512])
    assert inspect(good) == (0, P, 2, 3, 5 + 9)
    assert inspect(_wire(P, 1, 0, [256]))[0] == 0  # empty batch
    assert inspect(good[:-8])[0] == N.HM_ERR_INVALID_LENGTH
    assert inspect(good + b"\0" * 8)[0] == N.HM_ERR_INVALID_LENGTH
    assert inspect(_wire(P, 2, 3, [256, 512], magic=b"HMB2"))[0] == N.HM_ERR_INVALID_ARGUMENT
    assert inspect(good[:10])[0] == N.HM_ERR_INVALID_ARGUMENT
    # n chosen so that n * value_words * 8 wraps to 0 (value_words = 1): 2^61 values and an empty body
    wrap = b"HMB1" + struct.pack("<4HIQ", *P, 1, 1 << 61) + struct.pack("<Q", 0)
    assert inspect(wrap)[0] == N.HM_ERR_INVALID_LENGTH
    wrap2 = b"HMB1" + struct.pack("<4HIQ", *P, 2, (1 << 64) // 16) + struct.pack("<2Q", 0, 0)  # vw = 2: n * 16 = 2^64
    assert inspect(wrap2)[0] == N.HM_ERR_INVALID_LENGTH
    # degree bounds beyond what the kernels' 32-bit slot offsets can describe
    assert inspect(_wire(P, 1, 0, [(1 << 31) + 1], body_words=0))[0] == N.HM_ERR_INVALID_ARGUMENT
    assert inspect(_wire(P, 1, 0, [1 << 40], body_words=0))[0] == N.HM_ERR_INVALID_ARGUMENT
    big = [1 << 31] * 128  # each allowed, the sum of widths (128 * 2^25 = 2^32) is not
    assert inspect(_wire(P, 128, 0, big, body_words=0))[0] == N.HM_ERR_INVALID_ARGUMENT
    # L = 0, L > 128, truncated bound table
    assert inspect(b"HMB1" + struct.pack("<4HIQ", *P, 0, 0))[0] == N.HM_ERR_INVALID_ARGUMENT
    assert inspect(b"HMB1" + struct.pack("<4HIQ", *P, 129, 0) + b"\0" * (129 * 8))[0] == N.HM_ERR_INVALID_ARGUMENT
    assert inspect(b"HMB1" + struct.pack("<4HIQ", *P, 4, 0) + b"\0" * 8)[0] == N.HM_ERR_INVALID_LENGTH
    assert lib.hm_batch_wire_inspect(None, 0, None, None, None, None) == N.HM_ERR_INVALID_ARGUMENT


def test_result_slot_bounds_follow_the_reference_degrees():
    """hm_result_slot_bounds = the degree recurrences of common.rs:37-105, no GPU needed."""
    lib = hm.lib()
    u64p = C.POINTER(C.c_uint64)

    def bounds(op, da, db):
        a, b = np.asarray(da, dtype=np.uint64), np.asarray(db, dtype=np.uint64)
        o = np.zeros(a.size, dtype=np.uint64)
        rc = lib.hm_result_slot_bounds(op, a.size, a.ctypes.data_as(u64p), b.ctypes.data_as(u64p), o.ctypes.data_as(u64p))
        return rc, [int(x) for x in o]

    D = 256
    assert bounds(N.HM_OP_XOR, [D, 100], [7, 300]) == (0, [D, 300])
    assert bounds(N.HM_OP_AND, [D] * 3, [D] * 3) == (0, [2 * D] * 3)
    assert bounds(N.HM_OP_OR, [D], [2 * D]) == (0, [3 * D])
    rc, add = bounds(N.HM_OP_ADD, [D] * 32, [D] * 32)
    assert rc == 0 and add == [D, 2 * D] + [(3 * k - 1) * D for k in range(2, 32)]
    rc, mul = bounds(N.HM_OP_MUL, [D] * 8, [D] * 8)
    assert rc == 0 and [m // D for m in mul] == [2, 2, 4, 6, 10, 18, 32, 56]
    assert bounds(N.HM_OP_MUL, [D] * 32, [D] * 32)[0] == N.HM_ERR_UNSUPPORTED  # u32 multiplication: infeasible growth
    assert bounds(N.HM_OP_NOT, [D], [D])[0] == N.HM_ERR_INVALID_ARGUMENT
    assert bounds(N.HM_OP_ADD, [1 << 40], [1])[0] == N.HM_ERR_INVALID_ARGUMENT


def test_shard_range_matches_the_python_rule():
    """hm_shard_range (C ABI, used by the device groups) == sharding.shard_range (used under torchrun): contiguous ranges in
    rank order whose sizes differ by at most one."""
    from homomorph_rust_b200.sharding import shard_range

    lib = hm.lib()
    for n in (0, 1, 7, 8, 9, 100, 262144, 262145):
        for world in (1, 2, 3, 4, 8):
            end = 0
            for r in range(world):
                f, c = C.c_size_t(), C.c_size_t()
                assert lib.hm_shard_range(n, r, world, C.byref(f), C.byref(c)) == 0
                assert (f.value, f.value + c.value) == shard_range(n, r, world)
                assert f.value == end
                end += c.value
            assert end == n
    f, c = C.c_size_t(), C.c_size_t()
    assert lib.hm_shard_range(10, 2, 2, C.byref(f), C.byref(c)) == N.HM_ERR_INVALID_ARGUMENT
    assert lib.hm_shard_range(10, 0, 0, C.byref(f), C.byref(c)) == N.HM_ERR_INVALID_ARGUMENT
    assert lib.hm_group_size(None) == 0 and lib.hm_group_create(128, 128, 1, 128, None, 0, None) == N.HM_ERR_INVALID_ARGUMENT


def test_philox_stream_independent_restatement():
    """The documented seeded-mask stream (include/hmgpu.h) computed by the library's host twin == a numpy restatement that
    shares no code with it (tests/helpers.philox_masks) — the GPU tests use the latter to check stream positions > 0."""
    import numpy as np

    from helpers import philox_masks
    from homomorph_rust_b200 import _native as N

    lib = N.lib()
    for tau, units, seed in ((128, 100, 12345678901234567), (200, 7, 99), (5, 9, 3), (256, 33, 2**64 - 1)):
        mb = (tau + 7) // 8
        out = np.zeros(units * mb, dtype=np.uint8)
        assert lib.hm_masks_generate_host(tau, units, seed, out.ctypes.data) == 0
        np.testing.assert_array_equal(philox_masks(units, seed, 0, mb), out)
    # a later stream position is the tail of a longer stream from 0
    full = philox_masks(50, 7, 0, 16)
    np.testing.assert_array_equal(philox_masks(20, 7, 30, 16), full[30 * 16:])
