"""ctypes binding of libhmgpu.so (include/hmgpu.h).

The library is the product; this module only declares its C ABI.  There is no CPU
fallback: if the shared library is missing the import fails loudly, and every compute
entry point returns HM_ERR_CUDA when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhmgpu.so")
CSRC = os.path.join(_HERE, "csrc")

HM_OK = 0
HM_ERR_INVALID_PARAMETERS = -1
HM_ERR_PUBLIC_KEY_UNSET = -2
HM_ERR_SECRET_KEY_UNSET = -3
HM_ERR_OPERATION_REQUIREMENT = -4
HM_ERR_INVALID_LENGTH = -5
HM_ERR_CUDA = -6
HM_ERR_UNSUPPORTED = -7
HM_ERR_INVALID_ARGUMENT = -8
HM_ERR_DIVIDE_BY_ZERO = -9
HM_ERR_OUT_OF_MEMORY = -10

HM_OP_AND, HM_OP_OR, HM_OP_XOR, HM_OP_NOT, HM_OP_ADD, HM_OP_MUL = range(6)


def build(force: bool = False) -> str:
    """Compile csrc/ into libhmgpu.so for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in ("hmgpu.cu", "kernels_b.cu", "kernels_adder.cu", "kernels_adder.h", "kernels_mul.cu", "kernels_mul.h", "kernels_enc_umma.cu", "kernels_enc_umma.h", "probes.cu", "probes.h", "hmgroup.cu", "kernels.cuh", "gf2_blocks.cuh", "gf2host.hpp")]
    srcs.append(os.path.join(_HERE, "..", "include", "hmgpu.h"))
    stale = (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(s) for s in srcs)
    if force or stale:
        r = subprocess.run(["make", "-C", CSRC] + (["-B"] if force else []), capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building libhmgpu.so failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA engine has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C homomorph_rust_b200/csrc`). "
            "There is no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    vp, sz, u8p, u32p, u64p = C.c_void_p, C.c_size_t, C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    u16 = C.c_uint16
    sig = {
        "hm_status_string": (C.c_char_p, [C.c_int]),
        "hm_last_error": (C.c_char_p, [vp]),
        "hm_device_count": (C.c_int, []),
        "hm_context_create": (C.c_int, [u16, u16, u16, u16, C.c_int, C.POINTER(vp)]),
        "hm_context_destroy": (None, [vp]),
        "hm_context_parameters": (C.c_int, [vp, C.POINTER(u16), C.POINTER(u16), C.POINTER(u16), C.POINTER(u16)]),
        "hm_context_set_stream": (C.c_int, [vp, vp]),
        "hm_context_stream": (vp, [vp]),
        "hm_context_synchronize": (C.c_int, [vp]),
        "hm_context_kernel_launches": (C.c_uint64, [vp]),
        "hm_context_device": (C.c_int, [vp]),
        "hm_set_tuning": (C.c_int, [C.c_char_p, C.c_long]),
        "hm_set_secret_key": (C.c_int, [vp, vp, sz]),
        "hm_set_public_key": (C.c_int, [vp, C.POINTER(vp), C.POINTER(sz), sz]),
        "hm_generate_keys_seeded": (C.c_int, [vp, C.c_uint64]),
        "hm_key_stream_host": (C.c_int, [C.c_uint64, C.c_uint32, sz, vp]),
        "hm_secret_key_bytes": (C.c_int, [vp, vp, sz, C.POINTER(sz)]),
        "hm_public_key_bytes": (C.c_int, [vp, sz, vp, sz, C.POINTER(sz)]),
        "hm_has_secret_key": (C.c_int, [vp]),
        "hm_has_public_key": (C.c_int, [vp]),
        "hm_batch_len": (sz, [vp]),
        "hm_batch_bits": (C.c_uint32, [vp]),
        "hm_batch_value_words": (sz, [vp]),
        "hm_batch_slot_words": (C.c_int, [vp, u32p]),
        "hm_batch_slot_degree_bounds": (C.c_int, [vp, u64p]),
        "hm_batch_device_ptr": (vp, [vp]),
        "hm_batch_free": (None, [vp, vp]),
        "hm_batch_upload": (C.c_int, [vp, sz, C.c_uint32, u32p, vp, C.POINTER(vp)]),
        "hm_batch_upload_bounded": (C.c_int, [vp, sz, C.c_uint32, u64p, vp, C.POINTER(vp)]),
        "hm_batch_download": (C.c_int, [vp, vp, vp]),
        "hm_batch_download_range": (C.c_int, [vp, vp, sz, sz, vp]),
        "hm_batch_clone": (C.c_int, [vp, vp, C.POINTER(vp)]),
        "hm_batch_slice": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32, C.POINTER(vp)]),
        "hm_batch_concat": (C.c_int, [vp, C.POINTER(vp), C.c_size_t, C.POINTER(vp)]),
        "hm_apply2_fields": (C.c_int, [vp, C.c_int, vp, vp, C.POINTER(C.c_uint32), C.c_size_t, C.POINTER(vp)]),
        "hm_batch_serialized_size": (sz, [vp]),
        "hm_batch_serialize": (C.c_int, [vp, vp, vp, sz]),
        "hm_batch_deserialize": (C.c_int, [vp, vp, sz, C.POINTER(vp)]),
        "hm_batch_download_canonical": (C.c_int, [vp, vp, vp, sz, C.POINTER(sz)]),
        "hm_batch_upload_canonical": (C.c_int, [vp, sz, C.c_uint32, vp, sz, u64p, C.POINTER(vp)]),
        "hm_batch_wire_inspect": (C.c_int, [vp, sz, C.POINTER(u16), u32p, u64p, u64p]),
        "hm_host_alloc": (vp, [sz]),
        "hm_host_free": (None, [vp]),
        "hm_encrypt": (C.c_int, [vp, vp, sz, C.c_uint32, vp, C.POINTER(vp)]),
        "hm_encrypt_device": (C.c_int, [vp, vp, sz, C.c_uint32, vp, C.POINTER(vp)]),
        "hm_encrypt_device_into": (C.c_int, [vp, vp, sz, C.c_uint32, vp, vp]),
        "hm_encrypt_device_seeded_into": (C.c_int, [vp, vp, sz, C.c_uint32, C.c_uint64, C.c_uint64, vp]),
        "hm_masks_generate_host": (C.c_int, [u16, sz, C.c_uint64, vp]),
        "hm_masks_generate_device": (C.c_int, [vp, sz, C.c_uint64, vp]),
        "hm_encrypt_seeded": (C.c_int, [vp, vp, sz, C.c_uint32, C.c_uint64, C.POINTER(vp)]),
        "hm_encrypt_seeded_at": (C.c_int, [vp, vp, sz, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(vp)]),
        "hm_encrypt_async": (C.c_int, [vp, vp, sz, C.c_uint32, vp, C.POINTER(vp)]),
        "hm_decrypt_async": (C.c_int, [vp, vp, vp]),
        "hm_shard_range": (C.c_int, [sz, C.c_int, C.c_int, C.POINTER(sz), C.POINTER(sz)]),
        "hm_group_create": (C.c_int, [u16, u16, u16, u16, C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]),
        "hm_group_destroy": (None, [vp]),
        "hm_group_size": (C.c_int, [vp]),
        "hm_group_context": (vp, [vp, C.c_int]),
        "hm_group_set_secret_key": (C.c_int, [vp, vp, sz]),
        "hm_group_set_public_key": (C.c_int, [vp, C.POINTER(vp), C.POINTER(sz), sz]),
        "hm_group_synchronize": (C.c_int, [vp]),
        "hm_group_encrypt": (C.c_int, [vp, vp, sz, C.c_uint32, vp, C.POINTER(vp)]),
        "hm_group_encrypt_seeded": (C.c_int, [vp, vp, sz, C.c_uint32, C.c_uint64, C.POINTER(vp)]),
        "hm_group_apply2": (C.c_int, [vp, C.c_int, vp, vp, C.POINTER(vp)]),
        "hm_group_apply2_into": (C.c_int, [vp, C.c_int, vp, vp, vp]),
        "hm_group_apply1": (C.c_int, [vp, C.c_int, vp]),
        "hm_group_decrypt": (C.c_int, [vp, vp, vp]),
        "hm_group_batch_download": (C.c_int, [vp, vp, vp]),
        "hm_group_batch_len": (sz, [vp]),
        "hm_group_batch_bits": (C.c_uint32, [vp]),
        "hm_group_batch_part": (vp, [vp, C.c_int]),
        "hm_group_batch_free": (None, [vp]),
        "hm_decrypt": (C.c_int, [vp, vp, vp]),
        "hm_decrypt_device": (C.c_int, [vp, vp, vp]),
        "hm_apply2": (C.c_int, [vp, C.c_int, vp, vp, C.POINTER(vp)]),
        "hm_apply1": (C.c_int, [vp, C.c_int, vp]),
        "hm_apply2_unchecked": (C.c_int, [vp, C.c_int, vp, vp, C.POINTER(vp)]),
        "hm_apply2_generic": (C.c_int, [vp, C.c_int, vp, vp, C.POINTER(vp)]),
        "hm_apply2_into": (C.c_int, [vp, C.c_int, vp, vp, vp]),
        "hm_measure_alu_peak": (C.c_int, [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "hm_measure_kara8_peak": (C.c_int, [vp, C.POINTER(C.c_double)]),
        "hm_measure_pipe_peaks": (C.c_int, [vp, C.c_double, C.POINTER(C.c_double)]),
        "hm_op_min_d_over_delta": (C.c_int, [C.c_int]),
        "hm_result_slot_words": (C.c_int, [vp, C.c_int, C.c_uint32, u32p, u32p, u32p]),
        "hm_apply2_host": (C.c_int, [vp, C.c_int, sz, C.c_uint32, u32p, vp, u32p, vp, vp]),
        "hm_apply2_host_bounded": (C.c_int, [vp, C.c_int, sz, C.c_uint32, u64p, vp, u64p, vp, vp]),
        "hm_result_slot_bounds": (C.c_int, [C.c_int, C.c_uint32, u64p, u64p, u64p]),
        "hm_poly_add": (C.c_int, [vp, vp, vp, C.POINTER(vp)]),
        "hm_poly_mul": (C.c_int, [vp, vp, vp, C.POINTER(vp)]),
        "hm_poly_rem": (C.c_int, [vp, vp, C.POINTER(vp)]),
        "hm_poly_mulrem": (C.c_int, [vp, vp, vp, C.POINTER(vp)]),
        "hm_poly_mulrem_into": (C.c_int, [vp, vp, vp, vp]),
        "hm_fresh_slot_words": (C.c_uint32, [vp]),
        "hm_decrypt_vector": (C.c_int, [vp, sz, u64p]),
        "hm_poly_degree": (sz, [u64p, sz]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here = header/library mismatch, fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


SYMBOLS = None


def exported_symbols():
    """Names declared in include/hmgpu.h (parsed), for the ABI test."""
    import re

    hdr = open(os.path.join(_HERE, "..", "include", "hmgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)  # declarations only, not prose
    return sorted(set(re.findall(r"\b(hm_[a-z0-9_]+)\s*\(", hdr)))
