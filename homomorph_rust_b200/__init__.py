"""homomorph_rust_b200 — B200-native batched ciphertext engine for the GF(2)[X] hot path of
mathisbot/homomorph-rust.  The product is libhmgpu.so (C ABI in include/hmgpu.h, CUDA kernels in
csrc/); this package is the host-side mirror of the reference's public API over it."""
from ._native import build, lib, LIB_PATH  # noqa: F401
from .api import (  # noqa: F401
    Ciphered,
    CipherError,
    Context,
    ContextCryptoError,
    EngineError,
    HomomorphicAddition,
    HomomorphicAndGate,
    HomomorphicMultiplication,
    HomomorphicNotGate,
    HomomorphicOrGate,
    HomomorphicXorGate,
    InvalidCipheredLength,
    OperationError,
    Parameters,
    PublicKey,
    PublicKeyUnset,
    SecretKey,
    SecretKeyUnset,
)
