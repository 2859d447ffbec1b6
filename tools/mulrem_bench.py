"""Times hm_poly_mulrem_into on fresh pairs (config A) and checks a sample against the oracle."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import homomorph_rust_b200 as hm
from oracle import hmoracle as orc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
cfg = (512, 512, 8, 256) if len(sys.argv) > 2 and sys.argv[2] == "B" else (128, 128, 1, 128)
MB = (cfg[3] + 7) // 8
rng = np.random.default_rng(3)
sk, pk = orc.keygen(*cfg, rng)
ctx = hm.Context(hm.Parameters(*cfg))
ctx.set_secret_key(hm.SecretKey.from_bytes(sk.words(0).astype("<u8").tobytes()))
ctx.set_public_key(hm.PublicKey.from_bytes([pk.words(i).astype("<u8").tobytes() for i in range(cfg[3])]))
lib = hm.lib()
a = rng.integers(0, 2**32, size=n, dtype=np.uint32)
b = rng.integers(0, 2**32, size=n, dtype=np.uint32)
ma = np.frombuffer(rng.bytes(n * 32 * MB), dtype=np.uint8)
mb = np.frombuffer(rng.bytes(n * 32 * MB), dtype=np.uint8)
ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
mr = ctx.poly_mulrem(ca, cb)
stream = torch.cuda.ExternalStream(lib.hm_context_stream(ctx._h))
into = lib.hm_poly_mulrem_into(ctx._h, ca._h, cb._h, mr._h) == 0
def run():
    global mr
    if into:
        lib.hm_poly_mulrem_into(ctx._h, ca._h, cb._h, mr._h)
    else:
        mr.free()
        mr = ctx.poly_mulrem(ca, cb)
for _ in range(3):
    run()
ctx.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record(stream)
for _ in range(reps):
    run()
e1.record(stream)
ctx.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"mulrem {cfg}: {n * 32} pairs in {ms:.3f} ms -> {n * 32 / ms / 1e6:.2f} G/s  (mode {os.environ.get('HM_MULREM_MODE', 'default')})")
# parity on the first 64 values
k = 64
oa = orc.encrypt(pk, np.frombuffer(a[:k].astype('<u4').tobytes(), dtype=np.uint8), 4, ma[: k * 32 * MB])[0]
ob = orc.encrypt(pk, np.frombuffer(b[:k].astype('<u4').tobytes(), dtype=np.uint8), 4, mb[: k * 32 * MB])[0]
want, _ = orc.poly_mulrem(oa, ob, sk)
W = (cfg[0] - 1) // 64 + 1
got = mr.to_host()[:k].reshape(-1, W)
exp = want.padded(W)
print("parity:", bool(np.array_equal(got, exp)))
