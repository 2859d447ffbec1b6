"""One u8 homomorphic multiply over 2^14 pairs (config 4) — for an ncu launch list of the circuit."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import homomorph_rust_b200 as hm
ctx = hm.Context(hm.Parameters(128, 128, 1, 128))
rng = np.random.default_rng(1)
sk = hm.SecretKey.random(128, rng); ctx.set_secret_key(sk); ctx.set_public_key(hm.PublicKey.random(128, 1, 128, sk, rng))
n = 1 << 14
a = rng.integers(0, 256, size=n, dtype=np.uint8); b = rng.integers(0, 256, size=n, dtype=np.uint8)
ca, cb = ctx.encrypt(a, seed=1), ctx.encrypt(b, seed=2)
for i in range(2):
    t0 = time.perf_counter(); p = ctx.apply2(hm.HomomorphicMultiplication, ca, cb); ctx.synchronize(); t1 = time.perf_counter()
    print(f"u8 mul x {n}: {(t1 - t0) * 1e3:.2f} ms", (ctx.decrypt(p) == a * b).mean())
