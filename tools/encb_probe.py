"""Config B (d=d'=512, tau=256, delta=8): which encrypt/decrypt kernels run and how fast (device-resident masks)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import homomorph_rust_b200 as hm
ctx = hm.Context(hm.Parameters(512, 512, 8, 256))
rng = np.random.default_rng(1)
sk = hm.SecretKey.random(512, rng); ctx.set_secret_key(sk); ctx.set_public_key(hm.PublicKey.random(512, 8, 256, sk, rng))
n = 1 << 18
v = rng.integers(0, 2**32, size=n, dtype=np.uint32)
for rep in range(3):
    ctx.synchronize(); t0 = time.perf_counter(); c = ctx.encrypt(v, seed=5 + rep); ctx.synchronize(); t1 = time.perf_counter()
    d = ctx.decrypt(c); t2 = time.perf_counter()
    print(f"seeded encrypt {n} u32: {(t1 - t0) * 1e3:.2f} ms ({n * 32 * 136 / (t1 - t0) / 1e9:.0f} GB/s of ciphertext), decrypt {(t2 - t1) * 1e3:.2f} ms, ok={bool((d == v).all())}")
    c.free()
