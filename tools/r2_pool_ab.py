"""A/B of the allocation pool on the u8 multiplier circuit (2^14 pairs): per-call wall clock, 12 calls."""
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
ts = []
p = None
for i in range(12):
    if p is not None:
        p.free()
    t0 = time.perf_counter(); p = ctx.apply2(hm.HomomorphicMultiplication, ca, cb); ctx.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print("HM_PRIVATE_POOL=%s HM_POOL_MAX_MB=%s  ms per call: %s  ok=%s" % (os.environ.get("HM_PRIVATE_POOL", "1"), os.environ.get("HM_POOL_MAX_MB", "default"),
      [round(x, 2) for x in ts], (ctx.decrypt(p) == a * b).mean()))
