import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import homomorph_rust_b200 as hm
ctx = hm.Context(hm.Parameters(128, 128, 1, 128))
rng = np.random.default_rng(5)
sk = hm.SecretKey.random(128, rng); ctx.set_secret_key(sk); ctx.set_public_key(hm.PublicKey.random(128, 1, 128, sk, rng))
n = 16384
a = rng.integers(0, 256, size=n, dtype=np.uint8); b = rng.integers(0, 256, size=n, dtype=np.uint8)
ca, cb = ctx.encrypt(a, seed=1), ctx.encrypt(b, seed=2)
ref = None
bad = 0
for it in range(25):
    p = ctx.apply2(hm.HomomorphicMultiplication, ca, cb)
    if it % 3 == 0:
        q = ctx.apply2(hm.HomomorphicAddition, ca, cb)  # interleave other work on the same context
        q.free()
    h = np.bitwise_xor.reduce(p.to_host().reshape(-1, 8), axis=0)
    d = ctx.decrypt(p)
    if ref is None: ref = h
    if not np.array_equal(h, ref) or not np.array_equal(d, a * b): bad += 1
    p.free()
print("iterations with a differing result:", bad)
