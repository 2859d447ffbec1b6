import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import homomorph_rust_b200 as hm
ctx = hm.Context(hm.Parameters(128, 128, 1, 128))
rng = np.random.default_rng(1)
sk = hm.SecretKey.random(128, rng); ctx.set_secret_key(sk); ctx.set_public_key(hm.PublicKey.random(128, 1, 128, sk, rng))
n = 1 << 14
a = rng.integers(0, 256, size=n, dtype=np.uint8); b = rng.integers(0, 256, size=n, dtype=np.uint8)
ca, cb = ctx.encrypt(a, seed=1), ctx.encrypt(b, seed=2)
for mode, chunk in ((0, 32), (0, 24), (1, 32), (0, 32), (0, 24)):
    hm.lib().hm_set_tuning(b"mul_circuit_sequential", mode)
    hm.lib().hm_set_tuning(b"mul_thread_chunk", chunk)
    ts = []
    for i in range(15):
        ctx.synchronize()
        t0 = time.perf_counter(); p = ctx.apply2(hm.HomomorphicMultiplication, ca, cb); t1 = time.perf_counter(); ctx.synchronize(); t2 = time.perf_counter()
        ts.append((round((t1 - t0) * 1e3, 2), round((t2 - t0) * 1e3, 2)))
        p.free()
    tt = sorted(t[1] for t in ts[1:])
    print("sequential" if mode else "batched", chunk, "min %.2f median %.2f ms" % (tt[0], tt[len(tt) // 2]), "enqueue median %.2f" % sorted(t[0] for t in ts)[len(ts) // 2])
