"""A/B of the two generic-adder plans (literal common.rs:44-53 vs regrouped) at d=d'=512 (config B), u32."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import homomorph_rust_b200 as hm
ctx = hm.Context(hm.Parameters(512, 512, 8, 256))
rng = np.random.default_rng(1)
sk = hm.SecretKey.random(512, rng); ctx.set_secret_key(sk); ctx.set_public_key(hm.PublicKey.random(512, 8, 256, sk, rng))
for n in (256, 4096, 16384):
    a = rng.integers(0, 2**32, size=n, dtype=np.uint32); b = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    ca, cb = ctx.encrypt(a, seed=1), ctx.encrypt(b, seed=2)
    for mode in (0, 1):
        hm.lib().hm_set_tuning(b"adder_generic_sequential", mode)
        ts = []
        for i in range(4):
            ctx.synchronize()
            t0 = time.perf_counter(); p = ctx.apply2(hm.HomomorphicAddition, ca, cb); ctx.synchronize(); t1 = time.perf_counter()
            ts.append(round((t1 - t0) * 1e3, 2))
            ok = bool((ctx.decrypt(p) == a + b).all()) if i == 0 else True
            p.free()
        print("n", n, "literal" if mode else "regrouped", ts, "correct" if ok else "WRONG", f"-> {n / (min(ts) * 1e-3):.0f} u32 adds/s")
