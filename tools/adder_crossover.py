"""Warp-per-value vs thread-per-value adder as a function of the batch size (config A, u32): where to switch."""
import os, sys, time, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import homomorph_rust_b200 as hm
ctx = hm.Context(hm.Parameters(128, 128, 1, 128))
rng = np.random.default_rng(1)
sk = hm.SecretKey.random(128, rng); ctx.set_secret_key(sk); ctx.set_public_key(hm.PublicKey.random(128, 1, 128, sk, rng))
lib = hm.lib()
for n in (2048, 4096, 8192, 12288, 16384, 24576, 32768, 49152, 75776):
    a = rng.integers(0, 2**32, size=n, dtype=np.uint32); b = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    ca, cb = ctx.encrypt(a, seed=1), ctx.encrypt(b, seed=2)
    res = []
    for thr in (1 << 40, 0):
        lib.hm_set_tuning(b"adder_thread_min", thr)
        out = ctx.apply2(hm.HomomorphicAddition, ca, cb); ctx.synchronize()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); assert lib.hm_apply2_into(ctx._h, 4, ca._h, cb._h, out._h) == 0; ctx.synchronize(); ts.append(time.perf_counter() - t0)
        res.append(n / min(ts))
        out.free()
    lib.hm_set_tuning(b"adder_thread_min", -1)
    print(f"n={n:6d}  warp kernel {res[0] / 1e6:5.2f} M/s   thread kernel {res[1] / 1e6:5.2f} M/s")
    ca.free(); cb.free()
