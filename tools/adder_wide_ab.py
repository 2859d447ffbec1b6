"""Config B (D = 1024) u32 adds: the fused thread-per-value chain (adder_chain_wide_kernel) against the regrouped generic plan."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import homomorph_rust_b200 as hm
ctx = hm.Context(hm.Parameters(512, 512, 8, 256)); ctx.generate_keys_seeded(5)
lib = hm.lib()
rng = np.random.default_rng(1)
sizes = [int(x) for x in sys.argv[1:]] or [4096, 16384, 37888, 65536]
for n in sizes:
    a = rng.integers(0, 2**32, size=n, dtype=np.uint32); b = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    ca, cb = ctx.encrypt(a, seed=1), ctx.encrypt(b, seed=2)
    line = f"n={n:6d}"
    for name, wide, phases in (("generic", -1, 0), ("wide p1", 1, 1), ("wide p8", 1, 8), ("wide auto", 1, 0)):
        lib.hm_set_tuning(b"adder_wide_min", wide); lib.hm_set_tuning(b"adder_phases", phases)
        out = ctx.apply2(hm.HomomorphicAddition, ca, cb); ctx.synchronize()
        ts = []
        for _ in range(2):
            t0 = time.perf_counter(); assert lib.hm_apply2_into(ctx._h, 4, ca._h, cb._h, out._h) == 0; ctx.synchronize(); ts.append(time.perf_counter() - t0)
        ok = bool((ctx.decrypt(out) == (a + b)).all())
        line += f"  {name} {n / min(ts) / 1e3:7.1f} k/s ({min(ts) * 1e3:.0f} ms{'' if ok else ' WRONG'})"
        out.free()
    lib.hm_set_tuning(b"adder_wide_min", 0); lib.hm_set_tuning(b"adder_phases", 0)
    print(line, flush=True)
    ca.free(); cb.free()
