"""u8 multiplier circuit at config A: the fused one-launch column multiplier (kernels_mul.cu) against the column-batched plan."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import homomorph_rust_b200 as hm
ctx = hm.Context(hm.Parameters(128, 128, 1, 128)); ctx.generate_keys_seeded(3)
lib = hm.lib()
rng = np.random.default_rng(2)
for n in [int(x) for x in sys.argv[1:]] or [64, 1024, 4096, 16384, 65536]:
    a = rng.integers(0, 256, size=n, dtype=np.uint8); b = rng.integers(0, 256, size=n, dtype=np.uint8)
    ca, cb = ctx.encrypt(a, seed=1), ctx.encrypt(b, seed=2)
    line = f"n={n:6d}"
    for name, fused in (("column plan", 0), ("fused", 1)):
        lib.hm_set_tuning(b"mul_circuit_fused", fused)
        out = ctx.apply2(hm.HomomorphicMultiplication, ca, cb); ctx.synchronize()
        ts = []
        for _ in range(4):
            l0 = ctx.kernel_launches(); t0 = time.perf_counter()
            assert lib.hm_apply2_into(ctx._h, 5, ca._h, cb._h, out._h) == 0; ctx.synchronize()
            ts.append(time.perf_counter() - t0)
        ok = bool((ctx.decrypt(out) == a * b).all())
        line += f"   {name}: {min(ts) * 1e3:8.3f} ms = {n / min(ts) / 1e6:6.3f} M muls/s, {ctx.kernel_launches() - l0} launches{'' if ok else ' WRONG'}"
        out.free()
    lib.hm_set_tuning(b"mul_circuit_fused", 1)
    print(line, flush=True)
    ca.free(); cb.free()
