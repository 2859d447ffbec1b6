"""Config A encryption kernels side by side (HM_ENC_MODE is read once per process, so each mode runs in its own process):
device-timed encrypt with masks in HBM, seeded encrypt (masks from Philox), and seeded end to end from host plaintexts."""
import os, subprocess, sys, time
MODES = {"3": "tensor cores at config B: encrypt_umma_b_kernel (tcgen05 kind::i8, one launch); config A as mode 2", "2": "encrypt_tab4_kernel (+ Philox in the kernel)", "1": "encrypt_tab6b_kernel + mask_fill_kernel"}
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import ctypes as C
    import numpy as np, torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import homomorph_rust_b200 as hm
    cfg = (512, 512, 8, 256) if os.environ.get("ENC_AB_CFG") == "B" else (128, 128, 1, 128)
    ctx = hm.Context(hm.Parameters(*cfg)); ctx.generate_keys_seeded(1)
    lib = hm.lib()
    n, L = 1 << 18, 32
    rng = np.random.default_rng(3)
    a = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    dv = torch.from_numpy(a.view(np.uint8).copy()).cuda()
    dm = torch.from_numpy(np.frombuffer(rng.bytes(n * L * (cfg[3] // 8)), dtype=np.uint8).copy()).cuda()
    ce = ctx.encrypt(a, seed=1)
    def timed(fn, reps=20):
        fn(); ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = torch.cuda.ExternalStream(ctx.stream_handle()) if hasattr(ctx, "stream_handle") else None
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        ctx.synchronize()
        return (time.perf_counter() - t0) / reps
    s1 = timed(lambda: lib.hm_encrypt_device_into(ctx._h, dv.data_ptr(), n, L, dm.data_ptr(), ce._h))
    s2 = timed(lambda: lib.hm_encrypt_device_seeded_into(ctx._h, dv.data_ptr(), n, L, 12345, 0, ce._h))
    hv = torch.from_numpy(a.copy()).pin_memory()
    def e2e():
        o = C.c_void_p()
        assert lib.hm_encrypt_seeded(ctx._h, hv.data_ptr(), n, L, 12345, C.byref(o)) == 0
        lib.hm_batch_free(ctx._h, o)
    s3 = timed(e2e, reps=10)
    ok = bool((ctx.decrypt(ce) == a).all())
    print(f"  masks in HBM {s1 * 1e6:7.1f} us ({n / s1 / 1e9:.2f} G u32/s)   seeded {s2 * 1e6:7.1f} us ({n / s2 / 1e9:.2f} G u32/s)   "
          f"seeded from host plaintexts {s3 * 1e6:7.1f} us ({n / s3 / 1e9:.2f} G u32/s)   decrypts back: {ok}", flush=True)
else:
    for cfg in ("A", "B"):
        print(f"config {cfg} ({'d=dp=128, tau=128' if cfg == 'A' else 'd=dp=512, tau=256'}), 2^18 u32 = 8 388 608 bit-ciphertexts", flush=True)
        for mode, name in MODES.items():
            print(f" HM_ENC_MODE={mode}: {name if cfg == 'A' else name.replace('tab4_', 'tab4b_').replace('encrypt_tab6b_kernel', 'encrypt_tab_kernel<17,8,4,256>')}", flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=dict(os.environ, HM_ENC_MODE=mode, ENC_AB_CFG=cfg), check=False)
