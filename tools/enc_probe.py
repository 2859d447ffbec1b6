import ctypes as C, time, sys, os
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import homomorph_rust_b200 as hm
lib = hm.lib()
ctx = hm.Context(hm.Parameters(128,128,1,128))
rng = np.random.default_rng(1)
sk = hm.SecretKey.random(128, rng); ctx.set_secret_key(sk); ctx.set_public_key(hm.PublicKey.random(128,1,128,sk,rng))
n, L = 1<<18, 32
a = rng.integers(0, 2**32, size=n, dtype=np.uint32)
hv = torch.from_numpy(a.copy()).pin_memory()
def T(label, fn, reps=5):
    fn(); ctx.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ctx.synchronize(); ts.append((time.perf_counter()-t0)*1e3)
    print(label, ["%.2f" % t for t in ts])
def seeded():
    o = C.c_void_p(); assert lib.hm_encrypt_seeded(ctx._h, hv.data_ptr(), n, L, 12345, C.byref(o)) == 0; lib.hm_batch_free(ctx._h, o)
T("encrypt_seeded", seeded)
dm = torch.empty(n*L*16, dtype=torch.uint8, device="cuda")
T("masks_generate_device", lambda: lib.hm_masks_generate_device(ctx._h, n*L, 5, dm.data_ptr()))
dv = torch.from_numpy(a.view(np.uint8).copy()).cuda()
def encdev():
    o = C.c_void_p(); assert lib.hm_encrypt_device(ctx._h, dv.data_ptr(), n, L, dm.data_ptr(), C.byref(o)) == 0; lib.hm_batch_free(ctx._h, o)
T("encrypt_device(alloc+kernel+free)", encdev)
ce = ctx.encrypt(a[:n], seed=1)
T("encrypt_device_into", lambda: lib.hm_encrypt_device_into(ctx._h, dv.data_ptr(), n, L, dm.data_ptr(), ce._h))
hmk = torch.from_numpy(np.frombuffer(np.random.default_rng(2).bytes(n * L * 16), dtype=np.uint8).copy()).pin_memory()
def hostmasks():
    o = C.c_void_p(); assert lib.hm_encrypt(ctx._h, hv.data_ptr(), n, L, hmk.data_ptr(), C.byref(o)) == 0; lib.hm_batch_free(ctx._h, o)
T("encrypt(host masks, pinned)", hostmasks, reps=8)
big = ctx.encrypt(a, seed=3)
s = ctx.apply2(hm.HomomorphicAddition, big, big)   # 12 GB result, then freed: does the pool still behave?
s.free()
T("encrypt(host masks) after a 12 GB alloc/free", hostmasks, reps=8)
T("encrypt_seeded after", seeded, reps=5)
