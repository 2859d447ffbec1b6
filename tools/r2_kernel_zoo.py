"""Runs every non-adder kernel of the bench once at its bench size, for ncu captures (tools/r2_run4.sh):
encrypt_tab6b (hm_encrypt_device_into), decrypt_uniform, decrypt_value_tma (after a small add), xor_flat, mul_small<8,8>,
mulrem_fresh_a (config A) and, with argument B, encrypt_tab<17,8,4,256> + mulrem_fresh32q (config B)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import homomorph_rust_b200 as hm
from homomorph_rust_b200 import _native as N

cfg_b = len(sys.argv) > 1 and sys.argv[1] == "B"
cfg = (512, 512, 8, 256) if cfg_b else (128, 128, 1, 128)
lib = hm.lib()
ctx = hm.Context(hm.Parameters(*cfg))
rng = np.random.default_rng(11)
sk = hm.SecretKey.random(cfg[0], rng); ctx.set_secret_key(sk); ctx.set_public_key(hm.PublicKey.random(cfg[1], cfg[2], cfg[3], sk, rng))
if cfg_b:
    n = 1 << 17
    a = rng.integers(0, 256, size=n, dtype=np.uint8); b = rng.integers(0, 256, size=n, dtype=np.uint8)
    ca, cb = ctx.encrypt(a, seed=1), ctx.encrypt(b, seed=2)   # mask_fill + encrypt_tab_kernel<17,8,4,256>
    mr = ctx.poly_mulrem(ca, cb)                              # mulrem_fresh32q_kernel, 2^20 pairs
    for _ in range(2):
        assert lib.hm_poly_mulrem_into(ctx._h, ca._h, cb._h, mr._h) == 0
    ctx.synchronize()
    # the fused D = 1024 adder chain on one exact wave of resident threads (2 CTAs x 128 threads per SM)
    nw = 148 * 256
    wa = rng.integers(0, 2**32, size=nw, dtype=np.uint32); wb = rng.integers(0, 2**32, size=nw, dtype=np.uint32)
    xa, xb = ctx.encrypt(wa, seed=5), ctx.encrypt(wb, seed=6)
    ws = ctx.apply2(hm.HomomorphicAddition, xa, xb)          # adder_chain_wide_kernel<2>
    ctx.synchronize()
    print("wide adder ok:", bool((ctx.decrypt(ws) == wa + wb).all()))
    print("zoo B done", ctx.kernel_launches())
    sys.exit(0)
n, L = 1 << 18, 32
a = rng.integers(0, 2**32, size=n, dtype=np.uint32); b = rng.integers(0, 2**32, size=n, dtype=np.uint32)
ca, cb = ctx.encrypt(a, seed=1), ctx.encrypt(b, seed=2)
dm = torch.from_numpy(np.frombuffer(rng.bytes(n * L * 16), dtype=np.uint8).copy()).cuda()
dv = torch.from_numpy(a.view(np.uint8).copy()).cuda()
dout = torch.empty(n * 4, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
for _ in range(2):
    assert lib.hm_encrypt_device_into(ctx._h, dv.data_ptr(), n, L, dm.data_ptr(), ca._h) == 0     # encrypt_tab4_kernel<false> (HM_ENC_MODE=1: encrypt_tab6b_kernel)
    assert lib.hm_encrypt_device_seeded_into(ctx._h, dv.data_ptr(), n, L, 12345, 0, ca._h) == 0    # encrypt_tab4_kernel<true> (Philox in the kernel)
    assert lib.hm_decrypt_device(ctx._h, ca._h, dout.data_ptr()) == 0                               # decrypt_uniform_kernel
xo = ctx.apply2(hm.HomomorphicXorGate, ca, cb)                                                       # xor_flat_kernel
ao = ctx.apply2(hm.HomomorphicAndGate, ca, cb)                                                       # mul_small_kernel<8,8>
mr = ctx.poly_mulrem(ca, cb)                                                                         # mulrem_fresh_a_kernel<1024>
for _ in range(2):
    assert lib.hm_apply2_into(ctx._h, N.HM_OP_XOR, ca._h, cb._h, xo._h) == 0
    assert lib.hm_apply2_into(ctx._h, N.HM_OP_AND, ca._h, cb._h, ao._h) == 0
    assert lib.hm_poly_mulrem_into(ctx._h, ca._h, cb._h, mr._h) == 0
# decrypt after add on 16 384 values (warp-per-value adder_fused_kernel, then decrypt_value_tma_kernel)
lib.hm_set_tuning(b"adder_thread_min", 1 << 40)
sa = ctx.encrypt(a[:16384], seed=3); sb = ctx.encrypt(b[:16384], seed=4)
s = ctx.apply2(hm.HomomorphicAddition, sa, sb)
d2 = torch.empty(16384 * 4, dtype=torch.uint8, device="cuda")
for _ in range(2):
    assert lib.hm_decrypt_device(ctx._h, s._h, d2.data_ptr()) == 0
ctx.synchronize()
print("zoo A done", ctx.kernel_launches(), bool((d2.cpu().numpy().view(np.uint32) == a[:16384] + b[:16384]).all()))
