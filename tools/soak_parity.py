"""Parity soak: the CUDA engine against the CPU oracle on larger seeded samples than the unit tests use, every output
word compared.  Prints one line per check; exit code 1 on any mismatch.  (python tools/soak_parity.py on a B200)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import homomorph_rust_b200 as hm  # noqa: E402
from helpers import engine_context, expected_padded, keys, oracle_encrypt, philox_masks  # noqa: E402
from oracle import hmoracle as orc  # noqa: E402

A = (128, 128, 1, 128)
ok = True
SEED = int(os.environ.get("SOAK_SEED", "42"))  # keys, plaintexts and masks all derive from it


def check(name, got, want, t0):
    global ok
    same = bool(np.array_equal(got, want))
    ok &= same
    print(f"{'OK  ' if same else 'FAIL'} {name:58s} {got.size:>12d} words compared  {time.time() - t0:6.1f} s", flush=True)


def rb(seed, n):
    return np.frombuffer(np.random.default_rng(seed + 1000 * SEED).bytes(n), dtype=np.uint8)


print(f"# soak seed {SEED}", flush=True)
sk, pk, skb, pkb = keys(orc, *A, 777 + SEED)
ctx = engine_context(hm, *A, skb, pkb)
lib = hm.lib()
rng = np.random.default_rng(SEED)
T = orc.max_threads()

# encrypt / decrypt
t0 = time.time(); n = 100_000
v = rng.integers(0, 2**32, size=n, dtype=np.uint32); m = rb(1, n * 512)
ct = ctx.encrypt(v, m); want = oracle_encrypt(orc, pk, v, m)
check("encrypt u32 (table kernel)", ct.to_host(), expected_padded(want, n, [5] * 32), t0)
check("decrypt fresh u32", ctx.decrypt(ct).view(np.uint8), orc.decrypt(sk, want, 32, threads=T)[0], t0)

# seeded encrypt: Philox masks drawn inside encrypt_tab4_kernel, from a stream position past 2^32
t0 = time.time(); n = 60_000
import torch  # noqa: E402
v = rng.integers(0, 2**32, size=n, dtype=np.uint32)
first = (1 << 32) + 12345 * SEED
out = ctx.encrypt(np.zeros(n, dtype=np.uint32), np.zeros(n * 512, dtype=np.uint8))
dv = torch.from_numpy(v.view(np.uint8).copy()).cuda(); torch.cuda.synchronize()
assert lib.hm_encrypt_device_seeded_into(ctx._h, dv.data_ptr(), n, 32, 99 + SEED, first, out._h) == 0
check("seeded encrypt u32 (Philox in the kernel, numpy stream)", out.to_host(), expected_padded(oracle_encrypt(orc, pk, v, philox_masks(n * 32, 99 + SEED, first)), n, [5] * 32), t0)
del out

# u32 add: thread-per-value kernel and warp-per-value kernel
for name, tune, n in (("u32 add, thread-per-value Karatsuba kernel", 0, 20_000), ("u32 add, warp-per-value comb kernel", 1 << 40, 6_000)):
    t0 = time.time()
    a = rng.integers(0, 2**32, size=n, dtype=np.uint32); b = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    ma, mb = rb(2 + n, n * 512), rb(3 + n, n * 512)
    lib.hm_set_tuning(b"adder_thread_min", tune)
    s = ctx.apply2(hm.HomomorphicAddition, ctx.encrypt(a, ma), ctx.encrypt(b, mb))
    lib.hm_set_tuning(b"adder_thread_min", -1)
    ws, _ = orc.apply(orc.OP_ADD, oracle_encrypt(orc, pk, a, ma), oracle_encrypt(orc, pk, b, mb), 32, threads=T)
    check(name, s.to_host(), expected_padded(ws, n, s.slot_words()), t0)
    check("  decrypt after add", ctx.decrypt(s).view(np.uint8), orc.decrypt(sk, ws, 32, threads=T)[0], t0)
    del s

# fused mul+rem
t0 = time.time(); n = 40_000
a = rng.integers(0, 2**32, size=n, dtype=np.uint32); b = rng.integers(0, 2**32, size=n, dtype=np.uint32)
ma, mb = rb(5, n * 512), rb(6, n * 512)
r = ctx.poly_mulrem(ctx.encrypt(a, ma), ctx.encrypt(b, mb))
wr, _ = orc.poly_mulrem(oracle_encrypt(orc, pk, a, ma), oracle_encrypt(orc, pk, b, mb), sk, threads=T)
check("mul+rem on fresh pairs (1.28 M pairs)", r.to_host(), expected_padded(wr, n, [2] * 32), t0)

# u8 multiplier circuit: the fused one-launch column multiplier (default), then the column-batched plan on a slice
t0 = time.time(); n = 20_000
a = rng.integers(0, 256, size=n, dtype=np.uint8); b = rng.integers(0, 256, size=n, dtype=np.uint8)
ma, mb = rb(7, n * 128), rb(8, n * 128)
p = ctx.apply2(hm.HomomorphicMultiplication, ctx.encrypt(a, ma), ctx.encrypt(b, mb))
wp, _ = orc.apply(orc.OP_MUL, oracle_encrypt(orc, pk, a, ma), oracle_encrypt(orc, pk, b, mb), 8, threads=T)
check("u8 multiply circuit (fused column multiplier)", p.to_host(), expected_padded(wp, n, p.slot_words()), t0)
check("  decrypt after mul", ctx.decrypt(p), orc.decrypt(sk, wp, 8, threads=T)[0], t0)
lib.hm_set_tuning(b"mul_circuit_fused", 0)
p2 = ctx.apply2(hm.HomomorphicMultiplication, ctx.encrypt(a[:4000], ma[: 4000 * 128]), ctx.encrypt(b[:4000], mb[: 4000 * 128]))
lib.hm_set_tuning(b"mul_circuit_fused", 1)
check("u8 multiply circuit (column-batched plan) == fused", p2.to_host(), p.to_host()[:4000], t0)
del p, p2

# AND / OR gates
t0 = time.time(); n = 30_000
a = rng.integers(0, 2**32, size=n, dtype=np.uint32); b = rng.integers(0, 2**32, size=n, dtype=np.uint32)
ma, mb = rb(9, n * 512), rb(10, n * 512)
ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
oa, ob = oracle_encrypt(orc, pk, a, ma), oracle_encrypt(orc, pk, b, mb)
for op, oop, nm in ((hm.HomomorphicAndGate, orc.OP_AND, "AND"), (hm.HomomorphicOrGate, orc.OP_OR, "OR")):
    g = ctx.apply2(op, ca, cb)
    wg, _ = orc.apply(oop, oa, ob, 32, threads=T)
    check(f"{nm} gate on fresh u32 batches", g.to_host(), expected_padded(wg, n, g.slot_words()), t0)

# config B (d = d' = 512, tau = 256, delta = 8): table encrypt, fused mul+rem, regrouped generic adder
del ctx
B = (512, 512, 8, 256)
sk, pk, skb, pkb = keys(orc, *B, 778 + SEED)
ctx = engine_context(hm, *B, skb, pkb)
t0 = time.time(); n = 20_000
a = rng.integers(0, 256, size=n, dtype=np.uint8); b = rng.integers(0, 256, size=n, dtype=np.uint8)
ma, mb = rb(11, n * 8 * 32), rb(12, n * 8 * 32)
ca, cb = ctx.encrypt(a, ma), ctx.encrypt(b, mb)
oa, ob = oracle_encrypt(orc, pk, a, ma), oracle_encrypt(orc, pk, b, mb)
check("config B: encrypt u8 (tensor-core encrypt_umma_b_kernel, host masks)", ca.to_host(), expected_padded(oa, n, [17] * 8), t0)
check("config B: decrypt fresh u8", ctx.decrypt(ca), orc.decrypt(sk, oa, 8, threads=T)[0], t0)
r = ctx.poly_mulrem(ca, cb)
wr, _ = orc.poly_mulrem(oa, ob, sk, threads=T)
check("config B: mul+rem on fresh pairs (160 k pairs)", r.to_host(), expected_padded(wr, n, [8] * 8), t0)
t0 = time.time(); n = 600
s = ctx.apply2(hm.HomomorphicAddition, ctx.encrypt(a[:n], ma[: n * 256]), ctx.encrypt(b[:n], mb[: n * 256]))
ws, _ = orc.apply(orc.OP_ADD, oracle_encrypt(orc, pk, a[:n], ma[: n * 256]), oracle_encrypt(orc, pk, b[:n], mb[: n * 256]), 8, threads=T)
check("config B: u8 add (regrouped generic plan)", s.to_host(), expected_padded(ws, n, s.slot_words()), t0)
check("  decrypt after add", ctx.decrypt(s), orc.decrypt(sk, ws, 8, threads=T)[0], t0)

# config B u32 add through the fused D = 1024 chain (forced on a batch the oracle can follow), seeded config-B encrypt
t0 = time.time(); n = 160
a = rng.integers(0, 2**32, size=n, dtype=np.uint32); b = rng.integers(0, 2**32, size=n, dtype=np.uint32)
ma, mb = philox_masks(n * 32, 5 + SEED, 7, 32), philox_masks(n * 32, 6 + SEED, 9, 32)
ea = ctx.encrypt(np.zeros(n, dtype=np.uint32), np.zeros(n * 32 * 32, dtype=np.uint8)); eb = ea.clone()
for batch, vals, seed, fu in ((ea, a, 5 + SEED, 7), (eb, b, 6 + SEED, 9)):
    dv = torch.from_numpy(vals.view(np.uint8).copy()).cuda(); torch.cuda.synchronize()
    assert lib.hm_encrypt_device_seeded_into(ctx._h, dv.data_ptr(), n, 32, seed, fu, batch._h) == 0
oa, ob = oracle_encrypt(orc, pk, a, ma), oracle_encrypt(orc, pk, b, mb)
check("config B: seeded encrypt u32 (tensor-core encrypt_umma_b_kernel)", ea.to_host(), expected_padded(oa, n, [17] * 32), t0)
lib.hm_set_tuning(b"adder_wide_min", 1)
s = ctx.apply2(hm.HomomorphicAddition, ea, eb)
lib.hm_set_tuning(b"adder_wide_min", 0)
ws, _ = orc.apply(orc.OP_ADD, oa, ob, 32, threads=T)
check("config B: u32 add (adder_chain_wide_kernel)", s.to_host(), expected_padded(ws, n, s.slot_words()), t0)
check("  decrypt after add", ctx.decrypt(s).view(np.uint8), orc.decrypt(sk, ws, 32, threads=T)[0], t0)

print("ALL OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
