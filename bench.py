#!/usr/bin/env python
"""bench.py — BASELINE.json's headline metric on B200: u32 homomorphic adds/s (config 3: ripple-carry
XOR/AND circuit on 2^18 encrypted pairs per GPU, d=d'=128, delta=1, tau=128), plus the second half of the metric
(GF(2)[X] mul+rem/s at d=d'=128, and the d=d'=512 sweep of config 5) and the secondary lines in `extra`.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One JSON line on stdout (rank 0).  A "step" is one pass of the fused adder over one batch of 2^18 synthetic
encrypted pairs per GPU (inputs 2 x 320 MiB, result 11.45 GiB: far larger than the 126 MB L2, so no L2 flush
is needed between iterations).  `value` is timed with CUDA events on the engine's stream with inputs resident in
HBM.  `e2e` goes through the host-buffer C-ABI call hm_apply2_host_bounded (pinned host CIPHERTEXTS in, pinned host
result out, copies inside the timed region — the drop-in for Context::apply2 on host-resident Ciphered<T>);
`e2e_circuit` is the flow of the reference's own bench (benches/u32.rs:8-50) with the ciphertexts staying in HBM:
host plaintexts -> encrypt -> add -> decrypt -> host plaintexts.  `--impl reference` times the CPU restatement of the
reference (oracle/, all host threads) on a bounded sample of the same workload — the reference itself is Rust and
cannot be built in this image (DESIGN.md "Oracle").
"""
from __future__ import annotations

import argparse
import ctypes as C
import faulthandler
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

faulthandler.enable()
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D, DP, DELTA, TAU = 128, 128, 1, 128
L = 32
METRIC = "u32 hom. adds/s"
UNIT = "adds/s"
WORKLOAD = ("configs[2]: u32 homomorphic add (ripple-carry XOR/AND circuit) on 2^18 encrypted pairs per GPU, "
            "d=dp=128, delta=1, tau=128")
# SURVEY.md §8(d) / §A.2 — algorithmic work of one u32 homomorphic add at D = d+d' = 256
BITMACS_PER_ADD = 2.751e8        # schoolbook AND-XOR pairs over the 93 reference multiplications
BYTES_PER_ADD = 2560 + 46912     # 2 x 32 x 5 words in, 5 864 words out
BITMACS_PER_MULREM = 66049 + 49665
BYTES_PER_MULREM = 96
# Executed instructions and DRAM traffic of the dominant kernel, from the ncu capture of the shipped adder_chain_kernel<8,0,4>
# at the bench size (profiles/r02_adder_chain_ncu.txt: one launch over 2^18 pairs, 58.12 ms under ncu), divided by 2^18.
# Warp-level instructions per add (one thread = one add, 32 adds per warp instruction).
NCU = {
    "source": "profiles/r02_adder_chain_ncu.txt (ncu --set full + pipe instruction counts, adder_chain_kernel<8,0,4>, 2^18 pairs, one launch)",
    "fmaheavy_warp_instr_per_add": 10643432704 / 262144,  # smsp__inst_executed_pipe_fmaheavy.sum: 40 601 (2 881 products x 432 IMAD.WIDE / 32 = 38 894, + moves)
    "alu_warp_instr_per_add": 20703528846 / 262144,       # smsp__inst_executed_pipe_alu.sum: 78 978 (LOP3 & co)
    "all_warp_instr_per_add": 32334274325 / 262144,       # smsp__inst_executed.sum: 123 345
    "dram_bytes_per_add": (15.329295e9 + 16.664368e9) / 262144,  # dram__bytes_read.sum + dram__bytes_write.sum: 122.0 KB (algorithmic: 49.5 KB)
    "pipe_pct_at_capture": {"alu": 61.24, "fmaheavy": 61.66, "adds_per_s_at_capture": 262144 / 58.120576e-3,
                            "note": "sm__pipe_{alu,fmaheavy}_cycles_active.avg.pct_of_peak_sustained_elapsed of that capture (4.51 M adds/s); ncu's 100 % "
                                    "is 2.0 (ALU) / 1.0 (IMAD.WIDE) warp-instructions per clock per SM; scaled by adds/s for this run"},
}
KARA8_PER_ADD = 1 + 30 * 3 + 6 * 465  # per bit: g (k=0), g, g_lo*p, g_hi*p (k=1..30); chain: sum_k k = 465 chunks x 6


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t_from=None, t_to=None):
        """Summarises the samples taken inside [t_from, t_to] (the timed region); the process itself is started before the
        warm-up steps so that nvidia-smi's start-up (NVML initialisation touches every GPU of the box) is not in the region."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (ts, r) in self.rows if t_from is None or (t_from <= ts <= t_to + 0.15)] or [r for (_, r) in self.rows]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for nme, val in zip(names, r[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def make_keys(hm, seed=2026):
    rng = np.random.default_rng(seed)
    sk = hm.SecretKey.random(D, rng)
    pk = hm.PublicKey.random(DP, DELTA, TAU, sk, rng)
    return sk, pk


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """CPU restatement of the reference's add_internal (oracle/hm_oracle.c), all host threads."""
    if rank != 0:
        return
    from oracle import hmoracle as orc

    threads = orc.max_threads()
    rng = np.random.default_rng(7)
    sk, pk = orc.keygen(D, DP, DELTA, TAU, rng)

    def raw(v):
        return np.frombuffer(v.astype("<u4").tobytes(), dtype=np.uint8)

    def enc(n):
        v = rng.integers(0, 2**32, size=n, dtype=np.uint32)
        m = rng.integers(0, 256, size=n * L * 16, dtype=np.uint8)
        return orc.encrypt(pk, raw(v), 4, m, threads=threads)[0]

    # calibrate: one add per thread
    a, b = enc(threads), enc(threads)
    _, sec = orc.apply(orc.OP_ADD, a, b, L, threads=threads)
    rate = threads / max(sec, 1e-9)
    total_steps = args.steps + args.warmup
    per_step_s = min(8.0, max(1.0, 120.0 / max(total_steps, 1)))
    n = max(threads, int(rate * per_step_s) // threads * threads)
    a, b = enc(n), enc(n)
    for _ in range(args.warmup):
        orc.apply(orc.OP_ADD, a, b, L, threads=threads)
    t = 0.0
    for _ in range(args.steps):
        _, sec = orc.apply(orc.OP_ADD, a, b, L, threads=threads)
        t += sec
    value = n * args.steps / t
    # the whole circuit of benches/u32.rs on the same cores: plaintexts -> encrypt x2 -> add -> decrypt -> plaintexts
    nc = max(threads, n // 2 // threads * threads)
    va = rng.integers(0, 2**32, size=nc, dtype=np.uint32)
    vb = rng.integers(0, 2**32, size=nc, dtype=np.uint32)
    ma = rng.integers(0, 256, size=nc * L * 16, dtype=np.uint8)
    mb = rng.integers(0, 256, size=nc * L * 16, dtype=np.uint8)
    t0 = time.perf_counter()
    ca, _ = orc.encrypt(pk, raw(va), 4, ma, threads=threads)
    cb, _ = orc.encrypt(pk, raw(vb), 4, mb, threads=threads)
    cs, _ = orc.apply(orc.OP_ADD, ca, cb, L, threads=threads)
    dec, _ = orc.decrypt(sk, cs, L, threads=threads)
    circuit_s = time.perf_counter() - t0
    circuit_ok = float(np.mean(dec.view("<u4") == (va + vb)))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 words of GF(2)[X]", "data": "synthetic (seeded keys, plaintexts and subset masks)",
        "config": {"workload": WORKLOAD, "pairs_per_gpu": 1 << 18, "bits": L, "pairs_per_step": n,
                   "note": "each step is a bounded sample of the 2^18-pair workload on the host cores; CPU port of the reference's "
                           "algorithms (the Rust reference cannot be built in this image)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n} pairs per step x {args.steps} steps, {threads} threads over independent values"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "e2e_circuit": {"value": nc / circuit_s, "unit": UNIT, "pairs": nc, "seconds": circuit_s, "correct_frac": circuit_ok,
                        "flow": "plaintexts -> encrypt x2 -> add -> decrypt -> plaintexts, all host threads (benches/u32.rs:8-50)"},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args, rank, local_rank, world):
    import torch

    import homomorph_rust_b200 as hm
    from homomorph_rust_b200 import _native as N

    lib = hm.lib()
    if lib.hm_device_count() <= local_rank:
        raise SystemExit("bench.py: no CUDA device for this rank — the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # pin this rank to the CPUs of its GPU's NUMA node so that pinned host buffers are local to the PCIe root
    # (matters at N > 1 on a two-socket box); restored before the CPU baseline runs
    all_cpus = os.sched_getaffinity(0)
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node >= 0:
            cpus = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            cpus &= all_cpus
            if cpus:
                os.sched_setaffinity(0, cpus)
    except Exception:
        pass
    ctx = hm.Context(hm.Parameters(D, DP, DELTA, TAU), device=local_rank)
    sk, pk = make_keys(hm)  # the public key is replicated on every GPU; no collective on the hot path
    ctx.set_secret_key(sk)
    ctx.set_public_key(pk)
    stream = torch.cuda.ExternalStream(lib.hm_context_stream(ctx._h), device=dev)

    n = args.pairs
    rng = np.random.default_rng(1000 + rank)
    a = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    b = rng.integers(0, 2**32, size=n, dtype=np.uint32)

    def enc(v, seed):
        g = np.random.default_rng(seed)
        # masks are host generated (reproducible): 16 B per bit-ciphertext
        m = np.frombuffer(g.bytes(v.size * L * 16), dtype=np.uint8)
        return ctx.encrypt(v, m)

    ca, cb = enc(a, 5000 + rank), enc(b, 6000 + rank)
    out = ctx.apply2(hm.HomomorphicAddition, ca, cb)  # allocates the 11.45 GiB result once; also a warm-up

    def step():
        rc = lib.hm_apply2_into(ctx._h, N.HM_OP_ADD, ca._h, cb._h, out._h)
        if rc != 0:
            raise RuntimeError(f"hm_apply2_into failed: {rc} {lib.hm_last_error(ctx._h).decode()}")

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    ctx.synchronize()
    barrier()
    t_region0 = time.time()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    l_before = ctx.kernel_launches()
    ev[0].record(stream)
    for i in range(args.steps):
        step()
        ev[i + 1].record(stream)
    ctx.synchronize()
    barrier()
    clocks = sampler.stop(t_region0, time.time()) if rank == 0 else None
    l_after = ctx.kernel_launches()
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms_max = max_over_ranks(total_ms)
    value = world * n * args.steps / (total_ms_max * 1e-3)

    # sanity outside the timed region: the decrypted sums (probabilistically exact at delta=1)
    dec = ctx.decrypt(out)
    frac_ok = float(np.mean(dec == (a + b)))

    # ---- end to end through the host-buffer C-ABI call (ciphertexts in host memory) ------------------------------
    ne = min(args.e2e_pairs, n)
    fresh = np.full(L, D + DP, dtype=np.uint64)
    wo = out.slot_words().astype(np.uint32)
    vwo = int(wo.sum())
    h_a = lib.hm_host_alloc(ne * 160 * 8)
    h_b = lib.hm_host_alloc(ne * 160 * 8)
    h_o = lib.hm_host_alloc(ne * vwo * 8)
    if not (h_a and h_b and h_o):
        raise RuntimeError("pinned host allocation failed")
    for dst, batch in ((h_a, ca), (h_b, cb)):  # pinned copies of the first `ne` encrypted pairs
        host = batch.to_host()  # keep the array alive while it is copied
        C.memmove(dst, host.ctypes.data, ne * 160 * 8)
        del host
    u64p = C.POINTER(C.c_uint64)

    def e2e_step():
        rc = lib.hm_apply2_host_bounded(ctx._h, N.HM_OP_ADD, ne, L, fresh.ctypes.data_as(u64p), h_a, fresh.ctypes.data_as(u64p), h_b, h_o)
        if rc != 0:
            raise RuntimeError(f"hm_apply2_host_bounded failed: {rc} {lib.hm_last_error(ctx._h).decode()}")

    for _ in range(4):  # first DMAs into freshly pinned pages are slow (seen: 79, 73, then 58 ms per step), more so with 8 ranks at once
        e2e_step()
    barrier()
    e2e_steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    e2e_step_ms = []
    for _ in range(e2e_steps):
        ts = time.perf_counter()
        e2e_step()
        e2e_step_ms.append((time.perf_counter() - ts) * 1e3)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * ne * e2e_steps / e2e_s
    # the host result of the last e2e step equals the device result of the timed steps (same inputs)
    host_out = np.ctypeslib.as_array(C.cast(h_o, C.POINTER(C.c_uint64)), shape=(ne, vwo))
    probe = ctx.upload(np.ascontiguousarray(host_out[:64]), [int(x) for x in wo], [int(x) for x in out.slot_degree_bounds()])
    e2e_matches = bool(np.array_equal(ctx.decrypt(probe, np.uint32), dec[:64]))
    probe.free()
    lib.hm_host_free(h_a); lib.hm_host_free(h_b); lib.hm_host_free(h_o)

    # ---- end to end, whole circuit: host plaintexts -> encrypt (device-side masks) -> add -> decrypt -> host plaintexts ----
    nc = min(args.circuit_pairs, n)
    hva = torch.from_numpy(a[:nc].copy()).pin_memory()
    hvb = torch.from_numpy(b[:nc].copy()).pin_memory()
    hres = torch.zeros(nc, dtype=torch.int32).pin_memory()

    def circuit_step(seed):
        ea, eb, es = C.c_void_p(), C.c_void_p(), C.c_void_p()
        rc = lib.hm_encrypt_seeded(ctx._h, hva.data_ptr(), nc, L, seed, C.byref(ea))
        rc = rc or lib.hm_encrypt_seeded(ctx._h, hvb.data_ptr(), nc, L, seed + 1, C.byref(eb))
        rc = rc or lib.hm_apply2(ctx._h, N.HM_OP_ADD, ea, eb, C.byref(es))
        rc = rc or lib.hm_decrypt(ctx._h, es, hres.data_ptr())
        for h in (ea, eb, es):
            if h.value:
                lib.hm_batch_free(ctx._h, h)
        if rc != 0:
            raise RuntimeError(f"circuit step failed: {rc} {lib.hm_last_error(ctx._h).decode()}")

    circuit_step(100); circuit_step(200)
    ctx.synchronize()
    barrier()
    c_steps = max(1, min(args.steps, 3))
    lc0 = ctx.kernel_launches()
    t0 = time.perf_counter()
    c_ms = []
    for i in range(c_steps):
        ts = time.perf_counter()
        circuit_step(1000 + 2 * i)
        c_ms.append((time.perf_counter() - ts) * 1e3)
    barrier()
    c_s = max_over_ranks(time.perf_counter() - t0)
    circuit_value = world * nc * c_steps / c_s
    circuit_launches = int(ctx.kernel_launches() - lc0)
    circuit_ok = float(np.mean(hres.numpy().view(np.uint32) == (a[:nc] + b[:nc])))
    del hva, hvb, hres

    # ---- the box's PCIe ceiling with every rank copying at once (what bounds `e2e` at N > 1) -----------------------------
    pcie = {}
    hp = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
    dp_ = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    for name, src, dst in (("h2d", hp, dp_), ("d2h", dp_, hp)):
        dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        mine = time.perf_counter() - t0
        pcie[name + "_GBps_this_rank"] = 3 * (1 << 30) / mine / 1e9
        pcie[name + "_GBps_all_ranks_concurrent"] = world * 3 * (1 << 30) / max_over_ranks(mine) / 1e9
    del hp, dp_
    pcie["note"] = ("every rank copies 1 GiB of pinned memory at the same time (barrier before, max over ranks): the measured "
                    "aggregate host<->device rate of this box, the ceiling of any host-buffer call at this N")
    e2e_bytes_per_s = e2e_value * (46912 + 2560)
    pcie["e2e_d2h_frac_of_concurrent_ceiling"] = e2e_value * 46912 / 1e9 / pcie["d2h_GBps_all_ranks_concurrent"]

    # ---- second half of BASELINE's metric on EVERY rank: fused (a*b) mod S on fresh pairs --------------------------------
    extra = {}
    hbm_peak, hbm_src = measured_peaks()

    def timed(fn, reps=5, strm=None, sync=None):
        strm = strm or stream
        sync = sync or ctx.synchronize
        fn(); sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(strm)
        for _ in range(reps):
            fn()
        e1.record(strm); sync()
        return e0.elapsed_time(e1) * 1e-3 / reps

    if not args.no_extra:
        mr = ctx.poly_mulrem(ca, cb)  # every slot of (ca, cb) is one pair -> n*32 pairs per launch

        def mulrem():
            rc = lib.hm_poly_mulrem_into(ctx._h, ca._h, cb._h, mr._h)
            assert rc == 0, rc

        barrier()
        s_mr = max_over_ranks(timed(mulrem))
        pairs = n * L
        extra["mulrem"] = {"metric": "GF(2)[X] mul+rem/s (d=d'=128)", "value": world * pairs / s_mr, "unit": "mul+rem/s", "n_gpus": world,
                           "kernel": "mulrem_fresh_a_kernel<1024>", "pairs_per_launch_per_gpu": pairs, "ms": s_mr * 1e3,
                           "hbm_GBps_per_gpu": pairs * BYTES_PER_MULREM / s_mr / 1e9, "hbm_frac": pairs * BYTES_PER_MULREM / s_mr / 1e9 / hbm_peak,
                           "Tbitmac_per_s_per_gpu": pairs * BITMACS_PER_MULREM / s_mr / 1e12,
                           "note": "(a b) mod S computed as ((a mod S)(b mod S)) mod S: same remainder bit for bit, a third of the leaf products; "
                                   "every rank runs it on its own shard, time = max over ranks, value = aggregate"}
        mr.free()
        # config 5 (stress): d=d'=512, tau=256, delta=8 fused mul+rem, swept over 2^10 .. 2^22 fresh pairs per GPU
        rb = np.random.default_rng(55)
        ctxb = hm.Context(hm.Parameters(512, 512, 8, 256), device=local_rank)
        skb = hm.SecretKey.random(512, rb)
        ctxb.set_secret_key(skb)
        ctxb.set_public_key(hm.PublicKey.random(512, 8, 256, skb, rb))
        streamb = torch.cuda.ExternalStream(lib.hm_context_stream(ctxb._h), device=dev)
        sweep = []
        top = args.sweep_max_log2
        nb = (1 << top) // 8  # u8 values -> 8 pairs each
        vb1 = rb.integers(0, 256, size=nb, dtype=np.uint8)
        vb2 = rb.integers(0, 256, size=nb, dtype=np.uint8)
        cb1 = ctxb.encrypt(vb1, seed=77 + rank)
        cb2 = ctxb.encrypt(vb2, seed=99 + rank)
        for lg in range(10, top + 1, 2):
            cnt = (1 << lg) // 8
            if cnt != nb:  # a prefix of the same plaintexts, encrypted again (cheap; batches have no sub-range views)
                s1 = ctxb.encrypt(vb1[:cnt], seed=77 + rank)
                s2 = ctxb.encrypt(vb2[:cnt], seed=99 + rank)
            else:
                s1, s2 = cb1, cb2
            mrb = ctxb.poly_mulrem(s1, s2)

            def mulrem_b():
                assert lib.hm_poly_mulrem_into(ctxb._h, s1._h, s2._h, mrb._h) == 0

            barrier()
            reps = 20 if lg <= 16 else 5
            sb = max_over_ranks(timed(mulrem_b, reps=reps, strm=streamb, sync=ctxb.synchronize))
            sweep.append({"log2_pairs_per_gpu": lg, "value": world * (cnt * 8) / sb, "unit": "mul+rem/s", "ms": sb * 1e3})
            mrb.free()
            if cnt != nb:
                s1.free(); s2.free()
        big = sweep[-1]
        extra["mulrem_config_b"] = {"metric": "GF(2)[X] mul+rem/s (d=d'=512, tau=256, delta=8)", "value": big["value"], "unit": "mul+rem/s", "n_gpus": world,
                                    "kernel": "mulrem_fresh32q_kernel", "pairs_per_launch_per_gpu": 1 << top, "ms": big["ms"],
                                    "Tbitmac_per_s_per_gpu": (1 << top) * (1050625 + 788481) / (big["ms"] * 1e-3) / 1e12,
                                    "sweep": sweep,
                                    "note": "configs[4]: carry-less mul + rem sweep; operands reduced mod S first (sliding-window table folds from shared "
                                            "memory), then a 16-word product (three 8x8-word Karatsubas) and one more fold; aggregate over ranks, "
                                            "small batches are launch-latency bound"}
        if rank == 0:
            # config-B encryption on the tensor cores: 2^22 bit-ciphertexts (the sweep's largest operand) encrypted again in place
            try:
                dvb = torch.from_numpy(vb1.copy()).to(dev)
                torch.cuda.synchronize()

                def enc_b():
                    assert lib.hm_encrypt_device_seeded_into(ctxb._h, dvb.data_ptr(), nb, 8, 77 + rank, 0, cb1._h) == 0

                sb = timed(enc_b, reps=10, strm=streamb, sync=ctxb.synchronize)
                units_b = nb * 8
                extra["encrypt_config_b"] = {"value": units_b / sb, "unit": "bit-ciphertexts/s", "ms": sb * 1e3, "units": units_b,
                                             "kernel": "encrypt_umma_b_kernel<true> x 2 passes (tcgen05.mma kind::i8, M128 N256 K32; accumulators in TMEM; "
                                                       "parity-and-pack epilogue; Philox masks drawn in the kernel)",
                                             "hbm_GBps": units_b * 136 / sb / 1e9,
                                             "tensor": {"int8_MAC_per_s": units_b * 256 * 1024 / sb,
                                                        "frac_of_probe_rate": (units_b * 256 * 1024 / sb) / (256 * 128 / 4.9 * 148 * 1.965e9),
                                                        "note": "against the MMA-only stream of tools/umma_encrypt_probe.cu (4.9 clk per bit-ciphertext of a 128x256x128 tile per SM = 6.7 k int8 MAC per clk per SM)"},
                                             "note": "masks[n x 256] * PK[256 x 1025] over GF(2) as an int8 GEMM; the table kernel encrypt_tab4b_kernel takes "
                                                     "1.47x as long (HM_ENC_MODE=2)"}
                del dvb
            except Exception as e:
                extra["encrypt_config_b"] = {"error": repr(e)}
        for ob in (cb1, cb2):
            ob.free()
        ctxb.close()

    # ---- secondary lines (rank 0 only, short) ------------------------------------------------------------------------------
    if rank == 0 and not args.no_extra:
        dout = torch.empty(n * 4, dtype=torch.uint8, device=dev)
        s = timed(lambda: lib.hm_decrypt_device(ctx._h, out._h, dout.data_ptr()))
        extra["decrypt_after_add"] = {"value": n / s, "unit": "u32/s", "kernel": "decrypt_value_tma_kernel<2,256> x 2 CTAs/SM", "ms": s * 1e3,
                                      "hbm_GBps": n * 46912 / s / 1e9, "hbm_frac": n * 46912 / s / 1e9 / hbm_peak}
        s = timed(lambda: lib.hm_decrypt_device(ctx._h, ca._h, dout.data_ptr()), reps=20)
        extra["decrypt_fresh"] = {"value": n / s, "unit": "u32/s", "kernel": "decrypt_uniform_kernel", "ms": s * 1e3,
                                  "hbm_GBps": n * 1280 / s / 1e9, "hbm_frac": n * 1280 / s / 1e9 / hbm_peak,
                                  "note": "320 MiB of ciphertext per launch (> 126 MB L2)"}
        # encrypt with values + masks already in HBM, into an existing batch
        g = np.random.default_rng(1)
        dm = torch.from_numpy(np.frombuffer(g.bytes(n * L * 16), dtype=np.uint8).copy()).to(dev)
        dv = torch.from_numpy(a.view(np.uint8).copy()).to(dev)
        ce = ca.clone()

        def encd():
            rc = lib.hm_encrypt_device_into(ctx._h, dv.data_ptr(), n, L, dm.data_ptr(), ce._h)
            assert rc == 0, rc

        s = timed(encd, reps=20)
        props = torch.cuda.get_device_properties(local_rank)
        sm_count, sm_clk = props.multi_processor_count, getattr(props, "clock_rate", 1965000) * 1e3
        extra["encrypt"] = {"value": n / s, "unit": "u32/s", "kernel": "encrypt_tab4_kernel<false> (masks read from HBM)", "ms": s * 1e3,
                            "hbm_GBps": n * 1792 / s / 1e9, "hbm_frac": n * 1792 / s / 1e9 / hbm_peak,
                            "smem_frac": (n * L / s) * 4.0 / (sm_count * sm_clk),
                            "note": "table lookups: 16 x 32 B of shared-memory reads per 56 B of HBM traffic (4 LDS wavefronts per "
                                    "bit-ciphertext), so the shared-memory pipe and instruction issue bind before HBM does "
                                    "(smem_frac = LDS wavefront cycles needed / available)"}

        def encs():
            rc = lib.hm_encrypt_device_seeded_into(ctx._h, dv.data_ptr(), n, L, 12345, 0, ce._h)
            assert rc == 0, rc

        s = timed(encs, reps=20)
        extra["encrypt_seeded"] = {"value": n / s, "unit": "u32/s", "kernel": "encrypt_tab4_kernel<true> (Philox masks drawn in the kernel)",
                                   "ms": s * 1e3, "hbm_GBps": n * 1284 / s / 1e9, "hbm_frac": n * 1284 / s / 1e9 / hbm_peak,
                                   "smem_frac": (n * L / s) * 4.0 / (sm_count * sm_clk),
                                   "note": "no mask buffer: 40 B written per bit-ciphertext, 4 B of plaintext read per u32"}
        ce.free()
        del dm, dv
        # end-to-end encryption, host plaintexts in -> ciphertexts resident in HBM: (i) host-generated masks cross PCIe
        # (16 B per bit), (ii) masks generated on the device from a seed (Philox4x32-10), only 4 B per u32 cross
        hv = torch.from_numpy(a.copy()).pin_memory()
        hm_ = torch.from_numpy(np.frombuffer(np.random.default_rng(2).bytes(n * L * 16), dtype=np.uint8).copy()).pin_memory()

        def e2e_enc_masks():
            o = C.c_void_p()
            assert lib.hm_encrypt(ctx._h, hv.data_ptr(), n, L, hm_.data_ptr(), C.byref(o)) == 0
            lib.hm_batch_free(ctx._h, o)

        def e2e_enc_seed():
            o = C.c_void_p()
            assert lib.hm_encrypt_seeded(ctx._h, hv.data_ptr(), n, L, 12345, C.byref(o)) == 0
            lib.hm_batch_free(ctx._h, o)

        for name, fn in (("encrypt_e2e_host_masks", e2e_enc_masks), ("encrypt_e2e_seeded", e2e_enc_seed)):
            fn(); ctx.synchronize()
            times = []
            for _ in range(5):
                t0 = time.perf_counter()
                fn()
                ctx.synchronize()
                times.append(time.perf_counter() - t0)
            dt = float(np.median(times))
            extra[name] = {"value": n / dt, "unit": "u32/s", "ms": dt * 1e3, "ms_each": [round(x * 1e3, 3) for x in times]}
        del hv, hm_
        # configs[1]: u32 batched encryption + decryption of 2^20 values, host plaintexts in -> ciphertexts in HBM -> host plaintexts
        # out (Context::encrypt / Context::decrypt of benches/u32.rs on one batch); seeded masks, so 4 B per value cross PCIe each way
        try:
            n1 = 1 << 20
            v1 = torch.from_numpy(np.random.default_rng(31).integers(0, 2**32, size=n1, dtype=np.uint32)).pin_memory()
            r1 = torch.empty(n1, dtype=torch.int32).pin_memory()

            def config1():
                o = C.c_void_p()
                assert lib.hm_encrypt_seeded(ctx._h, v1.data_ptr(), n1, L, 777, C.byref(o)) == 0
                assert lib.hm_decrypt(ctx._h, o, r1.data_ptr()) == 0
                lib.hm_batch_free(ctx._h, o)

            config1(); ctx.synchronize()
            t1s = []
            for _ in range(5):
                t0 = time.perf_counter()
                config1()
                ctx.synchronize()
                t1s.append(time.perf_counter() - t0)
            dt = float(np.median(t1s))
            extra["config1_encrypt_decrypt"] = {"value": n1 / dt, "unit": "u32 round trips/s", "values": n1, "ms": dt * 1e3,
                                                "ms_each": [round(x * 1e3, 3) for x in t1s],
                                                "round_trip_exact": bool((r1.numpy().view(np.uint32) == v1.numpy()).all()),
                                                "hbm_GBps": n1 * (1284 + 1284) / dt / 1e9,
                                                "note": "configs[1]: hm_encrypt_seeded -> hm_decrypt through the C ABI with host plaintexts, wall clock incl. "
                                                        "copies and the 1.34 GB batch allocation; 40 B written + 40 B read per bit-ciphertext"}
            del v1, r1
        except Exception as e:  # a secondary line must not take the headline down
            extra["config1_encrypt_decrypt"] = {"error": repr(e)}
        # gates on fresh u32 batches (gate_xor / gate_and, common.rs:5-27): n*32 bit-ciphertext pairs per launch
        xo = ctx.apply2(hm.HomomorphicXorGate, ca, cb)
        s = timed(lambda: lib.hm_apply2_into(ctx._h, N.HM_OP_XOR, ca._h, cb._h, xo._h), reps=20)
        extra["xor_gate"] = {"value": n * L / s, "unit": "bit-ciphertext xors/s", "kernel": "xor_flat_kernel", "ms": s * 1e3,
                             "hbm_GBps": n * L * 120 / s / 1e9, "hbm_frac": n * L * 120 / s / 1e9 / hbm_peak}
        xo.free()
        ao = ctx.apply2(hm.HomomorphicAndGate, ca, cb)
        s = timed(lambda: lib.hm_apply2_into(ctx._h, N.HM_OP_AND, ca._h, cb._h, ao._h), reps=5)
        extra["and_gate"] = {"value": n * L / s, "unit": "bit-ciphertext ands/s", "kernel": "mul_small_kernel<8,8>", "ms": s * 1e3,
                             "hbm_GBps": n * L * 152 / s / 1e9, "Tbitmac_per_s": n * L * 66049 / s / 1e12}
        s = timed(lambda: lib.hm_apply2_into(ctx._h, N.HM_OP_OR, ca._h, cb._h, ao._h), reps=5)
        extra["or_gate"] = {"value": n * L / s, "unit": "bit-ciphertext ors/s", "kernel": "mul_small_kernel<8,8> (a + b + a*b fused)", "ms": s * 1e3,
                            "hbm_GBps": n * L * 152 / s / 1e9}
        ao.free()
        # config 4: u8 homomorphic multiply (column circuit, common.rs:66-105) on 2^14 pairs, then decrypt
        n8 = 1 << 14
        g8 = np.random.default_rng(8)
        a8 = g8.integers(0, 256, size=n8, dtype=np.uint8)
        b8 = g8.integers(0, 256, size=n8, dtype=np.uint8)
        c8a = ctx.encrypt(a8, np.frombuffer(g8.bytes(n8 * 8 * 16), dtype=np.uint8))
        c8b = ctx.encrypt(b8, np.frombuffer(g8.bytes(n8 * 8 * 16), dtype=np.uint8))
        l0 = ctx.kernel_launches()
        t0 = time.perf_counter()
        p8 = ctx.apply2(hm.HomomorphicMultiplication, c8a, c8b)
        ctx.synchronize()
        t1 = time.perf_counter()
        l1 = ctx.kernel_launches()
        t8 = []
        p8b = None
        for it in range(7):  # two more warm-up calls (the stream-ordered pool is still growing), then five timed
            if p8b is not None:
                p8b.free()
            ts = time.perf_counter()
            p8b = ctx.apply2(hm.HomomorphicMultiplication, c8a, c8b)
            ctx.synchronize()
            if it >= 2:
                t8.append(time.perf_counter() - ts)
        t2 = t1 + float(np.median(t8))
        d8 = ctx.decrypt(p8b)
        extra["u8_mul"] = {"value": n8 / (t2 - t1), "unit": "u8 muls/s", "pairs": n8, "ms": (t2 - t1) * 1e3, "first_call_ms": (t1 - t0) * 1e3,
                           "ms_each": [round(x * 1e3, 3) for x in t8],
                           "kernel_launches": int(l1 - l0), "correct_frac": float(np.mean(d8 == a8 * b8)),
                           "kernel": "mul_circuit_fused_kernel<8> (one launch: a warp per value, partial products, prefixes and carries in shared memory)",
                           "note": "configs[3] at L = 8 (u32 multiplication is infeasible by construction, SURVEY.md A.3); wall clock incl. the launch, "
                                   "median of 5 after 3 warm-up calls; 2 196 8x8-word block products per multiply (32-word chunking) against the "
                                   "measured 13.4 G block products/s of the GPU = 6.1 M muls/s"}
        for o8 in (p8, p8b, c8a, c8b):
            o8.free()

    # ---- roofline of the dominant kernel (one adder_chain_kernel launch per step) --------------------------------------------
    roofline = roofline_hbm = None
    if rank == 0:
        props = torch.cuda.get_device_properties(local_rank)
        sm_count = props.multi_processor_count
        launch_s = float(np.mean(step_ms)) * 1e-3
        adds_per_s = n / launch_s
        sm_max = (clocks or {}).get("sm_max_mhz") or 1965.0
        pp = (C.c_double * 12)()
        probes = None
        probe_clocks = None
        for attempt in range(2):  # probes whose nvidia-smi clock was below 95 % of the maximum SM clock are taken again once
            ps = ClockSampler(local_rank)
            ps.start()
            time.sleep(0.3)
            tp0 = time.time()
            rcp = lib.hm_measure_pipe_peaks(ctx._h, 150.0, pp)
            probe_clocks = ps.stop(tp0, time.time())
            if rcp != 0:
                break
            probes = {name: {"warp_instr_per_s": pp[4 * i], "ms": pp[4 * i + 1], "warps_per_sm": int(pp[4 * i + 2]),
                             "warp_instr_per_clk_per_sm": pp[4 * i] / sm_count / ((probe_clocks.get("sm_mhz") or sm_max) * 1e6),
                             "cycle_counter_mhz_not_the_sm_clock": pp[4 * i + 3]}
                      for i, name in enumerate(("imad_wide_rr", "lop3_rrr", "mix_1w_2l"))}
            if (probe_clocks.get("sm_mhz") or 0) >= 0.95 * sm_max:
                break
        k8 = C.c_double(0.0)
        lib.hm_measure_kara8_peak(ctx._h, C.byref(k8))
        if probes:
            fma_ach = adds_per_s * NCU["fmaheavy_warp_instr_per_add"]
            alu_ach = adds_per_s * NCU["alu_warp_instr_per_add"]
            fma_frac = fma_ach / probes["imad_wide_rr"]["warp_instr_per_s"]
            alu_frac = alu_ach / probes["lop3_rrr"]["warp_instr_per_s"]
            busier = "fmaheavy" if fma_frac >= alu_frac else "alu"
            mix = probes["mix_1w_2l"]
            roofline = {
                "kernel": "adder_chain_kernel<8,0,4> (thread-per-value Karatsuba chain on IMAD.WIDE + LOP3, dynamically scheduled work units)",
                "bound": "alu", "bound_detail": "integer pipes of the SM: FMA-heavy (IMAD.WIDE, the 32x32->64 products) and ALU (LOP3); neither HBM nor tensor",
                "busier_pipe": busier,
                "achieved": (fma_ach if busier == "fmaheavy" else alu_ach) / 1e9,
                "peak": (probes["imad_wide_rr"] if busier == "fmaheavy" else probes["lop3_rrr"])["warp_instr_per_s"] / 1e9,
                "unit": "G warp-instructions/s on the busier pipe", "frac": max(fma_frac, alu_frac),
                "pipes": {"fmaheavy": {"achieved_G_warp_instr_per_s": fma_ach / 1e9, "peak": probes["imad_wide_rr"]["warp_instr_per_s"] / 1e9, "frac": fma_frac,
                                       "warp_instr_per_add": NCU["fmaheavy_warp_instr_per_add"]},
                          "alu": {"achieved_G_warp_instr_per_s": alu_ach / 1e9, "peak": probes["lop3_rrr"]["warp_instr_per_s"] / 1e9, "frac": alu_frac,
                                  "warp_instr_per_add": NCU["alu_warp_instr_per_add"]}},
                "peak_source": "measured in this run by hm_measure_pipe_peaks: IMAD.WIDE (both operands in registers) and LOP3 issue probes, >= 150 ms each "
                               "(fastest of three), every CTA resident; SM clock during the probes sampled with nvidia-smi (`probe_clocks`)",
                "probe_clocks": probe_clocks,
                "instruction_counts_source": NCU["source"],
                "ncu_pipe_pct_scaled_to_this_rate": {k: NCU["pipe_pct_at_capture"][k] * adds_per_s / NCU["pipe_pct_at_capture"]["adds_per_s_at_capture"] / 100.0
                                                     for k in ("alu", "fmaheavy")},
                "probes": probes,
                "joint_issue_ceiling": {
                    "what": "the two pipes do not issue independently: the probe with the kernel's own mix (1 IMAD.WIDE : 2 LOP3, 8 chains/thread) sustains "
                            "less than the sum of the single-pipe peaks; this is the ceiling for THIS instruction mix",
                    "mix_probe_G_warp_instr_per_s": mix["warp_instr_per_s"] / 1e9, "kernel_G_warp_instr_per_s": (fma_ach + alu_ach) / 1e9,
                    "mix_probe_warp_instr_per_clk_per_sm": mix["warp_instr_per_clk_per_sm"],
                    "kernel_warp_instr_per_clk_per_sm": (fma_ach + alu_ach) / sm_count / ((clocks or {}).get("sm_mhz") or sm_max) / 1e6,
                    "frac_of_mix_ceiling": (fma_ach + alu_ach) / mix["warp_instr_per_s"]},
                "kara8": {"achieved": adds_per_s * KARA8_PER_ADD / 1e9, "peak": k8.value / 1e9, "unit": "G 8x8-word products/s",
                          "frac": adds_per_s * KARA8_PER_ADD / k8.value if k8.value else None,
                          "note": "round-1 figure, kept as a sub-object: the 8x8-word Karatsuba product (432 IMAD.WIDE + 854 LOP3) timed in isolation; "
                                  "2 881 such products per add"},
                "traffic": n * NCU["dram_bytes_per_add"],
                "traffic_source": NCU["source"] + "; algorithmic bytes per launch = %d" % (n * BYTES_PER_ADD),
            }
        if k8.value and isinstance(extra.get("u8_mul"), dict) and extra["u8_mul"].get("value"):
            # the multiplier circuit against the same product-pipe measurement: 2 196 8x8-word block products per u8 multiply
            # (32-word chunking of the 56 carry products + 36 partial products)
            ach8 = extra["u8_mul"]["value"] * 2196
            extra["u8_mul"]["roofline"] = {"bound": "alu", "achieved": ach8 / 1e9, "peak": k8.value / 1e9, "unit": "G 8x8-word products/s",
                                           "frac": ach8 / k8.value,
                                           "peak_source": "hm_measure_kara8_peak: the 8x8-word Karatsuba product (432 IMAD.WIDE + 854 LOP3) timed alone in this run; "
                                                          "it runs at the joint-issue ceiling of its instruction mix (roofline.joint_issue_ceiling)",
                                           "note": "78 block-product rounds per value on 32 lanes against 68.6 if every lane were always busy; "
                                                   "2 warps per scheduler (25 KB of shared memory per value)"}
        gbs = n * BYTES_PER_ADD / launch_s / 1e9
        roofline_hbm = {"kernel": "adder_chain_kernel<8,0,4>", "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                        "frac": gbs / hbm_peak, "traffic": n * NCU["dram_bytes_per_add"], "peak_source": hbm_src,
                        "note": "the same kernel against the HBM roofline: far from it, the kernel is integer-pipe bound"}

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port on the box's host cores ------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import hmoracle as orc

        os.sched_setaffinity(0, all_cpus)
        threads = orc.max_threads()
        g = np.random.default_rng(7)
        osk, opk = orc.keygen(D, DP, DELTA, TAU, g)

        def oenc(cnt):
            v = g.integers(0, 2**32, size=cnt, dtype=np.uint32)
            m = g.integers(0, 256, size=cnt * L * 16, dtype=np.uint8)
            return orc.encrypt(opk, np.frombuffer(v.astype("<u4").tobytes(), dtype=np.uint8), 4, m, threads=threads)[0]

        oa, ob = oenc(threads), oenc(threads)
        _, sec = orc.apply(orc.OP_ADD, oa, ob, L, threads=threads)
        cnt = max(threads, int(threads / max(sec, 1e-9) * 12.0) // threads * threads)
        oa, ob = oenc(cnt), oenc(cnt)
        _, sec = orc.apply(orc.OP_ADD, oa, ob, L, threads=threads)
        _, sec1 = orc.apply(orc.OP_ADD, orc.PolyVec.from_words([oa.words(i) for i in range(4 * L)]),
                            orc.PolyVec.from_words([ob.words(i) for i in range(4 * L)]), L, threads=1)
        # the second metric of BASELINE.json beside it: mul + rem on fresh pairs (bit-serial mul, long-division rem)
        mcnt = 65536 * threads // L * L
        ma_, mb_ = oenc(mcnt // L), oenc(mcnt // L)
        _, msec = orc.poly_mulrem(ma_, mb_, osk, threads=threads)
        cpu_mulrem = {"value": mcnt / msec, "unit": "mul+rem/s", "cores": threads, "kind": "port", "sample": f"{mcnt} pairs, {msec:.2f} s"}
        # the rest of the reference's own bench list (benches/u32.rs, benches/u8.rs) on small samples: 1 thread is what
        # the README's criterion numbers are, all threads is the most the reference's design admits
        table = {}
        cnt_small = 64 * threads
        vs = g.integers(0, 2**32, size=cnt_small, dtype=np.uint32)
        ms_ = g.integers(0, 256, size=cnt_small * L * 16, dtype=np.uint8)
        raw = np.frombuffer(vs.astype("<u4").tobytes(), dtype=np.uint8)
        for th in (1, threads):
            k = cnt_small if th > 1 else 64
            enc_ct, s_enc = orc.encrypt(opk, raw[: 4 * k], 4, ms_[: k * L * 16], threads=th)
            _, s_dec = orc.decrypt(osk, enc_ct, L, threads=th)
            ka = 8 if th == 1 else threads
            sa_, sb_ = oenc(ka), oenc(ka)
            sum_ct, s_add = orc.apply(orc.OP_ADD, sa_, sb_, L, threads=th)
            _, s_dadd = orc.decrypt(osk, sum_ct, L, threads=th)
            v8 = g.integers(0, 256, size=ka, dtype=np.uint8)
            m8 = g.integers(0, 256, size=ka * 8 * 16, dtype=np.uint8)
            c8 = orc.encrypt(opk, v8, 1, m8, threads=th)[0]
            prod_ct, s_mul = orc.apply(orc.OP_MUL, c8, c8, 8, threads=th)
            _, s_dmul = orc.decrypt(osk, prod_ct, 8, threads=th)
            table[f"{th}_threads"] = {"u32_encrypt_per_s": k / s_enc, "u32_decrypt_per_s": k / s_dec, "u32_add_per_s": ka / s_add,
                                      "u32_decrypt_after_add_per_s": ka / s_dadd, "u8_mul_per_s": ka / s_mul,
                                      "u8_decrypt_after_mul_per_s": ka / s_dmul}
        cpu = {"value": cnt / sec, "unit": UNIT, "cores": threads, "kind": "port", "mulrem": cpu_mulrem, "table": table,
               "published_table": "README.md:73-77 (Ryzen 7 7800X3D, 1 thread): encrypt 76.0 us, decrypt 12.5 us, add 950 us, "
                                  "decrypt after add 1.03 ms per u32",
               "sample": f"{cnt} pairs, {threads} threads over independent values, {sec:.1f} s",
               "single_thread_value": 4 / sec1,
               "published_reference": "README.md:75: 950 us per add = 1053 adds/s, 1 thread of a Ryzen 7 7800X3D"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 words of GF(2)[X] (bit-packed; carry-less products on IMAD.WIDE + LOP3)",
            "data": "synthetic (seeded keys, uniform u32 plaintexts, host-generated subset masks)",
            "config": {"workload": WORKLOAD, "pairs_per_gpu": n, "bits": L,
                       "l2": "inputs 2 x %.0f MiB + result %.2f GiB per step >> 126 MB L2 (no flush needed)" % (n * 1280 / 2**20, n * 46912 / 2**30),
                       "sharding": "independent values split by index across ranks; no collective on the data path"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": ne * 2560, "d2h_bytes_per_step": ne * vwo * 8,
                    "pairs_per_step": ne, "pairs_per_step_note": "a quarter of the headline batch (2^16 of 2^18 pairs per GPU per call): 3.07 GB of pinned "
                                                                  "result per call; the rate is D2H-bound and does not depend on the batch size",
                    "steps": e2e_steps, "step_ms": e2e_step_ms,
                    "call": "hm_apply2_host_bounded (pinned host ciphertexts in/out, 3-stream chunked pipeline)",
                    "matches_device_result": e2e_matches, "bytes_per_s": e2e_bytes_per_s},
            "e2e_circuit": {"value": circuit_value, "unit": UNIT, "pairs_per_step": nc, "steps": c_steps, "step_ms": c_ms,
                            "h2d_bytes_per_step": nc * 8, "d2h_bytes_per_step": nc * 4, "gpu_launches": circuit_launches,
                            "correct_frac": circuit_ok,
                            "flow": "host plaintexts -> hm_encrypt_seeded x2 -> hm_apply2(ADD) -> hm_decrypt -> host plaintexts; ciphertexts stay in HBM "
                                    "(benches/u32.rs:8-50 as one batch); allocation and release of the 11.45 GiB result inside the timed region"},
            "pcie": pcie,
            "gpu_launches": int(l_after - l_before),
            "kernels_in_step": ["adder_chain_kernel<8,0,4>"],
            "clocks": clocks,
            "roofline": roofline, "roofline_hbm": roofline_hbm,
            "cpu_baseline": cpu,
            "decrypted_sums_correct_frac": frac_ok,
            "step_ms": step_ms,
            "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ our arm, one process for N GPUs
def run_group(args):
    """`python bench.py --gpus N` WITHOUT torchrun: one host thread drives N GPUs through the device-group C ABI
    (hm_group_*, include/hmgpu.h): keys replicated, values split by index, no collective.  Same metric and workload as the
    torchrun path (2^18 pairs per GPU); device-timed with CUDA events on every device's stream, max over devices."""
    import torch

    import homomorph_rust_b200 as hm
    from homomorph_rust_b200 import _native as N
    from homomorph_rust_b200.api import ContextGroup

    lib = hm.lib()
    world = args.gpus
    if lib.hm_device_count() < world:
        raise SystemExit(f"bench.py: {world} GPUs asked, {lib.hm_device_count()} visible — the engine has no CPU fallback")
    grp = ContextGroup(hm.Parameters(D, DP, DELTA, TAU), list(range(world)))
    sk, pk = make_keys(hm)
    grp.set_secret_key(sk)
    grp.set_public_key(pk)
    n = args.pairs * world
    rng = np.random.default_rng(1000)
    a = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    b = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    ga, gb = grp.encrypt(a, seed=5000), grp.encrypt(b, seed=6000)
    out = grp.apply2(hm.HomomorphicAddition, ga, gb)
    grp.synchronize()
    streams = [torch.cuda.ExternalStream(lib.hm_context_stream(lib.hm_group_context(grp._h, i)), device=torch.device("cuda", i)) for i in range(world)]

    def step():
        rc = lib.hm_group_apply2_into(grp._h, N.HM_OP_ADD, ga._h, gb._h, out._h)
        if rc != 0:
            raise RuntimeError(f"hm_group_apply2_into failed: {rc}")

    sampler = ClockSampler(0)
    sampler.start()
    for _ in range(args.warmup):
        step()
    grp.synchronize()
    t_region0 = time.time()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(world)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(world)]
    launches0 = sum(lib.hm_context_kernel_launches(lib.hm_group_context(grp._h, i)) for i in range(world))
    for i in range(world):
        with torch.cuda.device(i):
            ev0[i].record(streams[i])
    for _ in range(args.steps):
        step()
    for i in range(world):
        with torch.cuda.device(i):
            ev1[i].record(streams[i])
    grp.synchronize()
    clocks = sampler.stop(t_region0, time.time())
    launches = sum(lib.hm_context_kernel_launches(lib.hm_group_context(grp._h, i)) for i in range(world)) - launches0
    ms = max(ev0[i].elapsed_time(ev1[i]) for i in range(world))
    value = n * args.steps / (ms * 1e-3)
    dec = grp.decrypt(out)
    frac_ok = float(np.mean(dec == (a + b)))
    # whole circuit through the group calls: host plaintexts -> encrypt (device masks) -> add -> decrypt -> host plaintexts
    def circuit(seed):
        ea, eb = grp.encrypt(a, seed=seed), grp.encrypt(b, seed=seed + 1)
        es = grp.apply2(hm.HomomorphicAddition, ea, eb)
        r = grp.decrypt(es)
        for x in (ea, eb, es):
            x.free()
        return r
    circuit(1); circuit(3)
    t0 = time.perf_counter()
    c_steps = max(1, min(args.steps, 3))
    for i in range(c_steps):
        r = circuit(10 + 2 * i)
    c_s = time.perf_counter() - t0
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 words of GF(2)[X] (bit-packed; carry-less products on IMAD.WIDE + LOP3)",
        "data": "synthetic (seeded keys, uniform u32 plaintexts, Philox-seeded subset masks)",
        "config": {"workload": WORKLOAD, "pairs_per_gpu": args.pairs, "bits": L,
                   "mode": "ONE process, one host thread, device-group C ABI (hm_group_*); the driver's torchrun path is run_ours",
                   "sharding": "independent values split by index across devices; no collective on the data path"},
        "e2e": {"value": n * c_steps / c_s, "unit": UNIT, "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": n * 4, "pairs_per_step": n,
                "call": "hm_group_encrypt_seeded x2 -> hm_group_apply2 -> hm_group_decrypt (host plaintexts in/out)",
                "correct_frac": float(np.mean(r == (a + b)))},
        "gpu_launches": int(launches), "kernels_in_step": ["adder_chain_kernel<8,0,4> on every device"],
        "clocks": clocks, "decrypted_sums_correct_frac": frac_ok,
    }
    print(json.dumps(line), flush=True)
    grp.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1 << 18, help="encrypted u32 pairs per GPU per step")
    ap.add_argument("--e2e-pairs", type=int, default=1 << 16, help="pairs per end-to-end step (pinned host ciphertext buffers)")
    ap.add_argument("--circuit-pairs", type=int, default=1 << 18, help="pairs per whole-circuit end-to-end step (host plaintexts in/out)")
    ap.add_argument("--sweep-max-log2", type=int, default=22, help="largest batch of the config-5 mul+rem sweep, log2 of pairs per GPU")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif world == 1 and args.gpus > 1:  # no torchrun: one process drives all GPUs through the device-group ABI
        run_group(args)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
