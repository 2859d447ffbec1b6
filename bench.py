#!/usr/bin/env python
"""bench.py — BASELINE.json's headline metric on B200: u32 homomorphic adds/s (config 3: ripple-carry
XOR/AND circuit on 2^18 encrypted pairs per GPU, d=d'=128, delta=1, tau=128), plus the secondary lines
(GF(2)[X] mul+rem/s, batched encrypt / decrypt) in `extra`.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One JSON line on stdout (rank 0).  A "step" is one pass of the fused adder over one batch of 2^18 synthetic
encrypted pairs per GPU (inputs 2 x 320 MiB, result 11.45 GiB: far larger than the 126 MB L2, so no L2 flush
is needed between iterations).  `value` is timed with CUDA events on the engine's stream with inputs resident in
HBM; `e2e` goes through the host-buffer C-ABI call hm_apply2_host (pinned host ciphertexts in, pinned host
result out, copies inside the timed region).  `--impl reference` times the CPU restatement of the reference
(oracle/, all host threads) on a bounded sample of the same workload — the reference itself is Rust and cannot
be built in this image (DESIGN.md "Oracle").
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
# SURVEY.md §8(d) / §A.2 — algorithmic work of one u32 homomorphic add at D = d+d' = 256
BITMACS_PER_ADD = 2.751e8        # schoolbook AND-XOR pairs over the 93 reference multiplications
BYTES_PER_ADD = 2560 + 46912     # 2 x 32 x 5 words in, 5 864 words out
BITMACS_PER_MULREM = 66049 + 49665
BYTES_PER_MULREM = 96
# dram__bytes_read.sum + dram__bytes_write.sum of adder_fused_kernel from the ncu --set full capture in profiles/
# (r01_adder_ncu_details.txt: 42.3 MB + 711.4 MB for 16 384 adds), per add
NCU_DRAM_BYTES_PER_ADD = (4.277474e9 + 4.852116e9) / 75776  # ncu capture of adder_thread_smem_kernel<4>, profiles/r01_adder_thread_smem_ncu_details.txt


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

    def enc(n):
        v = rng.integers(0, 2**32, size=n, dtype=np.uint32)
        m = rng.integers(0, 256, size=n * L * 16, dtype=np.uint8)
        return orc.encrypt(pk, np.frombuffer(v.astype("<u4").tobytes(), dtype=np.uint8), 4, m, threads=threads)[0]

    # calibrate: one add per thread
    a, b = enc(threads), enc(threads)
    _, sec = orc.apply(orc.OP_ADD, a, b, L, threads=threads)
    rate = threads / max(sec, 1e-9)
    total_steps = args.steps + args.warmup
    per_step_s = min(8.0, max(1.0, 150.0 / max(total_steps, 1)))
    n = max(threads, int(rate * per_step_s) // threads * threads)
    a, b = enc(n), enc(n)
    for _ in range(args.warmup):
        orc.apply(orc.OP_ADD, a, b, L, threads=threads)
    t = 0.0
    for _ in range(args.steps):
        _, sec = orc.apply(orc.OP_ADD, a, b, L, threads=threads)
        t += sec
    value = n * args.steps / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 GF(2) words", "data": "synthetic (seeded keys, plaintexts and subset masks)",
        "config": {"workload": "configs[2]: u32 homomorphic add (ripple-carry XOR/AND circuit) on 2^18 encrypted pairs per GPU, "
                               "d=dp=128, delta=1, tau=128", "pairs_per_gpu": 1 << 18, "bits": L,
                   "pairs_per_step": n, "note": "each step is a bounded sample of the 2^18-pair workload on the host cores; CPU port "
                                                "of the reference's algorithms (the Rust reference cannot be built in this image)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n} pairs per step x {args.steps} steps, {threads} threads over independent values"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

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
    stream = torch.cuda.ExternalStream(lib.hm_context_stream(ctx._h), device=torch.device("cuda", local_rank))

    n = args.pairs
    rng = np.random.default_rng(1000 + rank)
    a = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    b = rng.integers(0, 2**32, size=n, dtype=np.uint32)

    def enc(v, seed):
        g = np.random.default_rng(seed)
        out = None
        # masks are host generated (reproducible): 16 B per bit-ciphertext, uploaded in slices
        m = np.frombuffer(g.bytes(v.size * L * 16), dtype=np.uint8)
        out = ctx.encrypt(v, m)
        return out

    ca, cb = enc(a, 5000 + rank), enc(b, 6000 + rank)
    launches0 = ctx.kernel_launches()
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
    t = torch.tensor([total_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * n * args.steps / (total_ms_max * 1e-3)

    # sanity outside the timed region: the decrypted sums (probabilistically exact at delta=1)
    dec = ctx.decrypt(out)
    frac_ok = float(np.mean(dec == (a + b)))

    # ---- end to end through the host-buffer C-ABI call -------------------------------------------------
    ne = min(args.e2e_pairs, n)
    wa = np.full(L, 5, dtype=np.uint32)
    wo = out.slot_words().astype(np.uint32)
    vwo = int(wo.sum())
    h_a = lib.hm_host_alloc(ne * 160 * 8)
    h_b = lib.hm_host_alloc(ne * 160 * 8)
    h_o = lib.hm_host_alloc(ne * vwo * 8)
    if not (h_a and h_b and h_o):
        raise RuntimeError("pinned host allocation failed")
    # pinned copies of the first `ne` encrypted pairs
    for dst, batch in ((h_a, ca), (h_b, cb)):
        host = batch.to_host()  # keep the array alive while it is copied
        C.memmove(dst, host.ctypes.data, ne * 160 * 8)
        del host
    u32p = C.POINTER(C.c_uint32)

    def e2e_step():
        rc = lib.hm_apply2_host(ctx._h, N.HM_OP_ADD, ne, L, wa.ctypes.data_as(u32p), h_a, wa.ctypes.data_as(u32p), h_b, h_o)
        if rc != 0:
            raise RuntimeError(f"hm_apply2_host failed: {rc} {lib.hm_last_error(ctx._h).decode()}")

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
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local_rank}")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * ne * e2e_steps / float(te.item())
    # the host result of the last e2e step equals the device result of the timed steps (same inputs)
    host_out = np.ctypeslib.as_array(C.cast(h_o, C.POINTER(C.c_uint64)), shape=(ne, vwo))
    probe = ctx.upload(np.ascontiguousarray(host_out[:64]), [int(x) for x in wo], [int(x) for x in out.slot_degree_bounds()])
    e2e_matches = bool(np.array_equal(ctx.decrypt(probe, np.uint32), dec[:64]))
    probe.free()
    lib.hm_host_free(h_a); lib.hm_host_free(h_b); lib.hm_host_free(h_o)

    # ---- secondary lines (same run, short) ----------------------------------------------------------------
    extra = {}
    if rank == 0 and not args.no_extra:
        def timed(fn, reps=5):
            fn(); ctx.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                fn()
            e1.record(stream); ctx.synchronize()
            return e0.elapsed_time(e1) * 1e-3 / reps

        hbm_peak, _ = measured_peaks()
        # mul+rem on fresh pairs: every slot of (ca, cb) is one pair -> n*32 pairs per launch
        mr = ctx.poly_mulrem(ca, cb)
        def mulrem():
            rc = lib.hm_poly_mulrem_into(ctx._h, ca._h, cb._h, mr._h)
            assert rc == 0, rc
        s = timed(mulrem)
        pairs = n * L
        lane_ops = C.c_double(0.0)
        lib.hm_measure_alu_peak(ctx._h, C.byref(lane_ops), None)
        extra["mulrem"] = {"metric": "GF(2)[X] mul+rem/s (d=d'=128)", "value": pairs / s, "unit": "mul+rem/s",
                           "kernel": "mulrem_fresh_a_kernel<1024>", "pairs_per_launch": pairs, "ms": s * 1e3,
                           "hbm_GBps": pairs * BYTES_PER_MULREM / s / 1e9, "hbm_frac": pairs * BYTES_PER_MULREM / s / 1e9 / hbm_peak,
                           "Tbitmac_per_s": pairs * BITMACS_PER_MULREM / s / 1e12,
                           "alu_frac": pairs * BITMACS_PER_MULREM / s / (lane_ops.value * 32.0)}
        k8m = C.c_double(0.0)
        lib.hm_measure_kara8_peak(ctx._h, C.byref(k8m))
        # the kernel reduces both operands mod S first, so its product is 4x4 words = 9 leaf products (an 8x8-word Karatsuba is 27):
        # what is left is mostly the three table folds (14 words x 4 lookups per pair)
        extra["mulrem"]["product_pipe_frac"] = pairs / 3.0 / s / k8m.value
        extra["mulrem"]["note"] = ("(a b) mod S computed as ((a mod S)(b mod S)) mod S: same remainder bit for bit, a third of the leaf "
                                   "products; 10.3 G/s with the full 8x8-word product first")
        mr.free()
        # decrypt after add (HBM bound: 46 912 B per value)
        dout = torch.empty(n * 4, dtype=torch.uint8, device=f"cuda:{local_rank}")
        s = timed(lambda: lib.hm_decrypt_device(ctx._h, out._h, dout.data_ptr()))
        extra["decrypt_after_add"] = {"value": n / s, "unit": "u32/s", "kernel": "decrypt_value_tma_kernel<2,256> x 2 CTAs/SM", "ms": s * 1e3,
                                      "hbm_GBps": n * 46912 / s / 1e9, "hbm_frac": n * 46912 / s / 1e9 / hbm_peak}
        s = timed(lambda: lib.hm_decrypt_device(ctx._h, ca._h, dout.data_ptr()), reps=20)
        extra["decrypt_fresh"] = {"value": n / s, "unit": "u32/s", "kernel": "decrypt_uniform_kernel", "ms": s * 1e3,
                                  "hbm_GBps": n * 1280 / s / 1e9, "hbm_frac": n * 1280 / s / 1e9 / hbm_peak,
                                  "note": "320 MiB of ciphertext per launch (> 126 MB L2)"}
        # encrypt with values + masks already in HBM, into an existing batch
        g = np.random.default_rng(1)
        dm = torch.from_numpy(np.frombuffer(g.bytes(n * L * 16), dtype=np.uint8).copy()).to(f"cuda:{local_rank}")
        dv = torch.from_numpy(a.view(np.uint8).copy()).to(f"cuda:{local_rank}")
        ce = ca.clone()
        def encd():
            rc = lib.hm_encrypt_device_into(ctx._h, dv.data_ptr(), n, L, dm.data_ptr(), ce._h)
            assert rc == 0, rc
        s = timed(encd, reps=20)
        # shared-memory side of the same kernel: 15 rotated + 1 plain warp-wide LDS.64 per 6 bit-ciphertexts, 2 wavefronts
        # (128 B each) per LDS.64 -> 32 / 6 shared-memory cycles per bit-ciphertext per SM at best
        props = torch.cuda.get_device_properties(local_rank)
        sm_count, sm_clk = props.multi_processor_count, getattr(props, "clock_rate", 1965000) * 1e3  # max SM clock; the run's clocks are in "clocks"
        extra["encrypt"] = {"value": n / s, "unit": "u32/s", "kernel": "encrypt_tab6b_kernel", "ms": s * 1e3,
                            "hbm_GBps": n * 1792 / s / 1e9, "hbm_frac": n * 1792 / s / 1e9 / hbm_peak,
                            "smem_frac": (n * L / s) * (32.0 / 6.0) / (sm_count * sm_clk),
                            "note": "table lookups: 640 B of shared-memory reads per 56 B of HBM traffic, so the binding roofline is the "
                                    "shared-memory pipe (smem_frac = LDS wavefront cycles needed / available), not HBM"}
        ce.free()
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
        # gates on fresh u32 batches (gate_xor / gate_and, common.rs:5-27): n*32 bit-ciphertext pairs per launch
        xo = ctx.apply2(hm.HomomorphicXorGate, ca, cb)
        s = timed(lambda: lib.hm_apply2_into(ctx._h, N.HM_OP_XOR, ca._h, cb._h, xo._h), reps=20)
        extra["xor_gate"] = {"value": n * L / s, "unit": "bit-ciphertext xors/s", "kernel": "xor_flat_kernel", "ms": s * 1e3,
                             "hbm_GBps": n * L * 120 / s / 1e9, "hbm_frac": n * L * 120 / s / 1e9 / hbm_peak}
        xo.free()
        ao = ctx.apply2(hm.HomomorphicAndGate, ca, cb)
        s = timed(lambda: lib.hm_apply2_into(ctx._h, N.HM_OP_AND, ca._h, cb._h, ao._h), reps=5)
        extra["and_gate"] = {"value": n * L / s, "unit": "bit-ciphertext ands/s", "kernel": "mul_small_kernel<8,8>", "ms": s * 1e3,
                             "hbm_GBps": n * L * 152 / s / 1e9, "Tbitmac_per_s": n * L * 66049 / s / 1e12,
                             "alu_frac": n * L * 66049 / s / (lane_ops.value * 32.0)}
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
                           "note": "column-batched circuit: per column one prefix-XOR launch + one batch of independent carry products (mul_small on a side stream, mul_thread32 for the big ones) over a per-value arena in HBM; wall clock incl. launches, median of 5 after 3 warm-up calls"}
        for o8 in (p8, p8b, c8a, c8b):
            o8.free()
        # config 5 (stress): d=d'=512, tau=256, delta=8 fused mul+rem on 2^20 fresh pairs
        rb = np.random.default_rng(55)
        ctxb = hm.Context(hm.Parameters(512, 512, 8, 256), device=local_rank)
        skb = hm.SecretKey.random(512, rb)
        ctxb.set_secret_key(skb)
        ctxb.set_public_key(hm.PublicKey.random(512, 8, 256, skb, rb))
        nb = 1 << 17  # u8 values -> 2^20 pairs
        vb1 = rb.integers(0, 256, size=nb, dtype=np.uint8)
        vb2 = rb.integers(0, 256, size=nb, dtype=np.uint8)
        cb1 = ctxb.encrypt(vb1, np.frombuffer(rb.bytes(nb * 8 * 32), dtype=np.uint8))
        cb2 = ctxb.encrypt(vb2, np.frombuffer(rb.bytes(nb * 8 * 32), dtype=np.uint8))
        mrb = ctxb.poly_mulrem(cb1, cb2)
        streamb = torch.cuda.ExternalStream(lib.hm_context_stream(ctxb._h), device=torch.device("cuda", local_rank))
        ctxb.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(streamb)
        for _ in range(5):
            assert lib.hm_poly_mulrem_into(ctxb._h, cb1._h, cb2._h, mrb._h) == 0
        e1.record(streamb)
        ctxb.synchronize()
        sb = e0.elapsed_time(e1) * 1e-3 / 5
        extra["mulrem_config_b"] = {"metric": "GF(2)[X] mul+rem/s (d=d'=512, tau=256, delta=8)", "value": nb * 8 / sb, "unit": "mul+rem/s",
                                    "kernel": "mulrem_fresh32q_kernel", "pairs_per_launch": nb * 8, "ms": sb * 1e3,
                                    "Tbitmac_per_s": nb * 8 * (1050625 + 788481) / sb / 1e12,
                                    "alu_frac": nb * 8 * (1050625 + 788481) / sb / (lane_ops.value * 32.0),
                                    "product_pipe_frac": 3 * nb * 8 / sb / k8m.value,  # three 8x8-word products per pair (+ the table folds)
                                    "note": "operands reduced mod S first (sliding-window table folds from shared memory), then a 16-word product "
                                            "(three 8x8-word Karatsubas) and one more fold; the fully unrolled first kernel did 275 M/s, the rolled "
                                            "32-word product followed by one fold 720 M/s, reduce-first with conflicting table lookups 1.2-1.35 G/s"}
        for ob in (cb1, cb2, mrb):
            ob.free()
        ctxb.close()
        # PCIe copy rates of this box (the ceiling of every host-buffer call)
        hp = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
        dp_ = torch.empty(1 << 30, dtype=torch.uint8, device=f"cuda:{local_rank}")
        for name, src, dst in (("h2d_GBps", hp, dp_), ("d2h_GBps", dp_, hp)):
            dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            extra[name] = 3 * (1 << 30) / (time.perf_counter() - t0) / 1e9
        del hp, dp_
        extra["e2e_pcie_frac"] = (e2e_value / world) * 46912 / 1e9 / extra["d2h_GBps"]

    # ---- roofline of the dominant kernel (adder_fused_kernel: one launch per step) ---------------------------
    roofline = roofline_hbm = None
    if rank == 0:
        lane_ops = C.c_double(0.0); mhz = C.c_double(0.0)
        lib.hm_measure_alu_peak(ctx._h, C.byref(lane_ops), C.byref(mhz))
        launch_s = float(np.mean(step_ms)) * 1e-3
        peak_bitmac = lane_ops.value * 32.0
        ach = n * BITMACS_PER_ADD / launch_s
        k8 = C.c_double(0.0)
        lib.hm_measure_kara8_peak(ctx._h, C.byref(k8))
        KARA8_PER_ADD = 1 + 30 * 3 + 6 * 465  # per bit: g (k=0), g, g_lo*p, g_hi*p (k=1..30); chain: sum_k ceil((24k-7)/24) = 465 chunks x 6
        k8_ach = n * KARA8_PER_ADD / launch_s
        # The binding unit of this kernel is the integer multiplier (FMA-heavy pipe): the adder is a chain of 8x8-word
        # Karatsuba products, 432 IMAD.WIDE each.  frac = products/s achieved / the same product timed in isolation.
        roofline = {"kernel": "adder_thread_smem_kernel<4> (thread-per-value Karatsuba on IMAD.WIDE + LOP3, operands staged in shared memory by cp.async)", "bound": "alu",
                    "bound_detail": "fma-heavy pipe (integer multiplier): 432 IMAD.WIDE per 8x8-word Karatsuba product, 2 881 products per add",
                    "achieved": k8_ach / 1e9, "peak": k8.value / 1e9, "unit": "G 8x8-word products/s", "frac": k8_ach / k8.value, "traffic": None,
                    "peak_source": "measured in this run: hm_measure_kara8_peak (the product in isolation, 16 warps/SM)",
                    "pipes_busy_ncu": {"alu": 0.58, "fmaheavy": 0.58, "issue_slots": 0.46,
                                       "source": "profiles/r01_adder_thread_smem_ncu_details.txt (75 776 adds)"},
                    "lop3_equivalent": {"achieved": ach / 1e12, "peak": peak_bitmac / 1e12, "unit": "Tbit-MAC/s", "frac": ach / peak_bitmac,
                                        "peak_source": f"measured in this run: LOP3 issue-rate probe, {lane_ops.value / 1e12:.2f} T lane-ops/s at ~{mhz.value:.0f} MHz (x32 bits)",
                                        "note": "the reference's schoolbook AND-XOR pairs (SURVEY.md A.2) against the LOP3-only issue rate, the "
                                                "roofline of the first (comb) kernel; Karatsuba on the multiplier does fewer bit operations, "
                                                "so this ratio exceeds 1"}}
        roofline["product_pipe"] = {k: roofline[k] for k in ("achieved", "peak", "unit", "frac", "peak_source")}  # earlier name of the same figures
        hbm_peak, src = measured_peaks()
        gbs = n * BYTES_PER_ADD / launch_s / 1e9
        roofline_hbm = {"kernel": "adder_thread_smem_kernel<4>", "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                        "frac": gbs / hbm_peak, "traffic": (n * NCU_DRAM_BYTES_PER_ADD) if NCU_DRAM_BYTES_PER_ADD else None,
                        "traffic_source": "ncu --set full capture (profiles/r01_adder_thread_smem_ncu_details.txt), scaled per add; "
                                          "algorithmic bytes per launch = %d" % (n * BYTES_PER_ADD),
                        "peak_source": src}

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
            "config": {"workload": "configs[2]: u32 homomorphic add (ripple-carry XOR/AND circuit) on 2^18 encrypted pairs per GPU, "
                                   "d=dp=128, delta=1, tau=128", "pairs_per_gpu": n, "bits": L,
                       "l2": "inputs 2 x %.0f MiB + result %.2f GiB per step >> 126 MB L2 (no flush needed)" % (n * 1280 / 2**20, n * 46912 / 2**30),
                       "sharding": "independent values split by index across ranks; no collective on the data path"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": ne * 2560, "d2h_bytes_per_step": ne * vwo * 8,
                    "pairs_per_step": ne, "steps": e2e_steps, "step_ms": e2e_step_ms, "call": "hm_apply2_host (pinned host ciphertexts in/out, 3-stream chunked pipeline)",
                    "matches_device_result": e2e_matches},
            "gpu_launches": int(l_after - l_before),
            "kernels_in_step": ["adder_thread_smem_kernel<4>"],
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1 << 18, help="encrypted u32 pairs per GPU per step")
    ap.add_argument("--e2e-pairs", type=int, default=1 << 16, help="pairs per end-to-end step (pinned host buffers)")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
