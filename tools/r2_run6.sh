#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2_pytest6.txt
cat gpurun_out/r2_pytest6.txt
python - <<'PY'
import numpy as np, time
import homomorph_rust_b200 as hm
ctx = hm.Context(hm.Parameters(128, 128, 1, 128)); ctx.generate_keys_seeded(1)
rng = np.random.default_rng(1)
for n in (4, 256, 2048):
    a = rng.integers(0, 65536, size=n, dtype=np.uint16); b = rng.integers(0, 65536, size=n, dtype=np.uint16)
    ca, cb = ctx.encrypt(a, seed=1), ctx.encrypt(b, seed=2)
    for it in range(2):
        l0 = ctx.kernel_launches(); t0 = time.perf_counter()
        p = ctx.apply2(hm.HomomorphicMultiplication, ca, cb); ctx.synchronize()
        dt = time.perf_counter() - t0
        print(f"u16 mul n={n}: {dt*1e3:.1f} ms, {ctx.kernel_launches()-l0} launches, {n/dt:.1f} muls/s")
        p.free()
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
tail -c 300 gpurun_out/r2_bench_final.err
python - <<'PY'
import json
r = json.loads(open('gpurun_out/r2_bench_ref.json').read().strip().splitlines()[-1])
d = json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
print("ref", r["value"], r["e2e_circuit"]["value"], "cores", r["cpu_baseline"]["cores"])
print("ours value %.4g e2e %.4g circuit %.4g" % (d["value"], d["e2e"]["value"], d["e2e_circuit"]["value"]))
ro = d["roofline"]; print("frac", ro["frac"], ro["busier_pipe"], {k: (v["frac"], v["peak"]) for k, v in ro["pipes"].items()}, ro["joint_issue_ceiling"]["frac_of_mix_ceiling"], ro["probe_clocks"], ro["ncu_pipe_pct_scaled_to_this_rate"])
print({k: v.get("value") for k, v in d["extra"].items() if isinstance(v, dict)})
PY
