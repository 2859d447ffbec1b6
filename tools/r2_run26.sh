#!/bin/bash
# final validation of the round: full parity suite, soak (2 fresh seeds), bench + reference arm, smoke
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_pytest26.txt
cat gpurun_out/r2_pytest26.txt
for seed in 314 2718; do SOAK_SEED=$seed timeout 1200 python tools/soak_parity.py 2>&1 | tail -28; done > gpurun_out/r2_soak2.txt
grep -c "^OK" gpurun_out/r2_soak2.txt; grep "ALL\|MISMATCH\|FAIL" gpurun_out/r2_soak2.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref3.json 2> gpurun_out/r2_bench_ref3.err
timeout 900 python bench.py > gpurun_out/r2_bench_g.json 2> gpurun_out/r2_bench_g.err
tail -c 200 gpurun_out/r2_bench_g.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_g.json').read().strip().splitlines()[-1])
print("ours value %.4g e2e %.4g circuit %.4g ms/step %.2f frac %.3f" % (d["value"], d["e2e"]["value"], d["e2e_circuit"]["value"], d["ms_per_step"], d["roofline"]["frac"]))
for k, v in d["extra"].items():
    if isinstance(v, dict): print(" ", k, v.get("value"), v.get("unit"), v.get("ms"), v.get("error", ""), (v.get("roofline") or {}).get("frac", (v.get("tensor") or {}).get("frac_of_probe_rate")))
r = json.loads(open('gpurun_out/r2_bench_ref3.json').read().strip().splitlines()[-1])
print("ref", r["value"], r["e2e"]["value"], r["cpu_baseline"]["cores"])
PY
python -c "import __graft_entry__ as g; g.smoke()"
