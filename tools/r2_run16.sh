#!/bin/bash
# final-kernels session: full parity suite, bench, reference arm, launch list of the bench under ncu
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_pytest16.txt
cat gpurun_out/r2_pytest16.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref2.json 2> gpurun_out/r2_bench_ref2.err
timeout 900 python bench.py > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err
tail -c 300 gpurun_out/r2_bench_e.err
BARGS="--steps 2 --warmup 3 --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches_bench2.csv python bench.py $BARGS > gpurun_out/r2_launch_ncu2.log 2>&1
tail -2 gpurun_out/r2_launch_ncu2.log
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_e.json').read().strip().splitlines()[-1])
print("ours value %.4g e2e %.4g circuit %.4g ms/step %.2f" % (d["value"], d["e2e"]["value"], d["e2e_circuit"]["value"], d["ms_per_step"]))
for k, v in d["extra"].items():
    if isinstance(v, dict): print(" ", k, v.get("value"), v.get("unit"), v.get("ms"), v.get("error", ""), (v.get("roofline") or {}).get("frac"))
r = json.loads(open('gpurun_out/r2_bench_ref2.json').read().strip().splitlines()[-1])
print("ref", r["value"], r["e2e"]["value"], r["cpu_baseline"]["cores"])
PY
