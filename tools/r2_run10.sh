#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err
tail -c 400 gpurun_out/r2_bench_d.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_d.json').read().strip().splitlines()[-1])
print("ours value %.4g e2e %.4g circuit %.4g ms/step %.2f" % (d["value"], d["e2e"]["value"], d["e2e_circuit"]["value"], d["ms_per_step"]))
ro = d["roofline"]; print("frac", ro["frac"], ro.get("busier_pipe"))
for k, v in d["extra"].items():
    if isinstance(v, dict): print(" ", k, v.get("value"), v.get("unit"), v.get("ms"), v.get("error", ""))
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
python -c "import __graft_entry__ as g; g.smoke()"
