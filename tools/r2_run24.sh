#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err
tail -c 300 gpurun_out/r2_bench_f.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_f.json').read().strip().splitlines()[-1])
print("ours value %.4g e2e %.4g circuit %.4g ms/step %.2f" % (d["value"], d["e2e"]["value"], d["e2e_circuit"]["value"], d["ms_per_step"]))
for k, v in d["extra"].items():
    if isinstance(v, dict): print(" ", k, v.get("value"), v.get("unit"), v.get("ms"), v.get("error", ""), (v.get("roofline") or v.get("tensor") or {}).get("frac", (v.get("tensor") or {}).get("frac_of_probe_rate")))
PY
python -c "import __graft_entry__ as g; g.smoke()"
