#!/bin/bash
# GPU session 3: pool A/B, clocks during the pipe probes, parity tests (incl. device groups), bench with the new defaults.
mkdir -p gpurun_out
for pp in "1 16384" "0 16384" "1 1024" "0 1024"; do set -- $pp; HM_PRIVATE_POOL=$1 HM_POOL_MAX_MB=$2 python tools/r2_pool_ab.py; done > gpurun_out/r2_pool_ab.txt 2>&1
cat gpurun_out/r2_pool_ab.txt
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv -lms 100 > gpurun_out/r2_probe_clocks.csv &
SMI=$!
sleep 0.5
./tools/ubench2 400 2>&1 | head -12 > gpurun_out/r2_ubench2c.txt
kill $SMI
cat gpurun_out/r2_ubench2c.txt
awk -F, 'NR>1{print $1","$3","$4","$5}' gpurun_out/r2_probe_clocks.csv | sort | uniq -c | sort -rn | head -20
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest3.txt
cat gpurun_out/r2_pytest3.txt
timeout 900 python bench.py --no-cpu > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
tail -c 400 gpurun_out/r2_bench_b.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_b.json').read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "circuit", d["e2e_circuit"]["value"], d["e2e_circuit"]["step_ms"], "u8", d["extra"]["u8_mul"]["value"], d["extra"]["u8_mul"]["ms_each"])
PY
