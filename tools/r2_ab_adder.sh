#!/bin/bash
# Round-2 A/B of the adder kernels on one B200: HM_ADDER_CHAIN (0 = round-1 kernel) x HM_ADDER_PHASES, device-timed adds/s.
mkdir -p gpurun_out
out=gpurun_out/r2_ab_adder.txt
: > $out
for cfg in "0 4" "4 1" "4 2" "4 4" "4 8" "3 4" "13 4" "12 4"; do
  set -- $cfg
  line=$(HM_ADDER_CHAIN=$1 HM_ADDER_PHASES=$2 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu --e2e-pairs 4096 2>gpurun_out/r2_ab_err_$1_$2.log | tail -1)
  echo "chain=$1 phases=$2 $(echo "$line" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("adds/s=%.4g ms/step=%.3f ok=%s clocks=%s" % (d["value"], d["ms_per_step"], d["decrypted_sums_correct_frac"], d["clocks"]))')" >> $out
done
cat $out
