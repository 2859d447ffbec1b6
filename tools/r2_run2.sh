#!/bin/bash
# GPU session 2 of round 2: probes, parity tests, the full bench line, adder / pool A/B, ncu capture of the adder chain.
mkdir -p gpurun_out
./tools/ubench2 120 > gpurun_out/r2_ubench2b.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest2.txt
cat gpurun_out/r2_pytest2.txt
timeout 900 python bench.py > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
tail -c 600 gpurun_out/r2_bench_a.err
out=gpurun_out/r2_ab_adder2.txt
: > $out
for cfg in "4 4 1024 96" "4 8 1024 96" "3 8 1024 96" "13 8 1024 96" "4 6 1024 96" "4 8 16384 96" "4 8 16384 256"; do
  set -- $cfg
  line=$(HM_ADDER_CHAIN=$1 HM_ADDER_PHASES=$2 HM_POOL_MAX_MB=$3 HM_HOST_CHUNK_MB=$4 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu 2>gpurun_out/r2_ab2_err_$1_$2_$3_$4.log | tail -1)
  echo "chain=$1 phases=$2 pool_mb=$3 chunk_mb=$4 $(echo "$line" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("adds/s=%.4g e2e=%.4g e2e_circuit=%.4g circuit_ms=%s ok=%s" % (d["value"], d["e2e"]["value"], d["e2e_circuit"]["value"], [round(x,1) for x in d["e2e_circuit"]["step_ms"]], d["decrypted_sums_correct_frac"]))')" >> $out
done
cat $out
ARGS="--pairs 75776 --e2e-pairs 4096 --circuit-pairs 4096 --steps 1 --warmup 3 --no-extra --no-cpu"
python bench.py $ARGS > gpurun_out/r2_ads_plain.log 2>&1 && \
ncu --set full --metrics smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_fmalite.sum,smsp__inst_executed_pipe_lsu.sum,smsp__inst_executed.sum \
    --clock-control none --import-source on -k regex:adder_chain -s 3 -c 1 -o gpurun_out/r02_adder_chain python bench.py $ARGS > gpurun_out/r2_ads_ncu.log 2>&1
tail -3 gpurun_out/r2_ads_ncu.log
