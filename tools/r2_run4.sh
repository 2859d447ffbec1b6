#!/bin/bash
# GPU session 4: parity tests, then ncu evidence for round 2 — adder chain at the bench size, the other kernels, the bench's
# launch list.  The .ncu-rep files are summarised here (details + selected raw metrics) and only the adder's is kept: gpurun
# brings back at most 64 MiB.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest4.txt
cat gpurun_out/r2_pytest4.txt
summarise() { # $1 = report, $2 = output prefix
  ncu -i $1 --page details > $2_details.txt 2>/dev/null
  ncu -i $1 --page raw --csv > $2_raw.csv 2>/dev/null
}
ARGS="--steps 1 --warmup 3 --no-extra --no-cpu --e2e-pairs 4096 --circuit-pairs 4096"
python bench.py $ARGS > gpurun_out/r2_adsfull_plain.log 2>&1 && \
ncu --set full --metrics smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_lsu.sum,smsp__inst_executed.sum \
    --clock-control none --import-source on -k regex:adder_chain -s 3 -c 1 -o gpurun_out/r02_adder_chain_full python bench.py $ARGS > gpurun_out/r2_adsfull_ncu.log 2>&1
tail -2 gpurun_out/r2_adsfull_ncu.log
summarise gpurun_out/r02_adder_chain_full.ncu-rep gpurun_out/r02_adder_chain_full
python tools/r2_kernel_zoo.py > gpurun_out/r2_zoo_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:"encrypt_tab6b|decrypt_uniform|xor_flat|mul_small|mulrem_fresh_a|decrypt_value_tma|mask_fill" -c 30 -o /tmp/r02_zoo_a python tools/r2_kernel_zoo.py > gpurun_out/r2_zoo_ncu.log 2>&1
tail -2 gpurun_out/r2_zoo_ncu.log
summarise /tmp/r02_zoo_a.ncu-rep gpurun_out/r02_zoo_a
python tools/r2_kernel_zoo.py B > gpurun_out/r2_zoob_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:"encrypt_tab_kernel|mulrem_fresh32q|mask_fill" -c 8 -o /tmp/r02_zoo_b python tools/r2_kernel_zoo.py B > gpurun_out/r2_zoob_ncu.log 2>&1
tail -2 gpurun_out/r2_zoob_ncu.log
summarise /tmp/r02_zoo_b.ncu-rep gpurun_out/r02_zoo_b
BARGS="--steps 2 --warmup 3 --no-cpu"
python bench.py $BARGS > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py $BARGS > gpurun_out/r2_launch_ncu.log 2>&1
tail -2 gpurun_out/r2_launch_ncu.log
du -sh gpurun_out; ls -la gpurun_out | head -40
