#!/bin/bash
mkdir -p gpurun_out
python tools/u8mul_fused_ab.py 16384 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mul_circuit_fused -s 1 -c 1 -o /tmp/r02_k7 python tools/u8mul_fused_ab.py 16384 > gpurun_out/r2_k7_ncu.log 2>&1
tail -2 gpurun_out/r2_k7_ncu.log
ncu -i /tmp/r02_k7.ncu-rep --page raw --csv > gpurun_out/r02_k7_raw.csv 2>/dev/null
ncu -i /tmp/r02_k7.ncu-rep --page details > gpurun_out/r02_k7_details.txt 2>/dev/null
ncu -i /tmp/r02_k7.ncu-rep --page source --csv > gpurun_out/r02_k7_source.csv 2>/dev/null
ls -la gpurun_out/r02_k7*
