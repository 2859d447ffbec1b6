#!/bin/bash
mkdir -p gpurun_out
summarise() { ncu -i $1 --page raw --csv > $2_raw.csv 2>/dev/null; }
python tools/u8mul_fused_ab.py 16384 > /dev/null 2>&1 && \
ncu --set full --clock-control none -k regex:mul_circuit_fused -s 1 -c 1 -o /tmp/r02_k7b python tools/u8mul_fused_ab.py 16384 > gpurun_out/r2_k7b_ncu.log 2>&1
summarise /tmp/r02_k7b.ncu-rep gpurun_out/r02_k7b
ncu --set full --clock-control none -k regex:"encrypt_tab4b" -c 2 -o /tmp/r02_enc4b python tools/r2_kernel_zoo.py B > gpurun_out/r2_enc4b_ncu.log 2>&1
summarise /tmp/r02_enc4b.ncu-rep gpurun_out/r02_enc4b
ls -la gpurun_out/r02_k7b_raw.csv gpurun_out/r02_enc4b_raw.csv
