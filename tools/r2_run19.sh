#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/umma_encrypt_probe > gpurun_out/r2_umma_probe.txt 2>&1
cat gpurun_out/r2_umma_probe.txt
timeout 300 ncu --set full --clock-control none -k regex:umma_encrypt -c 8 -o /tmp/r02_umma ./tools/umma_encrypt_probe > gpurun_out/r2_umma_ncu.log 2>&1
tail -3 gpurun_out/r2_umma_ncu.log
ncu -i /tmp/r02_umma.ncu-rep --page raw --csv > gpurun_out/r02_umma_raw.csv 2>/dev/null
ls -la gpurun_out/r02_umma_raw.csv
