#!/bin/bash
mkdir -p gpurun_out
ENC_AB_CFG=B HM_ENC_MODE=3 timeout 300 ncu --set full --clock-control none -k regex:encrypt_umma_b -s 2 -c 2 -o /tmp/r02_ummab python tools/enc_ab.py child > gpurun_out/r2_ummab_ncu.log 2>&1
tail -3 gpurun_out/r2_ummab_ncu.log
ncu -i /tmp/r02_ummab.ncu-rep --page raw --csv > gpurun_out/r02_ummab_raw.csv 2>/dev/null
ls -la gpurun_out/r02_ummab_raw.csv
