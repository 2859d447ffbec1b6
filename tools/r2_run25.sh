#!/bin/bash
mkdir -p gpurun_out
ENC_AB_CFG=B HM_ENC_MODE=3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:encrypt_umma_b -s 2 -c 1 -o /tmp/r02_ummab3 python tools/enc_ab.py child > gpurun_out/r2_ummab_ncu.log 2>&1
ncu -i /tmp/r02_ummab3.ncu-rep --page source --csv > gpurun_out/r02_ummab_source.csv 2>/dev/null
ls -la gpurun_out/r02_ummab_source.csv
