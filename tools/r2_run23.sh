#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_pytest23.txt
cat gpurun_out/r2_pytest23.txt
SOAK_SEED=11 timeout 900 python tools/soak_parity.py 2>&1 | grep -i "config B\|seeded\|ALL\|MISMATCH"
timeout 600 python tools/enc_ab.py 2>&1 | tee gpurun_out/r2_enc_ab3.txt
ENC_AB_CFG=B HM_ENC_MODE=3 timeout 300 ncu --set full --clock-control none -k regex:encrypt_umma_b -s 2 -c 1 -o /tmp/r02_ummab2 python tools/enc_ab.py child > gpurun_out/r2_ummab_ncu.log 2>&1
ncu -i /tmp/r02_ummab2.ncu-rep --page raw --csv > gpurun_out/r02_ummab_raw.csv 2>/dev/null
