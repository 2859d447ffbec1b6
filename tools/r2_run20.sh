#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "encrypt or seeded or entry_points" 2>&1 | tail -4 > gpurun_out/r2_pytest20.txt
cat gpurun_out/r2_pytest20.txt
ENC_AB_ONLY=B timeout 600 python tools/enc_ab.py 2>&1 | tee gpurun_out/r2_enc_ab3.txt
