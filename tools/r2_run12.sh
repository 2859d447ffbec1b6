#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "encrypt or entry_points or group or seeded or config5 or mulrem_fresh" 2>&1 | tail -15 > gpurun_out/r2_pytest12.txt
cat gpurun_out/r2_pytest12.txt
timeout 600 python tools/enc_ab.py 2>&1 | tee gpurun_out/r2_enc_ab2.txt
