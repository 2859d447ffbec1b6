#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "encrypt or entry_points or group or seeded" 2>&1 | tail -15 > gpurun_out/r2_pytest8.txt
cat gpurun_out/r2_pytest8.txt
timeout 600 python tools/enc_ab.py 2>&1 | tee gpurun_out/r2_enc_ab.txt
