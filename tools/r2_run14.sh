#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "mul" 2>&1 | tail -4 > gpurun_out/r2_pytest14.txt
cat gpurun_out/r2_pytest14.txt
timeout 600 python tools/u8mul_fused_ab.py 16 256 1024 4096 16384 65536 2>&1 | tee gpurun_out/r2_u8mul_fused_ab.txt
