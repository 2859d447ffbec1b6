#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "wide" 2>&1 | tail -15 > gpurun_out/r2_pytest7.txt
cat gpurun_out/r2_pytest7.txt
timeout 600 python tools/adder_wide_ab.py 2>&1 | tee gpurun_out/r2_wide_ab.txt
