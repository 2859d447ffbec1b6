#!/bin/bash
mkdir -p gpurun_out
for seed in 42 2027; do
  SOAK_SEED=$seed timeout 1200 python tools/soak_parity.py 2>&1 | tail -30
done | tee gpurun_out/r2_soak.txt
