#!/bin/bash
# SM clock, power and throttle reasons while the tensor-core encrypt kernel (config B) and the table kernel run back to back for a few seconds each
mkdir -p gpurun_out
for mode in 3 2; do
  nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active --format=csv,noheader -lms 200 > gpurun_out/r2_umma_clocks_mode$mode.csv &
  SMI=$!
  ENC_AB_CFG=B HM_ENC_MODE=$mode python - <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import homomorph_rust_b200 as hm
ctx = hm.Context(hm.Parameters(512, 512, 8, 256)); ctx.generate_keys_seeded(1)
lib = hm.lib()
n, L = 1 << 18, 32
a = np.random.default_rng(3).integers(0, 2**32, size=n, dtype=np.uint32)
dv = torch.from_numpy(a.view(np.uint8).copy()).cuda()
ce = ctx.encrypt(a, seed=1); ctx.synchronize()
t0 = time.perf_counter(); reps = 0
while time.perf_counter() - t0 < 4.0:
    for _ in range(50):
        assert lib.hm_encrypt_device_seeded_into(ctx._h, dv.data_ptr(), n, L, 12345, 0, ce._h) == 0
    ctx.synchronize(); reps += 50
dt = time.perf_counter() - t0
print(f"HM_ENC_MODE={os.environ['HM_ENC_MODE']}: {dt / reps * 1e6:.1f} us per 2^18 u32 sustained over {dt:.1f} s")
PY
  kill $SMI
  echo "mode $mode clocks (MHz, max, W, reasons) — median of the loaded samples:"
  sort -t, -k3 -n -r gpurun_out/r2_umma_clocks_mode$mode.csv | head -12 | sort | uniq -c | sort -rn | head -4
done
