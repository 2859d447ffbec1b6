#!/bin/bash
# GPU session 5 (2 GPUs): device groups on two real devices, torchrun N=2 bench, single-process group bench.
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q -k "group or keys_on_device or u16_column or multiword" 2>&1 | tail -8 > gpurun_out/r2_pytest5.txt
cat gpurun_out/r2_pytest5.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -c 300 gpurun_out/r2_bench_n2.err
timeout 600 python bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_group2.json 2> gpurun_out/r2_bench_group2.err
tail -c 300 gpurun_out/r2_bench_group2.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_n2.json", "gpurun_out/r2_bench_group2.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "circuit", d.get("e2e_circuit", {}).get("value"), d.get("pcie"), {k: v.get("value") for k, v in d.get("extra", {}).items() if isinstance(v, dict) and "mulrem" in k})
    except Exception as e:
        print(f, "ERR", e)
PY
