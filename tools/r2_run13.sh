#!/bin/bash
# 8 GPUs: torchrun N=8 bench (ours + reference arm) for the committed scaling evidence
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -c 300 gpurun_out/r2_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_n8.json 2> gpurun_out/r2_bench_ref_n8.err
tail -c 300 gpurun_out/r2_bench_ref_n8.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_n8.json", "gpurun_out/r2_bench_ref_n8.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "circuit", d.get("e2e_circuit", {}).get("value"), d.get("pcie"), {k: v.get("value") for k, v in d.get("extra", {}).items() if isinstance(v, dict) and "mulrem" in k}, d.get("cpu_baseline", {}).get("cores"))
    except Exception as e:
        print(f, "ERR", e)
PY
