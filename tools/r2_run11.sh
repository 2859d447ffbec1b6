#!/bin/bash
# 2 GPUs: full-size tests incl. the new config-B one-wave adder, torchrun N=2 bench, reference arm under torchrun
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_pytest11.txt
cat gpurun_out/r2_pytest11.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2b.json 2> gpurun_out/r2_bench_n2b.err
tail -c 300 gpurun_out/r2_bench_n2b.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_n2.json 2> gpurun_out/r2_bench_ref_n2.err
tail -c 300 gpurun_out/r2_bench_ref_n2.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_n2b.json", "gpurun_out/r2_bench_ref_n2.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "circuit", d.get("e2e_circuit", {}).get("value"), d.get("pcie"), {k: v.get("value") for k, v in d.get("extra", {}).items() if isinstance(v, dict) and "mulrem" in k})
    except Exception as e:
        print(f, "ERR", e)
PY
