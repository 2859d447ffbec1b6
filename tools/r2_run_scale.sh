#!/bin/bash
# usage: r2_run_scale.sh N  — torchrun N-rank bench (ours + reference arm) of the final build
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_final_n$N.json 2> gpurun_out/r2_final_n$N.err
tail -c 200 gpurun_out/r2_final_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r2_final_ref_n$N.json 2> gpurun_out/r2_final_ref_n$N.err
python - <<PY
import json
for f in ("gpurun_out/r2_final_n$N.json", "gpurun_out/r2_final_ref_n$N.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "circuit", (d.get("e2e_circuit") or {}).get("value"), (d.get("pcie") or {}).get("d2h_GBps_all_ranks_concurrent"), {k: v.get("value") for k, v in d.get("extra", {}).items() if isinstance(v, dict) and "mulrem" in k})
    except Exception as e:
        print(f, "ERR", e)
PY
