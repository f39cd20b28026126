#!/bin/bash
# 4-GPU strong scaling line of the final state (checksums must equal the 1- and 2-GPU runs)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_r02w_n8.json 2> gpurun_out/bench_r02w_n8.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_r02w_n8.json").readline())
    e=d.get("e2e") or {}
    print(d["n_gpus"], d["scaling"], "ms/step %.2f value %.3e events %d checksum %s e2e %s h2d %s" % (d["ms_per_step"], d["value"], d["events"], d["result_checksum"], e.get("value"), e.get("h2d_gbs_this_rank")))
except Exception as ex:
    print("FAILED", ex); print(open("gpurun_out/bench_r02w_n8.err").read()[-1500:])
PY
