#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r02y_n2.json 2> gpurun_out/bench_r02y_n2.err
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-api > gpurun_out/bench_r02y_n1.json 2> gpurun_out/bench_r02y_n1.err
for f in n1 n2; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_r02y_$f.json").readline())
    print(d["n_gpus"], "ms/step %.2f value %.3e extra %s" % (d["ms_per_step"], d["value"], d["warmup_extra_steps"]), d["clocks"], d["result_checksum"])
except Exception as ex:
    print("FAILED", ex); print(open("gpurun_out/bench_r02y_$f.err").read()[-1500:])
PY
done
