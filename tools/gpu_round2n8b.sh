#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r02x_n8.json 2> gpurun_out/bench_r02x_n8.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r02x_n8.json").readline())
print(d["n_gpus"], "ms/step %.2f value %.3e" % (d["ms_per_step"], d["value"]), d["clocks"])
print({k.replace('xmhw_',''):round(v,2) for k,v in d['kernel_ms'].items()})
PY
nvidia-smi --query-gpu=index,clocks.sm,power.draw,power.limit --format=csv,noheader
