#!/bin/bash
# final state on 2 GPUs: strong scaling + checksums (N = 1 vs 2), weak mode, config 5 chunked; 4-GPU strong line when the box has 4
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu --no-api > gpurun_out/bench_r02w_n1.json 2> gpurun_out/bench_r02w_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_r02w_n2.json 2> gpurun_out/bench_r02w_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --scaling weak --no-e2e > gpurun_out/bench_r02w_n2_weak.json 2> gpurun_out/bench_r02w_n2_weak.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu --no-e2e --workload global010_30yr > gpurun_out/bench_r02w_cfg5_n2.json 2> gpurun_out/bench_r02w_cfg5_n2.err
for f in n1 n2 n2_weak cfg5_n2; do echo "== $f"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_r02w_$f.json").readline())
    e=d.get("e2e") or {}
    print(d["n_gpus"], d["scaling"], "ms/step %.2f value %.3e events %d checksum %s e2e %s h2d %s chunks %s" % (d["ms_per_step"], d["value"], d["events"], d["result_checksum"], e.get("value"), e.get("h2d_gbs_this_rank"), d["partition"]["chunks_per_rank"]))
except Exception as ex:
    print("FAILED", ex); print(open("gpurun_out/bench_r02w_$f.err").read()[-1500:])
PY
done
