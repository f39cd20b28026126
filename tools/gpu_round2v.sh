#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02v_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02v_tests.log
tail -4 gpurun_out/r02v_tests.log
if ! grep -q "pytest exit 0" gpurun_out/r02v_tests.log; then exit 0; fi
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02v_n1.json 2> gpurun_out/bench_r02v_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r02v_n1.json').readline())
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'api', d['e2e'].get('api'))
print(d['roofline']); print(d['kernel_ms']); print(d['cpu_baseline']); print('launches', d['gpu_launches'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02v.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-api > gpurun_out/launches_r02v.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02v_ref.json 2> gpurun_out/bench_r02v_ref.err; tail -c 600 gpurun_out/bench_r02v_ref.json
