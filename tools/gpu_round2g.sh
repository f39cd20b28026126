#!/bin/bash
# round-2 final state: full GPU tests, bench line, ncu launch list (+ full capture of the dominant kernel)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02g_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02g_tests.log
tail -4 gpurun_out/r02g_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02g_n1.json 2> gpurun_out/bench_r02g_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r02g_n1.json').readline())
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'api', d['e2e'].get('api'))
print(d['cpu_baseline'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02g.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-api > gpurun_out/launches_r02g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:clim_sweep_kernel -s 1 -c 1 -o gpurun_out/sweep_r02g_config3 \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-api > gpurun_out/sweep_r02g.log 2>&1
ls -la gpurun_out | tail -6
