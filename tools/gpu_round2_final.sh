#!/bin/bash
# round-2 final state: GPU tests, smoke, default bench line, config-4 bench lines, ncu launch list + full capture of the sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02f_tests.log
tail -3 gpurun_out/r02f_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r02final_n1.json 2> gpurun_out/bench_r02final_n1.err
python bench.py --workload global025_skipna99 --steps 3 --warmup 3 --no-cpu --no-api > gpurun_out/bench_r02final_cfg4i.json 2> gpurun_out/bench_r02final_cfg4i.err
python bench.py --workload global025_pentad --steps 3 --warmup 3 --no-cpu --no-api > gpurun_out/bench_r02final_cfg4ii.json 2> gpurun_out/bench_r02final_cfg4ii.err
python bench.py --workload regional_40yr --steps 3 --warmup 3 --no-cpu --no-api > gpurun_out/bench_r02final_cfg2.json 2> gpurun_out/bench_r02final_cfg2.err
for f in n1 cfg4i cfg4ii cfg2; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_r02final_$f.json").readline())
    e=d.get("e2e") or {}
    print("$f", d["config"]["workload"], "ms/step %.2f value %.3e events %d e2e %s kernel %s frac %.3f launches %s" % (d["ms_per_step"], d["value"], d["events"], e.get("value"), d["roofline"]["kernel"], d["roofline"]["frac"], d["gpu_launches"]))
except Exception as ex:
    print("$f FAILED", ex); print(open("gpurun_out/bench_r02final_$f.err").read()[-800:])
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02final.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-api > gpurun_out/launches_r02final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:clim_sweep2_tm -s 1 -c 1 -o gpurun_out/sweep2tm_r02final_config3 \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-api > gpurun_out/sweep2tm_r02final.log 2>&1
ls -la gpurun_out | grep r02final
