#!/bin/bash
# Round profile artefacts (run on the GPU box): launch list of the bench command + one full capture of the
# dominant kernel at the bench configuration.  Outputs land in gpurun_out/.
tag=${1:-r01}
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/launches_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:clim_sweep -s 1 -c 1 -o gpurun_out/sweep_${tag}_config3 \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/sweep_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"exceed4|event_stats_cm|clim_finish_reg|events_stage" -c 4 -o gpurun_out/others_${tag} \
    python bench.py --workload global025_quarter --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/others_${tag}.log 2>&1
cat gpurun_out/bench_${tag}_n1.json
