#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:event_stats_cm_kernel -s 1 -c 1 -o gpurun_out/stats_r02am_config3 \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-api > gpurun_out/stats_r02am.log 2>&1
ls -la gpurun_out/stats_r02am_config3.ncu-rep
