#!/bin/bash
mkdir -p gpurun_out
for n in 2 4 8; do echo "== nslab $n"; timeout 300 python tools/overlap_probe.py $n 2>&1 | tail -2; done | tee gpurun_out/r02y_overlap.log
