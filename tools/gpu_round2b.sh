#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02b_tests.log
tail -5 gpurun_out/r02b_tests.log
for wl in global025_quarter global025_30yr; do
  echo "== $wl topk"; bash tools/kms.sh $wl
done 2>&1 | tee gpurun_out/r02b_kms.log
ncu --set full --clock-control none --import-source on -k regex:clim_sweep2 -s 1 -c 1 -o gpurun_out/sweep2_r02b_quarter \
    python bench.py --workload global025_quarter --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/sweep2_r02b.log 2>&1
