#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "synth_daily or config4 or oisst or pentad or leap" > gpurun_out/r02c_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c_tests.log
tail -3 gpurun_out/r02c_tests.log
for w in 1 2 4; do
  echo "== quarter WPB=$w"; XMHW_B200_SWEEP2_WPB=$w bash tools/kms.sh global025_quarter
done 2>&1 | tee gpurun_out/r02c_kms.log
echo "== full default"; bash tools/kms.sh global025_30yr 2>&1 | tee -a gpurun_out/r02c_kms.log
ncu --set full --clock-control none --import-source on -k regex:clim_sweep2 -s 1 -c 1 -o gpurun_out/sweep2_r02c_quarter \
    python bench.py --workload global025_quarter --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/sweep2_r02c.log 2>&1
