#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02f_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02f_tests.log
tail -4 gpurun_out/r02f_tests.log
echo "== full fused"; bash tools/kms.sh global025_30yr 2>&1 | tee gpurun_out/r02f_kms.log
echo "== full chain"; XMHW_B200_DETECT=chain bash tools/kms.sh global025_30yr 2>&1 | tee -a gpurun_out/r02f_kms.log
ncu --set full --clock-control none --import-source on -k regex:detect_fused -s 1 -c 1 -o gpurun_out/fused_r02f_quarter \
    python bench.py --workload global025_quarter --steps 1 --warmup 1 --no-e2e --no-cpu --no-api > gpurun_out/fused_r02f.log 2>&1
