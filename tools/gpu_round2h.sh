#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/r02h_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02h_tests.log
tail -3 gpurun_out/r02h_tests.log
echo "== quarter"; bash tools/kms.sh global025_quarter 2>&1 | tee gpurun_out/r02h_kms.log
echo "== full"; bash tools/kms.sh global025_30yr 2>&1 | tee -a gpurun_out/r02h_kms.log
ncu --set full --clock-control none --import-source on -k regex:clim_sweep_kernel -s 1 -c 1 -o gpurun_out/sweep_r02h_quarter \
    python bench.py --workload global025_quarter --steps 1 --warmup 1 --no-e2e --no-cpu --no-api > gpurun_out/sweep_r02h.log 2>&1
