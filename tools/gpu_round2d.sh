#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02d_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02d_tests.log
tail -4 gpurun_out/r02d_tests.log
echo "== skipna99 auto(topk)"; bash tools/kms.sh global025_skipna99 2>&1 | tee gpurun_out/r02d_kms.log
echo "== skipna99 general"; XMHW_B200_SWEEP=general bash tools/kms.sh global025_skipna99 2>&1 | tee -a gpurun_out/r02d_kms.log
echo "== pentad"; bash tools/kms.sh global025_pentad 2>&1 | tee -a gpurun_out/r02d_kms.log
python bench.py --steps 5 --warmup 3 --cpu-cells 16 > gpurun_out/bench_r02d_n1.json 2> gpurun_out/bench_r02d_n1.err
cat gpurun_out/bench_r02d_n1.json | cut -c1-3000
tail -3 gpurun_out/bench_r02d_n1.err
