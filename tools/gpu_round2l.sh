#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sweeps_forced or synth_daily or whole_warp or config4" > gpurun_out/r02l_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02l_tests.log
tail -3 gpurun_out/r02l_tests.log
run() { echo "== $*"; env "$@" bash tools/kms.sh global025_30yr 2>&1; }
{
run A=default
run XMHW_B200_SWEEP_L2=0
run XMHW_B200_SWEEP_NEAR=6
run XMHW_B200_SWEEP_NEAR=11
run XMHW_B200_SWEEP_MINB=14 XMHW_B200_POOL_ROWS=120
} | tee gpurun_out/r02l_kms.log
ncu --set full --clock-control none --import-source on -k regex:clim_sweep_kernel -s 1 -c 1 -o gpurun_out/sweep_r02l_config3 \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-api > gpurun_out/sweep_r02l.log 2>&1
