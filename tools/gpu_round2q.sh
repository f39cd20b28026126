#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sweeps_forced or synth_daily or whole_warp" > gpurun_out/r02q_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02q_tests.log
tail -5 gpurun_out/r02q_tests.log
if ! grep -q "pytest exit 0" gpurun_out/r02q_tests.log; then exit 0; fi
run() { echo "== $*"; env "$@" timeout 200 bash tools/kms.sh global025_30yr 2>&1; }
{
run XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_TMEM=1
run XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_TMEM=1 XMHW_B200_SWEEP2_ORDER=0
run XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_TMEM=1 XMHW_B200_SWEEP2_TM_SYNC=0
run XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_TMEM=1 XMHW_B200_SWEEP2_TM_SYNC=4
run XMHW_B200_SWEEP=topk
run XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_ORDER=0
} | tee gpurun_out/r02q_kms.log
XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_TMEM=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:clim_sweep2_tm -s 1 -c 1 -o gpurun_out/sweep2tm_r02q_config3 \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-api > gpurun_out/sweep2tm_r02q.log 2>&1
