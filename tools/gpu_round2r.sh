#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/r02r_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02r_tests.log
tail -5 gpurun_out/r02r_tests.log
if ! grep -q "pytest exit 0" gpurun_out/r02r_tests.log; then exit 0; fi
run() { wl=$1; shift; echo "== $wl $*"; env "$@" timeout 200 bash tools/kms.sh $wl 2>&1; }
{
run global025_30yr XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_TMEM=1
run global025_30yr XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_TMEM=1 XMHW_B200_SWEEP2_TM_SYNC=1
run global025_30yr XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_TMEM=1 XMHW_B200_SWEEP2_TM_SYNC=8
run global025_30yr XMHW_B200_SWEEP=topk
run quarter_w2 XMHW_B200_SWEEP=topk
run quarter_w2 XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_TMEM=1
run quarter_w1 XMHW_B200_SWEEP=topk
run quarter_w3 XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_TMEM=1
run global025_skipna99 XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_TMEM=1
run global025_skipna99 XMHW_B200_SWEEP=general
} | tee gpurun_out/r02r_kms.log
XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_TMEM=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:clim_sweep2_tm -s 1 -c 1 -o gpurun_out/sweep2tm_r02r_config3 \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-api > gpurun_out/sweep2tm_r02r.log 2>&1
