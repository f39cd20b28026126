#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02s_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02s_tests.log
tail -5 gpurun_out/r02s_tests.log
if ! grep -q "pytest exit 0" gpurun_out/r02s_tests.log; then exit 0; fi
run() { wl=$1; shift; echo "== $wl $*"; env "$@" timeout 200 bash tools/kms.sh $wl 2>&1; }
{
run global025_30yr A=auto
run global025_30yr XMHW_B200_SWEEP2_TM_SYNC=8
run global025_30yr XMHW_B200_SWEEP2_TM_SYNC=16
run global025_30yr XMHW_B200_SWEEP2_TM_SYNC=0
run global025_30yr XMHW_B200_SWEEP2_TMEM=0 XMHW_B200_SWEEP=topk
run global025_30yr XMHW_B200_SWEEP=general
run quarter_w2 A=auto
run regional_40yr A=auto
run regional_40yr XMHW_B200_SWEEP=general
run global025_pentad A=auto
run global025_pentad XMHW_B200_SWEEP=general
} | tee gpurun_out/r02s_kms.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:clim_sweep2_tm -s 1 -c 1 -o gpurun_out/sweep2tm_r02s_config3 \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-api > gpurun_out/sweep2tm_r02s.log 2>&1
