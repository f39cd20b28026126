#!/bin/bash
mkdir -p gpurun_out
run() { wl=$1; shift; echo "== $wl $*"; env "$@" timeout 200 bash tools/kms.sh $wl 2>&1; }
{
run global025_30yr A=auto
run quarter_w2 A=auto
run quarter_w2 XMHW_B200_SWEEP2_SYNC=1
run quarter_w1 A=auto
run quarter_w1 XMHW_B200_SWEEP2_SYNC=1
run global025_skipna99 A=auto
run global025_pentad A=auto
run global025_pentad XMHW_B200_SWEEP2_SYNC=1
} | tee gpurun_out/r02t_kms.log
