#!/bin/bash
# A/B on ONE box: partial sums of the push conversion (1 = sequential, 2, 3)
mkdir -p gpurun_out
run() { wl=$1; shift; echo "== $wl $*"; env "$@" timeout 200 bash tools/kms.sh $wl 2>&1; }
{
for rep in 1 2; do
run global025_30yr A=sums1
run global025_30yr XMHW_B200_LIB=$PWD/xmhw_b200/_xmhw_b200_sums2.so
run global025_30yr XMHW_B200_LIB=$PWD/xmhw_b200/_xmhw_b200_sums3.so
done
} | tee gpurun_out/r02ac_kms.log
