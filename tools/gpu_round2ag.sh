#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run() { wl=$1; shift; echo "== $wl $*"; env "$@" timeout 200 bash tools/kms.sh $wl 2>&1; }
{
run global025_30yr A=auto
run global025_quarter A=auto
run global025_skipna99 A=auto
} | tee gpurun_out/r02ag_kms.log
