#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/r02z_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02z_tests.log
tail -3 gpurun_out/r02z_tests.log
if ! grep -q "pytest exit 0" gpurun_out/r02z_tests.log; then exit 0; fi
run() { wl=$1; shift; echo "== $wl $*"; env "$@" timeout 200 bash tools/kms.sh $wl 2>&1; }
{
run global025_30yr A=auto
run global025_skipna99 A=auto
} | tee gpurun_out/r02z_kms.log
