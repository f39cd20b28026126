#!/bin/bash
# the whole GPU suite with each sweep forced (the default "auto" run is tools/gpu_round2_final.sh)
mkdir -p gpurun_out
for mode in "XMHW_B200_SWEEP=general" "XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_TMEM=0" "XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_TMEM=1"; do
  echo "== $mode"
  env $mode timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
done | tee gpurun_out/r02ab_forced.log
