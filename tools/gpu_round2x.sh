#!/bin/bash
# compute-sanitizer memcheck of the top-K sweep kernels (shared-memory and tensor-memory variants) on small cases
mkdir -p gpurun_out
for tm in 0 1; do
  echo "== memcheck XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_TMEM=$tm"
  XMHW_B200_SWEEP=topk XMHW_B200_SWEEP2_TMEM=$tm timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/sanitize_r02x_tm$tm.log 2>&1
  echo "exit $?"; tail -6 gpurun_out/sanitize_r02x_tm$tm.log
done
echo "== 30-year case, default selection (tensor-memory kernel), memcheck"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python - > gpurun_out/sanitize_r02x_30yr.log 2>&1 <<'PY'
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from xmhw_b200 import core, synth
tm = synth.daily_time(1982, 2011); doy = synth.doy366(tm)
land = synth.land_mask(8, 40).ravel()
ts = core.synth_sst_device(len(tm), 320, synth.season_table(tm), land=land, nan_ppm=3000)
core.TRACE = []
th, se = core.threshold_arrays(ts, doy, 366)
torch.cuda.synchronize()
print([n for n, _, _ in core.TRACE], "nan cols", int(torch.isnan(th).all(0).sum()))
PY
echo "exit $?"; tail -5 gpurun_out/sanitize_r02x_30yr.log
