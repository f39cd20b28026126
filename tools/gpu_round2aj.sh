#!/bin/bash
# A/B on one box: tensor-memory sweep in blocks of 8 warps (one per SM) vs 4 warps (two per SM)
mkdir -p gpurun_out
BW4=$PWD/xmhw_b200/_xmhw_b200_bw4.so
XMHW_B200_LIB=$BW4 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sweeps_forced or persistent or synth_daily" 2>&1 | tail -2
run() { wl=$1; shift; echo "== $wl $*"; env "$@" timeout 200 bash tools/kms.sh $wl 2>&1 | head -2 | cut -c1-120; }
{
run global025_30yr A=bw8
run global025_30yr XMHW_B200_LIB=$BW4
run global025_quarter A=bw8
run global025_quarter XMHW_B200_LIB=$BW4
} | tee gpurun_out/r02aj_kms.log
for lib in "" "$BW4"; do
echo "== rank 0 of 8 (ocean-balanced range), lib=$lib"
XMHW_B200_LIB=$lib python - <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
from xmhw_b200 import core, synth, shard
tm = synth.daily_time(1982, 2011); doy = synth.doy366(tm); T = len(tm)
nlat, nlon = 720, 1440
land = synth.land_mask(nlat, nlon, 0.33).ravel()
ranges = shard.balanced_ranges((land == 0).astype(np.int64), 8)
for r in (0, 3):
    a, b = ranges[r]
    ts = core.synth_sst_device(T, b - a, synth.season_table(tm), land=land[a:b], cell0=a)
    for _ in range(2): core.threshold_arrays(ts, doy, 366)
    core.TRACE = []
    for _ in range(3): core.threshold_arrays(ts, doy, 366)
    torch.cuda.synchronize()
    ms = [x.elapsed_time(y) for n, x, y in core.TRACE if n == "xmhw_clim_sweep2_f32"]
    core.TRACE = None
    print("rank", r, "cells", b - a, "sweep ms %.3f" % np.mean(ms))
    del ts
PY
done
