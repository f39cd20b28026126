#!/bin/bash
mkdir -p gpurun_out
run() { wl=$1; shift; echo "== $wl $*"; env "$@" timeout 200 bash tools/kms.sh $wl 2>&1 | head -1; }
{
run global025_30yr XMHW_B200_SWEEP2_PERSIST=0
run global025_30yr XMHW_B200_SWEEP2_PERSIST=1
run global025_quarter XMHW_B200_SWEEP2_PERSIST=0
run global025_quarter XMHW_B200_SWEEP2_PERSIST=1
} | tee gpurun_out/r02ah_kms.log
for n in 1 2 4; do
for pz in 0 1; do
echo "== eighth grid (what one of 8 ranks holds), persist=$pz"
XMHW_B200_SWEEP2_PERSIST=$pz python - <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
from xmhw_b200 import core, synth
tm = synth.daily_time(1982, 2011); doy = synth.doy366(tm); T = len(tm)
nlat, nlon = 720, 1440
land = synth.land_mask(nlat, nlon, 0.33).ravel()
w = nlat * nlon // 8
ts = core.synth_sst_device(T, w, synth.season_table(tm), land=land[:w], cell0=0)
for _ in range(2): core.threshold_arrays(ts, doy, 366)
core.TRACE = []
for _ in range(3): core.threshold_arrays(ts, doy, 366)
torch.cuda.synchronize()
ms = [a.elapsed_time(b) for n, a, b in core.TRACE if n == "xmhw_clim_sweep2_f32"]
print("sweep ms", np.mean(ms))
PY
done
break
done
