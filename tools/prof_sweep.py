"""Run only the climatology sweep (for ncu)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from xmhw_b200 import core, synth
from xmhw_b200._cabi import lib, check
nlat, nlon = int(sys.argv[1]), int(sys.argv[2]); y0, y1 = 1982, int(sys.argv[3]) if len(sys.argv) > 3 else 2011
tm = synth.daily_time(y0, y1); doy = synth.doy366(tm); T = len(tm); ngrid = nlat * nlon
ts = core.synth_sst_device(T, ngrid, synth.season_table(tm), land=synth.land_mask(nlat, nlon).ravel())
th, se = core.threshold_arrays(ts, doy, 366)
ev = core.detect_arrays(ts, doy, 366, th, se)
torch.cuda.synchronize()
print("done", len(ev))
