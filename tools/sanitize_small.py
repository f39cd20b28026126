"""Small end-to-end run for compute-sanitizer (memcheck / racecheck)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from xmhw_b200 import core, synth
for years, ncell, ndoy_mode in (((2001, 2006), 70, "daily"), ((2001, 2004), 33, "pentad")):
    if ndoy_mode == "daily":
        tm = synth.daily_time(*years); doy = synth.doy366(tm); ndoy = 366; sea = synth.season_table(tm); T = len(tm)
        kw = {}
    else:
        doy = np.tile(np.arange(1, 74), 6); ndoy = 73; T = len(doy); sea = synth.season_table(T)
        kw = dict(smoothPercentileWidth=5, feb29=False)
    land = np.zeros(ncell, np.uint8); land[3] = 1
    ts = core.synth_sst_device(T, ncell, sea, land=land, nan_ppm=5000)
    th, se = core.threshold_arrays(ts, doy, ndoy, **kw)
    ev = core.detect_arrays(ts, doy, ndoy, th, se)
    torch.cuda.synchronize()
    print(ndoy_mode, "events", len(ev), "nan thresh cols", int(torch.isnan(th).all(0).sum()))
host = torch.from_numpy(synth.synth_sst(731, 96, synth.season_table(731))).pin_memory()
res = core.threshold_detect_host(host, synth.doy366(synth.daily_time(2003, 2004)), 366, slabs=3)
print("host path events", res["n_events"])
