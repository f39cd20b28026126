"""Development probe: does the ALU-bound climatology sweep of one column block overlap with the memory-bound
detect chain of another when they run on two streams?  (python tools/overlap_probe.py [nslab])"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.build()
from xmhw_b200 import core, synth
nslab = int(sys.argv[1]) if len(sys.argv) > 1 else 4
nlat, nlon = 720, 1440
tm = synth.daily_time(1982, 2011); doy = synth.doy366(tm); T = len(tm)
land = synth.land_mask(nlat, nlon, 0.33).ravel()
sea = synth.season_table(tm)
ngrid = nlat * nlon
w = ngrid // nslab
slabs = [core.synth_sst_device(T, w, sea, land=land[i * w:(i + 1) * w], cell0=i * w) for i in range(nslab)]
def seq():
    for ts in slabs:
        th, se = core.threshold_arrays(ts, doy, 366)
        core.detect_arrays(ts, doy, 366, th, se)
def piped():
    main = torch.cuda.current_stream()
    s2 = torch.cuda.Stream()
    def detect_on_s2(ts, th, se, ready):
        with torch.cuda.stream(s2):
            s2.wait_event(ready)
            for x in (ts, th, se):
                x.record_stream(s2)
            core.detect_arrays(ts, doy, 366, th, se)          # its one host sync only waits for stream s2
    pend = None
    for ts in slabs:
        th, se = core.threshold_arrays(ts, doy, 366)          # sweep + finish of THIS block enqueued on main first
        ready = torch.cuda.Event(); ready.record(main)
        if pend is not None:
            detect_on_s2(*pend)                               # detect of the PREVIOUS block runs beside it
        pend = (ts, th, se, ready)
    detect_on_s2(*pend)
    main.wait_stream(s2)
for name, fn in (("sequential", seq), ("two streams", piped)):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize(); print(name, "%.2f ms/step" % ((time.perf_counter() - t0) / 3 * 1e3))
