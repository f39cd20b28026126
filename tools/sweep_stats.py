"""Development aid: walk statistics of the sweep on synthetic cells (host emulation)."""
import ctypes as C, sys, numpy as np
sys.path.insert(0, ".")
from xmhw_b200 import plan as P, synth as S, _cabi
lib = C.CDLL("/tmp/_ss.so")
keep = int(sys.argv[1]) if len(sys.argv) > 1 else 6
ncell = int(sys.argv[2]) if len(sys.argv) > 2 else 256
tm = S.daily_time(1982, 2011); doy = S.doy366(tm)
ts = S.synth_sst(len(tm), ncell, S.season_table(tm))
hp = P.build_clim_plan(doy, 366, 5, 0.9, keep=keep)
s, _k = _cabi.numpy_plan_struct(hp)
out = np.zeros(72)
lib.sweep_stats(ts.ctypes.data_as(C.c_void_p), C.c_int64(ncell), C.byref(s), out.ctypes.data_as(C.c_void_p))
names = ["lane |d0|", "lane pops", "lane scans", "warp scans", "warp pop iters", "lane scratch pops", "warp-steps w/ scratch", "warp max|d0|"]
h = out[8:40]; print("rank hist %:", np.round(100 * h / h.sum(), 1))
o = out[40:72]; print("offset (rank - ptr at step start) hist %:", {i - 16: round(100 * o[i] / o.sum(), 2) for i in range(32) if o[i]})
print("keep", keep, " ".join("%s=%.2f" % (n, v) for n, v in zip(names, out)))
