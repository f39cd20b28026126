"""Per-kernel CUDA-event timing of the hot path on synthetic data (development aid)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from xmhw_b200 import core, synth, _cabi
from xmhw_b200._cabi import lib, check

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

def run(nlat, nlon, y0, y1, land_frac=0.33, reps=2):
    tm = synth.daily_time(y0, y1); doy = synth.doy366(tm); T = len(tm); ngrid = nlat * nlon
    land = synth.land_mask(nlat, nlon, land_frac).ravel() if land_frac else None
    ts = core.synth_sst_device(T, ngrid, synth.season_table(tm), land=land)
    torch.cuda.synchronize()
    nocean = ngrid - (int(land.sum()) if land is not None else 0)
    print(f"== {nlat}x{nlon} T={T} ngrid={ngrid} ocean={nocean} bytes={T*ngrid*4/1e9:.2f} GB")
    dp = core.device_plan(doy, 366, 5, 0.9, ts.device)
    print("plan: scratch_rows", dp.host.scratch_rows, "pool_rows", dp.host.pool_rows, "smem KB", dp.host.smem_bytes()/1024, "max_lists", dp.host.max_lists)
    st = torch.cuda.current_stream().cuda_stream
    for r in range(reps):
        raw_t = torch.empty((366, ngrid), dtype=torch.float64, device="cuda"); raw_s = torch.empty_like(raw_t)
        out_t = torch.empty_like(raw_t); out_s = torch.empty_like(raw_t)
        e0 = ev()
        scratch = torch.empty(max(1, ((ngrid + 31) // 32) * dp.host.scratch_rows * 32), dtype=torch.int32, device="cuda")
        torch.cuda.synchronize(); e0 = ev()
        check(lib.xmhw_clim_sweep_f32(ts.data_ptr(), T, ngrid, dp.struct, raw_t.data_ptr(), raw_s.data_ptr(), scratch.data_ptr(), st), "sweep")
        e1 = ev()
        check(lib.xmhw_clim_finish2_f64(raw_t.data_ptr(), out_t.data_ptr(), raw_s.data_ptr(), out_s.data_ptr(), 366, ngrid, 1, 31, st), "fin")
        e2 = ev()
        del raw_t, raw_s
        ptr, tidx, doy32 = core._doy_tables(doy, 366, ts.device)
        ncg = (ngrid + 31) // 32
        mask = torch.empty((ncg, T), dtype=torch.int32, device="cuda"); nvalid = torch.zeros(ngrid, dtype=torch.int32, device="cuda")
        e3 = ev()
        check(lib.xmhw_exceed_mask_f32(ts.data_ptr(), T, ngrid, ptr.data_ptr(), tidx.data_ptr(), 366, out_t.data_ptr(), mask.data_ptr(), nvalid.data_ptr(), st), "exc")
        e4 = ev()
        counts = torch.empty(ngrid, dtype=torch.int32, device="cuda")
        check(lib.xmhw_events_count(mask.data_ptr(), T, ngrid, 5, 1, 2, counts.data_ptr(), st), "cnt")
        offsets = torch.empty(ngrid + 1, dtype=torch.int64, device="cuda"); scratch = torch.empty(ngrid // 1024 + 2, dtype=torch.int64, device="cuda")
        check(lib.xmhw_exclusive_scan_i32(counts.data_ptr(), ngrid, offsets.data_ptr(), scratch.data_ptr(), st), "scan")
        e5 = ev()
        nev = int(offsets[-1].item())
        ei = torch.empty((_cabi.EI_COUNT, nev), dtype=torch.int32, device="cuda"); ef = torch.empty((_cabi.EF_COUNT, nev), dtype=torch.float64, device="cuda")
        e6 = ev()
        check(lib.xmhw_events_fill(mask.data_ptr(), T, ngrid, 5, 1, 2, offsets.data_ptr(), nev, ei.data_ptr(), st), "fill")
        e7 = ev()
        check(lib.xmhw_event_stats_f32(ts.data_ptr(), T, ngrid, doy32.data_ptr(), out_t.data_ptr(), out_s.data_ptr(), nev, nev, ei.data_ptr(), ef.data_ptr(), st), "stats")
        e8 = ev()
        torch.cuda.synchronize()
        t = lambda a, b: a.elapsed_time(b)
        tot = t(e0, e2) + t(e3, e5) + t(e6, e8)
        balg = ngrid * T * 4 + nocean * 2 * 366 * 8 + nocean * 4 + nev * 180
        print(f"rep{r}: sweep {t(e0,e1):.2f} finish {t(e1,e2):.2f} exceed {t(e3,e4):.2f} count+scan {t(e4,e5):.2f} fill {t(e6,e7):.2f} stats {t(e7,e8):.2f} | total {tot:.2f} ms | events {nev} ({nev/max(nocean,1)/((y1-y0+1)):.2f}/cell-yr) | {nocean*(y1-y0+1)/tot*1e3:.3e} cell-yr/s | B_alg {balg/1e9:.2f} GB -> {balg/tot/1e6:.1f} GB/s = {balg/tot/1e6/6550.4*100:.2f}% of 6550")
        del mask, ei, ef, out_t, out_s
    del ts
    torch.cuda.empty_cache()

if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "c3q"]
    if "c2" in which: run(160, 240, 1982, 2021, land_frac=0)
    if "c3q" in which: run(180, 1440, 1982, 2011)
    if "c3" in which: run(720, 1440, 1982, 2011)
