"""Downstream statistics of the reference's `xmhw/stats.py` on the compact event table.

`block_average` (stats.py:27-183) bins the events of every cell into blocks of calendar years and
aggregates their properties (agg_mhw, stats.py:322-368), optionally with the series' mean / max / min
(agg_ts, :400-428) and the number of event days per category (agg_cats, :371-398) of every block;
`mhw_rank` (stats.py:446-510) ranks the events of a cell by every property.

What is followed.  Upstream marks this module unfinished, and two of its lines cannot be meant:
`intensity_mean_abs` / `intensity_cumulative_abs` aggregate the NON-abs columns (stats.py:358-359) and
`mhw_rank` hard-codes a record of 14245 days (stats.py:475).  Both follow Eric Oliver's
`marineHeatWaves.blockAverage` / `rank` here (the code xmhw ports): the `_abs` block means come from
the `_abs` event columns, `intensity_var*` block means are included, and the return period uses the
length of the analysed record, (nYears + 1) / rank.  Everything else is upstream's: bins
`range(period[0], period[1] + blockLength + 1, blockLength)` closed on the left (pd.cut right=False),
events assigned by the year of `mtime` (start or peak), NaN-skipping means, `ecount`, `total_icum`,
`intensity_max_max`, category days counted in the year of each day, rank = len - argsort(argsort(x)).
`removeMissing` / `split` are upstream stubs (no effect there) and are not offered.
"""
import numpy as np
import torch

from . import _cabi
from .core import EF_FIELDS, _call, _ptr, _require_cuda, _stream

BLOCK_MEAN_FIELDS = ("duration", "intensity_max", "intensity_mean", "intensity_var", "intensity_cumulative",
                     "intensity_max_relThresh", "intensity_mean_relThresh", "intensity_var_relThresh",
                     "intensity_cumulative_relThresh", "intensity_max_abs", "intensity_mean_abs", "intensity_var_abs",
                     "intensity_cumulative_abs", "severity_mean", "severity_cumulative", "rate_onset", "rate_decline")
BA_NCOL = 20
RANK_FIELDS = ("duration", "category", "duration_moderate", "duration_strong", "duration_severe",
               "duration_extreme") + tuple(EF_FIELDS)


def year_blocks(years, period=None, blockLength=1):
    """(block index of every time step [T] int32 with -1 outside, first year of every block)."""
    years = np.asarray(years, np.int64)
    if period is None:
        period = (int(years[0]), int(years[-1]))
    bins = np.arange(int(period[0]), int(period[1]) + blockLength + 1, blockLength)      # stats.py:131
    nblocks = len(bins) - 1
    b = (years - bins[0]) // blockLength
    b = np.where((years >= bins[0]) & (b < nblocks), b, -1).astype(np.int32)
    return b, bins[:-1].astype(np.int64)


def block_average_arrays(ev, years, period=None, blockLength=1, mtime="time_start", ts=None, doy=None,
                         thresh=None, seas=None):
    """Block statistics of an EventTable (core.detect_arrays).  `years` [T] = calendar year of every
    time step.  Returns (dict name -> CUDA tensor [nblocks, ngrid], block start years).  With `ts`
    (CUDA float32 [T, ngrid]) the series statistics are added; with `doy`, `thresh`, `seas` as well
    the event days per category and `total_days`."""
    if mtime not in ("time_start", "time_peak"):
        raise ValueError("mtime must be 'time_start' or 'time_peak'")
    dev = ev.i32.device
    block_of_t, starts = year_blocks(years, period, blockLength)
    nblocks, ngrid = len(starts), ev.ngrid
    bt = torch.from_numpy(block_of_t).to(dev)
    out = {}
    with torch.cuda.device(dev):
        st = _stream()
        buf = torch.empty((BA_NCOL, nblocks, ngrid), dtype=torch.float64, device=dev)
        _call("xmhw_block_average", _ptr(ev.i32), _ptr(ev.f64), ev.i32.shape[1], _ptr(ev.offsets), ngrid, _ptr(bt),
              nblocks, int(mtime == "time_peak"), _ptr(buf), st)
        out["ecount"] = buf[0]
        for k, f in enumerate(BLOCK_MEAN_FIELDS):
            out[f] = buf[1 + k]
        out["intensity_max_max"] = buf[18]
        out["total_icum"] = buf[19]
        if ts is not None:
            _require_cuda(ts, "ts", torch.float32)
            T = ts.shape[0]
            tb = torch.empty((3, nblocks, ngrid), dtype=torch.float64, device=dev)
            _call("xmhw_block_ts_f32", _ptr(ts), T, ngrid, _ptr(bt), nblocks, _ptr(tb), st)
            out["ts_mean"], out["ts_max"], out["ts_min"] = tb[0], tb[1], tb[2]
            if thresh is not None and seas is not None and doy is not None:
                d32 = torch.from_numpy(np.asarray(doy, np.int32)).to(dev)
                days = torch.empty((4, nblocks, ngrid), dtype=torch.int32, device=dev)
                _call("xmhw_block_cat_days_f32", _ptr(ts), T, ngrid, _ptr(d32), _ptr(thresh), _ptr(seas), _ptr(ev.i32),
                      ev.n, ev.i32.shape[1], _ptr(bt), nblocks, _ptr(days), st)
                for k, f in enumerate(("moderate_days", "strong_days", "severe_days", "extreme_days")):
                    out[f] = days[k]
                out["total_days"] = days.sum(0)                                  # stats.py:304-310
                torch.cuda.current_stream().synchronize()                        # d32 stays alive until here
    return out, starts


def mhw_rank_arrays(ev, nyears, fields=RANK_FIELDS):
    """(rank, return period) of every event within its cell for every property: dicts of CUDA float64 [n]."""
    dev = ev.i32.device
    rank, period = {}, {}
    n = ev.n
    with torch.cuda.device(dev):
        st = _stream()
        cells = ev.i32[0, :max(n, 1)].contiguous()
        for f in fields:
            col = ev.column(f).to(torch.float64).contiguous() if n else torch.empty(0, dtype=torch.float64, device=dev)
            r = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
            _call("xmhw_event_rank_f64", _ptr(col), _ptr(cells), _ptr(ev.offsets), n, _ptr(r), st)
            rank[f] = r[:n]
            period[f] = (float(nyears) + 1.0) / r[:n]
        torch.cuda.current_stream().synchronize()
    return rank, period


def block_average(ev, time, period=None, blockLength=1, mtime="time_start", ts=None, doy=None, thresh=None, seas=None):
    """`block_average` for a detect result kept on the device: `ev` an EventTable, `time` the datetime64
    axis of the series.  Returns a labeled.Dataset on (years, cell) (host)."""
    from . import labeled
    from .identify import _ymd
    years = _ymd(np.asarray(time))[0]
    out, starts = block_average_arrays(ev, years, period, blockLength, mtime, ts, doy, thresh, seas)
    ds = labeled.Dataset(coords={"years": starts, "cell": np.arange(ev.ngrid)})
    for k, v in out.items():
        ds[k] = labeled.DataArray(v.cpu().numpy(), ("years", "cell"))
    return ds


def mhw_rank(ev, time):
    """`mhw_rank` for an EventTable: (rank, return_period) labeled.Datasets on the event rows."""
    from . import labeled
    t = np.asarray(time)
    nyears = float((t[-1] - t[0]) / np.timedelta64(1, "D") + 1) / 365.25
    rank, period = mhw_rank_arrays(ev, nyears)
    a, b = labeled.Dataset(coords={"row": np.arange(ev.n)}), labeled.Dataset(coords={"row": np.arange(ev.n)})
    for k in rank:
        a[k] = labeled.DataArray(rank[k].cpu().numpy(), ("row",))
        b[k] = labeled.DataArray(period[k].cpu().numpy(), ("row",))
    return a, b
