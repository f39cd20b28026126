"""Array-level API of the B200 hot path (device tensors in, device tensors out).

These two functions replace the reference's per-cell Python loops

    xmhw/xmhw.py:184-197   for c in ts.cell: calc_clim(...)     -> threshold_arrays
    xmhw/xmhw.py:440-454   for c in ts.cell: define_events(...) -> detect_arrays

with a handful of batched kernel launches over the whole (time, cell) array.
PyTorch is plumbing only (device memory, streams); all arithmetic happens in
the CUDA library reached through the C ABI (xmhw_b200/_cabi.py).  There is no
CPU path: tensors must live on a CUDA device.
"""
import hashlib
from dataclasses import dataclass

import numpy as np
import torch

from . import _cabi
from . import plan as _plan
from . import plan2 as _plan2
from ._cabi import EF_COUNT, EF_FIELDS, EI_COUNT, EI_FIELDS, check, lib

_plan_cache = {}

# Optional per-kernel timing (bench.py): when TRACE is a list, every library call is
# bracketed by CUDA events on the launch stream and (name, start, end) is appended.
TRACE = None
LAUNCHES = {"n": 0}
_KERNELS_PER_CALL = {"xmhw_exclusive_scan_i32": 3, "xmhw_intermediate_f32": 2, "xmhw_clim_finish2_f64": 2}


def _call(name, *args):
    fn = getattr(lib, name)
    LAUNCHES["n"] += _KERNELS_PER_CALL.get(name, 1)
    if TRACE is None:
        check(fn(*args), name)
        return
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    check(fn(*args), name)
    e1.record()
    TRACE.append((name, e0, e1))


def _ptr(t):
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t, name, dtype):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError("%s must be a CUDA tensor (xmhw_b200 has no CPU path)" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)


@dataclass
class DevicePlan:
    host: _plan.ClimPlanHost
    tensors: dict
    struct: _cabi.ClimPlanStruct


def device_plan(doy, ndoy, w, q, device):
    """Build (or fetch from cache) the climatology sweep plan on `device`."""
    doy = np.ascontiguousarray(doy, dtype=np.int64)
    key = (hashlib.sha1(doy.tobytes()).hexdigest(), int(ndoy), int(w), float(q), str(device), _plan.default_keep(), _plan.default_pool_rows())
    hit = _plan_cache.get(key)
    if hit is not None:
        return hit
    host = _plan.build_clim_plan(doy, ndoy, w, q)
    if host.smem_bytes() > 227 * 1024:
        raise ValueError("climatology window needs %d KB of shared memory per warp (> 227 KB)"
                         % (host.smem_bytes() // 1024))
    tensors = {n: torch.from_numpy(np.ascontiguousarray(getattr(host, n))).to(device)
               for n in _cabi.PLAN_ARRAYS}
    struct = _cabi.plan_struct(host, {n: _ptr(t) for n, t in tensors.items()})
    dp = DevicePlan(host, tensors, struct)
    if len(_plan_cache) > 16:
        _plan_cache.clear()
    _plan_cache[key] = dp
    return dp


@dataclass
class DevicePlan2:
    host: _plan2.ClimPlan2Host
    struct: _cabi.ClimPlan2Struct
    exc_rows: object            # device int32: window rows of the exceptional doys (CSR by host.exc_off)


# Which sweep "auto" picks.  Measured on B200 (30-year daily series, profiles/kernel_ms_r02o_sweep_occupancy.txt,
# profiles/kernel_ms_r02r_sweep2_tmem.txt): the two-stack top-K sweep is bound by the warps its unit slots leave
# room for.  With 8 warps per SM it beats the general sorted-list sweep at every window measured -- in shared
# memory alone for narrow windows (windowHalfWidth <= 2: <= 25 KB per warp), with the slots that do not fit
# shared memory in TENSOR MEMORY otherwise (default window: 55 KB per warp; 39 ms against 52 ms on the global
# grid).  "auto" therefore takes the top-K sweep whenever its plan exists and 8 warps fit one way or the
# other, and the general sweep for the rest (large top-K capacity, very wide windows, > 48-year lists).
TOPK_AUTO_MIN_WARPS = 8
TM_WARPS, TM_COLS_PER_WARP = 8, 256          # csrc/xmhw_kernels.cu: the tensor-memory kernel's block and column budget


def _topk_warps_per_sm(host_plan):
    return (227 * 1024) // (host_plan.pool_rows * 128 + 256)


def _topk_tmem_fits(host_plan):
    """The launcher's test (xmhw_clim_sweep2_f32): the slots beyond TM_COLS_PER_WARP tensor-memory columns per warp
    must fit 8 warps' share of the shared memory."""
    max_smem_slots = (227 * 1024 - 1024) // (TM_WARPS * host_plan.slot_rows * 128)
    return host_plan.nslots - TM_COLS_PER_WARP // host_plan.slot_rows <= max_smem_slots


def sweep2_kernel_name(host_plan):
    """Which kernel xmhw_clim_sweep2_f32 launches for this plan (the launcher's rule, for reports)."""
    import os
    tm = int(os.environ.get("XMHW_B200_SWEEP2_TMEM", "-1"))
    use_tm = (tm == 1 or (tm < 0 and _topk_warps_per_sm(host_plan) < TM_WARPS)) and _topk_tmem_fits(host_plan)
    return "clim_sweep2_tm_kernel" if use_tm else "clim_sweep2_kernel"


def _group_order(ts):
    """Processing order of the 32-cell groups for the top-K sweep (xmhw_group_order_f32): groups that look like
    land (all 32 cells NaN in three probe rows) last, so a block's warps carry equal work.  Any permutation is
    correct; this one only costs three row reads."""
    T, ngrid = ts.shape
    ncg = (ngrid + 31) // 32
    flags = torch.empty(ncg, dtype=torch.uint8, device=ts.device)
    order = torch.empty(ncg + 1, dtype=torch.int32, device=ts.device)       # + the sweep's work ticket
    _call("xmhw_group_order_f32", _ptr(ts), T, ngrid, _ptr(flags), _ptr(order), _stream())
    return order


def sweep_mode():
    """XMHW_B200_SWEEP: "auto" (default, see TOPK_AUTO_MIN_WARPS), "topk" (the two-stack top-K sweep
    whenever the calendar fits) or "general" (always the sorted-list sweep of plan.py)."""
    import os
    return os.environ.get("XMHW_B200_SWEEP", "auto")


def device_plan2(doy, ndoy, w, q, device):
    """Plan of the two-stack top-K sweep on `device`, or None when the calendar / quantile needs
    the general sweep (plan2.build_clim_plan2 returns None)."""
    doy = np.ascontiguousarray(doy, dtype=np.int64)
    key = ("topk", hashlib.sha1(doy.tobytes()).hexdigest(), int(ndoy), int(w), float(q), str(device))
    if key in _plan_cache:
        return _plan_cache[key]
    host = _plan2.build_clim_plan2(doy, ndoy, w, q)
    dp = None
    if host is not None:
        dp = DevicePlan2(host, _cabi.plan2_struct(host),
                         torch.from_numpy(np.ascontiguousarray(host.exc_rows)).to(device))
    if len(_plan_cache) > 16:
        _plan_cache.clear()
    _plan_cache[key] = dp
    return dp


_csr_cache = {}


def _doy_tables(doy, ndoy, device):
    doy = np.ascontiguousarray(doy, dtype=np.int64)
    key = (hashlib.sha1(doy.tobytes()).hexdigest(), int(ndoy), str(device))
    hit = _csr_cache.get(key)
    if hit is None:
        ptr, tidx = _plan.doy_csr(doy, ndoy)
        hit = (torch.from_numpy(ptr).to(device), torch.from_numpy(tidx).to(device),
               torch.from_numpy(doy.astype(np.int32)).to(device))
        if len(_csr_cache) > 16:
            _csr_cache.clear()
        _csr_cache[key] = hit
    return hit


def _relabel_keeps_feb(labels):
    """feb29 uses positions 59,60,61: valid after compaction only if labels 1..61 are all present."""
    return len(labels) >= 61 and bool((labels[:61] == np.arange(1, 62)).all())


def threshold_arrays(ts, doy, ndoy, pctile=90, windowHalfWidth=5, smoothPercentile=True,
                     smoothPercentileWidth=31, feb29=True, return_raw=False, return_nempty=False):
    """Climatological threshold and seasonal mean for every cell of ts.

    ts   CUDA float32 [T, ngrid] (time-major, the reference's (time, cell) stack)
    doy  host int array [T], 1-based day-of-year labels (identify.py:28-79)
    Returns (thresh, seas): CUDA float64 [ndoy, ngrid]; NaN where a (cell, doy)
    has no sample (land).  Semantics: identify.py:184-270 + :137-181, xmhw.py:250-307.
    return_nempty appends the int32 [ngrid] number of doys without any sample per cell (== ndoy: land).
    """
    _require_cuda(ts, "ts", torch.float32)
    if ts.dim() != 2:
        raise ValueError("ts must be [T, ngrid]")
    T, ngrid = ts.shape
    if len(doy) != T:
        raise ValueError("doy must have one label per time step")
    if smoothPercentile and smoothPercentileWidth % 2 == 0:
        raise ValueError("smoothPercentileWidth should be odd")
    # A doy label that never occurs (doy 60 in a series without a leap year) does not exist in
    # the reference's groupby output, so feb29/runavg act on the compacted doy axis
    # (identify.py:233-241, :175-180): relabel to 1..n, run, scatter back with NaN rows.
    labels = np.unique(np.asarray(doy))
    if len(labels) < ndoy and not return_raw:
        sub = threshold_arrays(ts, np.searchsorted(labels, doy) + 1, len(labels), pctile, windowHalfWidth,
                               smoothPercentile, smoothPercentileWidth,
                               feb29=bool(feb29) and ndoy >= 61 and bool(np.isin([59, 60, 61], labels).all())
                               and _relabel_keeps_feb(labels), return_nempty=return_nempty)
        rows = torch.from_numpy(labels - 1).to(ts.device)
        outs = []
        for a in sub[:2]:
            full = torch.full((ndoy, ngrid), float("nan"), dtype=torch.float64, device=ts.device)
            full[rows] = a
            outs.append(full)
        if return_nempty:
            outs.append(sub[2] + (ndoy - len(labels)))
        return tuple(outs)
    with torch.cuda.device(ts.device):
        st = _stream()
        q = pctile / 100.0
        ncg = (ngrid + 31) // 32
        nempty = torch.empty(ngrid, dtype=torch.int32, device=ts.device)
        mode = sweep_mode()
        dp2 = device_plan2(doy, ndoy, windowHalfWidth, q, ts.device) if mode in ("topk", "auto") else None
        if dp2 is not None and mode == "auto" and _topk_warps_per_sm(dp2.host) < TOPK_AUTO_MIN_WARPS \
                and not _topk_tmem_fits(dp2.host):
            dp2 = None
        if dp2 is not None:
            # two-stack top-K sweep; rows of the doys it does not cover (absent labels) stay NaN
            h = dp2.host
            full = h.nsteps + len(h.exc_doy) == ndoy
            alloc = torch.empty if full else (lambda *a, **k: torch.full(*a, float("nan"), **k))
            raw_t = alloc((ndoy, ngrid), dtype=torch.float64, device=ts.device)
            raw_s = alloc((ndoy, ngrid), dtype=torch.float64, device=ts.device)
            import os
            order = _group_order(ts) if os.environ.get("XMHW_B200_SWEEP2_ORDER", "1") != "0" else None
            _call("xmhw_clim_sweep2_f32", _ptr(ts), T, ngrid, dp2.struct, _ptr(raw_t), _ptr(raw_s), _ptr(nempty),
                  _ptr(order) if order is not None else None, st)
            for k, d in enumerate(h.exc_doy):
                a, b = int(h.exc_off[k]), int(h.exc_off[k + 1])
                _call("xmhw_clim_direct_f32", _ptr(ts), T, ngrid, _ptr(dp2.exc_rows) + 4 * a, b - a, h.kp, float(q),
                      _ptr(raw_t) + (int(d) - 1) * ngrid * 8, _ptr(raw_s) + (int(d) - 1) * ngrid * 8, _ptr(nempty), st)
            if not full:
                nempty += ndoy - h.nsteps - len(h.exc_doy)
        else:
            dp = device_plan(doy, ndoy, windowHalfWidth, q, ts.device)
            raw_t = torch.empty((ndoy, ngrid), dtype=torch.float64, device=ts.device)
            raw_s = torch.empty((ndoy, ngrid), dtype=torch.float64, device=ts.device)
            scratch = torch.empty(max(1, ncg * dp.host.scratch_rows * 32), dtype=torch.int32, device=ts.device)
            _call("xmhw_clim_sweep_f32", _ptr(ts), T, ngrid, dp.struct, _ptr(raw_t), _ptr(raw_s), _ptr(nempty),
                  _ptr(scratch), st)
        W = int(smoothPercentileWidth) if smoothPercentile else 1
        do_feb = bool(feb29) and ndoy >= 61
        if W <= 1 and not do_feb:
            if return_nempty:
                return raw_t, raw_s, nempty
            return (raw_t, raw_s, raw_t, raw_s) if return_raw else (raw_t, raw_s)
        out_t = torch.empty_like(raw_t)
        out_s = torch.empty_like(raw_s)
        _call("xmhw_clim_finish2_f64", _ptr(raw_t), _ptr(out_t), _ptr(raw_s), _ptr(out_s), ndoy, ngrid,
              int(do_feb), W, _ptr(nempty), st)
    if return_nempty:
        return out_t, out_s, nempty
    return (out_t, out_s, raw_t, raw_s) if return_raw else (out_t, out_s)


def interp_gaps_(ts, max_pad):
    """In-place pre-step `maxPadLength` (xmhw.py:159-160, :409-410): NaN runs of at most
    `max_pad` steps between two valid samples are filled by linear interpolation along time."""
    _require_cuda(ts, "ts", torch.float32)
    T, ngrid = ts.shape
    with torch.cuda.device(ts.device):
        _call("xmhw_interp_gaps_f32", _ptr(ts), T, ngrid, int(max_pad), _stream())
    return ts


# staging capacity of the one-pass event finder (events per cell and year; overflow falls back
# to the exact two-pass path)
STAGE_EVENTS_PER_YEAR = 4


class EventTable:
    """Compact event table on the device (struct of arrays).

    i32 [EI_COUNT, n] int32 and f64 [EF_COUNT, n] float64 hold the columns named in
    `_cabi.EI_FIELDS` / `_cabi.EF_FIELDS` (the reference's per-event variables,
    features.py:114-152, :181-189, :290-291), ordered by cell then start index.
    `nvalid` [ngrid] int32 is the number of non-NaN samples per cell (land_check).
    """

    def __init__(self, i32, f64, n, offsets, nvalid, T, ngrid):
        self.i32, self.f64, self.n = i32, f64, n
        self.offsets, self.nvalid, self.T, self.ngrid = offsets, nvalid, T, ngrid

    def __len__(self):
        return self.n

    def column(self, name):
        if name in EI_FIELDS:
            return self.i32[EI_FIELDS.index(name), :self.n]
        return self.f64[EF_FIELDS.index(name), :self.n]

    def to_numpy(self):
        i32 = self.i32[:, :self.n].cpu().numpy()
        f64 = self.f64[:, :self.n].cpu().numpy()
        out = {f: i32[k].astype(np.int64) for k, f in enumerate(EI_FIELDS)}
        out.update({f: f64[k] for k, f in enumerate(EF_FIELDS)})
        return out


def detect_arrays(ts, doy, ndoy, thresh, seas, minDuration=5, joinGaps=True, maxGap=2):
    """Marine-heatwave events of every cell of ts given the climatologies.

    ts CUDA float32 [T, ngrid]; thresh, seas CUDA float64 [ndoy, ngrid]; doy host
    int [T].  Semantics: identify.py:328-479 (define_events, mhw_filter, join_gaps)
    and features.py:22-295.  Returns an EventTable.
    """
    _require_cuda(ts, "ts", torch.float32)
    _require_cuda(thresh, "thresh", torch.float64)
    _require_cuda(seas, "seas", torch.float64)
    T, ngrid = ts.shape
    if thresh.shape != (ndoy, ngrid) or seas.shape != (ndoy, ngrid):
        raise ValueError("thresh/seas must be [ndoy, ngrid]")
    if len(doy) != T:
        raise ValueError("doy must have one label per time step")
    if maxGap >= minDuration:
        raise ValueError("Maximum gap between mhw events should be smaller than event minimum duration")
    dev = ts.device
    if detect_mode() == "fused":
        ev = _detect_fused(ts, doy, ndoy, thresh, seas, minDuration, joinGaps, maxGap)
        if ev is not None:
            return ev
    with torch.cuda.device(dev):
        st = _stream()
        ptr, tidx, doy32 = _doy_tables(doy, ndoy, dev)
        ncg = (ngrid + 31) // 32
        mask = torch.empty((ncg, T), dtype=torch.int32, device=dev)
        nvalid = torch.zeros(ngrid, dtype=torch.int32, device=dev)
        _call("xmhw_exceed_mask_f32", _ptr(ts), T, ngrid, _ptr(ptr), _ptr(tidx), ndoy, _ptr(thresh),
                                       _ptr(mask), _ptr(nvalid), st)
        counts = torch.empty(ngrid, dtype=torch.int32, device=dev)
        # the count pass parks each cell's (start, end) pairs in a staging table sized for
        # STAGE_EVENTS_PER_YEAR events per year, so the mask is scanned once; a cell with more
        # events sets the overflow flag and the exact second pass over the mask runs instead
        years = max(1, -(-T // max(ndoy, 1)))
        stage_cap = int(min(max(8, STAGE_EVENTS_PER_YEAR * years), max(8, T // max(1, minDuration + 1) + 1)))
        stage = torch.empty((ncg, 2, stage_cap, 32), dtype=torch.int32, device=dev)
        offsets = torch.empty(ngrid + 2, dtype=torch.int64, device=dev)     # [ngrid + 1] = overflow flag
        overflow = offsets[ngrid + 1:].view(torch.int32)
        overflow.zero_()
        _call("xmhw_events_count_stage", _ptr(mask), T, ngrid, int(minDuration), int(bool(joinGaps)), int(maxGap),
                                          _ptr(counts), _ptr(stage), stage_cap, _ptr(overflow), st)
        scratch = torch.empty(ngrid // 1024 + 2, dtype=torch.int64, device=dev)
        _call("xmhw_exclusive_scan_i32", _ptr(counts), ngrid, _ptr(offsets), _ptr(scratch), st)
        # the one host sync: event total (sizes the table) + overflow flag.  The copy is asynchronous and the
        # cell-major {thresh, seas} pairs (an event's consecutive days become one contiguous run; independent of
        # the events) are enqueued BEHIND it, so the device keeps working while the host waits and launches
        tail_h = torch.empty(2, dtype=torch.int64, pin_memory=True)
        tail_h.copy_(offsets[ngrid:], non_blocking=True)
        got_tail = torch.cuda.Event()
        got_tail.record()
        clim_cm = torch.empty((ngrid, ndoy, 2), dtype=torch.float64, device=dev)
        _call("xmhw_clim_cellmajor_f64", _ptr(thresh), _ptr(seas), ndoy, ngrid, _ptr(clim_cm), st)
        got_tail.synchronize()
        nev, overflowed = int(tail_h[0]), bool(int(tail_h[1]) & 0xffffffff)
        offsets = offsets[:ngrid + 1]
        cap = max(nev, 1)
        ev_i32 = torch.empty((EI_COUNT, cap), dtype=torch.int32, device=dev)
        ev_f64 = torch.empty((EF_COUNT, cap), dtype=torch.float64, device=dev)
        if nev:
            if overflowed:
                _call("xmhw_events_fill", _ptr(mask), T, ngrid, int(minDuration), int(bool(joinGaps)), int(maxGap),
                                           _ptr(offsets), cap, _ptr(ev_i32), st)
            else:
                _call("xmhw_events_gather", _ptr(stage), stage_cap, _ptr(counts), _ptr(offsets), ngrid, cap,
                                             _ptr(ev_i32), st)
            del stage
            _call("xmhw_event_stats_cm_f32", _ptr(ts), T, ngrid, _ptr(doy32), ndoy, _ptr(clim_cm), nev, cap,
                                              _ptr(ev_i32), _ptr(ev_f64), st)
    return EventTable(ev_i32, ev_f64, nev, offsets, nvalid, T, ngrid)


def detect_mode():
    """XMHW_B200_DETECT: "chain" (default: exceedance mask -> staged run finding -> gather -> per-event
    statistics) or "fused" (one time-major pass, xmhw_detect_fused_f32).  Measured on B200 at the global
    0.25 deg grid: chain 27.9 ms, fused 80.8 + 13.9 ms (profiles/ncu_r02f_detect_fused_quarter.txt): with the
    thresholds of 32 cells in shared memory only two blocks fit an SM and the per-cell run rules of a
    block are one warp's serial work, so the other warps wait at the block barrier (57 % of the stall
    samples); the chain's kernels already issue ~2 instructions per cycle, i.e. the work is as much
    issue-bound as memory-bound and a fused pass cannot drop below their sum by much."""
    import os
    return os.environ.get("XMHW_B200_DETECT", "chain")


# staging capacity of the fused pass in events per cell and year (real SST: ~2-3); an overflow falls
# back to the kernel chain, which sizes its table exactly
FUSED_EVENTS_PER_YEAR = 3.5


def _detect_fused(ts, doy, ndoy, thresh, seas, minDuration, joinGaps, maxGap):
    """detect_arrays by the fused time-major pass; None when its staging table overflowed."""
    T, ngrid = ts.shape
    dev = ts.device
    with torch.cuda.device(dev):
        st = _stream()
        _, _, doy32 = _doy_tables(doy, ndoy, dev)
        clim_cm = torch.empty((ngrid, ndoy, 2), dtype=torch.float64, device=dev)
        _call("xmhw_clim_cellmajor_f64", _ptr(thresh), _ptr(seas), ndoy, ngrid, _ptr(clim_cm), st)
        years = max(1.0, T / max(ndoy, 1))
        scap = int(ngrid * years * FUSED_EVENTS_PER_YEAR) + 4096
        stage_i = torch.empty((EI_COUNT + 1, scap), dtype=torch.int32, device=dev)
        stage_f = torch.empty((EF_COUNT, scap), dtype=torch.float64, device=dev)
        counts = torch.zeros(ngrid, dtype=torch.int32, device=dev)
        nvalid = torch.zeros(ngrid, dtype=torch.int32, device=dev)
        offsets = torch.empty(ngrid + 2, dtype=torch.int64, device=dev)     # [ngrid + 1] = staged count | overflow flag
        counter = offsets[ngrid + 1:].view(torch.int32)
        _call("xmhw_detect_fused_f32", _ptr(ts), T, ngrid, _ptr(doy32), ndoy, _ptr(thresh), _ptr(clim_cm),
              int(minDuration), int(bool(joinGaps)), int(maxGap), _ptr(counts), _ptr(nvalid), _ptr(stage_i),
              _ptr(stage_f), scap, _ptr(counter), st)
        del clim_cm
        scratch = torch.empty(ngrid // 1024 + 2, dtype=torch.int64, device=dev)
        _call("xmhw_exclusive_scan_i32", _ptr(counts), ngrid, _ptr(offsets), _ptr(scratch), st)
        tail = offsets[ngrid:].cpu()           # the one host sync: event total + (staged count, overflow flag)
        nev = int(tail[0])
        staged, overflowed = int(tail[1]) & 0xffffffff, (int(tail[1]) >> 32) != 0
        if overflowed or staged != nev:
            return None
        offsets = offsets[:ngrid + 1]
        cap = max(nev, 1)
        ev_i32 = torch.empty((EI_COUNT, cap), dtype=torch.int32, device=dev)
        ev_f64 = torch.empty((EF_COUNT, cap), dtype=torch.float64, device=dev)
        if nev:
            _call("xmhw_events_scatter", _ptr(stage_i), _ptr(stage_f), scap, nev, _ptr(offsets), cap, _ptr(ev_i32),
                  _ptr(ev_f64), st)
    return EventTable(ev_i32, ev_f64, nev, offsets, nvalid, T, ngrid)


def synth_sst_device(T, ngrid, season, land=None, cell0=0, seed=None, nan_ppm=0, device="cuda", out=None,
                     coherent=1):
    """Device twin of synth.synth_sst (bit-identical): float32 [T, ngrid] on `device`."""
    from . import synth
    dev = torch.device(device)
    with torch.cuda.device(dev):
        ts = out if out is not None else torch.empty((T, ngrid), dtype=torch.float32, device=dev)
        sea = torch.from_numpy(np.ascontiguousarray(season, np.float64)).to(dev)
        if sea.numel() < T + 366:
            raise ValueError("season table must have T + 366 entries")
        ld = None if land is None else torch.from_numpy(np.ascontiguousarray(land, np.uint8).ravel()).to(dev)
        _call("xmhw_synth_sst_f32", _ptr(ts), T, ngrid, int(cell0), 0 if ld is None else _ptr(ld), _ptr(sea),
                                     synth.SEED if seed is None else int(seed), synth.RHO, synth.SIGMA,
                                     synth.NOISE_SCALE, int(nan_ppm), int(coherent), _stream())
        torch.cuda.current_stream().synchronize()   # keep `sea`/`ld` alive until the kernel is done
    return ts


def intermediate_arrays(ts, doy, ndoy, thresh, seas, events):
    """Dense per-timestep fields of the reference's `intermediate=True` dataset
    (identify.py:404-411; mhw_df, features.py:22-69) for an EventTable from detect_arrays.
    Returns {name: CUDA tensor [T, ngrid]} (float64, `mabs` float32, flags/bthresh bool)."""
    _require_cuda(ts, "ts", torch.float32)
    T, ngrid = ts.shape
    dev = ts.device
    dt = {"f8": torch.float64, "f4": torch.float32, "u1": torch.uint8}
    with torch.cuda.device(dev):
        out = {n: torch.empty((T, ngrid), dtype=dt[k], device=dev) for n, k in _cabi.INTERMEDIATE_FIELDS}
        out["events"].fill_(float("nan"))
        st = _cabi.IntermediateStruct(**{n: _ptr(out[n]) for n, _ in _cabi.INTERMEDIATE_FIELDS})
        _, _, doy32 = _doy_tables(doy, ndoy, dev)
        _call("xmhw_intermediate_f32", _ptr(ts), T, ngrid, _ptr(doy32), _ptr(thresh), _ptr(seas),
              _ptr(events.i32), events.n, events.i32.shape[1], st, _stream())
    for n, k in _cabi.INTERMEDIATE_FIELDS:
        if k == "u1":
            out[n] = out[n].bool()
    return out


def threshold_detect_host(ts_host, doy, ndoy, pctile=90, windowHalfWidth=5, smoothPercentile=True,
                          smoothPercentileWidth=31, feb29=True, minDuration=5, joinGaps=True, maxGap=2,
                          device="cuda", out=None, slabs=24):
    """Host-buffer entry point for threshold + detect in one pass over the host series (what a
    reference-side binding calls): see host_pipeline."""
    return host_pipeline(ts_host, doy, ndoy, do_threshold=True, do_detect=True, pctile=pctile,
                         windowHalfWidth=windowHalfWidth, smoothPercentile=smoothPercentile,
                         smoothPercentileWidth=smoothPercentileWidth, feb29=feb29, minDuration=minDuration,
                         joinGaps=joinGaps, maxGap=maxGap, device=device, out=out, slabs=slabs)


def _pin(t):
    """Page-lock a host tensor in place for the duration of a call (cudaHostRegister) so that the
    strided column-block copies run asynchronously at full link speed; returns an unregister callable."""
    if t.is_pinned() or t.numel() * t.element_size() < (64 << 20):
        return lambda: None
    rt = torch.cuda.cudart()
    try:
        rc = rt.cudaHostRegister(t.data_ptr(), t.numel() * t.element_size(), 0)
        if int(rc) != 0:
            return lambda: None
    except Exception:
        return lambda: None
    return lambda: rt.cudaHostUnregister(t.data_ptr())


def host_pipeline(ts_host, doy, ndoy, do_threshold=True, do_detect=True, th_host=None, se_host=None,
                  pctile=90, windowHalfWidth=5, smoothPercentile=True, smoothPercentileWidth=31, feb29=True,
                  minDuration=5, joinGaps=True, maxGap=2, negate=False, max_pad=0, anynans=False,
                  device="cuda", out=None, slabs=24, th_dev=None, se_dev=None, clim_on_device=False):
    """The hot path on a HOST series: `ts_host` float32 [T, ngrid] (pinned, or page-locked here for the
    call).  The grid is cut into `slabs` column blocks; the strided host->device copy of block i+1,
    the kernels of block i and the device->host copy of the results of block i-1 overlap on three
    streams (cells are independent, so blocks are).  The copy in dominates (PCIe); what the pipeline
    adds is the tail after the last block's copy, so more, smaller blocks are better: 920 / 877 /
    866 ms with 8 / 16 / 24 blocks at config 3.

    do_threshold   climatologies from the series (else `th_host` / `se_host` float64 [ndoy, ngrid] are uploaded per
                   block, or `th_dev` / `se_dev`, the same arrays already on the device, are sliced per block)
    clim_on_device the climatologies stay on the device: the result holds `thresh_dev` / `seas_dev` [ndoy, ngrid]
                   instead of the host copies (the public API drops land rows / columns there and copies once)
    do_detect      event table (needs climatologies from either source)
    negate, max_pad, anynans   the public functions' pre-steps on the device block: coldSpells sign flip
                   (xmhw.py:153-154, :412-413), maxPadLength gap interpolation (:159-160, :409-410), and
                   cells with any NaN masked out (identify.py:522-525)
    Returns dict(thresh, seas [ndoy, ngrid] host (do_threshold), nvalid [ngrid] (samples per cell AFTER the
    pre-steps), nempty [ngrid] (doys without samples, do_threshold), ev_i32, ev_f64, n_events, byte counts);
    `out` may hold preallocated pinned result tensors (thresh, seas, nvalid, ev_i32, ev_f64)."""
    dev = torch.device(device)
    if isinstance(ts_host, np.ndarray):
        ts_host = torch.from_numpy(ts_host)
    if ts_host.dtype != torch.float32 or ts_host.dim() != 2 or not ts_host.is_contiguous():
        raise TypeError("ts_host must be a contiguous float32 [T, ngrid] host tensor")
    clim_dev = th_dev is not None and se_dev is not None
    if not do_threshold and do_detect and not clim_dev and (th_host is None or se_host is None):
        raise ValueError("detect without threshold needs th_host and se_host (or th_dev and se_dev)")
    T, ngrid = ts_host.shape
    out = {} if out is None else out
    unpin = [_pin(ts_host)]
    pin = ts_host.is_pinned() or True

    def host_buf(name, shape, dtype, exact=True):
        t = out.get(name)
        if t is not None and t.dtype == dtype and (tuple(t.shape) == tuple(shape) if exact else
                                                   (tuple(t.shape[:-1]) == tuple(shape[:-1]) and t.shape[-1] >= shape[-1])):
            return t
        return torch.empty(shape, dtype=dtype, pin_memory=True)

    to_host = do_threshold and not clim_on_device
    th_h = host_buf("thresh", (ndoy, ngrid), torch.float64) if to_host else None
    se_h = host_buf("seas", (ndoy, ngrid), torch.float64) if to_host else None
    th_d = se_d = None
    nv_h = host_buf("nvalid", (ngrid,), torch.int32)
    ne_h = host_buf("nempty", (ngrid,), torch.int32) if do_threshold else None
    if clim_dev:
        if tuple(th_dev.shape) != (ndoy, ngrid) or tuple(se_dev.shape) != (ndoy, ngrid) or th_dev.dtype != torch.float64:
            raise ValueError("th_dev / se_dev must be float64 [ndoy, ngrid]")
    elif not do_threshold and do_detect:
        th_src = th_host if isinstance(th_host, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(th_host, np.float64))
        se_src = se_host if isinstance(se_host, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(se_host, np.float64))
        if th_src.shape != (ndoy, ngrid) or se_src.shape != (ndoy, ngrid):
            raise ValueError("th_host / se_host must be [ndoy, ngrid]")
        unpin += [_pin(th_src), _pin(se_src)]
    w = -(-ngrid // max(1, int(slabs)))
    w = -(-w // 32) * 32
    ranges = [(a, min(ngrid, a + w)) for a in range(0, ngrid, w)]
    ev_parts, ev_parts_all = [], []
    nev = 0
    ei_h = ef_h = None
    try:
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream()
            s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
            bufs = [torch.empty((T, w), dtype=torch.float32, device=dev) for _ in range(2)]
            cbufs = None
            if do_threshold and clim_on_device:
                th_d = torch.empty((ndoy, ngrid), dtype=torch.float64, device=dev)
                se_d = torch.empty((ndoy, ngrid), dtype=torch.float64, device=dev)
            if not do_threshold and do_detect and not clim_dev:
                cbufs = [[torch.empty((ndoy, w), dtype=torch.float64, device=dev) for _ in range(2)] for _ in range(2)]
            s_in.wait_stream(main)          # the buffers were allocated on `main`
            s_out.wait_stream(main)
            free_ev = [None, None]
            loaded = {}

            def block_view(buf, rows, n):
                return buf[:, :n] if n == w else buf.view(-1)[:rows * n].view(rows, n)

            def start_load(i):
                a, b = ranges[i]
                with torch.cuda.stream(s_in):
                    if free_ev[i % 2] is not None:
                        s_in.wait_event(free_ev[i % 2])
                    dst = block_view(bufs[i % 2], T, b - a)
                    check(lib.xmhw_copy2d_async(_ptr(dst), (b - a) * 4, ts_host.data_ptr() + a * 4, ngrid * 4,
                                                (b - a) * 4, T, 0, s_in.cuda_stream), "xmhw_copy2d_async")
                    clim = None
                    if cbufs is not None:
                        clim = []
                        for src, cb in ((th_src, cbufs[0][i % 2]), (se_src, cbufs[1][i % 2])):
                            d = block_view(cb, ndoy, b - a)
                            check(lib.xmhw_copy2d_async(_ptr(d), (b - a) * 8, src.data_ptr() + a * 8, ngrid * 8,
                                                        (b - a) * 8, ndoy, 0, s_in.cuda_stream), "xmhw_copy2d_async")
                            clim.append(d)
                    e = torch.cuda.Event()
                    e.record(s_in)
                    loaded[i] = (dst, clim, e)

            start_load(0)
            keep_alive = []
            # event tables stream to the host per block when the caller preallocated them (their
            # final size is only known at the end); otherwise they are copied after the last block
            stream_ev = do_detect and "ev_i32" in out and "ev_f64" in out
            pos = 0

            def copy_events(e, pos):
                if e.n:
                    for k in range(EI_COUNT):                   # row by row: contiguous DMAs
                        ei_h[k, pos:pos + e.n].copy_(e.i32[k, :e.n], non_blocking=True)
                    for k in range(EF_COUNT):
                        ef_h[k, pos:pos + e.n].copy_(e.f64[k, :e.n], non_blocking=True)

            if stream_ev:
                ei_h, ef_h = out["ev_i32"], out["ev_f64"]
            for i, (a, b) in enumerate(ranges):
                if i + 1 < len(ranges):
                    start_load(i + 1)
                ts, clim, e = loaded.pop(i)
                main.wait_event(e)
                # land_check comes first in the reference (xmhw.py:138, :399), on the data as given
                nvalid = None
                if anynans or not do_detect:
                    nvalid = count_valid(ts)
                if anynans:                                       # a cell with any NaN produces no output
                    ts.masked_fill_((nvalid != T).unsqueeze(0), float("nan"))
                    nvalid = torch.where(nvalid == T, nvalid, torch.zeros_like(nvalid))
                if negate:
                    ts.neg_()
                if max_pad:
                    interp_gaps_(ts, max_pad)
                th = se = nempty = None
                if do_threshold:
                    th, se, nempty = threshold_arrays(ts, doy, ndoy, pctile, windowHalfWidth, smoothPercentile,
                                                      smoothPercentileWidth, feb29, return_nempty=True)
                elif clim_dev:
                    th, se = th_dev[:, a:b].contiguous(), se_dev[:, a:b].contiguous()
                else:
                    th, se = clim if clim is not None else (None, None)
                if th_d is not None:
                    th_d[:, a:b].copy_(th)
                    se_d[:, a:b].copy_(se)
                ev = None
                if do_detect:
                    ev = detect_arrays(ts, doy, ndoy, th, se, minDuration, joinGaps, maxGap)
                    if ev.n:
                        ev.i32[0, :ev.n] += a                    # block-local -> global cell ids
                    if nvalid is None:
                        nvalid = ev.nvalid
                done = torch.cuda.Event()
                done.record(main)
                free_ev[i % 2] = done
                if stream_ev and pos + ev.n > ei_h.shape[1]:
                    stream_ev = False                            # preallocated table too small: defer
                    ev_parts = ev_parts_all[:]
                with torch.cuda.stream(s_out):
                    s_out.wait_event(done)
                    if to_host:
                        for src, dst in ((th, th_h), (se, se_h)):
                            check(lib.xmhw_copy2d_async(dst.data_ptr() + a * 8, ngrid * 8, _ptr(src), (b - a) * 8,
                                                        (b - a) * 8, ndoy, 1, s_out.cuda_stream), "xmhw_copy2d_async")
                    if do_threshold:
                        ne_h[a:b].copy_(nempty, non_blocking=True)
                    nv_h[a:b].copy_(nvalid, non_blocking=True)
                    if stream_ev:
                        copy_events(ev, pos)
                if do_detect:
                    ev_parts_all.append(ev)
                    if not stream_ev:
                        ev_parts.append(ev)
                    pos += ev.n
                keep_alive.append((th, se, nvalid, nempty))
            nev = pos
            if do_detect and not stream_ev:
                ei_h = host_buf("ev_i32", (EI_COUNT, nev), torch.int32, exact=False)
                ef_h = host_buf("ev_f64", (EF_COUNT, nev), torch.float64, exact=False)
                with torch.cuda.stream(s_out):
                    p2 = 0
                    for e in ev_parts:
                        copy_events(e, p2)
                        p2 += e.n
            s_out.synchronize()
            main.synchronize()
    finally:
        for u in unpin:
            u()
    res = {"nvalid": nv_h, "n_events": nev,
           "h2d_bytes": T * ngrid * 4 + (0 if do_threshold or not do_detect or clim_dev else 2 * ndoy * ngrid * 8),
           "d2h_bytes": (2 * ndoy * ngrid * 8 if to_host else 0) + (ngrid * 4 if do_threshold else 0) + ngrid * 4
           + nev * (EI_COUNT * 4 + EF_COUNT * 8)}
    if do_threshold:
        res.update(nempty=ne_h)
        if to_host:
            res.update(thresh=th_h, seas=se_h)
        else:
            res.update(thresh_dev=th_d, seas_dev=se_d)
    if do_detect:
        res.update(ev_i32=ei_h[:, :nev], ev_f64=ef_h[:, :nev])
    return res


def count_valid(ts):
    """Non-NaN samples per cell of a CUDA float32 [T, ngrid] block (land_check, identify.py:522-525)."""
    _require_cuda(ts, "ts", torch.float32)
    T, ngrid = ts.shape
    nvalid = torch.empty(ngrid, dtype=torch.int32, device=ts.device)
    with torch.cuda.device(ts.device):
        _call("xmhw_count_valid_f32", _ptr(ts), T, ngrid, _ptr(nvalid), _stream())
    return nvalid
