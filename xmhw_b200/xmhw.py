"""Drop-in public API: `threshold` and `detect` with the reference's signatures
(xmhw/xmhw.py:38-51 and :310-323), same keyword meaning, same XmhwException conditions.

    from xmhw_b200.xmhw import threshold, detect
    clim = threshold(sst)                                   # Dataset{thresh, seas} (doy, lat, lon)
    mhw  = detect(sst, clim['thresh'], clim['seas'])        # Dataset (events, lat, lon)

Inputs may be `xarray.DataArray`s or `xmhw_b200.labeled.DataArray`s (xarray is not part
of this image); outputs are xarray objects when xarray is importable and labeled ones
otherwise.  The per-cell loops of the reference (xmhw.py:184-197, :440-454) are replaced
by batched CUDA kernels (xmhw_b200.core); everything here is validation, layout
(stack sorted non-time dims into `cell`, identify.py:520) and result assembly.
"""
import warnings

import numpy as np

from . import labeled
from .exception import XmhwException
from .features import EVENT_VARIABLES, FLOAT32_VARIABLES, flip_cold
from .identify import add_doy, annotate_ds, get_calendar, land_check_shape

DENSE_LIMIT_BYTES = 8 << 30


def _coord(temp, name):
    c = temp.coords[name] if name in getattr(temp, "coords", {}) else None
    if c is None:
        return None
    return np.asarray(getattr(c, "values", c))


def _unpack(temp, tdim):
    """-> (data [T, ...sorted dims] float32, time values, sorted dims, their coords, attrs)"""
    dims = list(temp.dims)
    if tdim not in dims:
        raise XmhwException(f"{tdim} dimension not present, default"
                            + "is 'time' or pass as tdim='time_dimension_name'")
    data = np.asarray(temp.values)
    time = _coord(temp, tdim)
    if time is None:
        raise XmhwException(f"{tdim} coordinate values are required to build the day-of-year axis")
    other = sorted(d for d in dims if d != tdim)
    order = [dims.index(tdim)] + [dims.index(d) for d in other]
    data = np.transpose(data, order)
    if data.dtype != np.float32:
        if data.dtype == np.float64:
            warnings.warn("xmhw_b200 computes on float32 series; float64 input is rounded to float32")
        data = data.astype(np.float32)
    coords = {}
    for d, n in zip(other, data.shape[1:]):
        c = _coord(temp, d)
        coords[d] = np.arange(n) if c is None else c
    enc = getattr(getattr(temp, "coords", {}).get(tdim, None), "encoding", None) or getattr(temp, "encoding", {})
    tattrs = getattr(getattr(temp, "coords", {}).get(tdim, None), "attrs", None) or {}
    # attributes of every coordinate of the input (xmhw.py:130-131, :390-391): xarray coordinates carry
    # `.attrs`, the light containers an optional `coord_attrs` dict
    cattrs = {}
    for d in dims:
        a = getattr(getattr(temp, "coords", {}).get(d, None), "attrs", None)
        if not a:
            a = getattr(temp, "coord_attrs", {}).get(d)
        if a:
            cattrs[d] = dict(a)
    _unpack.last_coord_attrs = cattrs
    return np.ascontiguousarray(data), time, other, coords, dict(getattr(temp, "attrs", {})), enc, tattrs


def _pad_steps(maxPadLength, time):
    """Longest NaN run (in time steps) that `ts.interpolate_na(dim=tdim, max_gap=maxPadLength)` fills
    (xmhw.py:159-160, :409-410).  xarray measures a gap as the COORDINATE DISTANCE between the valid
    samples that bound it, so a run of k NaNs on a regular axis has length (k + 1) steps and is filled
    when (k + 1) * step <= max_gap.  A number counts time steps; a timedelta-like value (what current
    xarray requires on a datetime axis) is divided by the step of the axis."""
    if not maxPadLength:
        return 0
    n = maxPadLength
    if isinstance(n, np.timedelta64) or not isinstance(n, (int, float, np.integer, np.floating)):
        t = np.asarray(time)
        if not np.issubdtype(t.dtype, np.datetime64) or len(t) < 2:
            raise XmhwException("a timedelta-like maxPadLength needs a datetime time axis")
        step = np.median(np.diff(t)).astype("timedelta64[s]").astype(np.int64)
        try:
            secs = np.timedelta64(n).astype("timedelta64[s]").astype(np.int64)
        except (TypeError, ValueError):
            secs = int(n.total_seconds())
        n = secs / max(1, step)
    return max(0, int(np.floor(n)) - 1)


def _all_land(flat):
    """True when no sample at all is valid (identify.py:527-528), found without a pass over the array:
    the scan stops at the first time row that holds a valid value (row 0 for any real data set)."""
    for t0 in range(0, flat.shape[0], 64):
        if not np.isnan(flat[t0:t0 + 64]).all():
            return False
    return True


def _slabs(flat):
    """Column blocks of the host pipeline: ~1 GB of series per block, at least one, at most 32."""
    return int(min(32, max(1, flat.nbytes >> 30)))


def _to_host(t):
    """Device tensor -> numpy array through ONE pinned buffer (the array keeps the buffer alive)."""
    import torch
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return h.numpy()


def _take1(values, idx):
    """values[idx] for a 1-D coordinate / time vector and a torch integer index of tens of millions of
    entries: torch's index_select runs on all host cores (numpy's take on one)."""
    import torch
    v = np.ascontiguousarray(values)
    if v.dtype.kind in "mM":
        return torch.from_numpy(v.view(np.int64)).index_select(0, idx).numpy().view(v.dtype)
    if v.dtype.kind in "iuf" and v.dtype.itemsize in (4, 8):
        return torch.from_numpy(v).index_select(0, idx).numpy()
    return v[idx.numpy()]


def _wrap(ds, like):
    if labeled.is_xarray(like):
        try:
            return labeled.to_xarray(ds)
        except ImportError:
            pass
    return ds


def _present(mask_nd):
    """Index vectors of the coordinate values that survive `unstack` (xmhw.py:213-214):
    rows/columns without any ocean cell vanish."""
    keep = []
    for ax in range(mask_nd.ndim):
        other = tuple(a for a in range(mask_nd.ndim) if a != ax)
        keep.append(np.nonzero(mask_nd.any(axis=other))[0] if other else np.nonzero(mask_nd)[0])
    return keep


def threshold(temp, tdim="time", climatologyPeriod=[None, None], pctile=90, windowHalfWidth=5,
              smoothPercentile=True, smoothPercentileWidth=31, maxPadLength=None, coldSpells=False,
              tstep=False, anynans=False, skipna=False):
    """Calculate threshold and mean climatology (day-of-year) -- xmhw/xmhw.py:38-247.

    Same parameters as the reference.  `skipna` has no effect on results (the reference drops
    NaNs at identify.py:208 in either mode) and none on speed here.  Returns a Dataset with
    `thresh` and `seas` on (doy, <sorted non-time dims>), float64.
    """
    import torch
    from . import core

    if smoothPercentileWidth % 2 == 0:                         # xmhw.py:103-104
        raise XmhwException("smoothPercentileWidth should be odd")
    data, time, other, coords, attrs, enc, tattrs = _unpack(temp, tdim)   # xmhw.py:105-109
    cattrs = _unpack.last_coord_attrs
    if all(climatologyPeriod):                                 # xmhw.py:112-119
        from .identify import _ymd
        years = _ymd(time)[0]
        sel = (years >= int(climatologyPeriod[0])) & (years <= int(climatologyPeriod[1]))
        data, time = np.ascontiguousarray(data[sel]), time[sel]
    point = data.ndim == 1                                     # xmhw.py:122-126
    if not point:
        land_check_shape(data.shape, [tdim] + other, tdim)     # identify.py:509-516
    if get_calendar(time, enc, tattrs) == 360.0:               # xmhw.py:142-144
        tstep = True
    doy, ndoy = add_doy(time, keep_tstep=tstep)                # xmhw.py:145
    grid_shape = data.shape[1:]
    T = data.shape[0]
    flat = np.ascontiguousarray(data.reshape(T, -1))
    if not point and _all_land(flat):
        raise XmhwException("All points of grid are either land or NaN")   # identify.py:527-528
    if not torch.cuda.is_available():
        raise RuntimeError("xmhw_b200 needs a CUDA device (there is no CPU path)")
    # one pipelined pass over the host series (column blocks: copy in / kernels / copy out overlap);
    # the land census (identify.py:522-525) comes back from the device, no host pass over the array
    try:
        res = core.host_pipeline(torch.from_numpy(flat), doy, ndoy, do_threshold=True, do_detect=False,
                                 pctile=pctile, windowHalfWidth=windowHalfWidth, smoothPercentile=smoothPercentile,
                                 smoothPercentileWidth=smoothPercentileWidth, feb29=not tstep, negate=coldSpells,
                                 max_pad=_pad_steps(maxPadLength, time), anynans=anynans, slabs=_slabs(flat),
                                 clim_on_device=True)
    except NotImplementedError as exc:                       # a calendar / window the sweep plans cannot express
        raise XmhwException("threshold: %s" % exc)
    ocean = res["nvalid"].numpy() > 0
    if not point and not ocean.any():
        raise XmhwException("All points of grid are either land or NaN")   # identify.py:527-528
    # the climatologies are still on the device: rows / columns without ocean vanish and the coordinates
    # get sorted THERE (unstack, xmhw.py:213-214), then one copy to the host -- no host pass over 6 GB arrays
    th_d, se_d = res["thresh_dev"], res["seas_dev"]
    doy_coord = np.arange(1, ndoy + 1, dtype=np.int64)
    if not point:
        keep = _present(ocean.reshape(grid_shape))
        th_d = th_d.view((ndoy,) + tuple(grid_shape))
        se_d = se_d.view((ndoy,) + tuple(grid_shape))
        out_coords = {"doy": doy_coord}
        for ax, (d, k) in enumerate(zip(other, keep)):
            c = coords[d][k]
            srt = np.argsort(c, kind="stable")              # unstack returns sorted coordinate values
            idx = np.asarray(k)[srt]
            if len(idx) != grid_shape[ax] or not np.array_equal(idx, np.arange(grid_shape[ax])):
                idx_d = torch.from_numpy(np.ascontiguousarray(idx, np.int64)).to(th_d.device)
                th_d, se_d = th_d.index_select(ax + 1, idx_d), se_d.index_select(ax + 1, idx_d)
            out_coords[d] = c[srt]
        dims = ("doy",) + tuple(other)
    else:
        th_d, se_d = th_d[:, 0], se_d[:, 0]
        out_coords, dims = {"doy": doy_coord}, ("doy",)
    # a doy without samples disappears from the reference's groupby output (identify.py:233)
    present = (~torch.isnan(th_d).flatten(1).all(dim=1) if th_d.dim() > 1 else ~torch.isnan(th_d)).cpu().numpy()
    if not present.all():
        sel = torch.from_numpy(np.flatnonzero(present)).to(th_d.device)
        th_d, se_d = th_d.index_select(0, sel), se_d.index_select(0, sel)
        out_coords["doy"] = doy_coord[present]
    th_h, se_h = _to_host(th_d), _to_host(se_d)
    del th_d, se_d, res
    out_coords["quantile"] = np.float64(pctile / 100.0)
    ds = labeled.Dataset(coords=out_coords)
    ds["thresh"] = labeled.DataArray(th_h, dims, name="threshold")      # xmhw.py:215-216
    ds["seas"] = labeled.DataArray(se_h, dims, name="seasonal")
    annotate_ds(ds, dict({"ts": attrs}, **cattrs), "clim")
    from .identify import _ymd
    yrs = _ymd(time)[0]
    params = f"""Threshold calculated using:
    {pctile} percentile;
    climatology period is {yrs[0]}-{yrs[-1]}';
    window half width used for percentile is {windowHalfWidth}"""
    if skipna:
        params += """;
            NaNs where skipped in percentile and mean calculations"""
    if smoothPercentile:
        params += f""";
         width of moving average window to smooth percentile is
         {smoothPercentileWidth}"""
    if anynans:
        params += """;
            any grid point with even only 1 NaN along time
            axis has been removed from calculation"""
    ds.attrs["xmhw_parameters"] = params                       # xmhw.py:221-246
    return _wrap(ds, temp)


def _clim_to_grid(arr, other, grid_coords, grid_shape, ndoy, name, device):
    """Bring a (doy, ...) climatology onto the full (doy, cell) grid of `temp` by coordinate label (the
    reference matches cells positionally after land_check, xmhw.py:399-402).  The array is uploaded as it
    is and spread over the full grid ON THE DEVICE (NaN where it has no value): float64 [ndoy, ncell]."""
    import torch
    from . import core
    dims = list(arr.dims)
    if "doy" not in dims:
        raise XmhwException(f"{name} must have a 'doy' dimension")
    o = sorted(d for d in dims if d != "doy")
    if o != list(other):
        raise XmhwException(f"{name} dimensions {o} do not match the series dimensions {list(other)}")
    doyc = _coord(arr, "doy")
    # label -> position on the full grid, per axis (None: the axis already is the full one)
    index = []
    drow = (np.asarray(doyc, np.int64) - 1) if doyc is not None else None
    if drow is not None and (drow.min(initial=0) < 0 or drow.max(initial=0) >= ndoy):
        raise XmhwException(f"{name} has doy values outside 1..{ndoy}")
    index.append(None if drow is None or np.array_equal(drow, np.arange(ndoy)) else drow)
    for d in other:
        c = _coord(arr, d)
        if c is None or np.array_equal(c, grid_coords[d]):
            index.append(None)
            continue
        pos = {v: i for i, v in enumerate(grid_coords[d].tolist())}
        try:
            index.append(np.array([pos[v] for v in c.tolist()], np.int64))
        except KeyError:
            raise XmhwException(f"{name} has {d} values that are not on the series grid")
    host = torch.from_numpy(np.ascontiguousarray(np.asarray(arr.values, np.float64)))
    unpin = core._pin(host)
    try:
        data = host.to(device, non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
    finally:
        unpin()
    data = data.permute([dims.index("doy")] + [dims.index(d) for d in o])
    full_shape = (ndoy,) + tuple(grid_shape)
    for ax, ix in enumerate(index):
        if ix is None:
            if data.shape[ax] == full_shape[ax]:
                continue
            if data.shape[ax] > full_shape[ax]:
                raise XmhwException(f"{name} is larger than the series grid along {(['doy'] + list(other))[ax]}")
            ix = np.arange(data.shape[ax], dtype=np.int64)     # unlabeled and shorter: the leading positions
        shape = list(data.shape)
        shape[ax] = full_shape[ax]
        wide = torch.full(shape, float("nan"), dtype=torch.float64, device=device)
        wide.index_copy_(ax, torch.from_numpy(ix).to(device), data)
        data = wide
    return data.contiguous().view(ndoy, -1)


def detect(temp, th, se, tdim="time", minDuration=5, joinGaps=True, maxGap=2, maxPadLength=None,
           coldSpells=False, intermediate=False, anynans=False, tstep=False, compact=False):
    """Apply the Hobday et al. (2016) marine heat wave definition -- xmhw/xmhw.py:310-518.

    Same parameters as the reference plus `compact`: when True the result is the compact
    event table (one row per event with its cell coordinates) instead of the reference's
    dense (events, lat, lon) cube, which cannot exist at global scale.
    """
    import torch
    from . import core

    if maxGap >= minDuration:                                  # xmhw.py:373-377
        raise XmhwException("Maximum gap between mhw events should"
                            + " be smaller than event minimum duration")
    data, time, other, coords, attrs, enc, tattrs = _unpack(temp, tdim)
    cattrs = _unpack.last_coord_attrs
    point = data.ndim == 1
    if not point:
        land_check_shape(data.shape, [tdim] + other, tdim)
    doy, ndoy = add_doy(time, keep_tstep=tstep)                # xmhw.py:404
    grid_shape = data.shape[1:]
    T = data.shape[0]
    flat = np.ascontiguousarray(data.reshape(T, -1))
    if not point and _all_land(flat):
        raise XmhwException("All points of grid are either land or NaN")
    if not torch.cuda.is_available():
        raise RuntimeError("xmhw_b200 needs a CUDA device (there is no CPU path)")
    dev = torch.device("cuda", torch.cuda.current_device())
    thd = _clim_to_grid(th, other, coords, grid_shape, ndoy, "th", dev)
    sed = _clim_to_grid(se, other, coords, grid_shape, ndoy, "se", dev)
    ts = None
    if intermediate:
        # the per-timestep dataset needs the whole series on the device at once (small grids only)
        nbytes = T * flat.shape[1] * 80
        if nbytes > DENSE_LIMIT_BYTES:
            raise XmhwException("the per-timestep (intermediate) dataset would need %.1f GB; "
                                "split the grid (reference docs/dask.rst)" % (nbytes / 1e9))
        ts = torch.from_numpy(flat).cuda()
        nvalid = core.count_valid(ts)
        if anynans:
            ts[:, nvalid != T] = float("nan")
            nvalid = torch.where(nvalid == T, nvalid, torch.zeros_like(nvalid))
        if _pad_steps(maxPadLength, time):                     # xmhw.py:409-410
            core.interp_gaps_(ts, _pad_steps(maxPadLength, time))
        if coldSpells:                                         # xmhw.py:412-413
            ts = -ts
        ev = core.detect_arrays(ts, doy, ndoy, thd, sed, minDuration, joinGaps, maxGap)
        ocean = nvalid.cpu().numpy() > 0
        ei_t, ef_t = ev.i32[:, :ev.n].cpu(), ev.f64[:, :ev.n].cpu()
    else:
        # one pipelined pass over the host series; th / se already sit on the device, sliced per column block
        res = core.host_pipeline(torch.from_numpy(flat), doy, ndoy, do_threshold=False, do_detect=True,
                                 th_dev=thd, se_dev=sed, minDuration=minDuration, joinGaps=joinGaps,
                                 maxGap=maxGap, negate=coldSpells, max_pad=_pad_steps(maxPadLength, time), anynans=anynans,
                                 slabs=_slabs(flat))
        ocean = res["nvalid"].numpy() > 0
        ei_t, ef_t = res["ev_i32"], res["ev_f64"]
        del thd, sed
    # event table columns: int32 [EI_COUNT, n] / float64 [EF_COUNT, n] host tensors; conversions and gathers
    # of the (tens of millions of) rows go through torch, which uses all host cores
    icol = {f: ei_t[k] for k, f in enumerate(core.EI_FIELDS)}
    fcol = {f: ef_t[k] for k, f in enumerate(core.EF_FIELDS)}
    tab = {f: icol[f].numpy() for f in ("cell", "index_start", "index_end", "index_peak", "category")}
    if not point and not ocean.any():
        raise XmhwException("All points of grid are either land or NaN")
    inter = None
    if intermediate:                                           # identify.py:404-411
        inter = {k: v.cpu().numpy() for k, v in core.intermediate_arrays(ts, doy, ndoy, thd, sed, ev).items()}
        inter["ts"] = ts.cpu().numpy()
    n = len(tab["cell"])
    cols = {"event": icol["index_start"].to(torch.float64).numpy()}
    for f in ("index_start", "index_end", "index_peak", "duration", "category"):
        cols[f] = icol[f].to(torch.float64).numpy()            # integer-valued float64 in the reference
    cols["category"][tab["category"] < 0] = np.nan
    for f in ("duration_moderate", "duration_strong", "duration_severe", "duration_extreme"):
        cols[f] = icol[f].to(torch.int64).numpy()
    for f in core.EF_FIELDS:
        cols[f] = (fcol[f].to(torch.float32) if f in FLOAT32_VARIABLES else fcol[f]).numpy()
    time = np.asarray(time)
    cols["time_start"] = _take1(time, icol["index_start"])
    cols["time_end"] = _take1(time, icol["index_end"])
    cols["time_peak"] = _take1(time, icol["index_peak"])
    if coldSpells:                                             # xmhw.py:481-482
        cols = flip_cold(cols)
    params = f"MHW detected using: {minDuration} days of minimum duration"
    if joinGaps:
        params += f""";
            events separated by {maxGap} or less days were joined"""
    if coldSpells:
        params += """;
                cold events were detected instead of heat events"""
    if maxPadLength:
        params += f""";
            where original timeseries had missing values interpolation
            was used to fill them. Gaps > {maxPadLength} days long were
            left as NaNs;"""
    if anynans:
        params += """;
            any grid point with even only 1 NaN along time
            axis has been removed from calculation"""
    cell_idx = ()
    if not point:                                              # unravel_index on all host cores
        rem, parts = icol["cell"].to(torch.int64), []
        for size in reversed(grid_shape):
            parts.append(torch.remainder(rem, size))
            rem = torch.div(rem, size, rounding_mode="floor")
        cell_idx = tuple(reversed(parts))
    if compact or point:
        ds = labeled.Dataset(coords={"events": tab["index_start"].astype(np.int64) if point else np.arange(n)})
        dim = ("events",) if point else ("row",)
        if not point:
            ds.coords = {"row": np.arange(n)}
            for d, ix in zip(other, cell_idx):
                ds[d] = labeled.DataArray(_take1(coords[d], ix), dim)
        for v in EVENT_VARIABLES:
            ds[v] = labeled.DataArray(cols[v], dim)
    else:
        keep = _present(ocean.reshape(grid_shape))
        events = np.unique(tab["index_start"])
        shape = (len(events),) + tuple(len(k) for k in keep)
        nbytes = int(np.prod(shape)) * 8 * len(EVENT_VARIABLES)
        if nbytes > DENSE_LIMIT_BYTES:
            raise XmhwException("the dense (events, %s) result would need %.1f GB; call detect(..., compact=True) "
                                "or split the grid (reference docs/dask.rst)" % (", ".join(other), nbytes / 1e9))
        erow = np.searchsorted(events, tab["index_start"])
        out_coords = {"events": events}
        pos = []
        for d, k, ix in zip(other, keep, cell_idx):
            c = coords[d][k]
            srt = np.argsort(c, kind="stable")
            inv = np.empty(len(coords[d]), np.int64)
            inv[k[srt]] = np.arange(len(k))
            pos.append(inv[ix.numpy()])
            out_coords[d] = c[srt]
        ds = labeled.Dataset(coords=out_coords)
        dims = ("events",) + tuple(other)
        for v in EVENT_VARIABLES:
            col = cols[v]
            if np.issubdtype(col.dtype, np.datetime64):
                dense = np.full(shape, np.datetime64("NaT"), dtype=col.dtype)
            elif col.dtype == object:
                dense = np.full(shape, None, dtype=object)
            else:
                dense = np.full(shape, np.nan, dtype=np.float32 if col.dtype == np.float32 else np.float64)
            dense[(erow,) + tuple(pos)] = col
            ds[v] = labeled.DataArray(dense, dims)
    annotate_ds(ds, dict({"ts": attrs}, **cattrs), "mhw")
    ds.attrs["xmhw_parameters"] = params                       # xmhw.py:487-515
    if intermediate:                                           # xmhw.py:461-463, :471-478
        if point:
            di = labeled.Dataset(coords={tdim: time})
            for k, v in inter.items():
                di[k] = labeled.DataArray(v[:, 0], (tdim,))
        else:
            keep = _present(ocean.reshape(grid_shape))
            out_coords = {tdim: time}
            order = []
            for d, k in zip(other, keep):
                c = coords[d][k]
                srt = np.argsort(c, kind="stable")
                order.append(k[srt])
                out_coords[d] = c[srt]
            di = labeled.Dataset(coords=out_coords)
            for name, v in inter.items():
                a = v.reshape((T,) + tuple(grid_shape))
                for ax, ix in enumerate(order):
                    a = np.take(a, ix, axis=ax + 1)
                di[name] = labeled.DataArray(a, (tdim,) + tuple(other))
        return _wrap(ds, temp), _wrap(di, temp)
    return _wrap(ds, temp)
