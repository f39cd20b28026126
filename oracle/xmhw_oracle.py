"""numpy restatement of the xmhw threshold/detect hot path (float64 oracle).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Parity status: PINNED
(goldens: tests/test_oracle_golden.py; reference fuzz: tests/test_oracle_vs_reference.py).

Every function cites the reference lines it restates (paths relative to
/root/reference).  Arithmetic that the reference delegates to xarray/numpy/
pandas is written out with a *defined* evaluation order so the CUDA path can be
held to it:

* quantile     numpy 2.3 `method="linear"`: v=(n-1)q in f64, a=s[floor v],
               b=s[floor v + 1] (both = max when v >= n-1), d=b-a in the INPUT
               dtype (float32), result a+d*g (g<0.5) or b-d*(1-g) (g>=0.5) in
               f64, two roundings (numpy/lib/_function_base_impl.py:_quantile,
               _get_indexes, _get_gamma, _lerp).  Checked bit-for-bit against
               np.quantile / np.nanquantile in tests/test_oracle_golden.py.
* seasonal     f64 sum of the f32 samples in window order (k-major then t,
               identify.py:207) divided by n.
* feb29        sequential f64 sum over the doys present in {59,60,61} / count.
* runavg       fresh left-to-right f64 sum of W terms / W.
* detect       plain-rule run-length encoding equivalent to identify.py:441-470
               (fuzz-verified), per-event statistics in f64 with pandas'
               skip-NaN semantics; `*_abs` fields rounded to float32 like the
               reference's float32 `mabs` column.
"""
import math

import numpy as np

# ---------------------------------------------------------------------------
# calendar helpers (host-side in the reference too)
# ---------------------------------------------------------------------------


def add_doy(time, keep_tstep=False):
    """identify.py:28-79.  `time` is a numpy datetime64 array.

    Daily: doy = dayofyear + (not leap and month >= 3)  -> 366-day calendar
    (identify.py:73-76).  keep_tstep: steps/year = number of steps in the SECOND
    calendar year (identify.py:59-60); len(t) must be a multiple of it
    (identify.py:61-66); doy = 1..steps tiled (identify.py:67-70).
    """
    t = np.asarray(time).astype("datetime64[s]")
    years = t.astype("datetime64[Y]").astype(np.int64) + 1970
    if keep_tstep:
        uy = np.unique(years)
        steps = int(np.sum(years == uy[1]))
        if len(t) % steps != 0:
            raise ValueError("timeseries has to have complete years")
        return np.tile(np.arange(1, steps + 1, dtype=np.int64), len(t) // steps)
    days = t.astype("datetime64[D]")
    jan1 = days.astype("datetime64[Y]").astype("datetime64[D]")
    dayofyear = (days - jan1).astype(np.int64) + 1
    month = t.astype("datetime64[M]").astype(np.int64) % 12 + 1
    leap = (years % 4 == 0) & ((years % 100 != 0) | (years % 400 == 0))
    return dayofyear + ((~leap) & (month >= 3)).astype(np.int64)


def land_mask(ts, anynans=False):
    """identify.py:520-525: cells (columns of ts[T, ncell]) kept by land_check:
    not all-NaN (`how="all"`), or with no NaN at all when anynans."""
    nan = np.isnan(ts)
    return ~nan.any(axis=0) if anynans else ~nan.all(axis=0)


def interp_gaps(ts, max_pad):
    """xmhw.py:159-160, :409-410 `interpolate_na(dim=tdim, max_gap=maxPadLength)`: linear
    interpolation along time (axis 0) of interior NaN runs of at most max_pad steps
    (Oliver's `pad` semantics: run length in samples; np.interp arithmetic in float64)."""
    out = np.array(ts, np.float32, copy=True)
    T = out.shape[0]
    flat = out.reshape(T, -1)
    idx = np.arange(T)
    for c in range(flat.shape[1]):
        col = flat[:, c]
        nan = np.isnan(col)
        if not nan.any() or nan.all():
            continue
        filled = np.interp(idx, idx[~nan], col[~nan]).astype(np.float32)
        edges = np.diff(np.concatenate(([0], nan.view(np.int8), [0])))
        for s, e in zip(np.nonzero(edges == 1)[0], np.nonzero(edges == -1)[0]):
            if s > 0 and e < T and (e - s) <= max_pad:
                col[s:e] = filled[s:e]
    return out


# ---------------------------------------------------------------------------
# climatology
# ---------------------------------------------------------------------------


def window_index(doy, d, w):
    """identify.py:204-208: time indices of the samples pooled for day-of-year d,
    in the reference's stack order (window offset k major, then time), edge
    positions (t+k outside the series) dropped as `dropna(dim="z")` does."""
    T = len(doy)
    t = np.nonzero(np.asarray(doy) == d)[0]
    out = []
    for k in range(-w, w + 1):
        tk = t + k
        out.append(tk[(tk >= 0) & (tk < T)])
    return np.concatenate(out) if out else np.zeros(0, np.int64)


def quantile_table(nmax, q):
    """(floor(v), gamma) for n = 0..nmax, numpy 2.x linear method: v=(n-1)*q."""
    n = np.arange(nmax + 1, dtype=np.int64)
    v = (n - 1) * np.float64(q)
    lo = np.floor(v)
    gamma = v - lo
    above = v >= (n - 1)
    lo = np.where(above, n - 1, lo)
    gamma = np.where(above, 0.0, gamma)  # a == b == max there, gamma irrelevant
    lo = np.where(n == 0, 0, lo)
    return lo.astype(np.int64), gamma.astype(np.float64)


def lerp(a, b, g):
    """numpy `_lerp` on float32 a,b and float64 g."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    d = (b - a).astype(np.float32)              # subtract in the input dtype
    d64 = d.astype(np.float64)
    lo = a.astype(np.float64) + d64 * g
    hi = b.astype(np.float64) - d64 * (1.0 - g)
    return np.where(g >= 0.5, hi, lo)


def quantile_linear(x, q):
    """Quantile of the non-NaN entries of each column of x[n, ncell] (float32)
    -> float64[ncell]; NaN where a column has no sample.  identify.py:233-235
    via np.quantile / np.nanquantile (dropna at identify.py:208 already removed
    NaNs, so skipna never changes the result)."""
    x = np.asarray(x, np.float32)
    if x.ndim == 1:
        x = x[:, None]
    ncell = x.shape[1]
    if x.shape[0] == 0:
        return np.full(ncell, np.nan)
    s = np.sort(x, axis=0)                      # NaNs sort to the end
    n = np.sum(~np.isnan(x), axis=0)
    v = (n - 1) * np.float64(q)
    lo = np.floor(v)
    g = v - lo
    hi = lo + 1
    above = v >= (n - 1)
    lo = np.where(above, n - 1, lo)
    hi = np.where(above, n - 1, hi)
    lo = np.clip(lo, 0, None).astype(np.int64)
    hi = np.clip(hi, 0, None).astype(np.int64)
    cols = np.arange(ncell)
    a = s[lo, cols]
    b = s[hi, cols]
    out = lerp(a, b, g)
    return np.where(n == 0, np.nan, out)


def seasonal_mean(x):
    """identify.py:263: mean of the non-NaN entries of each column, f64 sum in
    row (= window stack) order."""
    x = np.asarray(x, np.float32)
    if x.ndim == 1:
        x = x[:, None]
    acc = np.zeros(x.shape[1], np.float64)
    n = np.zeros(x.shape[1], np.int64)
    for row in x:
        ok = ~np.isnan(row)
        acc = np.where(ok, acc + row.astype(np.float64), acc)
        n += ok
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(n > 0, acc / n, np.nan)


def feb29(x):
    """identify.py:137-151 (+ :237-240, :265-268): value for doy 60 = mean over the
    doys in {59,60,61} that are present (NaN-skipping), rows of x are doy 1..ndoy."""
    acc = np.zeros(x.shape[1:], np.float64)
    n = np.zeros(x.shape[1:], np.int64)
    for d in (59, 60, 61):
        v = x[d - 1]
        ok = ~np.isnan(v)
        acc = np.where(ok, acc + v, acc)
        n += ok
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(n > 0, acc / n, np.nan)


def runavg(x, w):
    """identify.py:154-181: circular centred moving mean of odd width w over the
    doy axis (axis 0); any NaN in a window gives NaN (rolling min_periods=w).

    The reference's rounding depends on the installed rolling backend (bottleneck
    running sum or numpy window mean), so the summation ORDER is defined here and the
    CUDA kernel follows it bit for bit: wrap the series, e[j] = x[(j - h) mod nd] for
    j = 0 .. nd + w - 2, cut e into blocks of w, and write window d = b w + p as the
    right-to-left suffix sum of block b from p plus the left-to-right prefix sum of
    block b + 1 up to p - 1 (p == 0: the suffix alone).  Any order is within ~1e-15
    relative of any other; this one costs 3 additions per output."""
    if w % 2 == 0:
        raise ValueError("Running average window should be odd")
    x = np.asarray(x, np.float64)
    nd = x.shape[0]
    h = (w - 1) // 2
    L = nd + w - 1
    e = x[(np.arange(L) - h) % nd]
    pre = np.empty_like(e)
    suf = np.empty_like(e)
    for lo in range(0, L, w):
        hi = min(L, lo + w)
        acc = e[lo].copy()
        pre[lo] = acc
        for j in range(lo + 1, hi):
            acc = acc + e[j]
            pre[j] = acc
        acc = e[hi - 1].copy()
        suf[hi - 1] = acc
        for j in range(hi - 2, lo - 1, -1):
            acc = e[j] + acc
            suf[j] = acc
    out = np.empty_like(x, dtype=np.float64)
    for d in range(nd):
        out[d] = (suf[d] if d % w == 0 else suf[d] + pre[d + w - 1]) / w
    return out


def threshold(ts, doy, ndoy, pctile=90, windowHalfWidth=5, smoothPercentile=True,
              smoothPercentileWidth=31, tstep=False):
    """xmhw.py:250-307 (calc_clim) for every column of ts[T, ncell] (float32).

    Returns (thresh, seas) float64 [ndoy, ncell]; a (cell, doy) without samples
    is NaN (the reference drops that doy from the groupby output)."""
    ts = np.asarray(ts, np.float32)
    if ts.ndim == 1:
        ts = ts[:, None]
    doy = np.asarray(doy)
    ncell = ts.shape[1]
    thresh = np.full((ndoy, ncell), np.nan)
    seas = np.full((ndoy, ncell), np.nan)
    q = pctile / 100.0
    for d in range(1, ndoy + 1):
        idx = window_index(doy, d, windowHalfWidth)
        if len(idx) == 0:
            continue
        x = ts[idx]
        thresh[d - 1] = quantile_linear(x, q)
        seas[d - 1] = seasonal_mean(x)
    # A (cell, doy) without samples is absent from that cell's groupby output (identify.py:233,
    # :263 after the dropna at :208), so feb29 acts on the doy LABELS that are present
    # (`where(doy != 60, feb29(...))`, identify.py:137-151, :237-240) and runavg pads / rolls over
    # the cell's own compacted doy axis (identify.py:175-180).  Cells are grouped by their
    # presence pattern (almost always "all present").
    pres = ~np.isnan(thresh)
    pat, inv = np.unique(pres.T, axis=0, return_inverse=True)
    inv = np.asarray(inv).reshape(-1)
    for k in range(len(pat)):
        present = pat[k]
        if not present.any():
            continue
        cells = np.nonzero(inv == k)[0]
        didx = np.nonzero(present)[0]
        t_c, s_c = thresh[np.ix_(didx, cells)], seas[np.ix_(didx, cells)]
        if not tstep and ndoy >= 61 and present[59]:
            members = [int(np.searchsorted(didx, d)) for d in (58, 59, 60) if present[d]]
            i60 = int(np.searchsorted(didx, 59))
            for arr in (t_c, s_c):
                acc = np.zeros(len(cells))
                for m_ in members:
                    acc = acc + arr[m_]
                arr[i60] = acc / len(members)
        if smoothPercentile:
            t_c = runavg(t_c, smoothPercentileWidth)
            s_c = runavg(s_c, smoothPercentileWidth)
        thresh[np.ix_(didx, cells)] = t_c
        seas[np.ix_(didx, cells)] = s_c
    return thresh, seas


# ---------------------------------------------------------------------------
# detection
# ---------------------------------------------------------------------------


def exceedance(ts, doy, thresh):
    """identify.py:367-372: thresh looked up by doy label, strict `>` evaluated in
    float64 (float32 ts is upcast exactly); NaN on either side -> False."""
    th_t = np.asarray(thresh)[np.asarray(doy) - 1]
    with np.errstate(invalid="ignore"):
        return np.asarray(ts, np.float64) > th_t


def find_events(b, minDuration=5, joinGaps=True, maxGap=2):
    """Plain-rule equivalent of mhw_filter/join_gaps (identify.py:415-479, :273-325):

    (i)   index 0 is never part of an event (`ffill().fillna(0)` at :441);
    (ii)  maximal runs of True of length >= minDuration qualify (:458); a run
          reaching the last index is closed there (:454-455);
    (iii) if joinGaps, consecutive QUALIFIED events with
          start_j - end_{j-1} - 1 <= maxGap are merged, chains allowed (:310-321).
    Returns int64 arrays (start, end) after joining."""
    b = np.asarray(b, bool).copy()
    T = len(b)
    if T == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    b[0] = False
    pad = np.concatenate(([False], b, [False]))
    diff = np.diff(pad.astype(np.int8))
    starts = np.nonzero(diff == 1)[0]
    ends = np.nonzero(diff == -1)[0] - 1
    keep = (ends - starts + 1) >= minDuration
    starts, ends = starts[keep], ends[keep]
    if joinGaps and len(starts) > 1:
        gap = starts[1:] - ends[:-1] - 1
        brk = gap > maxGap
        first = np.concatenate(([True], brk))
        last = np.concatenate((brk, [True]))
        starts, ends = starts[first], ends[last]
    return starts.astype(np.int64), ends.astype(np.int64)


INT_FIELDS = ("index_start", "index_end", "index_peak", "duration", "category",
              "duration_moderate", "duration_strong", "duration_severe", "duration_extreme")
F64_FIELDS = ("intensity_max", "intensity_mean", "intensity_cumulative", "intensity_var",
              "severity_max", "severity_mean", "severity_cumulative", "severity_var",
              "intensity_max_relThresh", "intensity_mean_relThresh",
              "intensity_cumulative_relThresh", "intensity_var_relThresh",
              "intensity_max_abs", "intensity_mean_abs", "intensity_cumulative_abs",
              "intensity_var_abs", "rate_onset", "rate_decline")
F32_FIELDS = ("intensity_max_abs", "intensity_mean_abs", "intensity_cumulative_abs",
              "intensity_var_abs")       # float32 in the reference (mabs column, features.py:68)


def _std1(x):
    """sqrt of the ddof=1 variance of non-NaN x (features.py:139-145, :182-187)."""
    x = x[~np.isnan(x)]
    if len(x) < 2:
        return np.nan
    m = np.sum(x) / len(x)
    return math.sqrt(np.sum((x - m) ** 2) / (len(x) - 1))


def event_stats(ts, thresh_t, seas_t, start, end):
    """features.py:22-69 (mhw_df), :97-193 (agg_df, properties), :225-295
    (get_period, get_edge, onset_decline) for one event [start, end] of one cell.

    ts float32[T]; thresh_t/seas_t float64[T] already looked up by doy.
    pandas reductions skip NaN; first/last are first/last non-null."""
    T = len(ts)
    ts64 = np.asarray(ts, np.float32).astype(np.float64)
    sl = slice(start, end + 1)
    x, th, se = ts64[sl], thresh_t[sl], seas_t[sl]
    with np.errstate(invalid="ignore", divide="ignore"):
        relS = x - se                                   # features.py:52,55
        relT = x - th                                   # :53,56
        ths = th - se                                   # :54
        norm = relT / ths                               # :57
        sev = relS / -(ths)                             # :59-61
        cats = np.floor(1.0 + norm)                     # :62
        anom = ts64 - seas_t                            # :44 (unmasked)
    out = {}
    ok = ~np.isnan(relS)
    if not ok.any():
        raise ValueError("event without a valid day")
    imax = int(np.nanargmax(relS))                      # :120, first occurrence, NaN skipped
    out["index_start"] = start
    out["index_end"] = end
    out["index_peak"] = start + imax                    # :181
    out["duration"] = end - start + 1                   # :189
    cm = np.nanmax(cats) if (~np.isnan(cats)).any() else np.nan
    out["category"] = min(cm, 4) if not np.isnan(cm) else np.nan   # :147,188
    out["duration_moderate"] = int(np.sum(cats == 1.0))  # :63,148
    out["duration_strong"] = int(np.sum(cats == 2.0))
    out["duration_severe"] = int(np.sum(cats == 3.0))
    out["duration_extreme"] = int(np.sum(cats >= 4.0))

    def agg(v, name, peakval):
        vv = v[~np.isnan(v)]
        out[name % "max"] = peakval
        out[name % "mean"] = np.sum(vv) / len(vv) if len(vv) else np.nan
        out[name % "cumulative"] = np.sum(vv) if len(vv) else 0.0
        out[name % "var"] = _std1(v)

    agg(relS, "intensity_%s", relS[imax])               # :133-135,140,182
    agg(sev, "severity_%s", np.nanmax(sev) if (~np.isnan(sev)).any() else np.nan)  # :136-139,183
    agg(relT, "intensity_%s_relThresh", relT[imax])     # :141-143,184,186
    agg(x, "intensity_%s_abs", x[imax])                 # :144-146,185,187
    for f in F32_FIELDS:
        out[f] = float(np.float32(out[f]))
    # onset / decline (features.py:225-295)
    p = imax
    onset_period = (p if p != 0 else 1) + (0.0 if start == 0 else 0.5)     # :259-260
    y = (end - start - p) if p != T - 1 else 1                              # :258,261
    decline_period = y + (0.0 if end == T - 1 else 0.5)                     # :262
    relS_first = relS[ok][0]
    relS_last = relS[ok][-1]
    ap = np.concatenate(([np.nan], anom[:-1]))[sl]      # anom_plus = anom.shift(+1), :45
    am = np.concatenate((anom[1:], [np.nan]))[sl]       # anom_minus = anom.shift(-1), :46
    apv, amv = ap[~np.isnan(ap)], am[~np.isnan(am)]
    anom_first = apv[0] if len(apv) else np.nan         # :129 first non-null
    anom_last = amv[-1] if len(amv) else np.nan         # :130 last non-null
    edge_s = 0.5 * (relS_first + (relS_first if start == 0 else anom_first))   # :220-221,287
    edge_e = 0.5 * (relS_last + (relS_last if end == T - 1 else anom_last))    # :288
    with np.errstate(invalid="ignore", divide="ignore"):
        out["rate_onset"] = float((np.float64(out["intensity_max"]) - edge_s) / np.float64(onset_period))
        out["rate_decline"] = float((np.float64(out["intensity_max"]) - edge_e) / np.float64(decline_period))
    return out


INTER_FIELDS = ("events", "seas", "thresh", "relSeas", "relThresh", "relThreshNorm", "severity", "cats",
                "mabs", "bthresh", "duration_moderate", "duration_strong", "duration_severe",
                "duration_extreme")


def intermediate(ts, doy, thresh, seas, minDuration=5, joinGaps=True, maxGap=2):
    """identify.py:404-411 + features.py:22-69 (mhw_df): per-timestep dataset returned with
    `intermediate=True`, for every column of ts[T, ncell].  Event days carry the event label
    (start index); seas/thresh/relSeas/... are NaN off events, flags False."""
    ts = np.asarray(ts, np.float32)
    if ts.ndim == 1:
        ts = ts[:, None]
        thresh = np.asarray(thresh)[:, None] if np.ndim(thresh) == 1 else thresh
        seas = np.asarray(seas)[:, None] if np.ndim(seas) == 1 else seas
    doy = np.asarray(doy)
    T, n = ts.shape
    th_t = np.asarray(thresh, np.float64)[doy - 1]
    se_t = np.asarray(seas, np.float64)[doy - 1]
    x = ts.astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        b = x > th_t                                            # identify.py:372
        events = np.full((T, n), np.nan)
        for c in range(n):
            s, e = find_events(b[:, c], minDuration, joinGaps, maxGap)
            for si, ei in zip(s, e):
                events[si:ei + 1, c] = si                       # identify.py:466-470, :534-535
        m = ~np.isnan(events)                                   # features.py:38
        relS, relT, ths = x - se_t, x - th_t, th_t - se_t       # :52-54
        norm, sev = relT / ths, relS / -(ths)                   # :57-61
        cats = np.floor(1.0 + norm)                             # :62
    nanw = lambda a: np.where(m, a, np.nan)
    return {"events": events, "seas": nanw(se_t), "thresh": nanw(th_t), "relSeas": nanw(relS),
            "relThresh": nanw(relT), "relThreshNorm": nanw(norm), "severity": nanw(sev), "cats": nanw(cats),
            "mabs": np.where(m, ts, np.float32(np.nan)).astype(np.float32), "bthresh": b,
            "duration_moderate": m & (cats == 1.0), "duration_strong": m & (cats == 2.0),
            "duration_severe": m & (cats == 3.0), "duration_extreme": m & (cats >= 4.0)}


def detect(ts, doy, thresh, seas, minDuration=5, joinGaps=True, maxGap=2):
    """xmhw.py:440-454 + identify.py:328-412 for every column of ts[T, ncell].

    thresh/seas are float64 [ndoy, ncell].  Returns a dict of arrays (one entry
    per event, ordered by cell then start): 'cell' + INT_FIELDS (int64) +
    F64_FIELDS (float64; the *_abs ones hold float32-rounded values)."""
    ts = np.asarray(ts, np.float32)
    if ts.ndim == 1:
        ts = ts[:, None]
        thresh = np.asarray(thresh)[:, None] if np.ndim(thresh) == 1 else thresh
        seas = np.asarray(seas)[:, None] if np.ndim(seas) == 1 else seas
    doy = np.asarray(doy)
    rows = []
    for c in range(ts.shape[1]):
        th_t = np.asarray(thresh[:, c], np.float64)[doy - 1]
        se_t = np.asarray(seas[:, c], np.float64)[doy - 1]
        with np.errstate(invalid="ignore"):
            b = ts[:, c].astype(np.float64) > th_t
        s, e = find_events(b, minDuration, joinGaps, maxGap)
        for si, ei in zip(s, e):
            r = event_stats(ts[:, c], th_t, se_t, int(si), int(ei))
            r["cell"] = c
            rows.append(r)
    out = {"cell": np.array([r["cell"] for r in rows], np.int64)}
    for f in INT_FIELDS:
        vals = [r[f] for r in rows]
        out[f] = np.array([-1 if (isinstance(v, float) and np.isnan(v)) else v for v in vals],
                          np.int64)
    for f in F64_FIELDS:
        out[f] = np.array([r[f] for r in rows], np.float64)
    return out


# ---------------------------------------------------------------------------
# downstream statistics (reference xmhw/stats.py; semantics stated in xmhw_b200/stats.py)
# ---------------------------------------------------------------------------
BLOCK_MEAN_FIELDS = ("duration", "intensity_max", "intensity_mean", "intensity_var", "intensity_cumulative",
                     "intensity_max_relThresh", "intensity_mean_relThresh", "intensity_var_relThresh",
                     "intensity_cumulative_relThresh", "intensity_max_abs", "intensity_mean_abs", "intensity_var_abs",
                     "intensity_cumulative_abs", "severity_mean", "severity_cumulative", "rate_onset", "rate_decline")


def block_average(ev, years, ncell, period=None, blockLength=1, mtime="index_start", ts=None, doy=None,
                  thresh=None, seas=None):
    """stats.py:27-183 with agg_mhw / agg_ts / agg_cats (:322-428) per cell, plain loops.
    ev: event dict of `detect`; years [T] calendar year per step.  Returns dict name -> [nblocks, ncell]."""
    years = np.asarray(years, np.int64)
    if period is None:
        period = (int(years[0]), int(years[-1]))
    bins = np.arange(period[0], period[1] + blockLength + 1, blockLength)
    nb = len(bins) - 1
    blk = (years - bins[0]) // blockLength
    blk = np.where((years >= bins[0]) & (blk < nb), blk, -1)
    out = {"ecount": np.zeros((nb, ncell)), "intensity_max_max": np.full((nb, ncell), np.nan),
           "total_icum": np.zeros((nb, ncell))}
    for f in BLOCK_MEAN_FIELDS:
        out[f] = np.full((nb, ncell), np.nan)
    eb = blk[ev[mtime]] if len(ev["cell"]) else np.zeros(0, np.int64)
    for c in range(ncell):
        for b in range(nb):
            m = (ev["cell"] == c) & (eb == b)
            if not m.any():
                continue
            out["ecount"][b, c] = m.sum()
            with np.errstate(all="ignore"):
                for f in BLOCK_MEAN_FIELDS:
                    v = ev[f][m].astype(np.float64)
                    if (~np.isnan(v)).any():
                        out[f][b, c] = np.nansum(v) / (~np.isnan(v)).sum()      # sequential-sum mean like pandas
                v = ev["intensity_max"][m]
                if (~np.isnan(v)).any():
                    out["intensity_max_max"][b, c] = np.nanmax(v)
                out["total_icum"][b, c] = np.nansum(ev["intensity_cumulative"][m])
    if ts is not None:
        ts = np.asarray(ts, np.float32)
        for f in ("ts_mean", "ts_max", "ts_min"):
            out[f] = np.full((nb, ncell), np.nan)
        for b in range(nb):
            x = ts[blk == b]
            ok = ~np.isnan(x)
            n = ok.sum(0)
            with np.errstate(all="ignore"):
                out["ts_mean"][b] = np.where(n > 0, np.where(ok, x.astype(np.float64), 0.0).sum(0) / np.maximum(n, 1), np.nan)
                out["ts_max"][b] = np.where(n > 0, np.nanmax(np.where(ok, x, -np.inf), 0), np.nan)
                out["ts_min"][b] = np.where(n > 0, np.nanmin(np.where(ok, x, np.inf), 0), np.nan)
        if thresh is not None and seas is not None and doy is not None:
            days = np.zeros((4, nb, ncell), np.int64)
            d = np.asarray(doy) - 1
            for i in range(len(ev["cell"])):
                c = int(ev["cell"][i])
                for t in range(int(ev["index_start"][i]), int(ev["index_end"][i]) + 1):
                    if blk[t] < 0:
                        continue
                    x = np.float64(ts[t, c])
                    th, se = thresh[d[t], c], seas[d[t], c]
                    with np.errstate(all="ignore"):
                        cat = np.floor(1.0 + (x - th) / (th - se))
                    if cat >= 1:
                        days[min(int(cat), 4) - 1, blk[t], c] += 1
            for k, f in enumerate(("moderate_days", "strong_days", "severe_days", "extreme_days")):
                out[f] = days[k]
            out["total_days"] = days.sum(0)
    return out, bins[:-1]


def rank_in_cell(ev, field):
    """stats.py:493-510 per cell: len - argsort(argsort(x)) (stable; NaN sorts last)."""
    r = np.zeros(len(ev["cell"]))
    for c in np.unique(ev["cell"]):
        m = np.flatnonzero(ev["cell"] == c)
        x = np.asarray(ev[field][m], np.float64)
        r[m] = len(x) - np.argsort(np.argsort(x, kind="stable"), kind="stable")
    return r
