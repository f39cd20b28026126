"""Run the UNMODIFIED reference detect arithmetic from /root/reference.

TEST INFRASTRUCTURE ONLY.  xarray and dask are not installed in this image, but
`xmhw/features.py` needs only numpy and `mhw_filter`/`join_gaps`/`join_events`
in `xmhw/identify.py` are pure pandas.  With stub `xarray`/`dask` modules in
`sys.modules` the three files `exception.py`, `features.py`, `identify.py`
import unchanged, and the reference's own code can be driven with the
DataFrame that `ds.to_dataframe()` would yield at `identify.py:377`.

Used to (a) pin oracle/xmhw_oracle.py (tests/test_oracle_vs_reference.py), (b) generate
golden event tables (oracle/make_golden.py), (c) time the reference's own pandas detect as
the CPU arm of bench.py.  /root/reference does not exist on the GPU box: there the
unmodified copies under oracle/_ref (oracle/build_ref.py) are loaded instead; callers
must check `available()`.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import pandas as pd

REF_ROOT = os.environ.get("XMHW_REFERENCE_ROOT", "/root/reference")
# the unmodified copies made by oracle/build_ref.py (git-ignored; they travel to the GPU box)
_COPY_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
if not os.path.isfile(os.path.join(REF_ROOT, "xmhw", "identify.py")) and \
        os.path.isfile(os.path.join(_COPY_ROOT, "xmhw", "identify.py")):
    REF_ROOT = _COPY_ROOT
_mods = None


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "xmhw", "identify.py"))


def _delayed(*args, **kwargs):
    # `@dask.delayed(nout=1)` and bare `@dask.delayed` both become identity
    if len(args) == 1 and callable(args[0]) and not kwargs:
        return args[0]
    return lambda f: f


def load():
    """Return (identify, features) reference modules, loaded once."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REF_ROOT)
    saved = {k: sys.modules.get(k) for k in ("xarray", "dask")}
    sys.modules["xarray"] = types.ModuleType("xarray")
    dask = types.ModuleType("dask")
    dask.delayed = _delayed
    dask.compute = lambda *a, **k: a
    sys.modules["dask"] = dask
    try:
        pkg = types.ModuleType("_xmhw_ref")
        pkg.__path__ = [os.path.join(REF_ROOT, "xmhw")]
        sys.modules["_xmhw_ref"] = pkg
        out = []
        for name in ("exception", "features", "identify"):
            spec = importlib.util.spec_from_file_location(
                "_xmhw_ref.%s" % name, os.path.join(REF_ROOT, "xmhw", "%s.py" % name))
            mod = importlib.util.module_from_spec(spec)
            sys.modules["_xmhw_ref.%s" % name] = mod
            spec.loader.exec_module(mod)
            out.append(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _mods = (out[2], out[1])
    return _mods


def ref_mhw_filter(bthresh, minDuration=5, joinGaps=True, maxGap=2):
    """Reference `mhw_filter` (identify.py:415) on a boolean vector.
    Returns float arrays (start, end, events) of length T with NaN gaps."""
    identify, _ = load()
    T = len(bthresh)
    time = pd.date_range("2001-01-01", periods=T)
    b = pd.Series(np.asarray(bthresh, bool), index=time)
    idxarr = pd.Series(np.arange(T), index=time)
    df = identify.mhw_filter(b, idxarr, minDuration, joinGaps, maxGap)
    return df["start"].values, df["end"].values, df["events"].values


def ref_define_events(ts, thresh_t, seas_t, minDuration=5, joinGaps=True, maxGap=2,
                      want_inter=False):
    """Reference detect arithmetic for one cell (identify.py:372-399), given the
    per-timestep threshold and seasonal series already looked up by doy.

    ts is float32 (as xarray hands it over), thresh_t/seas_t float64.
    Returns a DataFrame with one row per event (all reference columns), or
    None when the cell has no event (the reference raises at features.py:157
    under pandas 3; SURVEY 8a-7 treats that as "no events").
    """
    identify, features = load()
    T = len(ts)
    time = pd.date_range("2001-01-01", periods=T)
    ts = np.asarray(ts, np.float32)
    df = pd.DataFrame({
        "ts": ts,
        "seas": np.asarray(seas_t, np.float64),
        "thresh": np.asarray(thresh_t, np.float64),
    }, index=pd.Index(time, name="time"))
    df["bthresh"] = df["ts"] > df["thresh"]      # identify.py:372
    df["doy"] = np.arange(T) % 366 + 1
    df["lat"] = 0.0
    df["lon"] = 0.0
    idxarr = pd.Series(np.arange(T), index=time)
    dfev = identify.mhw_filter(df.bthresh, idxarr, minDuration, joinGaps, maxGap)
    df = features.mhw_df(pd.concat([df, dfev], axis=1))
    if not df.events.notna().any():
        return (None, df) if want_inter else None
    try:
        out = features.mhw_features(df, T - 1, "time", ["lat", "lon"])
    except ValueError:
        return (None, df) if want_inter else None
    return (out, df) if want_inter else out
