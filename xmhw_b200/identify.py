"""Host-side helpers of the hot path that stay in Python, named after the reference's
xmhw/identify.py functions they replace (calendar, day-of-year labels, land mask,
attributes).  The per-cell operators of that file (window_roll, calculate_thresh,
calculate_seas, runavg, define_events, mhw_filter, join_gaps) are CUDA kernels reached
through xmhw_b200.core."""
from datetime import date

import numpy as np

from .exception import XmhwException

NDAYS = {"standard": 365.25, "gregorian": 365.25, "proleptic_gregorian": 365.25, "all_leap": 366,
         "noleap": 365, "365_day": 365, "360_day": 360, "julian": 365.25}


def get_calendar(time_values, encoding=None, attrs=None):
    """identify.py:82-134: days per year from the time axis' calendar attribute."""
    calendar = ""
    if encoding and "calendar" in encoding:
        calendar = encoding["calendar"]
    elif attrs and "calendar" in attrs:
        calendar = attrs["calendar"]
    else:
        calendar = getattr(np.asarray(time_values).ravel()[0], "calendar", "")
    if calendar in ("360", "365", "366"):
        calendar = "%s_day" % calendar
    elif calendar == "leap":
        calendar = "standard"
    return NDAYS.get(calendar, 365.25)


def _ymd(time_values):
    t = np.asarray(time_values)
    if np.issubdtype(t.dtype, np.datetime64):
        d = t.astype("datetime64[D]")
        years = d.astype("datetime64[Y]").astype(np.int64) + 1970
        months = d.astype("datetime64[M]").astype(np.int64) % 12 + 1
        dayofyear = (d - d.astype("datetime64[Y]").astype("datetime64[D]")).astype(np.int64) + 1
        return years, months, dayofyear
    # cftime-like objects
    years = np.array([x.year for x in t], np.int64)
    months = np.array([x.month for x in t], np.int64)
    dayofyear = np.array([getattr(x, "dayofyr", None) or x.timetuple().tm_yday for x in t], np.int64)
    return years, months, dayofyear


def _is_leap(years, calendar):
    """`t.dt.is_leap_year` of the reference (identify.py:73) for the calendar of the time axis: cftime
    calendars without leap years never have one, all_leap always, julian every fourth year."""
    if calendar in ("noleap", "365_day", "360_day"):
        return np.zeros(len(years), bool)
    if calendar in ("all_leap", "366_day"):
        return np.ones(len(years), bool)
    if calendar == "julian":
        return years % 4 == 0
    return (years % 4 == 0) & ((years % 100 != 0) | (years % 400 == 0))


def add_doy(time_values, keep_tstep=False, calendar=None):
    """identify.py:28-79: 1-based day-of-year labels on a 366-day calendar (Feb 29 = 60
    exists only in leap years), or 1..steps-per-year tiled when keep_tstep.
    `calendar`: CF calendar name; default = the `.calendar` of the time objects (cftime), else standard.
    Returns (doy int64[T], ndoy)."""
    years, months, dayofyear = _ymd(time_values)
    if calendar is None:
        t = np.asarray(time_values).ravel()
        calendar = getattr(t[0], "calendar", "standard") if len(t) else "standard"
    T = len(years)
    if keep_tstep:
        uy = np.unique(years)
        ref_year = uy[1] if len(uy) > 1 else uy[0]
        steps = int(np.sum(years == ref_year))                       # identify.py:59-60
        if steps == 0 or T % steps != 0:
            raise XmhwException("To use original timestep as climatology base unit, "
                                "timeseries has to have complete years")
        return np.tile(np.arange(1, steps + 1, dtype=np.int64), T // steps), steps
    leap = _is_leap(years, calendar)
    return dayofyear + ((~leap) & (months >= 3)).astype(np.int64), 366  # identify.py:73-76


def land_check_shape(shape, dims, tdim):
    """identify.py:504-516: argument checks of land_check (raise before any launch)."""
    other = [d for d in dims if d != tdim]
    if len(other) == 0:
        raise XmhwException("Series has only time dimension use point=True option, exiting")
    for d, n in zip(dims, shape):
        if d != tdim and n == 0:
            raise XmhwException("Dimension %s has 0 lenght, exiting" % d)
    return sorted(other)


GITHUB = "https://github.com/coecms/xmhw"

# Per-variable CF attributes of the detect output (reference identify.py:596-684), as data:
# name -> (long_name, units) with units "T" = series units, "T day", "T day-1" or "1"; None = not set.
# The strings are the reference's own, byte for byte (including its spelling), so that files written
# from either implementation carry identical metadata.
_REL = {"": "relative to seasonal climatology", "_relThresh": "relative to threshold", "_abs": "absolute magnitude"}
MHW_VARIABLE_ATTRS = {
    "event": ("MHW event identifier: starting index", "1"),
    "duration": ("MHW duration in number of days", "1"),
    "rate_onset": ("MHW onset rate", "T day-1"),
    "rate_decline": ("MHW decline rate", "T day-1"),
    "category": ("MHW category based on peak intensity: 1: Moderate, 2: Strong, 3: Severe or 4: Extreme", None),
}
for _sfx, _rel in _REL.items():
    MHW_VARIABLE_ATTRS["intensity_max" + _sfx] = ("MHW maximum (peak) intensity " + _rel, "T")
    MHW_VARIABLE_ATTRS["intensity_mean" + _sfx] = ("MHW mean intensity " + _rel, "T")
    MHW_VARIABLE_ATTRS["intensity_var" + _sfx] = ("MHW intensity variability " + _rel, "T")
    MHW_VARIABLE_ATTRS["intensity_cumulative" + _sfx] = ("MHW cumulative intensity " + _rel, "T day")
MHW_VARIABLE_ATTRS["intensity_var_abs"] = ("MHW intensity variability abosulute magnitude", "T")     # sic, identify.py:657
for _k, _what in (("max", "maximum (peak)"), ("mean", "mean"), ("var", None), ("cumulative", "cumulative")):
    MHW_VARIABLE_ATTRS["severity_" + _k] = (
        "MHW severity variability relative to seasonal climatology" if _k == "var"
        else "MHW %s severity relative to seasonal climatology" % _what, "T day" if _k == "cumulative" else "T")
for _c in ("moderate", "strong", "severe", "extreme"):
    MHW_VARIABLE_ATTRS["duration_" + _c] = ("Number of days falling in category " + _c.capitalize(), "1")


def annotate_ds(ds, ds_attrs, kind):
    """identify.py:539-696: CF / provenance attributes of the output dataset (`kind` = "clim" or "mhw").

    `ds` is a labeled.Dataset (variables carry `.attrs`, coordinates `ds.coord_attrs`); `ds_attrs` maps
    "ts" and every coordinate name to the attributes of the input series.  Like the reference the units
    of the series are looked up under the key "temp" (identify.py:554), which its callers never set
    (xmhw.py:129, :389 store "ts"), so the units are always "degree_C" -- kept for identical output.
    """
    try:
        uts = ds_attrs["temp"]["units"]
        if any(t in uts for t in ("Celsius", "celsius")):
            uts = "degree_C"
    except Exception:
        uts = "degree_C"
    for c in list(ds.coords):                                   # identify.py:560-578
        if c == "doy":
            ds.coord_attrs[c] = {"units": "1", "long_name": "Day of the year"}
        elif c == "events":
            ds.coord_attrs[c] = {"units": "1", "long_name": "MHW event identifier: starting index"}
        elif c != "point" and isinstance(ds_attrs.get(c), dict):
            ds.coord_attrs.setdefault(c, {}).update(ds_attrs[c])
    ds.attrs["source"] = f"xmhw code: {GITHUB}"
    if kind == "clim":                                          # identify.py:580-594
        ds.attrs["title"] = ("Seasonal climatology and threshold " + "calculated to detect marine heatwaves following the "
                             + " Hobday et al. (2016) definition")
        for v in ("thresh", "seas"):
            if v in ds:
                ds[v].attrs["units"] = uts
    else:                                                       # identify.py:596-694
        for name, (long_name, units) in MHW_VARIABLE_ATTRS.items():
            if name not in ds:
                continue
            if units is not None:
                ds[name].attrs["units"] = units.replace("T", uts)
            ds[name].attrs["long_name"] = long_name
        ds.attrs["title"] = ("Marine heatwave events identified "
                             + "applying the Hobday et al. (2016) marine heat wave definition")
    ds.attrs["history"] = f"{date.today()}: calculated using xmhw code {GITHUB}"
    return ds
