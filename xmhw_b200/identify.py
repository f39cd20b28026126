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


def add_doy(time_values, keep_tstep=False):
    """identify.py:28-79: 1-based day-of-year labels on a 366-day calendar (Feb 29 = 60
    exists only in leap years), or 1..steps-per-year tiled when keep_tstep.
    Returns (doy int64[T], ndoy)."""
    years, months, dayofyear = _ymd(time_values)
    T = len(years)
    if keep_tstep:
        uy = np.unique(years)
        ref_year = uy[1] if len(uy) > 1 else uy[0]
        steps = int(np.sum(years == ref_year))                       # identify.py:59-60
        if steps == 0 or T % steps != 0:
            raise XmhwException("To use original timestep as climatology base unit, "
                                "timeseries has to have complete years")
        return np.tile(np.arange(1, steps + 1, dtype=np.int64), T // steps), steps
    leap = (years % 4 == 0) & ((years % 100 != 0) | (years % 400 == 0))
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


def annotate_ds(attrs_out, ds_attrs, kind):
    """identify.py:539-696 (condensed): provenance/CF global attributes."""
    attrs_out["source"] = "xmhw_b200 (B200-native implementation of the xmhw hot path)"
    attrs_out["title"] = ("Seasonal climatology and threshold calculated to detect marine heatwaves"
                          if kind == "clim" else "Marine heatwave events")
    attrs_out["history"] = "%s: calculated using xmhw_b200" % date.today().strftime("%Y-%m-%d")
    units = ds_attrs.get("ts", {}).get("units", "degree_C")
    attrs_out["units"] = units
    return attrs_out
