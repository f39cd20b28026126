"""Saving results (the caller-side step right after the hot path).

The reference hands back xarray Datasets and its docs save them with `to_netcdf`
(docs/gettingstarted.rst:153-178, float32 + zlib encoding).  xarray / netCDF4 are not part of
this image, so this module writes what `xmhw_b200.xmhw.threshold` / `detect` return -- a
`labeled.Dataset` (or anything with the same `.data_vars/.coords/.attrs` shape, e.g. an
xarray Dataset) -- as NetCDF-3 64-bit-offset files with `scipy.io.netcdf_file`, which
`xarray.open_dataset` reads back.  NetCDF-3 has no compression, no int64 and no datetime64:

* int64 / bool variables become int32 (range checked), datetime64 becomes float64
  "days since 1970-01-01" with CF `units`/`calendar` attributes;
* `float32=True` stores float64 statistics as float32 like the reference's recommended encoding
  (halves the file; off by default);
* the compact event table of `detect(..., compact=True)` (one row per event, global cell index)
  is the form to save at global scale -- the dense (events, lat, lon) cube of the reference is
  2.8 TB there.
"""
import numpy as np

from . import labeled

_EPOCH = np.datetime64("1970-01-01T00:00:00")


def _plain(a, float32):
    """-> (array NetCDF-3 can hold, extra attributes)."""
    a = np.asarray(a)
    if np.issubdtype(a.dtype, np.datetime64):
        days = (a.astype("datetime64[s]") - _EPOCH).astype(np.float64) / 86400.0
        days = np.where(np.isnat(a), np.nan, days)
        return days, {"units": "days since 1970-01-01 00:00:00", "calendar": "proleptic_gregorian"}
    if a.dtype == np.bool_:
        return a.astype(np.int8), {}
    if np.issubdtype(a.dtype, np.integer) and a.dtype.itemsize > 4:
        if a.size and (a.max() > np.iinfo(np.int32).max or a.min() < np.iinfo(np.int32).min):
            return a.astype(np.float64), {}
        return a.astype(np.int32), {}
    if np.issubdtype(a.dtype, np.unsignedinteger):
        return a.astype(np.int32), {}
    if a.dtype == np.float64 and float32:
        return a.astype(np.float32), {}
    if a.dtype.kind in "iuf":
        return a, {}
    raise TypeError("cannot store dtype %s in NetCDF-3" % a.dtype)


def _attr(v):
    if isinstance(v, (bool, np.bool_)):
        return int(v)
    if isinstance(v, (list, tuple)):
        return np.asarray(v)
    return v


def save_dataset(ds, path, float32=False):
    """Write a Dataset (labeled or xarray-like) to a NetCDF-3 64-bit-offset file."""
    from scipy.io import netcdf_file
    data_vars = dict(ds.data_vars)
    coords = {k: np.asarray(getattr(v, "values", v)) for k, v in dict(ds.coords).items()}
    sizes = {}
    for name, v in data_vars.items():
        for d, n in zip(v.dims, np.shape(v.values)):
            if sizes.setdefault(d, n) != n:
                raise ValueError("dimension %r has two lengths" % d)
    for d, c in coords.items():
        if c.ndim == 1:
            if sizes.setdefault(d, c.shape[0]) != c.shape[0]:
                raise ValueError("coordinate %r does not match its dimension" % d)
    with netcdf_file(path, "w", version=2) as f:
        for d, n in sizes.items():
            f.createDimension(d, int(n))
        for k, v in dict(getattr(ds, "attrs", {})).items():
            setattr(f, k, _attr(v))
        scalars = []
        for d, c in coords.items():
            if c.ndim == 0:
                scalars.append((d, c))
                continue
            arr, extra = _plain(c, False)
            var = f.createVariable(d, arr.dtype, (d,))
            var[:] = arr
            for k, a in extra.items():
                setattr(var, k, a)
        for d, c in scalars:                       # scalar coordinates (e.g. `quantile`) as global attributes
            setattr(f, "coord_" + d, float(c) if c.dtype.kind == "f" else int(c))
        for name, v in data_vars.items():
            arr, extra = _plain(v.values, float32)
            var = f.createVariable(name, arr.dtype, tuple(v.dims))
            var[...] = arr
            for k, a in dict(getattr(v, "attrs", {})).items():
                setattr(var, k, _attr(a))
            for k, a in extra.items():
                setattr(var, k, a)


def load_dataset(path):
    """Read a file written by save_dataset back into a labeled.Dataset (times stay numeric)."""
    from scipy.io import netcdf_file
    with netcdf_file(path, "r", mmap=False) as f:
        dims = set(f.dimensions)
        coords, data_vars = {}, {}
        for name, var in f.variables.items():
            arr = np.array(var[...])
            arr = arr.astype(arr.dtype.newbyteorder("="))       # NetCDF-3 is big-endian on disk
            attrs = {k: (v.decode() if isinstance(v, bytes) else v) for k, v in var._attributes.items()}
            if name in dims and arr.ndim == 1:
                coords[name] = arr
            else:
                data_vars[name] = labeled.DataArray(arr, var.dimensions, attrs=attrs, name=name)
        gattrs = {k: (v.decode() if isinstance(v, bytes) else v) for k, v in f._attributes.items()}
    out = labeled.Dataset(coords=coords, attrs=gattrs)
    for k, v in data_vars.items():
        out[k] = v
    return out
