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


# ---------------------------------------------------------------------------
# Compressed store: zarr v2 (directory store) written and read with the standard library + numpy.
# The reference's docs recommend `comp = dict(zlib=True, complevel=5, shuffle=True, dtype='float32')`
# for every variable (docs/gettingstarted.rst:160-178); netCDF4 / h5py are not in this image, and
# zarr v2 is the other container xarray reads natively (`xr.open_zarr`): a directory with one JSON
# header per array (`.zarray`, `.zattrs` incl. xarray's `_ARRAY_DIMENSIONS`) and one zlib-deflated
# file per chunk.  float32 + zlib is applied as the docs ask; times are CF-encoded int64 so that
# xarray decodes them to datetime64.
# ---------------------------------------------------------------------------
def _zarr_dtype(a):
    return a.dtype.str if a.dtype.byteorder != "|" else a.dtype.str


def _write_zarr_array(root, name, arr, dims, attrs, level, chunk_elems):
    import json
    import os
    import zlib
    arr = np.ascontiguousarray(arr)
    d = os.path.join(root, name)
    os.makedirs(d, exist_ok=True)
    shape = list(arr.shape)
    if arr.ndim == 0:
        chunks = []
    else:
        rows = max(1, min(shape[0], chunk_elems // max(1, int(np.prod(shape[1:])))))
        chunks = [rows] + shape[1:]
    fill = None
    if arr.dtype.kind == "f":
        fill = "NaN"
    meta = {"zarr_format": 2, "shape": shape, "chunks": chunks if chunks else [], "dtype": arr.dtype.str,
            "compressor": {"id": "zlib", "level": int(level)}, "fill_value": fill, "order": "C", "filters": None}
    with open(os.path.join(d, ".zarray"), "w") as f:
        json.dump(meta, f)
    za = {"_ARRAY_DIMENSIONS": list(dims)}
    for k, v in attrs.items():
        v = _attr(v)
        za[k] = v.tolist() if isinstance(v, np.ndarray) else (v.item() if isinstance(v, np.generic) else v)
    with open(os.path.join(d, ".zattrs"), "w") as f:
        json.dump(za, f)
    if arr.ndim == 0:
        with open(os.path.join(d, "0"), "wb") as f:
            f.write(zlib.compress(arr.tobytes(), level))
        return
    nchunk = -(-shape[0] // chunks[0]) if shape[0] else 0
    for c in range(nchunk):
        blk = arr[c * chunks[0]:(c + 1) * chunks[0]]
        if blk.shape[0] < chunks[0]:                      # zarr stores full-size edge chunks
            pad = np.full([chunks[0] - blk.shape[0]] + shape[1:], np.nan if arr.dtype.kind == "f" else 0, arr.dtype)
            blk = np.concatenate([blk, pad])
        key = ".".join([str(c)] + ["0"] * (arr.ndim - 1))
        with open(os.path.join(d, key), "wb") as f:
            f.write(zlib.compress(np.ascontiguousarray(blk).tobytes(), level))


def save_zarr(ds, path, float32=True, level=5, chunk_elems=1 << 20):
    """Write a Dataset (labeled or xarray-like) as a zlib-compressed zarr v2 directory store
    (readable with `xarray.open_zarr(path, consolidated=False)`).  `float32=True` stores float64
    variables as float32, the encoding the reference's docs recommend; integer and time variables
    keep their width (datetime64 -> int64 "seconds since 1970-01-01")."""
    import json
    import os
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, ".zgroup"), "w") as f:
        json.dump({"zarr_format": 2}, f)
    gattrs = {}
    for k, v in dict(getattr(ds, "attrs", {})).items():
        v = _attr(v)
        gattrs[k] = v.tolist() if isinstance(v, np.ndarray) else (v.item() if isinstance(v, np.generic) else v)
    with open(os.path.join(path, ".zattrs"), "w") as f:
        json.dump(gattrs, f)

    def encode(a, is_coord):
        a = np.asarray(a)
        extra = {}
        if np.issubdtype(a.dtype, np.datetime64):
            secs = (a.astype("datetime64[s]") - _EPOCH).astype(np.int64)
            secs = np.where(np.isnat(a), np.iinfo(np.int64).min, secs)
            return secs, {"units": "seconds since 1970-01-01 00:00:00", "calendar": "proleptic_gregorian",
                          "_FillValue": int(np.iinfo(np.int64).min)}
        if a.dtype == np.bool_:
            return a.astype(np.int8), extra
        if a.dtype == np.float64 and float32 and not is_coord:
            return a.astype(np.float32), extra
        if a.dtype.kind in "iuf":
            return a, extra
        raise TypeError("cannot store dtype %s" % a.dtype)

    cattrs = dict(getattr(ds, "coord_attrs", {}))
    for name, c in dict(ds.coords).items():
        arr, extra = encode(getattr(c, "values", c), True)
        at = dict(cattrs.get(name, {}))
        at.update(extra)
        _write_zarr_array(path, name, arr, (name,) if arr.ndim == 1 else (), at, level, chunk_elems)
    for name, v in dict(ds.data_vars).items():
        arr, extra = encode(v.values, False)
        at = dict(getattr(v, "attrs", {}))
        at.update(extra)
        _write_zarr_array(path, name, arr, tuple(v.dims), at, level, chunk_elems)


def load_zarr(path):
    """Read a store written by save_zarr back into a labeled.Dataset (CF times are decoded)."""
    import json
    import os
    import zlib
    coords, data_vars = {}, {}
    gattrs = json.load(open(os.path.join(path, ".zattrs"))) if os.path.isfile(os.path.join(path, ".zattrs")) else {}
    cattrs = {}
    for name in sorted(os.listdir(path)):
        d = os.path.join(path, name)
        if not os.path.isfile(os.path.join(d, ".zarray")):
            continue
        meta = json.load(open(os.path.join(d, ".zarray")))
        attrs = json.load(open(os.path.join(d, ".zattrs")))
        dims = tuple(attrs.pop("_ARRAY_DIMENSIONS", ()))
        dt = np.dtype(meta["dtype"])
        shape, chunks = meta["shape"], meta["chunks"]
        if not shape:
            arr = np.frombuffer(zlib.decompress(open(os.path.join(d, "0"), "rb").read()), dt).reshape(())
        else:
            parts = []
            for c in range(-(-shape[0] // chunks[0]) if shape[0] else 0):
                key = ".".join([str(c)] + ["0"] * (len(shape) - 1))
                parts.append(np.frombuffer(zlib.decompress(open(os.path.join(d, key), "rb").read()), dt).reshape(chunks))
            arr = np.concatenate(parts)[:shape[0]] if parts else np.zeros(shape, dt)
        if str(attrs.get("units", "")).startswith("seconds since 1970-01-01"):
            fillv = attrs.pop("_FillValue", None)
            t = _EPOCH + arr.astype("timedelta64[s]")
            if fillv is not None:
                t = np.where(arr == fillv, np.datetime64("NaT"), t)
            arr = t
            attrs.pop("units", None)
            attrs.pop("calendar", None)
        if len(dims) <= 1 and (dims == (name,) or dims == ()):
            coords[name] = arr
            if attrs:
                cattrs[name] = attrs
        else:
            data_vars[name] = labeled.DataArray(arr, dims, attrs=attrs, name=name)
    out = labeled.Dataset(coords=coords, attrs=gattrs)
    out.coord_attrs.update(cattrs)
    for k, v in data_vars.items():
        out[k] = v
    return out
