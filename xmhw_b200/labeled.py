"""Minimal labelled-array containers.

xarray is the reference's data model but is not installed in this image.  The public
functions in xmhw_b200/xmhw.py accept real `xarray.DataArray`s (duck-typed: .dims,
.values, .coords, .attrs) AND these light containers, and return xarray objects when
xarray is importable, these containers otherwise.  They carry exactly what the hot path
needs (dims, coords, attrs) and deliberately implement no arithmetic.
"""
import numpy as np


class DataArray:
    def __init__(self, data, dims, coords=None, attrs=None, name=None, encoding=None):
        self.values = np.asarray(data)
        self.dims = tuple(dims)
        if self.values.ndim != len(self.dims):
            raise ValueError("dims do not match data rank")
        self.coords = {k: np.asarray(v) for k, v in (coords or {}).items()}
        self.attrs = dict(attrs or {})
        self.encoding = dict(encoding or {})
        self.name = name

    @property
    def shape(self):
        return self.values.shape

    @property
    def dtype(self):
        return self.values.dtype

    def __getitem__(self, key):
        if isinstance(key, str):
            return DataArray(self.coords[key], (key,), attrs={})
        return self.values[key]

    def __repr__(self):
        return "<xmhw_b200.DataArray %s %s %s>" % (self.name, dict(zip(self.dims, self.shape)), self.dtype)


class Dataset:
    def __init__(self, data_vars=None, coords=None, attrs=None):
        self.data_vars = dict(data_vars or {})
        self.coords = {k: np.asarray(v) for k, v in (coords or {}).items()}
        self.attrs = dict(attrs or {})
        self.coord_attrs = {}             # coordinate name -> attributes (identify.annotate_ds)

    def __getitem__(self, k):
        return self.data_vars[k]

    def __setitem__(self, k, v):
        # a variable carries the coordinates of its own dimensions (like xarray)
        if isinstance(v, DataArray):
            for d in v.dims:
                if d not in v.coords and d in self.coords and np.ndim(self.coords[d]) == 1:
                    v.coords[d] = self.coords[d]
        self.data_vars[k] = v

    def __contains__(self, k):
        return k in self.data_vars

    def __getattr__(self, k):
        dv = self.__dict__.get("data_vars", {})
        if k in dv:
            return dv[k]
        raise AttributeError(k)

    def keys(self):
        return self.data_vars.keys()

    def __repr__(self):
        return "<xmhw_b200.Dataset %s>" % ", ".join(
            "%s%s" % (k, v.dims) for k, v in self.data_vars.items())


def is_xarray(obj):
    return type(obj).__module__.split(".")[0] == "xarray"


def to_xarray(ds):
    """Convert a labeled.Dataset to xarray.Dataset (only when xarray is importable)."""
    import xarray as xr
    out = xr.Dataset()
    for k, v in ds.data_vars.items():
        out[k] = xr.DataArray(v.values, dims=v.dims, coords={d: ds.coords[d] for d in v.dims if d in ds.coords},
                              attrs=v.attrs, name=v.name)
    for k, v in ds.coords.items():
        if k not in out.coords and np.ndim(v) == 0:
            out = out.assign_coords({k: v})
    for k, a in getattr(ds, "coord_attrs", {}).items():
        if k in out.coords:
            out[k].attrs.update(a)
    out.attrs.update(ds.attrs)
    return out
