"""GPU tests of the drop-in API (xmhw_b200.xmhw.threshold / detect) on labelled arrays:
reference test cube end to end against the oracle and the golden climatology files."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def api():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from xmhw_b200 import labeled, xmhw
    return xmhw, labeled


def test_oisst_end_to_end(api, oisst, clim_gold):
    xmhw, labeled = api
    from oracle import xmhw_oracle as O
    da = labeled.DataArray(oisst["sst"], ("time", "lat", "lon"),
                           {"time": oisst["time"], "lat": oisst["lat"], "lon": oisst["lon"]})
    clim = xmhw.threshold(da)
    th, se = clim["thresh"], clim["seas"]
    assert th.dims == ("doy", "lat", "lon") and th.values.dtype == np.float64
    ocean = ~np.isnan(oisst["sst"]).all(0)
    assert th.shape == (366, int(ocean.any(1).sum()), int(ocean.any(0).sum()))
    # golden points [1,2] and [5,3] (test_xmhw.py:24-66) -> positions in the land-trimmed grid
    lat_keep, lon_keep = np.nonzero(ocean.any(1))[0], np.nonzero(ocean.any(0))[0]
    smooth, _ = clim_gold
    for (iy, ix), n in (((1, 2), "1"), ((5, 3), "2")):
        jy, jx = int(np.searchsorted(lat_keep, iy)), int(np.searchsorted(lon_keep, ix))
        np.testing.assert_almost_equal(th.values[82:, jy, jx], smooth["thresh" + n][82:], decimal=6)
        np.testing.assert_almost_equal(se.values[82:, jy, jx], smooth["seas" + n][82:], decimal=4)
    mhw = xmhw.detect(da, th, se)
    assert mhw["intensity_max"].dims == ("events", "lat", "lon")
    doy = O.add_doy(oisst["time"])
    flat = oisst["sst"].reshape(len(doy), -1)
    oth, ose = O.threshold(flat, doy, 366)
    exp = O.detect(flat, doy, oth, ose)
    assert int(np.sum(~np.isnan(mhw["event"].values))) == len(exp["cell"])
    jy, jx = int(np.searchsorted(lat_keep, 1)), int(np.searchsorted(lon_keep, 2))
    col = mhw["index_start"].values[:, jy, jx]
    assert col[~np.isnan(col)].tolist() == [1, 75, 99, 114, 138, 172, 225, 613, 709]
    assert mhw["time_start"].values.dtype.kind == "M"
    comp = xmhw.detect(da, th, se, compact=True)
    assert len(comp["event"].values) == len(exp["cell"])
    np.testing.assert_allclose(np.sort(comp["intensity_cumulative"].values), np.sort(exp["intensity_cumulative"]),
                               rtol=1e-9)
    # the returned Datasets go to NetCDF-3 files and back (xmhw_b200/io.py)
    import os
    import tempfile
    from xmhw_b200 import io
    with tempfile.TemporaryDirectory() as tmp:
        for name, ds in (("clim", clim), ("events", comp), ("dense", mhw)):
            io.save_dataset(ds, os.path.join(tmp, name + ".nc"))
        back = io.load_dataset(os.path.join(tmp, "events.nc"))
        assert np.array_equal(back["index_start"].values, comp["index_start"].values)
        assert np.array_equal(back["intensity_max"].values, comp["intensity_max"].values)
        assert np.array_equal(io.load_dataset(os.path.join(tmp, "clim.nc"))["thresh"].values, th.values, equal_nan=True)


def test_point_series_and_cold_spells(api, oisst):
    """BASELINE config 1: single-point series (xmhw.py:122-126 point path) + coldSpells sign rules."""
    xmhw, labeled = api
    from oracle import xmhw_oracle as O
    ts = oisst["sst"][:, 1, 2]
    da = labeled.DataArray(ts, ("time",), {"time": oisst["time"]})
    clim = xmhw.threshold(da)
    doy = O.add_doy(oisst["time"])
    oth, ose = O.threshold(ts, doy, 366)
    assert np.array_equal(clim["thresh"].values, oth[:, 0]) and clim["thresh"].dims == ("doy",)
    mhw = xmhw.detect(da, clim["thresh"], clim["seas"])
    assert mhw["index_start"].values.tolist() == [1, 75, 99, 114, 138, 172, 225, 613, 709]
    cold = xmhw.threshold(da, coldSpells=True, pctile=90)
    oth_c, ose_c = O.threshold(-ts, doy, 366)
    assert np.array_equal(cold["thresh"].values, oth_c[:, 0])          # returned for the flipped series
    mc = xmhw.detect(da, cold["thresh"], cold["seas"], coldSpells=True)
    exp = O.detect(-ts, doy, oth_c, ose_c)
    assert mc["index_start"].values.tolist() == exp["index_start"].tolist()
    np.testing.assert_allclose(mc["intensity_max"].values, -exp["intensity_max"])     # flipped back
    np.testing.assert_allclose(mc["intensity_var"].values, exp["intensity_var"])      # "_var" not flipped
    np.testing.assert_allclose(mc["rate_onset"].values, exp["rate_onset"])


def test_climatology_period_and_anynans(api, oisst):
    xmhw, labeled = api
    from oracle import xmhw_oracle as O
    sst = oisst["sst"].copy()
    sst[5, 1, 2] = np.nan                 # one NaN -> the cell is dropped with anynans=True
    da = labeled.DataArray(sst, ("time", "lat", "lon"), {"time": oisst["time"], "lat": oisst["lat"], "lon": oisst["lon"]})
    clim = xmhw.threshold(da, anynans=True, climatologyPeriod=[2003, 2003], smoothPercentile=False)
    sel = oisst["time"] < np.datetime64("2004-01-01")
    doy = O.add_doy(oisst["time"][sel])
    flat = sst[sel].reshape(int(sel.sum()), -1)
    keep = ~np.isnan(flat).any(0)
    oth, _ = O.threshold(flat, doy, 366, smoothPercentile=False)
    got = clim["thresh"].values
    assert got.shape[0] == 365                     # doy 60 has no sample in 2003: dropped like the groupby does
    assert int(np.sum(~np.isnan(got[0]))) == int(keep.sum())


def test_intermediate_and_maxpad_api(api, oisst):
    """detect(..., intermediate=True) returns (mhw, mhw_inter) like xmhw.py:516-518; maxPadLength
    fills short gaps before detection (xmhw.py:409-410)."""
    xmhw, labeled = api
    from oracle import xmhw_oracle as O
    ts = oisst["sst"][:, 1, 2].copy()
    da = labeled.DataArray(ts, ("time",), {"time": oisst["time"]})
    clim = xmhw.threshold(da)
    mhw, inter = xmhw.detect(da, clim["thresh"], clim["seas"], intermediate=True)
    doy = O.add_doy(oisst["time"])
    exp = O.intermediate(ts, doy, clim["thresh"].values, clim["seas"].values)
    assert inter["events"].dims == ("time",)
    assert np.array_equal(inter["events"].values, exp["events"][:, 0], equal_nan=True)
    assert np.allclose(inter["relSeas"].values, exp["relSeas"][:, 0], equal_nan=True)
    assert np.array_equal(inter["duration_moderate"].values, exp["duration_moderate"][:, 0])
    sst = oisst["sst"].copy()
    sst[100:103, 1, 2] = np.nan                        # 3-day gap inside the first-year series: filled
    sst[300:305, 5, 3] = np.nan                        # 5-day gap: its bounding samples are 6 days apart, NOT filled
    dg = labeled.DataArray(sst, ("time", "lat", "lon"), {"time": oisst["time"], "lat": oisst["lat"], "lon": oisst["lon"]})
    c1 = xmhw.threshold(dg, maxPadLength=5)
    c2 = xmhw.threshold(dg, maxPadLength=np.timedelta64(5, "D"))       # what xarray wants on a datetime axis
    assert np.array_equal(c1["thresh"].values, c2["thresh"].values, equal_nan=True)
    # interpolate_na(max_gap=5): a run of k NaNs is filled iff k + 1 <= 5 (xarray measures coordinate distance)
    filled = O.interp_gaps(sst.reshape(len(doy), -1), 4)
    assert np.isnan(filled[300:305, 5 * 4 + 3]).all() and not np.isnan(filled[100:103, 1 * 4 + 2]).any()
    oth, _ = O.threshold(filled, doy, 366)
    ocean = ~np.isnan(oisst["sst"]).all(0)
    assert np.array_equal(c1["thresh"].values.reshape(366, -1),
                          oth.reshape(366, 8, 4)[:, ocean.any(1)][:, :, ocean.any(0)].reshape(366, -1), equal_nan=True)


def test_plan_limits_surface_as_xmhw_exception(api):
    """A window the climatology plans cannot express (wider than a year) leaves threshold() as the
    reference's exception type, not as a raw internal error."""
    xmhw, labeled = api
    from xmhw_b200.exception import XmhwException
    t = np.arange(np.datetime64("2001-01-01"), np.datetime64("2004-01-01"))[:36 * 30:30]      # 36 monthly-ish steps
    da = labeled.DataArray(np.random.default_rng(0).normal(15, 1, (36, 2, 2)).astype(np.float32),
                           ("time", "lat", "lon"), {"time": t, "lat": np.arange(2.0), "lon": np.arange(2.0)})
    with pytest.raises(XmhwException):
        xmhw.threshold(da, tstep=True, windowHalfWidth=7, smoothPercentileWidth=3)


def test_attributes_anynans_coldspells_through_pipeline(api, oisst):
    """The public functions run through the pipelined host path (core.host_pipeline): CF attributes of
    annotate_ds, anynans dropping cells BEFORE the gap interpolation like the reference (xmhw.py:138 vs
    :159-160), coldSpells sign handling."""
    xmhw, labeled = api
    from oracle import xmhw_oracle as O
    sst = oisst["sst"].copy()
    sst[50, 1, 2] = np.nan                                   # one missing day in an ocean cell
    da = labeled.DataArray(sst, ("time", "lat", "lon"), {"time": oisst["time"], "lat": oisst["lat"], "lon": oisst["lon"]},
                           attrs={"units": "degree_C"})
    da.coord_attrs = {"lat": {"units": "degrees_north"}}
    clim = xmhw.threshold(da, anynans=True, maxPadLength=3)
    assert clim["thresh"].attrs["units"] == "degree_C" and clim.coord_attrs["lat"] == {"units": "degrees_north"}
    assert clim.attrs["source"].endswith("github.com/coecms/xmhw")
    doy = O.add_doy(oisst["time"])
    flat = sst.reshape(len(doy), -1)
    keep = ~np.isnan(flat).any(0)
    assert keep.sum() < (~np.isnan(flat).all(0)).sum()       # the cell with the single NaN is dropped
    oth, _ = O.threshold(np.where(keep[None], flat, np.nan), doy, 366)
    got = clim["thresh"].values
    exp = oth.reshape(366, 8, 4)[:, keep.reshape(8, 4).any(1)][:, :, keep.reshape(8, 4).any(0)]
    assert np.array_equal(got, exp, equal_nan=True)
    cold = xmhw.threshold(da, coldSpells=True, pctile=90)
    oth_c, _ = O.threshold(-flat, doy, 366)
    ocean = ~np.isnan(flat).all(0)
    exp_c = oth_c.reshape(366, 8, 4)[:, ocean.reshape(8, 4).any(1)][:, :, ocean.reshape(8, 4).any(0)]
    assert np.array_equal(cold["thresh"].values, exp_c, equal_nan=True)
    mhw = xmhw.detect(da, cold["thresh"], cold["seas"], coldSpells=True, compact=True)
    assert (mhw["intensity_max"].values < 0).all()           # flip_cold: intensities negated back (features.py:298-315)
    assert mhw["intensity_max"].attrs["long_name"].startswith("MHW maximum (peak) intensity")


def test_real_xarray_dataarray_when_available(api, oisst):
    """A real xarray.DataArray goes in and real xarray Datasets come out (skipped where xarray is absent)."""
    xr = pytest.importorskip("xarray")
    xmhw, labeled = api
    da = xr.DataArray(oisst["sst"], dims=("time", "lat", "lon"),
                      coords={"time": oisst["time"], "lat": oisst["lat"], "lon": oisst["lon"]}, attrs={"units": "degC"})
    da["lat"].attrs["units"] = "degrees_north"
    clim = xmhw.threshold(da)
    assert isinstance(clim, xr.Dataset) and clim["thresh"].dims == ("doy", "lat", "lon")
    assert clim["lat"].attrs["units"] == "degrees_north" and clim["doy"].attrs["long_name"] == "Day of the year"
    mhw = xmhw.detect(da, clim["thresh"], clim["seas"])
    assert isinstance(mhw, xr.Dataset) and "events" in mhw.dims
    ref = xmhw.threshold(labeled.DataArray(oisst["sst"], ("time", "lat", "lon"),
                                           {"time": oisst["time"], "lat": oisst["lat"], "lon": oisst["lon"]}))
    assert np.array_equal(clim["thresh"].values, ref["thresh"].values, equal_nan=True)


def test_noleap_cftime_axis_through_api(api):
    """A noleap (365-day) model calendar as xarray decodes it (cftime-like objects): label 60 never occurs,
    so the climatology has 365 doys (identify.py:233: the empty group vanishes); events equal the oracle's
    on the same labels.  With tstep=True the 365 steps are the labels themselves."""
    xmhw, labeled = api
    from oracle import xmhw_oracle as O
    from tests.test_api_cpu import _cf_axis
    from xmhw_b200 import identify, synth
    t = _cf_axis("noleap", list(range(2001, 2011)))
    ts = synth.synth_sst(len(t), 12, synth.season_table(len(t)), nan_ppm=3000).reshape(len(t), 3, 4)
    da = labeled.DataArray(ts, ("time", "lat", "lon"), {"time": t, "lat": np.arange(3.0), "lon": np.arange(4.0)})
    clim = xmhw.threshold(da)
    assert clim["thresh"].shape == (365, 3, 4) and 60 not in clim["thresh"].coords["doy"].tolist()
    doy, ndoy = identify.add_doy(t)
    flat = ts.reshape(len(t), -1)
    oth, ose = O.threshold(flat, doy, ndoy)
    present = ~np.isnan(oth).all(1)
    assert present.sum() == 365
    assert np.array_equal(clim["thresh"].values.reshape(365, -1), oth[present])
    ev = xmhw.detect(da, clim["thresh"], clim["seas"], compact=True)
    exp = O.detect(flat, doy, oth, ose)
    assert len(exp["cell"]) > 20 and np.array_equal(ev["index_start"].values, exp["index_start"].astype(np.float64))
    assert ev["time_start"].values[0] is t[int(exp["index_start"][0])]        # cftime objects pass through
    clim_t = xmhw.threshold(da, tstep=True)
    assert clim_t["thresh"].shape == (365, 3, 4)
    oth_t, _ = O.threshold(flat, np.tile(np.arange(1, 366), 10), 365, tstep=True)
    assert np.array_equal(clim_t["thresh"].values.reshape(365, -1), oth_t)
