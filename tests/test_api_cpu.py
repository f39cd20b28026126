"""CPU-side tests of the public API layer: argument validation raises the reference's
XmhwException for the reference's conditions BEFORE any launch, calendar / doy helpers,
labelled containers, synthetic generator determinism."""
import numpy as np
import pytest

from oracle import xmhw_oracle as O
from xmhw_b200 import identify, labeled, synth
from xmhw_b200.exception import XmhwException
from xmhw_b200.xmhw import detect, threshold


def _cube(T=40, ny=3, nx=4):
    time = np.arange(np.datetime64("2001-01-01"), np.datetime64("2001-01-01") + np.timedelta64(T, "D"))
    data = np.random.default_rng(0).normal(15, 1, (T, ny, nx)).astype(np.float32)
    return labeled.DataArray(data, ("time", "lat", "lon"),
                             {"time": time, "lat": np.arange(ny) * 0.25, "lon": np.arange(nx) * 0.25})


def test_threshold_validation():
    da = _cube()
    with pytest.raises(XmhwException):            # xmhw.py:103-104
        threshold(da, smoothPercentileWidth=30)
    with pytest.raises(XmhwException):            # xmhw.py:105-109
        threshold(da, tdim="t")
    land = labeled.DataArray(np.full((40, 3, 4), np.nan, np.float32), da.dims, da.coords)
    with pytest.raises(XmhwException):            # identify.py:527-528
        threshold(land)
    empty = labeled.DataArray(np.zeros((40, 0, 4), np.float32), da.dims,
                              {"time": da.coords["time"], "lat": np.zeros(0), "lon": da.coords["lon"]})
    with pytest.raises(XmhwException):            # identify.py:514-516
        threshold(empty)
    with pytest.raises(XmhwException):            # identify.py:61-66 incomplete years with tstep
        threshold(_cube(T=400), tstep=True)


def test_detect_validation():
    da = _cube()
    th = labeled.DataArray(np.zeros((366, 3, 4)), ("doy", "lat", "lon"),
                           {"doy": np.arange(1, 367), "lat": da.coords["lat"], "lon": da.coords["lon"]})
    with pytest.raises(XmhwException):            # xmhw.py:373-377
        detect(da, th, th, maxGap=5, minDuration=5)
    with pytest.raises(XmhwException):
        detect(da, th, th, tdim="t")


def test_add_doy_and_calendar(oisst):
    doy, ndoy = identify.add_doy(oisst["time"])
    assert ndoy == 366 and np.array_equal(doy, O.add_doy(oisst["time"]))
    t5 = np.arange(np.datetime64("2001-01-01"), np.datetime64("2003-01-01"), np.timedelta64(5, "D"))[:146]
    doy5, n5 = identify.add_doy(t5, keep_tstep=True)
    assert n5 == 73 and np.array_equal(doy5, np.tile(np.arange(1, 74), 2))
    assert identify.get_calendar(t5, {"calendar": "360_day"}) == 360
    assert identify.get_calendar(t5, None, {"calendar": "noleap"}) == 365
    assert identify.get_calendar(t5) == 365.25
    assert identify.get_calendar(t5, {"calendar": "360"}) == 360      # identify.py:125-126


class _CfDate:
    """What xarray hands over for a non-standard calendar: cftime-like objects (year, month, day,
    dayofyr, calendar) in an object array -- cftime itself is not part of this image."""

    def __init__(self, year, month, day, dayofyr, calendar):
        self.year, self.month, self.day, self.dayofyr, self.calendar = year, month, day, dayofyr, calendar


def _cf_axis(calendar, years):
    mlen = {"noleap": [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31], "all_leap": [31, 29, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31],
            "360_day": [30] * 12}[calendar]
    out = []
    for y in years:
        n = 0
        for m, ml in enumerate(mlen, 1):
            for d in range(1, ml + 1):
                n += 1
                out.append(_CfDate(y, m, d, n, calendar))
    return np.array(out, dtype=object)


def test_add_doy_cftime_calendars():
    """Non-Gregorian time axes (identify.py:57-79, :82-134): `is_leap_year` follows the calendar of the
    time objects, not the year number -- 2000 is not a leap year on a noleap axis."""
    t = _cf_axis("noleap", [1999, 2000, 2001])
    assert identify.get_calendar(t) == 365
    doy, ndoy = identify.add_doy(t)
    assert ndoy == 366 and len(doy) == 3 * 365
    one = np.concatenate([np.arange(1, 60), np.arange(61, 367)])          # label 60 (Feb 29) never occurs
    assert np.array_equal(doy, np.tile(one, 3))
    doy, ndoy = identify.add_doy(t, keep_tstep=True)
    assert ndoy == 365 and np.array_equal(doy, np.tile(np.arange(1, 366), 3))
    t = _cf_axis("all_leap", [2001, 2002])
    assert identify.get_calendar(t) == 366
    doy, ndoy = identify.add_doy(t)
    assert np.array_equal(doy, np.tile(np.arange(1, 367), 2))             # every year has its Feb 29
    t = _cf_axis("360_day", [2001, 2002])
    assert identify.get_calendar(t) == 360                                 # threshold() then forces tstep (xmhw.py:142-144)
    doy, ndoy = identify.add_doy(t, keep_tstep=True)
    assert ndoy == 360 and np.array_equal(doy, np.tile(np.arange(1, 361), 2))
    doy, _ = identify.add_doy(t)                                           # detect() with tstep=False (xmhw.py:404):
    assert doy[58] == 59 and doy[60] == 62                                 # dayofyr + 1 from March on, no leap years


def test_synth_is_deterministic_and_shardable():
    tm = synth.daily_time(2001, 2002)
    sea = synth.season_table(tm)
    a = synth.synth_sst(len(tm), 64, sea)
    b = np.concatenate([synth.synth_sst(len(tm), 32, sea, cell0=0), synth.synth_sst(len(tm), 32, sea, cell0=32)], 1)
    assert np.array_equal(a, b)                      # a shard can be generated on its own
    assert np.allclose(a * 100, np.rint(a * 100), atol=1e-3)        # 0.01 degC quantisation
    m = synth.land_mask(40, 80)
    assert m[-1].all() and m[:, -1].all() and 0.25 < m.mean() < 0.45
    assert np.array_equal(synth.doy366(tm), O.add_doy(tm))


def test_flip_cold():
    # test_features.py:90-100
    from xmhw_b200.features import flip_cold
    y = np.array([1.0, 2.0, np.nan])
    out = flip_cold({"intensity_sum_dummy": y.copy(), "intensity_var_dummy": y.copy(), "dummy": y.copy()})
    assert np.array_equal(out["intensity_sum_dummy"], -y, equal_nan=True)
    assert np.array_equal(out["intensity_var_dummy"], y, equal_nan=True)
    assert np.array_equal(out["dummy"], y, equal_nan=True)


def test_save_and_load_dataset_roundtrip(tmp_path):
    """xmhw_b200.io: what threshold()/detect() return goes to a NetCDF-3 file and back."""
    from xmhw_b200 import io, labeled
    rng = np.random.default_rng(0)
    doy = np.arange(1, 367)
    lat, lon = np.array([-42.5, -42.25]), np.array([148.0, 148.25, 148.5])
    ds = labeled.Dataset(coords={"doy": doy, "lat": lat, "lon": lon, "quantile": np.float64(0.9)},
                         attrs={"xmhw_parameters": "pctile: 90; windowHalfWidth: 5", "smooth": True})
    ds["thresh"] = labeled.DataArray(rng.normal(18, 2, (366, 2, 3)), ("doy", "lat", "lon"), attrs={"units": "degree_C"})
    ds["seas"] = labeled.DataArray(rng.normal(16, 2, (366, 2, 3)), ("doy", "lat", "lon"))
    p = tmp_path / "clim.nc"
    io.save_dataset(ds, str(p))
    back = io.load_dataset(str(p))
    assert np.array_equal(back["thresh"].values, ds["thresh"].values) and back["thresh"].dims == ("doy", "lat", "lon")
    assert np.array_equal(back.coords["doy"], doy) and np.array_equal(back.coords["lat"], lat)
    assert back["thresh"].attrs["units"] == "degree_C" and back.attrs["coord_quantile"] == 0.9
    # compact event table: int64 indices, datetime64 times, float32 statistics
    ev = labeled.Dataset(coords={"event": np.arange(5)})
    ev["cell"] = labeled.DataArray(np.array([0, 0, 3, 3, 5], np.int64), ("event",))
    ev["index_start"] = labeled.DataArray(np.array([1, 75, 11, 52, 613], np.int64), ("event",))
    ev["time_start"] = labeled.DataArray(np.array(["2003-01-02", "2003-03-17", "2003-01-12", "2003-02-22", "NaT"],
                                                  "datetime64[D]"), ("event",))
    ev["intensity_max"] = labeled.DataArray(rng.normal(2, 0.5, 5), ("event",))
    p2 = tmp_path / "events.nc"
    io.save_dataset(ev, str(p2), float32=True)
    b2 = io.load_dataset(str(p2))
    assert b2["cell"].values.dtype == np.int32 and np.array_equal(b2["index_start"].values, ev["index_start"].values)
    assert b2["intensity_max"].values.dtype == np.float32
    t = b2["time_start"].values
    assert t[0] == 12054.0 and np.isnan(t[4]) and b2["time_start"].attrs["units"].startswith("days since 1970")
    with pytest.raises(TypeError):
        bad = labeled.Dataset()
        bad["s"] = labeled.DataArray(np.array(["a", "b"]), ("x",))
        io.save_dataset(bad, str(tmp_path / "bad.nc"))


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU stand-in for the reference, needs no GPU): exactly one
    JSON line on stdout with the contract keys, whatever libraries write to file descriptor 1."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-cells", "4", "--workload", "small"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "cell-years/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] in ("port", "reference-pandas+glue-port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"] == "small"


def test_annotate_ds_attribute_table():
    """identify.py:539-696: per-variable long_name / units, coordinate attributes and the provenance
    strings, byte for byte (incl. the reference's spelling); when the reference modules are present
    every long_name of the table is checked against the reference source text."""
    import re
    from datetime import date
    from xmhw_b200.features import EVENT_VARIABLES
    ds = labeled.Dataset(coords={"events": np.arange(3), "lat": np.zeros(2), "lon": np.zeros(2)})
    for v in EVENT_VARIABLES:
        ds[v] = labeled.DataArray(np.zeros((3, 2, 2)), ("events", "lat", "lon"))
    identify.annotate_ds(ds, {"ts": {"units": "K"}, "lat": {"units": "degrees_north"}}, "mhw")
    assert ds["intensity_cumulative"].attrs == {"units": "degree_C day",
                                                "long_name": "MHW cumulative intensity relative to seasonal climatology"}
    assert ds["rate_onset"].attrs["units"] == "degree_C day-1" and ds["duration_strong"].attrs["units"] == "1"
    assert ds["intensity_var_abs"].attrs["long_name"] == "MHW intensity variability abosulute magnitude"
    assert "units" not in ds["category"].attrs and ds["category"].attrs["long_name"].endswith("4: Extreme")
    assert ds.coord_attrs["events"]["long_name"] == "MHW event identifier: starting index"
    assert ds.coord_attrs["lat"] == {"units": "degrees_north"}
    assert ds.attrs["source"] == "xmhw code: https://github.com/coecms/xmhw"
    assert ds.attrs["title"] == ("Marine heatwave events identified applying the Hobday et al. (2016) "
                                 "marine heat wave definition")
    assert ds.attrs["history"] == f"{date.today()}: calculated using xmhw code https://github.com/coecms/xmhw"
    assert all(ds[v].attrs.get("long_name") for v in EVENT_VARIABLES if not v.startswith(("time_", "index_")))
    clim = labeled.Dataset(coords={"doy": np.arange(1, 4)})
    clim["thresh"] = labeled.DataArray(np.zeros(3), ("doy",))
    clim["seas"] = labeled.DataArray(np.zeros(3), ("doy",))
    identify.annotate_ds(clim, {"ts": {}}, "clim")
    assert clim["thresh"].attrs["units"] == "degree_C" and clim.coord_attrs["doy"]["long_name"] == "Day of the year"
    assert clim.attrs["title"] == ("Seasonal climatology and threshold calculated to detect marine heatwaves "
                                   "following the  Hobday et al. (2016) definition")
    from oracle import ref_harness as rh
    if rh.available():
        import os
        src = open(os.path.join(rh.REF_ROOT, "xmhw", "identify.py")).read()
        joined = re.sub(r'"\s*\n?\s*\+\s*"', "", src)            # adjacent string literals joined with +
        for name, (long_name, _) in identify.MHW_VARIABLE_ATTRS.items():
            assert long_name in joined, name


def test_save_zarr_compressed_roundtrip(tmp_path):
    """xmhw_b200.io.save_zarr: zlib + float32 encoding (docs/gettingstarted.rst:160-178) in a zarr v2
    directory store; sparse float tables shrink, integers / times / attributes survive the round trip."""
    import json
    import os
    from xmhw_b200 import io
    rng = np.random.default_rng(1)
    n = 4000
    ev = labeled.Dataset(coords={"row": np.arange(n)}, attrs={"xmhw_parameters": "MHW detected using: 5 days"})
    ev.coord_attrs["row"] = {"long_name": "event row"}
    ev["index_start"] = labeled.DataArray(rng.integers(1, 10000, n).astype(np.float64), ("row",))
    ev["duration_moderate"] = labeled.DataArray(rng.integers(0, 30, n), ("row",), attrs={"units": "1"})
    dense = np.full((50, 8, 10), np.nan)
    dense[rng.integers(0, 50, 300), rng.integers(0, 8, 300), rng.integers(0, 10, 300)] = rng.normal(2, 1, 300)
    cube = labeled.Dataset(coords={"events": np.arange(50), "lat": np.arange(8) * 0.25, "lon": np.arange(10) * 0.25})
    cube["intensity_max"] = labeled.DataArray(dense, ("events", "lat", "lon"), attrs={"units": "degree_C"})
    cube["time_peak"] = labeled.DataArray(np.array(["2003-01-02", "NaT"] * 25, "datetime64[ns]"), ("events",))
    p = str(tmp_path / "mhw.zarr")
    io.save_zarr(cube, p)
    meta = json.load(open(os.path.join(p, "intensity_max", ".zarray")))
    assert meta["compressor"] == {"id": "zlib", "level": 5} and meta["dtype"] == "<f4" and meta["zarr_format"] == 2
    assert json.load(open(os.path.join(p, "intensity_max", ".zattrs")))["_ARRAY_DIMENSIONS"] == ["events", "lat", "lon"]
    stored = sum(os.path.getsize(os.path.join(p, "intensity_max", f)) for f in os.listdir(os.path.join(p, "intensity_max")))
    assert stored < dense.nbytes / 10                       # sparse cube: NaNs deflate away
    back = io.load_zarr(p)
    assert np.array_equal(back["intensity_max"].values, dense.astype(np.float32), equal_nan=True)
    assert back["intensity_max"].attrs["units"] == "degree_C" and back["intensity_max"].dims == ("events", "lat", "lon")
    t = back["time_peak"].values
    assert t[0] == np.datetime64("2003-01-02") and np.isnat(t[1])
    p2 = str(tmp_path / "table.zarr")
    io.save_zarr(ev, p2, float32=False, chunk_elems=1024)   # several chunks, ragged last one
    b2 = io.load_zarr(p2)
    assert np.array_equal(b2["index_start"].values, ev["index_start"].values)
    assert np.array_equal(b2["duration_moderate"].values, ev["duration_moderate"].values)
    assert b2.attrs["xmhw_parameters"].startswith("MHW detected") and b2.coord_attrs["row"]["long_name"] == "event row"


def test_sweep_selection_rule_matches_the_launcher():
    """core's "auto" rule (which sweep, which kernel) for the calendars of the bench workloads: host logic only.
    The numbers mirror xmhw_clim_sweep2_f32's launcher (8 warps per SM: shared memory alone, or shared +
    tensor memory when the unit slots split)."""
    from xmhw_b200 import core, plan2
    tm = synth.daily_time(1982, 2011)
    doy = synth.doy366(tm)
    hp = plan2.build_clim_plan2(doy, 366, 5, 0.9)                     # config 3: 11 slots x 40 rows = 55 KB per warp
    assert hp.kp == 36 and core._topk_warps_per_sm(hp) == 4 and core._topk_tmem_fits(hp)
    assert core.sweep2_kernel_name(hp) == "clim_sweep2_tm_kernel"
    hp2 = plan2.build_clim_plan2(doy, 366, 2, 0.9)                    # narrow window: 8 warps fit shared memory
    assert core._topk_warps_per_sm(hp2) >= 8 and core.sweep2_kernel_name(hp2) == "clim_sweep2_kernel"
    tm40 = synth.daily_time(1982, 2021)
    hp40 = plan2.build_clim_plan2(synth.doy366(tm40), 366, 5, 0.9)    # 40-year lists: 48-key arrays, 52-row slots
    assert hp40 is None or (core._topk_warps_per_sm(hp40) < 8 and not core._topk_tmem_fits(hp40))   # -> general sweep
    pent = np.tile(np.arange(1, 74), 30)
    hpp = plan2.build_clim_plan2(pent, 73, 5, 0.9)
    assert hpp is not None and (core._topk_warps_per_sm(hpp) >= 8 or core._topk_tmem_fits(hpp))
