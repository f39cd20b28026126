"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs, against the committed golden fixtures, and size-independent
properties at BASELINE sizes.  Integer outputs bit-exact; thresh bit-exact; seas and
float statistics within max(1e-5 abs, 1 float32 ulp rel)."""
import numpy as np
import pytest

from tests.util import assert_events_match, bit_equal

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def core():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from xmhw_b200 import core as c
    return c


def _oracle():
    from oracle import xmhw_oracle as O
    return O


def _float_fields():
    from xmhw_b200._cabi import EF_FIELDS
    return EF_FIELDS


def _clim_check(core, ts_h, doy, ndoy, **kw):
    O = _oracle()
    ts = torch.from_numpy(ts_h).cuda()
    th, se = core.threshold_arrays(ts, doy, ndoy, **kw)
    okw = dict(pctile=kw.get("pctile", 90), windowHalfWidth=kw.get("windowHalfWidth", 5),
               smoothPercentile=kw.get("smoothPercentile", True),
               smoothPercentileWidth=kw.get("smoothPercentileWidth", 31), tstep=not kw.get("feb29", True))
    oth, ose = O.threshold(ts_h, doy, ndoy, **okw)
    th_h, se_h = th.cpu().numpy(), se.cpu().numpy()
    assert bit_equal(th_h, oth), "thresh not bit-equal, max diff %g" % np.nanmax(np.abs(th_h - oth))
    assert np.array_equal(np.isnan(se_h), np.isnan(ose))
    assert np.nanmax(np.abs(se_h - ose), initial=0) <= 1e-9
    return ts, th, se, th_h, se_h


def test_synth_device_matches_host(core):
    from xmhw_b200 import synth
    time = synth.daily_time(2001, 2003)
    land = synth.land_mask(6, 16).ravel()
    sea = synth.season_table(time)
    host = synth.synth_sst(len(time), 96, sea, land=land, cell0=1000, nan_ppm=3000)
    dev = core.synth_sst_device(len(time), 96, sea, land=land, cell0=1000, nan_ppm=3000).cpu().numpy()
    assert np.array_equal(host.view(np.int32), dev.view(np.int32))
    host = synth.synth_sst(len(time), 96, sea, cell0=1000, coherent=32)
    dev = core.synth_sst_device(len(time), 96, sea, cell0=1000, coherent=32).cpu().numpy()
    assert np.array_equal(host.view(np.int32), dev.view(np.int32))


def test_oisst_cube_threshold_and_detect(core, oisst, clim_gold):
    """BASELINE config 1 data (reference test cube, 8x4 grid with 20 land cells)."""
    O = _oracle()
    doy = O.add_doy(oisst["time"])
    ts_h = np.ascontiguousarray(oisst["sst"].reshape(len(doy), -1))
    ts, th, se, th_h, se_h = _clim_check(core, ts_h, doy, 366)
    smooth, _ = clim_gold
    for cell, n in ((1 * 4 + 2, "1"), (5 * 4 + 3, "2")):     # points [1,2] and [5,3] (test_xmhw.py:24-66)
        assert np.abs(th_h[82:, cell] - smooth["thresh" + n][82:]).max() < 1e-6
        assert np.abs(se_h[82:, cell] - smooth["seas" + n][82:]).max() < 1e-4
    ev = core.detect_arrays(ts, doy, 366, th, se)
    got = ev.to_numpy()
    exp = O.detect(ts_h, doy, th_h, se_h)
    assert_events_match(got, exp, _float_fields())
    m = got["cell"] == 6      # survey anchor: 9 events at point [1,2]
    assert list(zip(got["index_start"][m], got["index_end"][m], got["index_peak"][m])) == [
        (1, 7, 5), (75, 81, 79), (99, 103, 101), (114, 121, 118), (138, 150, 146), (172, 184, 182),
        (225, 229, 228), (613, 618, 615), (709, 714, 712)]
    nv = ev.nvalid.cpu().numpy()
    assert np.array_equal(nv > 0, ~np.isnan(ts_h).all(0))


def test_nosmooth_and_raw(core, oisst):
    O = _oracle()
    doy = O.add_doy(oisst["time"])
    ts_h = np.ascontiguousarray(oisst["sst"].reshape(len(doy), -1))
    _clim_check(core, ts_h, doy, 366, smoothPercentile=False)
    _clim_check(core, ts_h, doy, 366, smoothPercentile=False, feb29=False)
    _clim_check(core, ts_h, doy, 366, smoothPercentileWidth=5, windowHalfWidth=2, pctile=75)


@pytest.mark.parametrize("years,ncell,nan_ppm,pctile", [((1982, 2011), 100, 0, 90),
                                                          ((1982, 2011), 70, 20000, 99),
                                                          ((1982, 2021), 64, 0, 90)])
def test_synth_daily(core, years, ncell, nan_ppm, pctile):
    """30/40-year daily series incl. land, scattered NaNs, NaN blocks, ragged last warp."""
    from xmhw_b200 import synth
    O = _oracle()
    time = synth.daily_time(*years)
    doy = synth.doy366(time)
    land = np.zeros(ncell, np.uint8)
    land[[3, 40, 41]] = 1
    ts_h = synth.synth_sst(len(time), ncell, synth.season_table(time), land=land, nan_ppm=nan_ppm)
    if nan_ppm:
        ts_h[100:400, 7] = np.nan          # long block
        ts_h[5000:, 9] = np.nan            # series ends early
        for y in range(10):                # a NaN block every winter
            ts_h[365 * y + 20:365 * y + 100, 11] = np.nan
    ts, th, se, th_h, se_h = _clim_check(core, ts_h, doy, 366, pctile=pctile)
    ev = core.detect_arrays(ts, doy, 366, th, se)
    exp = O.detect(ts_h, doy, th_h, se_h)
    assert len(ev) > 50
    assert_events_match(ev.to_numpy(), exp, _float_fields())


def test_config4_skipna99_winter_blocks(core):
    """BASELINE config 4(i) exactly as SURVEY 8d states it, on 4608 cells (144 whole warps):
    1 % i.i.d. NaN, 2 % of the cells with a 60-120-day NaN block EVERY winter, pctile 99.  The block
    cells have doys without any sample: the reference smooths those cells on their own compacted
    doy axis (identify.py:175-180, :233-241) -- clim_finish_compact_kernel vs the oracle."""
    from oracle import parallel as OP
    from xmhw_b200 import synth
    time = synth.daily_time(1982, 2011)
    doy = synth.doy366(time)
    ncell = 4608
    land = synth.land_mask(48, 96, 0.33).ravel()
    ts_h = synth.synth_sst(len(time), ncell, synth.season_table(time), land=land, nan_ppm=10000)
    rng = np.random.default_rng(44)
    blocks = rng.choice(np.flatnonzero(land == 0), int(0.02 * ncell), replace=False)
    year0 = np.flatnonzero(doy == 1)
    for c in blocks:
        start, length = int(rng.integers(330, 366)), int(rng.integers(60, 121))     # from late Nov / Dec on
        for y0 in year0:
            ts_h[max(0, y0 + start - 365):max(0, y0 + start - 365 + length), c] = np.nan
        ts_h[year0[-1] + start:, c][:length] = np.nan
    ts = torch.from_numpy(ts_h).cuda()
    th, se = core.threshold_arrays(ts, doy, 366, pctile=99)
    ev = core.detect_arrays(ts, doy, 366, th, se)
    th_h, se_h = th.cpu().numpy(), se.cpu().numpy()
    oth, ose, exp = OP.threshold_detect(ts_h, doy, 366, tkw=dict(pctile=99))
    partial = np.isnan(oth[:, blocks]).any(0) & ~np.isnan(oth[:, blocks]).all(0)
    assert partial.sum() >= 0.9 * len(blocks)              # the block cells really lack some doys
    assert bit_equal(th_h, oth), "thresh differs in %d cells" % (np.abs(th_h - oth) > 0).any(0).sum()
    assert np.array_equal(np.isnan(se_h), np.isnan(ose)) and np.nanmax(np.abs(se_h - ose)) <= 1e-9
    # the doys next to a cell's empty season stay finite (the old behaviour made them NaN)
    c = blocks[np.flatnonzero(partial)[0]]
    empty = np.isnan(oth[:, c])
    edge = np.flatnonzero(empty & ~np.roll(empty, 1))[0]          # first doy of the empty season
    assert np.isfinite(th_h[(edge - 1) % 366, c]) and np.isfinite(th_h[(edge - 10) % 366, c])
    assert_events_match(ev.to_numpy(), exp, _float_fields())


@pytest.mark.parametrize("mode", ["topk", "topk_tmem", "topk_tmem_persist", "topk_sync", "general"])
def test_both_sweeps_forced(core, monkeypatch, mode):
    """The two climatology sweeps (two-stack top-K, csrc/xmhw_topk.h; general sorted lists, csrc/xmhw_lane.h)
    forced on the same 30-year series incl. land, NaNs and a ragged last warp: bit-equal to the oracle
    and therefore to each other (the default picks one by the top-K capacity)."""
    from xmhw_b200 import synth
    monkeypatch.setenv("XMHW_B200_SWEEP", mode.split("_")[0])
    if mode.startswith("topk_tmem"):   # 8 warps per SM, the unit slots beyond shared memory in tensor memory
        monkeypatch.setenv("XMHW_B200_SWEEP2_TMEM", "1")
    if mode == "topk_tmem_persist":    # one persistent block per SM, groups drawn from the ticket word
        monkeypatch.setenv("XMHW_B200_SWEEP2_PERSIST", "1")
    if mode == "topk_sync":            # shared-memory kernel with the lockstep barrier
        monkeypatch.setenv("XMHW_B200_SWEEP2_TMEM", "0")
        monkeypatch.setenv("XMHW_B200_SWEEP2_SYNC", "1")
    time = synth.daily_time(1982, 2011)
    doy = synth.doy366(time)
    ncell = 200
    land = synth.land_mask(5, 40).ravel()
    ts_h = synth.synth_sst(len(time), ncell, synth.season_table(time), land=land, nan_ppm=5000)
    ts_h[3000:3400, 17] = np.nan
    core.TRACE = []
    _clim_check(core, ts_h, doy, 366)
    names = [n for n, _, _ in core.TRACE]
    core.TRACE = None
    assert ("xmhw_clim_sweep2_f32" in names) == mode.startswith("topk")
    _clim_check(core, ts_h, doy, 366, pctile=75, windowHalfWidth=2, smoothPercentileWidth=5)


def test_pentad_tstep(core):
    """BASELINE config 4(ii): 73 steps/yr, windowHalfWidth=5, smoothPercentileWidth=5, maxGap=1."""
    from xmhw_b200 import synth
    O = _oracle()
    doy = np.tile(np.arange(1, 74), 30)
    ts_h = synth.synth_sst(len(doy), 48, synth.season_table(len(doy)))
    ts, th, se, th_h, se_h = _clim_check(core, ts_h, doy, 73, smoothPercentileWidth=5, feb29=False)
    ev = core.detect_arrays(ts, doy, 73, th, se, minDuration=3, maxGap=1)
    exp = O.detect(ts_h, doy, th_h, se_h, 3, True, 1)
    assert_events_match(ev.to_numpy(), exp, _float_fields())


def test_detect_options(core):
    from xmhw_b200 import synth
    O = _oracle()
    time = synth.daily_time(1990, 1999)
    doy = synth.doy366(time)
    ts_h = synth.synth_sst(len(time), 40, synth.season_table(time), nan_ppm=8000)
    ts = torch.from_numpy(ts_h).cuda()
    th, se = core.threshold_arrays(ts, doy, 366)
    th_h, se_h = th.cpu().numpy(), se.cpu().numpy()
    for minD, join, maxG in ((5, False, 2), (3, True, 2), (7, True, 4), (1, True, 0), (2, False, 0)):
        ev = core.detect_arrays(ts, doy, 366, th, se, minDuration=minD, joinGaps=join, maxGap=maxG)
        exp = O.detect(ts_h, doy, th_h, se_h, minD, join, maxG)
        assert_events_match(ev.to_numpy(), exp, _float_fields())
    with pytest.raises(ValueError):
        core.detect_arrays(ts, doy, 366, th, se, minDuration=3, maxGap=3)


def test_event_staging_overflow_falls_back_to_exact_fill(core, monkeypatch):
    """The one-pass event finder parks a bounded number of events per cell; a cell with more
    must take the exact two-pass path and give the same table."""
    from xmhw_b200 import synth
    monkeypatch.setenv("XMHW_B200_DETECT", "chain")
    time = synth.daily_time(1982, 2011)
    doy = synth.doy366(time)
    ts = torch.from_numpy(synth.synth_sst(len(time), 96, synth.season_table(time))).cuda()
    th, se = core.threshold_arrays(ts, doy, 366)
    a = core.detect_arrays(ts, doy, 366, th, se).to_numpy()
    assert np.bincount(a["cell"]).max() > 8
    monkeypatch.setattr(core, "STAGE_EVENTS_PER_YEAR", 0)        # staging capacity 8 events per cell
    core.TRACE = []
    b = core.detect_arrays(ts, doy, 366, th, se).to_numpy()
    names = [n for n, _, _ in core.TRACE]
    core.TRACE = None
    assert "xmhw_events_fill" in names and "xmhw_events_gather" not in names
    for f in a:
        assert np.array_equal(a[f], b[f], equal_nan=True), f


def test_fused_detect_equals_kernel_chain(core, monkeypatch):
    """The fused time-major detect pass (xmhw_detect_fused_f32 + xmhw_events_scatter) and the kernel chain
    (exceedance mask -> staged run finding -> gather -> statistics) give the identical table, bit for bit,
    for several option sets incl. land, NaNs, a ragged last warp and an event that ends the series; a
    staging table that is too small makes the fused path fall back to the chain."""
    from xmhw_b200 import synth
    time = synth.daily_time(1990, 2004)
    doy = synth.doy366(time)
    T, ngrid = len(time), 1000
    land = synth.land_mask(10, 100).ravel()
    ts_h = synth.synth_sst(T, ngrid, synth.season_table(time), land=land, nan_ppm=3000)
    ts_h[-40:, 7] += 6.0                                   # an event running into the end of the series
    ts_h[:30, 9] += 6.0                                    # and one from the very first step (index 0 never counts)
    ts = torch.from_numpy(ts_h).cuda()
    th, se = core.threshold_arrays(ts, doy, 366)
    monkeypatch.setattr(core, "FUSED_EVENTS_PER_YEAR", 60.0)        # minDuration 1 yields ~25 events per cell-year
    for minD, join, maxG in ((5, True, 2), (1, True, 0), (3, False, 1), (7, True, 5)):
        monkeypatch.setenv("XMHW_B200_DETECT", "fused")
        core.TRACE = []
        a = core.detect_arrays(ts, doy, 366, th, se, minD, join, maxG)
        names = [n for n, _, _ in core.TRACE]
        core.TRACE = None
        assert "xmhw_detect_fused_f32" in names and "xmhw_exceed_mask_f32" not in names
        monkeypatch.setenv("XMHW_B200_DETECT", "chain")
        b = core.detect_arrays(ts, doy, 366, th, se, minD, join, maxG)
        assert a.n == b.n and a.n > 1000
        assert torch.equal(a.offsets, b.offsets) and torch.equal(a.nvalid, b.nvalid)
        assert torch.equal(a.i32[:, :a.n], b.i32[:, :b.n])
        assert np.array_equal(a.f64[:, :a.n].cpu().numpy().view(np.int64), b.f64[:, :b.n].cpu().numpy().view(np.int64))
    monkeypatch.setenv("XMHW_B200_DETECT", "fused")
    monkeypatch.setattr(core, "FUSED_EVENTS_PER_YEAR", 0.01)
    core.TRACE = []
    c = core.detect_arrays(ts, doy, 366, th, se, 7, True, 5)
    names = [n for n, _, _ in core.TRACE]
    core.TRACE = None
    assert "xmhw_detect_fused_f32" in names and "xmhw_event_stats_cm_f32" in names and c.n == b.n
    assert torch.equal(c.i32[:, :c.n], b.i32[:, :b.n])


def test_event_stats_layouts_and_two_pass_fill_agree(core):
    """The exported doy-major statistics entry point and the two-pass count/fill give exactly
    the table of the path core.detect_arrays uses (cell-major climatology copy, staged events)."""
    from xmhw_b200 import synth, _cabi
    from xmhw_b200.core import _call, _ptr, _stream, _doy_tables
    time = synth.daily_time(1995, 2006)
    doy = synth.doy366(time)
    T, ngrid = len(time), 200
    ts = torch.from_numpy(synth.synth_sst(T, ngrid, synth.season_table(time), nan_ppm=3000)).cuda()
    th, se = core.threshold_arrays(ts, doy, 366)
    ref = core.detect_arrays(ts, doy, 366, th, se)
    st = _stream()
    ptr, tidx, doy32 = _doy_tables(doy, 366, ts.device)
    mask = torch.empty(((ngrid + 31) // 32, T), dtype=torch.int32, device="cuda")
    nvalid = torch.zeros(ngrid, dtype=torch.int32, device="cuda")
    _call("xmhw_exceed_mask_f32", _ptr(ts), T, ngrid, _ptr(ptr), _ptr(tidx), 366, _ptr(th), _ptr(mask), _ptr(nvalid), st)
    counts = torch.empty(ngrid, dtype=torch.int32, device="cuda")
    _call("xmhw_events_count", _ptr(mask), T, ngrid, 5, 1, 2, _ptr(counts), st)
    offsets = torch.empty(ngrid + 1, dtype=torch.int64, device="cuda")
    scratch = torch.empty(ngrid // 1024 + 2, dtype=torch.int64, device="cuda")
    _call("xmhw_exclusive_scan_i32", _ptr(counts), ngrid, _ptr(offsets), _ptr(scratch), st)
    nev = int(offsets[-1].item())
    assert nev == ref.n and torch.equal(offsets, ref.offsets)
    ei = torch.empty((_cabi.EI_COUNT, nev), dtype=torch.int32, device="cuda")
    ef = torch.empty((_cabi.EF_COUNT, nev), dtype=torch.float64, device="cuda")
    _call("xmhw_events_fill", _ptr(mask), T, ngrid, 5, 1, 2, _ptr(offsets), nev, _ptr(ei), st)
    _call("xmhw_event_stats_f32", _ptr(ts), T, ngrid, _ptr(doy32), _ptr(th), _ptr(se), nev, nev, _ptr(ei), _ptr(ef), st)
    torch.cuda.synchronize()
    assert torch.equal(ei, ref.i32[:, :nev])
    a, b = ef.cpu().numpy(), ref.f64[:, :nev].cpu().numpy()
    assert np.array_equal(a.view(np.int64), b.view(np.int64))


def test_reference_event_tables(core, ref_cases):
    """Event tables from the UNMODIFIED reference pandas code (tests/golden/ref_detect_cases.npz)."""
    from tests.util import F32_FIELDS
    nev = 0
    for c in ref_cases:
        T = len(c["ts"])
        minD, join, maxG = (int(v) for v in c["par"])
        doy = np.arange(1, T + 1)
        ts = torch.from_numpy(c["ts"][:, None].copy()).cuda()
        th = torch.from_numpy(c["th"][:, None].copy()).cuda()
        se = torch.from_numpy(c["se"][:, None].copy()).cuda()
        got = core.detect_arrays(ts, doy, T, th, se, minD, bool(join), maxG).to_numpy()
        assert len(got["cell"]) == len(c["index_start"])
        for f in ("index_start", "index_end", "index_peak", "duration", "category", "duration_moderate",
                  "duration_strong", "duration_severe", "duration_extreme"):
            assert np.array_equal(got[f], np.where(np.isnan(c[f]), -1, c[f]).astype(np.int64)), f
        for f in _float_fields():
            tol = 5e-6 if f in F32_FIELDS else 1e-9
            np.testing.assert_allclose(got[f], c[f], rtol=tol, atol=tol, equal_nan=True, err_msg=f)
        nev += len(got["cell"])
    assert nev > 500


def test_all_land_and_tiny(core):
    """Edge cases: all-NaN grid (no events, NaN climatology), single cell, T < window."""
    doy = np.tile(np.arange(1, 13), 3)
    ts = torch.full((36, 5), float("nan"), dtype=torch.float32).cuda()
    th, se = core.threshold_arrays(ts, doy, 12, windowHalfWidth=1, smoothPercentileWidth=3, feb29=False)
    assert bool(torch.isnan(th).all()) and bool(torch.isnan(se).all())
    ev = core.detect_arrays(ts, doy, 12, th, se)
    assert len(ev) == 0 and int(ev.nvalid.sum()) == 0
    O = _oracle()
    rng = np.random.default_rng(3)
    ts_h = rng.normal(10, 2, (36, 1)).astype(np.float32)
    _clim_check(core, ts_h, doy, 12, windowHalfWidth=1, smoothPercentileWidth=3, feb29=False)


def test_regional_properties(core):
    """BASELINE config 2 shape (240x160 cells, 1982-2021, T=14610, 2.2 GB): size-independent
    properties + oracle spot check on a random sample of cells."""
    from xmhw_b200 import synth
    O = _oracle()
    time = synth.daily_time(1982, 2021)
    doy = synth.doy366(time)
    T, ngrid = len(time), 240 * 160
    sea = synth.season_table(time)
    ts = core.synth_sst_device(T, ngrid, sea)
    th, se = core.threshold_arrays(ts, doy, 366)
    ev = core.detect_arrays(ts, doy, 366, th, se)
    torch.cuda.synchronize()
    assert not bool(torch.isnan(th).any())
    assert bool((th > se).all())                        # 90th percentile above the mean
    got = ev.to_numpy()
    n = len(got["cell"])
    assert 1.5 < n / (ngrid * 40) < 3.5                 # ~2.4 events per cell-year
    assert np.all(got["duration"] == got["index_end"] - got["index_start"] + 1)
    assert np.all(got["duration"] >= 5) and np.all(got["index_start"] >= 1) and np.all(got["index_end"] < T)
    assert np.all((got["index_peak"] >= got["index_start"]) & (got["index_peak"] <= got["index_end"]))
    same = got["cell"][1:] == got["cell"][:-1]
    assert np.all(np.diff(got["cell"]) >= 0)
    assert np.all((got["index_start"][1:] - got["index_end"][:-1] - 1)[same] > 2)     # gaps > maxGap after joining
    assert np.all(got["duration_moderate"] + got["duration_strong"] + got["duration_severe"]
                  + got["duration_extreme"] <= got["duration"])
    assert np.all(got["category"] >= 1) and np.all(got["intensity_max"] > 0)
    assert np.array_equal(np.bincount(got["cell"], minlength=ngrid), np.diff(ev.offsets.cpu().numpy()))
    # determinism / idempotence: a second run gives the identical table
    ev2 = core.detect_arrays(ts, doy, 366, th, se).to_numpy()
    for k in got:
        assert np.array_equal(got[k], ev2[k], equal_nan=True), k
    # oracle check on 4096 cells taken as WHOLE WARPS (128 groups of 32 adjacent cells incl. the first
    # and the last of the grid): lane interactions (ballots, shared slots) show per warp, not per cell
    from oracle import parallel as OP
    groups = np.unique(np.concatenate([[0, ngrid // 32 - 1], np.random.default_rng(1).choice(ngrid // 32, 126, replace=False)]))
    cells = (groups[:, None] * 32 + np.arange(32)[None, :]).ravel()
    idx = torch.from_numpy(cells).cuda()
    ts_h = ts[:, idx].cpu().numpy()
    oth, ose, exp = OP.threshold_detect(ts_h, doy, 366)
    th_h, se_h = th[:, idx].cpu().numpy(), se[:, idx].cpu().numpy()
    assert bit_equal(th_h, oth)
    assert np.abs(se_h - ose).max() <= 1e-9
    sel = np.isin(got["cell"], cells)
    sub = {k: v[sel] for k, v in got.items()}
    sub["cell"] = np.searchsorted(cells, sub["cell"])
    assert_events_match(sub, exp, _float_fields())


def test_global_config3_properties(core):
    """BASELINE config 3 at FULL size (1440x720 grid, 1982-2011, T=10957, 45.4 GB, 33 % land: the
    bench workload): size-independent properties on the device + oracle spot check on sampled cells."""
    from xmhw_b200 import synth
    if torch.cuda.get_device_properties(0).total_memory < 120e9:
        pytest.skip("needs ~90 GB of device memory")
    O = _oracle()
    time = synth.daily_time(1982, 2011)
    doy = synth.doy366(time)
    nlat, nlon = 720, 1440
    T, ngrid = len(time), nlat * nlon
    land = synth.land_mask(nlat, nlon, 0.33).ravel()
    ts = core.synth_sst_device(T, ngrid, synth.season_table(time), land=land)
    th, se = core.threshold_arrays(ts, doy, 366)
    ev = core.detect_arrays(ts, doy, 366, th, se)
    torch.cuda.synchronize()
    land_d = torch.from_numpy(land.astype(bool)).cuda()
    # land cells: NaN climatology, no valid sample, no event; ocean cells: finite, thresh above the mean
    assert bool(torch.isnan(th[:, land_d]).all()) and bool(torch.isnan(se[:, land_d]).all())
    assert not bool(torch.isnan(th[:, ~land_d]).any()) and bool((th[:, ~land_d] > se[:, ~land_d]).all())
    assert bool((ev.nvalid[land_d] == 0).all()) and bool((ev.nvalid[~land_d] == T).all())
    n = ev.n
    nocean = int((~land_d).sum())
    assert 1.5 < n / (nocean * 30) < 3.5                # ~2.2 events per ocean cell-year
    col = {f: ev.column(f) for f in ("cell", "index_start", "index_end", "index_peak", "duration", "category",
                                     "duration_moderate", "duration_strong", "duration_severe", "duration_extreme")}
    assert bool((col["duration"] == col["index_end"] - col["index_start"] + 1).all())
    assert bool((col["duration"] >= 5).all()) and bool((col["index_start"] >= 1).all()) and bool((col["index_end"] < T).all())
    assert bool(((col["index_peak"] >= col["index_start"]) & (col["index_peak"] <= col["index_end"])).all())
    assert bool((col["cell"][1:] >= col["cell"][:-1]).all())
    same = col["cell"][1:] == col["cell"][:-1]
    assert bool(((col["index_start"][1:] - col["index_end"][:-1] - 1)[same] > 2).all())     # gaps > maxGap after joining
    assert bool((col["duration_moderate"] + col["duration_strong"] + col["duration_severe"]
                 + col["duration_extreme"] <= col["duration"]).all())
    assert bool((col["category"] >= 1).all()) and bool((col["category"] <= 4).all())
    assert bool((ev.column("intensity_max") > 0).all())
    assert not bool(land_d[col["cell"].long()].any())                                       # no event on land
    counts = torch.bincount(col["cell"].long(), minlength=ngrid)
    assert torch.equal(counts, ev.offsets[1:] - ev.offsets[:-1])
    # oracle check on >= 4096 cells taken as WHOLE WARPS: 60 coast warps (land and ocean lanes mixed),
    # 60 all-ocean warps, 6 all-land warps, the first and the last warp of the grid
    from oracle import parallel as OP
    rng = np.random.default_rng(3)
    per_warp = (land.reshape(-1, 32) == 0).sum(1)
    groups = np.unique(np.concatenate([
        [0, ngrid // 32 - 1],
        rng.choice(np.flatnonzero((per_warp > 0) & (per_warp < 32)), 60, replace=False),
        rng.choice(np.flatnonzero(per_warp == 32), 60, replace=False),
        rng.choice(np.flatnonzero(per_warp == 0), 6, replace=False)]))
    cells = (groups[:, None] * 32 + np.arange(32)[None, :]).ravel()
    idx = torch.from_numpy(cells).cuda()
    ts_h = ts[:, idx].cpu().numpy()
    th_h, se_h = th[:, idx].cpu().numpy(), se[:, idx].cpu().numpy()
    oth, ose, exp = OP.threshold_detect(ts_h, doy, 366)
    assert len(cells) >= 4096 and bit_equal(th_h, oth)
    assert np.nanmax(np.abs(se_h - ose)) <= 1e-9 and np.array_equal(np.isnan(se_h), np.isnan(ose))
    sel = torch.isin(col["cell"].long(), idx)
    keep = sel.nonzero().squeeze(1)
    from xmhw_b200.core import EI_FIELDS, EF_FIELDS
    got = {f: ev.i32[k, :n][keep].cpu().numpy().astype(np.int64) for k, f in enumerate(EI_FIELDS)}
    got.update({f: ev.f64[k, :n][keep].cpu().numpy() for k, f in enumerate(EF_FIELDS)})
    got["cell"] = np.searchsorted(cells, got["cell"])
    assert_events_match(got, exp, _float_fields())
    del ts, th, se, ev
    torch.cuda.empty_cache()


def test_series_without_leap_year(core):
    """doy 60 never occurs (2001-2003): absent from the reference's groupby output, so feb29 /
    runavg act on the compacted 365-doy axis and doy 60 comes back NaN."""
    from xmhw_b200 import synth
    O = _oracle()
    time = synth.daily_time(2001, 2003)
    doy = synth.doy366(time)
    ts_h = synth.synth_sst(len(time), 40, synth.season_table(time))
    ts, th, se, th_h, se_h = _clim_check(core, ts_h, doy, 366)
    assert np.isnan(th_h[59]).all() and not np.isnan(th_h[58]).any()
    ev = core.detect_arrays(ts, doy, 366, th, se)
    assert_events_match(ev.to_numpy(), O.detect(ts_h, doy, th_h, se_h), _float_fields())


def test_host_buffer_entry_point(core):
    """threshold_detect_host (pinned host series in, host results out, column blocks pipelined on
    three streams) gives exactly the single-shot device result."""
    from xmhw_b200 import synth
    time = synth.daily_time(1995, 2004)
    doy = synth.doy366(time)
    land = synth.land_mask(5, 40).ravel()
    ts_h = synth.synth_sst(len(time), 200, synth.season_table(time), land=land, nan_ppm=3000)
    host = torch.from_numpy(ts_h).pin_memory()
    res = core.threshold_detect_host(host, doy, 366, slabs=3)
    ts = torch.from_numpy(ts_h).cuda()
    th, se = core.threshold_arrays(ts, doy, 366)
    ev = core.detect_arrays(ts, doy, 366, th, se)
    assert np.array_equal(res["thresh"].numpy(), th.cpu().numpy(), equal_nan=True)
    assert np.array_equal(res["seas"].numpy(), se.cpu().numpy(), equal_nan=True)
    assert np.array_equal(res["nvalid"].numpy(), ev.nvalid.cpu().numpy())
    assert res["n_events"] == ev.n
    assert np.array_equal(res["ev_i32"].numpy(), ev.i32[:, :ev.n].cpu().numpy())
    assert np.array_equal(res["ev_f64"].numpy(), ev.f64[:, :ev.n].cpu().numpy(), equal_nan=True)


def test_interp_gaps_pre_step(core):
    """maxPadLength pre-step (xmhw.py:159-160, :409-410): bit-equal to the np.interp oracle."""
    from xmhw_b200 import synth
    O = _oracle()
    ts_h = synth.synth_sst(500, 70, synth.season_table(500), nan_ppm=60000)
    ts_h[:3, 4] = np.nan           # leading gap: never filled
    ts_h[-2:, 5] = np.nan          # trailing gap: never filled
    ts_h[100:140, 6] = np.nan      # long gap: left as NaN
    ts_h[:, 7] = np.nan
    for max_pad in (1, 3, 10):
        got = core.interp_gaps_(torch.from_numpy(ts_h.copy()).cuda(), max_pad).cpu().numpy()
        exp = O.interp_gaps(ts_h, max_pad)
        assert np.array_equal(got.view(np.int32), exp.view(np.int32)), max_pad


def test_intermediate_dataset(core):
    """intermediate=True per-timestep fields (identify.py:404-411, features.py:22-69) vs oracle."""
    from xmhw_b200 import synth
    O = _oracle()
    time = synth.daily_time(1996, 2003)
    doy = synth.doy366(time)
    land = np.zeros(50, np.uint8)
    land[7] = 1
    ts_h = synth.synth_sst(len(time), 50, synth.season_table(time), land=land, nan_ppm=4000)
    ts = torch.from_numpy(ts_h).cuda()
    th, se = core.threshold_arrays(ts, doy, 366)
    ev = core.detect_arrays(ts, doy, 366, th, se)
    got = core.intermediate_arrays(ts, doy, 366, th, se, ev)
    exp = O.intermediate(ts_h, doy, th.cpu().numpy(), se.cpu().numpy())
    for k in O.INTER_FIELDS:
        g = got[k].cpu().numpy()
        if g.dtype == bool:
            assert np.array_equal(g, exp[k]), k
        else:
            assert np.array_equal(np.isnan(g), np.isnan(exp[k])), k
            assert np.allclose(g, exp[k], rtol=0, atol=1e-12, equal_nan=True), k


def test_group_order_is_stable_land_last_permutation(core):
    """xmhw_group_order_f32: the 32-cell groups with data in a probe row first, the all-NaN groups last, both in
    grid order (ragged last group included) -- compared with the same rule in numpy."""
    from xmhw_b200 import synth
    time = synth.daily_time(2001, 2003)
    for ncell, land_shape in ((5000, (50, 100)), (1024 * 33 + 7, None)):
        land = synth.land_mask(*land_shape).ravel() if land_shape else (np.arange(ncell) // 97 % 3 == 0).astype(np.uint8)
        ts_h = synth.synth_sst(len(time), ncell, synth.season_table(time), land=land[:ncell], nan_ppm=2000)
        ts_h[0, 64:96] = np.nan                     # a data gap in one probe row only: still a group with data
        ts = torch.from_numpy(ts_h).cuda()
        T = len(time)
        ncg = (ncell + 31) // 32
        order = core._group_order(ts).cpu().numpy()
        assert len(order) == ncg + 1                 # + the word the sweep uses as its work ticket
        order = order[:ncg]
        probe = np.isnan(ts_h[0]) & np.isnan(ts_h[T // 2]) & np.isnan(ts_h[T - 1])
        probe = np.concatenate([probe, np.ones(ncg * 32 - ncell, bool)]).reshape(ncg, 32).all(1)
        exp = np.concatenate([np.flatnonzero(~probe), np.flatnonzero(probe)])
        assert probe.any() and (~probe).any() and np.array_equal(order, exp)


def test_tmem_persistent_launch_equals_block_launch(core, monkeypatch):
    """The persistent launch mode of the tensor-memory sweep (one block per SM, warps draw groups from the ticket
    word) on a grid large enough that every warp sweeps several groups: bit-equal raw climatologies to the default
    launch (one block per 8 groups), which the other tests pin to the oracle."""
    from xmhw_b200 import synth
    time = synth.daily_time(1982, 2011)
    doy = synth.doy366(time)
    nlat, nlon = 60, 1000                              # 60 000 cells = 1 875 groups > 148 SMs x 8 warps
    land = synth.land_mask(nlat, nlon, 0.3).ravel()
    ts = core.synth_sst_device(len(time), nlat * nlon, synth.season_table(time), land=land, nan_ppm=2000)
    monkeypatch.setenv("XMHW_B200_SWEEP", "topk")
    monkeypatch.setenv("XMHW_B200_SWEEP2_TMEM", "1")
    out = {}
    for persist in ("0", "1"):
        monkeypatch.setenv("XMHW_B200_SWEEP2_PERSIST", persist)
        th, se, ne = core.threshold_arrays(ts, doy, 366, smoothPercentile=False, feb29=False, return_nempty=True)
        out[persist] = (th.clone(), se.clone(), ne.clone())
    for a, b in zip(out["0"], out["1"]):
        assert torch.equal(torch.nan_to_num(a.double(), nan=-1.0), torch.nan_to_num(b.double(), nan=-1.0))
    assert int(torch.isnan(out["0"][0]).all(0).sum()) > 1000      # land columns really are there
