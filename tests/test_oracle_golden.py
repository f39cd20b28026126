"""Pin the CPU oracle against every golden vector the reference's tests hold for the
threshold/detect path.  Literals are transcribed from
/root/reference/test/xmhw_fixtures.py and test_*.py (file:line cited per test)."""
import numpy as np
import pytest

from oracle import xmhw_oracle as O


def test_add_doy_366(oisst):
    # test_identify.py:38-49 + fixture oisst_doy (xmhw_fixtures.py:70-73): 2003 (non-leap) skips 60
    doy = O.add_doy(oisst["time"])
    exp = np.concatenate((np.arange(1, 60), np.arange(61, 367), np.arange(1, 367)))
    assert np.array_equal(doy, exp)


def test_add_doy_tstep():
    # test_identify.py:43-49: pentads 1..73 x2, monthly 1..12 x2
    t5 = np.arange(np.datetime64("2001-01-01"), np.datetime64("2003-01-01"), np.timedelta64(5, "D"))
    t5 = t5[:146]
    assert np.array_equal(O.add_doy(t5, keep_tstep=True), np.tile(np.arange(1, 74), 2))
    tm = np.arange(np.datetime64("2001-01"), np.datetime64("2003-01")).astype("datetime64[D]")
    assert np.array_equal(O.add_doy(tm, keep_tstep=True), np.tile(np.arange(1, 13), 2))
    with pytest.raises(ValueError):
        O.add_doy(tm[:-1], keep_tstep=True)          # identify.py:61-66 incomplete years


def test_feb29(oisst):
    # test_identify.py:52-59: mean over doy in {59,60,61} of the raw series at [1,2] = 18.13
    doy = O.add_doy(oisst["time"])
    ts = oisst["sst"][:, 1, 2]
    sel = np.isin(doy, (59, 60, 61))
    assert abs(float(np.mean(ts[sel])) - 18.13) < 1e-5


def test_runavg():
    # test_identify.py:62-77
    a = np.array([1, 2, 2, 4, 3, 2], float)[:, None]
    np.testing.assert_almost_equal(O.runavg(a, 3)[:, 0], [1.66667, 1.66667, 2.66667, 3.0, 3.0, 2.0], 5)
    np.testing.assert_almost_equal(O.runavg(a, 5)[:, 0], [2.0, 2.2, 2.4, 2.6, 2.4, 2.4], 5)
    with pytest.raises(ValueError):
        O.runavg(a, 2)


def test_runavg_block_order_is_a_plain_circular_mean():
    """The oracle fixes a block suffix/prefix summation order (shared with the CUDA kernel); it must
    agree with the direct window mean to rounding, for any width, wrap-around and NaN placement."""
    rng = np.random.default_rng(7)
    for nd, w in ((366, 31), (366, 5), (73, 5), (12, 3), (40, 39), (7, 1), (366, 61)):
        x = rng.normal(15.0, 5.0, (nd, 9))
        x[rng.integers(0, nd), 3] = np.nan
        got = O.runavg(x, w)
        h = (w - 1) // 2
        idx = (np.arange(nd)[:, None] + np.arange(-h, h + 1)[None, :]) % nd
        ref = x[idx].mean(axis=1)
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        ok = ~np.isnan(ref)
        assert np.abs(got[ok] - ref[ok]).max() < 1e-13 * 20


def test_window_roll_order(oisst):
    # test_identify.py:80-87 + fixture tstack (xmhw_fixtures.py:96-98): z order is window-major
    ts = oisst["sst"][:3, 1, 2]
    z = []
    for k in (-1, 0, 1):
        for t in range(3):
            if 0 <= t + k < 3:
                z.append(ts[t + k])
    np.testing.assert_almost_equal(z, [16.99, 17.39, 16.99, 17.39, 17.3, 17.39, 17.3], 5)
    idx = np.concatenate([O.window_index(np.array([1, 2, 3]), d, 1) for d in (1, 2, 3)])
    assert sorted(idx.tolist()) == sorted([0, 1, 0, 1, 2, 1, 2])


def test_threshold_golden(oisst, clim_gold):
    # test_xmhw.py:24-66: thresh to 6 decimals, seas to 4, ranges [60:] (no smoothing) / [82:] (smoothing)
    smooth, nosmooth = clim_gold
    doy = O.add_doy(oisst["time"])
    pts = np.stack([oisst["sst"][:, 1, 2], oisst["sst"][:, 5, 3]], 1)
    th, se = O.threshold(pts, doy, 366, smoothPercentile=False)
    for c, n in ((0, "1"), (1, "2")):
        assert np.abs(th[60:, c] - nosmooth["thresh" + n][60:]).max() < 1e-12
        assert np.abs(se[60:, c] - nosmooth["seas" + n][60:]).max() < 1e-4
    th, se = O.threshold(pts, doy, 366)
    for c, n in ((0, "1"), (1, "2")):
        assert np.abs(th[82:, c] - smooth["thresh" + n][82:]).max() < 1e-12
        assert np.abs(se[82:, c] - smooth["seas" + n][82:]).max() < 1e-4
    # survey anchor values for config 1
    np.testing.assert_allclose(th[[0, 59, 60, 182, 365], 0],
                               [16.7001285, 18.91856903, 18.924730338, 13.599483481, 16.616515633], rtol=1e-9)


def test_quantile_is_numpy_bitwise():
    rng = np.random.default_rng(0)
    for n in list(range(1, 40)) + [77, 110, 329, 330, 440]:
        x = np.round(rng.normal(15, 3, (n, 5)), 2).astype(np.float32)
        for q in (0.9, 0.99, 0.1, 0.5):
            ref = np.quantile(x, np.asarray([q]), axis=0)[0]
            assert ref.dtype == np.float64
            assert np.array_equal(ref.view(np.int64), O.quantile_linear(x, q).view(np.int64))
    x = np.round(rng.normal(15, 3, (60, 4)), 2).astype(np.float32)
    x[rng.integers(0, 60, 12), rng.integers(0, 4, 12)] = np.nan
    ref = np.nanquantile(x, np.asarray([0.9]), axis=0)[0]
    assert np.array_equal(ref.view(np.int64), O.quantile_linear(x, 0.9).view(np.int64))


FILTER_A = [0, 1, 1, 1, 1, 1, 0, 0, 1, 1, 0, 1, 1, 1, 1, 1, 1, 0, 0, 0, 1, 1, 1, 1, 1, 0, 0, 0, 0]


def test_mhw_filter_golden():
    # test_identify.py:110-122 + fixture filter_data (xmhw_fixtures.py:101-156)
    b = np.array(FILTER_A) == 1
    s, e = O.find_events(b, 5, False)
    assert s.tolist() == [1, 11, 20] and e.tolist() == [5, 16, 24]
    s, e = O.find_events(b, 5, True, 2)          # test_join_gaps: maxGap=2 joins nothing
    assert s.tolist() == [1, 11, 20] and e.tolist() == [5, 16, 24]
    s, e = O.find_events(b, 5, True, 3)          # maxGap=3 joins events 11 and 20
    assert s.tolist() == [1, 11] and e.tolist() == [5, 24]


def test_define_events_golden():
    # test_identify.py:158-190 + fixtures define_data/mhw_data (xmhw_fixtures.py:186-263)
    ts = np.array([15.6, 17.3, 18.2, 19.5, 19.4, 19.6, 18.1, 17.0, 15.2], np.float32)
    se = np.array([15.8, 16.0, 16.2, 16.5, 16.6, 16.4, 16.6, 16.7, 16.4])
    th = np.array([16.0, 16.7, 17.6, 17.9, 18.1, 18.2, 17.3, 17.2, 17.0])
    ev = O.detect(ts, np.arange(1, 10), th, se)
    exp = {"index_start": 1, "index_end": 6, "index_peak": 5, "duration": 6, "category": 2,
           "duration_moderate": 4, "duration_strong": 2, "duration_severe": 0, "duration_extreme": 0,
           "intensity_max": 3.2, "intensity_mean": 2.3, "intensity_cumulative": 13.8, "intensity_var": 0.809938,
           "severity_max": -1.42857, "severity_mean": -1.86931, "severity_cumulative": -11.215873,
           "severity_var": 0.265495, "intensity_max_relThresh": 1.40, "intensity_mean_relThresh": 1.05,
           "intensity_cumulative_relThresh": 6.30, "intensity_var_relThresh": 0.437035,
           "intensity_max_abs": 19.6, "intensity_mean_abs": 18.6834, "intensity_cumulative_abs": 112.1,
           "intensity_var_abs": 0.9495613, "rate_onset": 0.5888889, "rate_decline": 1.5333333}
    assert len(ev["cell"]) == 1
    for k, v in exp.items():
        np.testing.assert_allclose(ev[k][0], v, rtol=1e-5, err_msg=k)   # xr.testing.assert_allclose default


def test_rates_golden():
    # test_features.py:46-51 + fixture rates_data (xmhw_fixtures.py:170-182), get_period :63-79
    start, end, peak, T = 3, 10, 8, 322
    p = peak - start
    onset = (3.1 - 0.5 * (2.3 + 0.3)) / (p + 0.5)
    decline = (3.1 - 0.5 * (1.8 + 0.2)) / ((end - start - p) + 0.5)
    np.testing.assert_almost_equal([onset, decline], [0.32727273, 0.84])


def test_reference_event_tables(ref_cases):
    """Oracle vs event tables produced by the UNMODIFIED reference pandas code
    (tests/golden/ref_detect_cases.npz, generated by oracle/make_golden.py)."""
    nev = 0
    for c in ref_cases:
        T = len(c["ts"])
        minD, join, maxG = (int(v) for v in c["par"])
        ev = O.detect(c["ts"], np.arange(1, T + 1), c["th"], c["se"], minD, bool(join), maxG)
        assert len(ev["cell"]) == len(c["index_start"])
        for f in O.INT_FIELDS:
            ref = c[f]
            got = ev[f].astype(float)
            assert np.array_equal(np.where(np.isnan(ref), -1, ref), got), f
        for f in O.F64_FIELDS:
            tol = 5e-6 if f in O.F32_FIELDS else 1e-9
            np.testing.assert_allclose(ev[f], c[f], rtol=tol, atol=tol, equal_nan=True, err_msg=f)
        nev += len(ev["cell"])
    assert nev > 500


def test_intermediate_golden():
    # test_identify.py:158-190 with intermediate=True + fixture inter_data (xmhw_fixtures.py:267-332)
    ts = np.array([15.6, 17.3, 18.2, 19.5, 19.4, 19.6, 18.1, 17.0, 15.2], np.float32)
    se = np.array([15.8, 16.0, 16.2, 16.5, 16.6, 16.4, 16.6, 16.7, 16.4])
    th = np.array([16.0, 16.7, 17.6, 17.9, 18.1, 18.2, 17.3, 17.2, 17.0])
    r = O.intermediate(ts, np.arange(1, 10), th, se)
    nan = np.nan
    exp = {"seas": [nan, 16.0, 16.2, 16.5, 16.6, 16.4, 16.6, nan, nan],
           "thresh": [nan, 16.7, 17.6, 17.9, 18.1, 18.2, 17.3, nan, nan],
           "bthresh": [0, 1, 1, 1, 1, 1, 1, 0, 0], "events": [nan, 1, 1, 1, 1, 1, 1, nan, nan],
           "relSeas": [nan, 1.3, 2.0, 3.0, 2.79999, 3.2, 1.5, nan, nan],
           "relThresh": [nan, 0.6, 0.6, 1.6, 1.3, 1.4, 0.8, nan, nan],
           "relThreshNorm": [nan, 0.85714, 0.4285714, 1.142857, 0.866667, 0.77778, 1.142857, nan, nan],
           "severity": [nan, -1.857143, -1.42857, -2.142857, -1.8666667, -1.77778, -2.142857, nan, nan],
           "cats": [nan, 1, 1, 2, 1, 1, 2, nan, nan], "duration_moderate": [0, 1, 1, 0, 1, 1, 0, 0, 0],
           "duration_strong": [0, 0, 0, 1, 0, 0, 1, 0, 0], "duration_severe": [0] * 9, "duration_extreme": [0] * 9,
           "mabs": [nan, 17.3, 18.2, 19.5, 19.4, 19.6, 18.1, nan, nan]}
    for k, v in exp.items():
        np.testing.assert_allclose(np.asarray(r[k][:, 0], float), np.asarray(v, float), rtol=1e-5, atol=1e-5,
                                   equal_nan=True, err_msg=k)
