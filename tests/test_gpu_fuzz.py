"""Seeded fuzz of the CUDA path against the oracle over calendars and parameters the fixed
parity cases do not reach: window half-widths 0..12, percentiles 5..99, smoothing widths,
daily / pentad / 10-day / monthly calendars, leap patterns, NaN densities, ragged grids,
detection parameters.  Integer outputs bit-exact, thresh bit-exact, floats within tolerance."""
import numpy as np
import pytest

from tests.util import assert_events_match, bit_equal

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _case(seed):
    rng = np.random.default_rng(seed)
    from xmhw_b200 import synth
    kind = ["daily", "daily", "pentad", "dekad", "monthly"][seed % 5]
    if kind == "daily":
        y0 = int(rng.integers(1950, 2000))
        y1 = y0 + int(rng.integers(2, 36))
        tm = synth.daily_time(y0, y1)
        doy, ndoy, feb = synth.doy366(tm), 366, True
        sea = synth.season_table(tm)
    else:
        steps = {"pentad": 73, "dekad": 36, "monthly": 12}[kind]
        years = int(rng.integers(3, 40))
        doy, ndoy, feb = np.tile(np.arange(1, steps + 1), years), steps, False
        sea = synth.season_table(len(doy))
    T = len(doy)
    ncell = int(rng.integers(1, 90))
    wmax = min(12, (ndoy - 1) // 2)
    w = int(rng.integers(0, wmax + 1))
    pct = int(rng.choice([5, 10, 50, 75, 90, 95, 99]))
    smooth = bool(rng.integers(0, 2))
    sw = int(rng.choice([1, 3, 5, 11, 31]))
    nan_ppm = int(rng.choice([0, 0, 3000, 30000]))
    ts = synth.synth_sst(T, ncell, sea, cell0=int(rng.integers(0, 10 ** 6)), nan_ppm=nan_ppm)
    if ncell > 4:
        ts[:, int(rng.integers(0, ncell))] = np.nan                   # a land cell
        c = int(rng.integers(0, ncell))
        ts[int(rng.integers(0, T // 2)):, c] = np.nan                  # series ends early
    minD = int(rng.integers(1, 8))
    maxG = int(rng.integers(0, minD))
    join = bool(rng.integers(0, 2))
    return dict(ts=ts, doy=doy, ndoy=ndoy, feb=feb, w=w, pct=pct, smooth=smooth, sw=sw,
                minD=minD, maxG=maxG, join=join, kind=kind)


@pytest.mark.parametrize("seed", range(20))
def test_fuzz(seed):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import xmhw_oracle as O
    from xmhw_b200 import core
    from xmhw_b200._cabi import EF_FIELDS
    c = _case(seed)
    ts = torch.from_numpy(c["ts"]).cuda()
    th, se = core.threshold_arrays(ts, c["doy"], c["ndoy"], pctile=c["pct"], windowHalfWidth=c["w"],
                                   smoothPercentile=c["smooth"], smoothPercentileWidth=c["sw"], feb29=c["feb"])
    oth, ose = O.threshold(c["ts"], c["doy"], c["ndoy"], pctile=c["pct"], windowHalfWidth=c["w"],
                           smoothPercentile=c["smooth"], smoothPercentileWidth=c["sw"], tstep=not c["feb"])
    th_h, se_h = th.cpu().numpy(), se.cpu().numpy()
    assert bit_equal(th_h, oth), (c["kind"], c["w"], c["pct"])
    assert np.array_equal(np.isnan(se_h), np.isnan(ose))
    assert np.nanmax(np.abs(se_h - ose), initial=0) <= 1e-9
    ev = core.detect_arrays(ts, c["doy"], c["ndoy"], th, se, c["minD"], c["join"], c["maxG"])
    exp = O.detect(c["ts"], c["doy"], th_h, se_h, c["minD"], c["join"], c["maxG"])
    assert_events_match(ev.to_numpy(), exp, EF_FIELDS)
