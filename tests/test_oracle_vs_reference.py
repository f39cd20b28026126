"""Live fuzz of the numpy oracle against the UNMODIFIED reference pandas code (identify.mhw_filter /
join_gaps, features.mhw_df / mhw_features), loaded by oracle/ref_harness.py from /root/reference or
from the copies oracle/build_ref.py made (oracle/_ref, which travels to the GPU box).  Skipped where
neither exists.  Integer outputs bit-exact, floats within the contract of tests/util.py."""
import warnings

import numpy as np
import pytest

from oracle import ref_harness as rh
from oracle import xmhw_oracle as O

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference modules not present (oracle/build_ref.py)")


def _series(rng, T, nan_frac):
    x = np.zeros(T)
    e = rng.normal(0, 0.6, T)
    for t in range(1, T):
        x[t] = 0.85 * x[t - 1] + e[t]
    se = 15 + 3 * np.sin(np.arange(T) / 58.0)
    th = se + rng.uniform(0.2, 0.9) + 0.1 * np.sin(np.arange(T) / 9.0)
    ts = np.round(se + x, 2).astype(np.float32)
    if nan_frac:
        ts[rng.integers(0, T, size=max(1, int(T * nan_frac)))] = np.nan
    return ts, th, se


@pytest.mark.parametrize("seed", range(6))
def test_run_rules_match_reference_filter(seed):
    """mhw_filter + join_gaps (identify.py:415-479, :273-325) vs oracle.find_events on random masks."""
    rng = np.random.default_rng(100 + seed)
    warnings.filterwarnings("ignore")
    for _ in range(25):
        T = int(rng.integers(20, 400))
        b = rng.random(T) < rng.uniform(0.2, 0.8)
        # long runs: smooth the mask
        b = np.convolve(b, np.ones(3), mode="same") >= 2
        minD = int(rng.integers(1, 8))
        maxG = int(rng.integers(0, minD)) if minD > 1 else 0
        join = bool(rng.integers(0, 2))
        st, en, ev = rh.ref_mhw_filter(b, minD, join, maxG)
        s, e = O.find_events(b, minD, join, maxG)
        lab = np.full(T, np.nan)
        for a, z in zip(s, e):
            lab[a:z + 1] = a
        assert np.array_equal(np.isnan(ev), np.isnan(lab)) and np.array_equal(ev[~np.isnan(ev)], lab[~np.isnan(lab)])


@pytest.mark.parametrize("seed", range(4))
def test_event_tables_match_reference(seed):
    rng = np.random.default_rng(500 + seed)
    warnings.filterwarnings("ignore")
    nev = 0
    for _ in range(10):
        T = int(rng.integers(120, 900))
        ts, th, se = _series(rng, T, float(rng.choice([0.0, 0.0, 0.01, 0.03])))
        minD, join, maxG = int(rng.integers(2, 7)), bool(rng.integers(0, 2)), int(rng.integers(0, 2))
        df = rh.ref_define_events(ts, th, se, minD, join, maxG)
        doy = np.arange(1, T + 1)
        got = O.detect(ts[:, None], doy, th[:, None], se[:, None], minD, join, maxG)
        n = 0 if df is None else len(df)
        assert len(got["cell"]) == n
        if n == 0:
            continue
        for f in O.INT_FIELDS:
            if f == "cell":
                continue
            ref = df[f].to_numpy().astype(np.float64)
            assert np.array_equal(got[f], np.where(np.isnan(ref), -1, ref).astype(np.int64)), f
        for f in O.F64_FIELDS:
            ref = df[f].to_numpy().astype(np.float64)
            tol = 5e-6 if f.endswith("_abs") else 1e-9
            np.testing.assert_allclose(got[f], ref, rtol=tol, atol=tol, equal_nan=True, err_msg=f)
        nev += n
    assert nev > 20
