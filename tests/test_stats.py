"""Downstream statistics (reference xmhw/stats.py block_average / mhw_rank): the oracle against the
pandas groupby the reference uses (CPU), the CUDA kernels against the oracle (GPU)."""
import numpy as np
import pytest

from oracle import xmhw_oracle as O
from xmhw_b200 import synth


def _case(ncell=12, years=(1995, 2004), nan_ppm=4000):
    tm = synth.daily_time(*years)
    doy = synth.doy366(tm)
    ts = synth.synth_sst(len(tm), ncell, synth.season_table(tm), nan_ppm=nan_ppm)
    ts[:, 5] = np.nan
    return tm, doy, ts


def test_oracle_block_average_matches_pandas_groupby():
    """agg_mhw (stats.py:322-368) is a pandas groupby over pd.cut bins: same call here per cell (with the
    `_abs` block means taken from the `_abs` columns, see xmhw_b200/stats.py)."""
    import pandas as pd
    tm, doy, ts = _case()
    th, se = O.threshold(ts, doy, 366)
    ev = O.detect(ts, doy, th, se)
    yrs = tm.astype("datetime64[Y]").astype(np.int64) + 1970
    for L in (1, 3):
        got, starts = O.block_average(ev, yrs, ts.shape[1], blockLength=L)
        bins = range(int(yrs[0]), int(yrs[-1]) + L + 1, L)
        for c in (0, 3, 5, 11):
            m = ev["cell"] == c
            df = pd.DataFrame({k: ev[k][m] for k in ("index_start", "duration", "intensity_max", "intensity_cumulative",
                                                     "intensity_mean_abs", "rate_onset", "severity_mean")})
            grp = df.groupby(pd.cut(yrs[df["index_start"].to_numpy()], bins, right=False), observed=False).agg(
                ecount=("index_start", "count"), duration=("duration", "mean"), intensity_max=("intensity_max", "mean"),
                intensity_max_max=("intensity_max", "max"), total_icum=("intensity_cumulative", "sum"),
                intensity_mean_abs=("intensity_mean_abs", "mean"), rate_onset=("rate_onset", "mean"),
                severity_mean=("severity_mean", "mean"))
            assert len(grp) == len(starts)
            for f in grp.columns:
                np.testing.assert_allclose(got[f][:, c], grp[f].to_numpy(dtype=np.float64), rtol=1e-12, atol=0,
                                           equal_nan=True, err_msg=f)


@pytest.mark.gpu
def test_block_average_and_rank_kernels_match_oracle():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from xmhw_b200 import core, stats
    tm, doy, ts_h = _case(ncell=70)
    ts = torch.from_numpy(ts_h).cuda()
    th, se = core.threshold_arrays(ts, doy, 366)
    ev = core.detect_arrays(ts, doy, 366, th, se)
    evh = ev.to_numpy()
    yrs = tm.astype("datetime64[Y]").astype(np.int64) + 1970
    th_h, se_h = th.cpu().numpy(), se.cpu().numpy()
    for L, mt, per in ((1, "time_start", None), (2, "time_peak", (1996, 2003))):
        got, starts = stats.block_average_arrays(ev, yrs, per, L, mt, ts=ts, doy=doy, thresh=th, seas=se)
        exp, estarts = O.block_average(evh, yrs, ts_h.shape[1], per, L, "index_start" if mt == "time_start" else "index_peak",
                                       ts=ts_h, doy=doy, thresh=th_h, seas=se_h)
        assert np.array_equal(starts, estarts) and set(got) == set(exp)
        for f in exp:
            g = got[f].cpu().numpy()
            if f.endswith("_days") or f == "ecount":
                assert np.array_equal(g, exp[f]), f
            else:
                np.testing.assert_allclose(g, exp[f], rtol=1e-10, atol=1e-12, equal_nan=True, err_msg=f)
    ds = stats.block_average(ev, tm, ts=ts, doy=doy, thresh=th, seas=se)
    assert ds["ecount"].dims == ("years", "cell") and ds.coords["years"][0] == 1995
    tot = int(ds["total_days"].values.sum())                 # gap days of joined events have no category (features.py:62-66)
    assert 0.9 * evh["duration"].sum() < tot <= evh["duration"].sum()
    rank, rp = stats.mhw_rank_arrays(ev, 10.0, fields=("duration", "intensity_max", "rate_decline"))
    for f in rank:
        assert np.array_equal(rank[f].cpu().numpy(), O.rank_in_cell(evh, f)), f
        assert np.allclose(rp[f].cpu().numpy(), 11.0 / O.rank_in_cell(evh, f))
