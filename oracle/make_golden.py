"""Generate tests/golden/*.npz from the reference's own data and code.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden

Outputs (all small, committed):
  oisst_2003_2004.npz       the reference's test SST cube (test/testdata/oisst_2003_2004.nc)
  clim_oisst[_nosmooth].npz Eric Oliver's thresh/seas at two points (test/testdata/test_clim_oisst*.nc,
                            compared by test/test_xmhw.py:24-66)
  ref_detect_cases.npz      event tables produced by the UNMODIFIED reference pandas code
                            (identify.mhw_filter, features.mhw_df/mhw_features via oracle/ref_harness.py)
                            for the two OISST points and a set of seeded synthetic series
                            (NaNs, ties, joins on/off, events touching both ends of the series).
"""
import os
import warnings

import numpy as np

from . import ref_harness as rh
from . import xmhw_oracle as O
from .nc_reader import read_nc

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
DATA = os.path.join(rh.REF_ROOT, "test", "testdata")

FIELDS = O.INT_FIELDS + O.F64_FIELDS


def synth_series(rng, T, nan_frac=0.0):
    x = np.zeros(T)
    e = rng.normal(0, 0.6, T)
    for t in range(1, T):
        x[t] = 0.85 * x[t - 1] + e[t]
    se = 15 + 3 * np.sin(np.arange(T) / 58.0)
    th = se + rng.uniform(0.3, 0.9) + 0.1 * np.sin(np.arange(T) / 9.0)
    ts = np.round(se + x, 2).astype(np.float32)
    if nan_frac:
        ts[rng.integers(0, T, size=max(1, int(T * nan_frac)))] = np.nan
    return ts, th, se


def ref_table(ts, th, se, minD, join, maxG):
    df = rh.ref_define_events(ts, th, se, minD, join, maxG)
    if df is None:
        return {f: np.zeros(0) for f in FIELDS}
    return {f: df[f].to_numpy().astype(np.float64) for f in FIELDS}


def main():
    warnings.filterwarnings("ignore")
    os.makedirs(GOLD, exist_ok=True)
    o = read_nc(os.path.join(DATA, "oisst_2003_2004.nc"))
    np.savez_compressed(os.path.join(GOLD, "oisst_2003_2004.npz"),
                        sst=o["sst"], time=o["time"], lat=o["lat"], lon=o["lon"])
    for name in ("test_clim_oisst", "test_clim_oisst_nosmooth"):
        c = read_nc(os.path.join(DATA, name + ".nc"))
        np.savez_compressed(os.path.join(GOLD, name.replace("test_", "") + ".npz"),
                            **{k: c[k] for k in ("thresh1", "thresh2", "seas1", "seas2")})

    cases = {}
    ncase = 0

    def add(ts, th, se, minD, join, maxG):
        nonlocal ncase
        tab = ref_table(ts, th, se, minD, join, maxG)
        p = "c%03d_" % ncase
        cases[p + "ts"] = ts
        cases[p + "th"] = th
        cases[p + "se"] = se
        cases[p + "par"] = np.array([minD, int(join), maxG], np.int64)
        for f in FIELDS:
            cases[p + f] = tab[f]
        ncase += 1

    # the two OISST points of the reference's threshold test, default parameters
    time = np.datetime64("2003-01-01T12:00:00") + o["time"].astype("timedelta64[D]")
    doy = O.add_doy(time)
    pts = np.stack([o["sst"][:, 1, 2], o["sst"][:, 5, 3]], 1)
    th, se = O.threshold(pts, doy, 366)
    for c in range(2):
        add(pts[:, c], th[doy - 1, c], se[doy - 1, c], 5, True, 2)
    rng = np.random.default_rng(20160227)
    for trial in range(60):
        T = int(rng.integers(30, 600))
        minD = int(rng.integers(2, 7))
        maxG = int(rng.integers(0, minD))
        join = bool(trial % 4 != 3)
        ts, thv, sev = synth_series(rng, T, nan_frac=0.04 if trial % 3 == 0 else 0.0)
        if trial % 5 == 0:      # force an event at the very start and at the very end
            ts[:8] = (thv[:8] + 1.0).astype(np.float32)
            ts[-7:] = (thv[-7:] + 0.5).astype(np.float32)
        add(ts, thv, sev, minD, join, maxG)
    cases["ncase"] = np.array(ncase)
    np.savez_compressed(os.path.join(GOLD, "ref_detect_cases.npz"), **cases)
    print("wrote", ncase, "reference detect cases to", GOLD)


if __name__ == "__main__":
    main()
