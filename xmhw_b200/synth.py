"""Deterministic synthetic SST -- host twin of `xmhw_synth_sst_f32`.

sst(c, t) = m_c + A_c * season[t + phi_c] + x_c(t), x an AR(1) process, rounded
to 0.01 degC like OISST (so ties occur), float32, layout (time, cell) C-order
exactly as xarray would hand it over; land cells are NaN.  Every quantity comes
from a counter-based 64-bit hash of (seed, global cell id, t) and float64
arithmetic without fused multiply-add, so the CUDA generator and this numpy
generator agree bit for bit and any shard can be generated on its own GPU.

Used by bench.py (device generator) and by the tests (both, compared).
"""
import numpy as np

SEED = 20160227
RHO = 0.9
SIGMA = 0.35
NOISE_SCALE = float(1.0 / np.sqrt(4.0 * (65536.0 ** 2 - 1.0) / 12.0))
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    x = np.asarray(x, np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def _u01(h):
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def daily_time(start_year, end_year):
    """datetime64[D] axis start_year-01-01 .. end_year-12-31 (proleptic Gregorian)."""
    return np.arange(np.datetime64("%04d-01-01" % start_year), np.datetime64("%04d-01-01" % (end_year + 1)))


def doy366(time):
    """366-day day-of-year labels (reference xmhw/identify.py:73-76)."""
    t = np.asarray(time).astype("datetime64[D]")
    years = t.astype("datetime64[Y]").astype(np.int64) + 1970
    dayofyear = (t - t.astype("datetime64[Y]").astype("datetime64[D]")).astype(np.int64) + 1
    month = t.astype("datetime64[M]").astype(np.int64) % 12 + 1
    leap = (years % 4 == 0) & ((years % 100 != 0) | (years % 400 == 0))
    return (dayofyear + ((~leap) & (month >= 3))).astype(np.int64)


def season_table(time_or_T):
    """sin(2 pi dayofyear / 365.25) for T + 366 consecutive days (float64, host only)."""
    if np.ndim(time_or_T) == 0:
        j = np.arange(int(time_or_T) + 366, dtype=np.float64)
    else:
        t = np.asarray(time_or_T).astype("datetime64[D]")
        ext = np.arange(t[0], t[0] + np.timedelta64(len(t) + 366, "D"))
        j = (ext - ext.astype("datetime64[Y]").astype("datetime64[D]")).astype(np.float64) + 1.0
    return np.sin(2.0 * np.pi * j / 365.25)


def land_mask(nlat, nlon, frac=0.33):
    """Deterministic smooth land mask (uint8, 1 = land), ~`frac` land, with at least
    one all-land row and one all-land column (exercises the reference's vanishing
    rows/cols after unstack, xmhw.py:213-214)."""
    la = np.linspace(-1.0, 1.0, nlat)[:, None]
    lo = np.linspace(0.0, 2.0 * np.pi, nlon, endpoint=False)[None, :]
    f = (np.sin(3 * lo + 2.0 * la) * np.cos(2.5 * la * np.pi) + 0.6 * np.sin(5 * lo - 1.0) * np.sin(4 * la)
         + 0.4 * np.cos(7 * lo + 3 * la))
    thr = np.quantile(f, 1.0 - frac)
    m = (f > thr).astype(np.uint8)
    m[nlat - 1, :] = 1
    m[:, nlon - 1] = 1
    return m


def synth_sst(T, ngrid, season, land=None, cell0=0, seed=SEED, rho=RHO, sigma=SIGMA, nan_ppm=0, coherent=1):
    """numpy twin of the CUDA generator: float32 [T, ngrid].  `coherent` consecutive cells share
    mean / amplitude / phase (1 = independent cells, the benchmark default); noise is per cell."""
    gid = (np.arange(ngrid, dtype=np.uint64) + np.uint64(cell0))
    with np.errstate(over="ignore"):
        h0 = splitmix64(np.uint64(seed) ^ (gid * np.uint64(0xD1B54A32D192ED03)))
        pgid = (gid // np.uint64(coherent)) * np.uint64(coherent)
        hp = splitmix64(np.uint64(seed) ^ (pgid * np.uint64(0xD1B54A32D192ED03)))
        m = 28.0 * _u01(splitmix64(hp + np.uint64(1)))
        A = 1.0 + 5.0 * _u01(splitmix64(hp + np.uint64(2)))
        phi = (365.0 * _u01(splitmix64(hp + np.uint64(3)))).astype(np.int64)
    out = np.empty((T, ngrid), np.float32)
    x = np.zeros(ngrid, np.float64)
    season = np.asarray(season, np.float64)
    for t in range(T):
        with np.errstate(over="ignore"):
            r = splitmix64(h0 ^ (np.uint64(t) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0x1234567)))
        s16 = ((r & np.uint64(0xffff)).astype(np.int64) + ((r >> np.uint64(16)) & np.uint64(0xffff)).astype(np.int64)
               + ((r >> np.uint64(32)) & np.uint64(0xffff)).astype(np.int64) + (r >> np.uint64(48)).astype(np.int64))
        eps = (s16 - 131070).astype(np.float64) * NOISE_SCALE
        x = rho * x + sigma * eps
        v = (m + A * season[t + phi]) + x
        o = (np.rint(v * 100.0) / 100.0).astype(np.float32)
        if nan_ppm:
            q = splitmix64(r ^ np.uint64(0xA5A5A5A5A5A5A5A5))
            o[(q % np.uint64(1000000)).astype(np.int64) < nan_ppm] = np.nan
        out[t] = o
    if land is not None:
        out[:, np.asarray(land).ravel().astype(bool)] = np.nan
    return out
