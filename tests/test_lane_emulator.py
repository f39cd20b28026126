"""CPU tests of the host-side plan builder and of the per-lane device logic compiled for
the host (tests/lane_emulator, TEST INFRASTRUCTURE ONLY): plan + sweep selection, run
finding and event statistics against the oracle and the reference-generated goldens."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import xmhw_oracle as O
from tests.util import bit_equal
from xmhw_b200 import plan as P
from xmhw_b200 import synth as S

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def em(_built):
    from xmhw_b200 import _cabi
    lib = C.CDLL(os.path.join(HERE, "lane_emulator", "_lane_emul.so"))
    lib.emul_find_events.restype = C.c_int
    return lib, _cabi


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def sweep(em, ts, doy, ndoy, w, q, keep=None):
    lib, cabi = em
    hp = P.build_clim_plan(doy, ndoy, w, q, keep=keep)
    s, keep = cabi.numpy_plan_struct(hp)
    ts = np.ascontiguousarray(ts, np.float32)
    T, ng = ts.shape
    thr = np.empty((ndoy, ng))
    se = np.empty((ndoy, ng))
    lib.emul_clim_sweep(_vp(ts), C.c_int64(T), C.c_int64(ng), C.byref(s), _vp(thr), _vp(se))
    return hp, thr, se


CASES = [
    ("30yr", (1982, 2011), 24, 0, 5, 90),
    ("30yr_nan_p99", (1982, 2011), 24, 20000, 5, 99),
    ("40yr_48key_lists", (1982, 2021), 12, 0, 5, 90),
    ("70yr_two_pieces", (1950, 2019), 8, 15000, 5, 90),
    ("noleap_w2", (2001, 2003), 12, 0, 2, 90),
    ("w0_median", (2001, 2003), 8, 0, 0, 50),
    ("w15_p10", (2001, 2004), 8, 0, 15, 10),
]


@pytest.mark.parametrize("keep", [16, 3, 1])
@pytest.mark.parametrize("name,years,ncell,nan_ppm,w,pct", CASES, ids=[c[0] for c in CASES])
def test_sweep_matches_oracle(em, name, years, ncell, nan_ppm, w, pct, keep):
    """keep = key rows per list in shared memory; small values force the exact global-memory
    tail path (tail_key / tail_count) on almost every move."""
    tm = S.daily_time(*years)
    doy = S.doy366(tm)
    assert np.array_equal(doy, O.add_doy(tm))
    ts = S.synth_sst(len(tm), ncell, S.season_table(tm), nan_ppm=nan_ppm)
    if nan_ppm:
        ts[100:300, 3] = np.nan
        ts[:, 5] = np.nan
        ts[700:, 7] = np.nan
    hp, thr, se = sweep(em, ts, doy, 366, w, pct / 100.0, keep=keep)
    oth, ose = O.threshold(ts, doy, 366, pctile=pct, windowHalfWidth=w, smoothPercentile=False, tstep=True)
    assert bit_equal(thr, oth)
    assert np.nanmax(np.abs(se - ose), initial=0) <= 1e-12
    assert hp.rows_loaded < len(doy) * 1.1 and hp.max_size <= 48


def sweep2(em, ts, doy, ndoy, w, q, phased=False):
    """two-stack top-K sweep (csrc/xmhw_topk.h) + the direct selection of the exceptional doys;
    phased: the step with separate push / flip loops (step_phased, what the tensor-memory kernel runs)"""
    from xmhw_b200 import plan2 as P2
    lib, cabi = em
    lib.emul_set_phased(1 if phased else 0)
    hp = P2.build_clim_plan2(doy, ndoy, w, q)
    if hp is None:
        return None, None, None, None
    s = cabi.plan2_struct(hp)
    ts = np.ascontiguousarray(ts, np.float32)
    T, ng = ts.shape
    thr = np.full((ndoy, ng), np.nan)
    se = np.full((ndoy, ng), np.nan)
    nz = np.zeros(ng, np.int32)
    assert lib.emul_clim_sweep2(_vp(ts), C.c_int64(T), C.c_int64(ng), C.byref(s), _vp(thr), _vp(se), _vp(nz)) == 0
    for k, d in enumerate(hp.exc_doy):
        rows = np.ascontiguousarray(hp.exc_rows[hp.exc_off[k]:hp.exc_off[k + 1]], np.int32)
        off = (int(d) - 1) * ng * 8
        assert lib.emul_clim_direct(_vp(ts), C.c_int64(ng), _vp(rows), C.c_int(len(rows)), C.c_int(hp.kp),
                                    C.c_double(q), C.c_void_p(thr.ctypes.data + off),
                                    C.c_void_p(se.ctypes.data + off), _vp(nz)) == 0
    return hp, thr, se, nz


CASES2 = CASES + [
    ("30yr_p95_w7", (1982, 2011), 16, 8000, 7, 95),
    ("13yr_from_leap", (1984, 1996), 16, 0, 5, 90),
    ("30yr_w3_p80", (1990, 2019), 8, 3000, 3, 80),
]


@pytest.mark.parametrize("phased", [False, True], ids=["step", "step_phased"])
@pytest.mark.parametrize("name,years,ncell,nan_ppm,w,pct", CASES2, ids=[c[0] for c in CASES2])
def test_topk_sweep_matches_oracle(em, name, years, ncell, nan_ppm, w, pct, phased):
    """The two-stack top-K sweep is bit-equal to the oracle wherever its plan accepts the calendar;
    the cases it declines (huge K, very wide windows) are the general sweep's (test above)."""
    tm = S.daily_time(*years)
    doy = S.doy366(tm)
    ts = S.synth_sst(len(tm), ncell, S.season_table(tm), nan_ppm=nan_ppm)
    if nan_ppm:
        ts[100:300, 3] = np.nan
        ts[:, 5] = np.nan
        ts[700:, 7] = np.nan
    hp, thr, se, nz = sweep2(em, ts, doy, 366, w, pct / 100.0, phased=phased)
    if name in ("70yr_two_pieces", "w15_p10"):
        assert hp is None
        return
    assert hp is not None and hp.smem_bytes() <= 227 * 1024
    oth, ose = O.threshold(ts, doy, 366, pctile=pct, windowHalfWidth=w, smoothPercentile=False, tstep=True)
    assert bit_equal(thr, oth)
    assert np.nanmax(np.abs(se - ose), initial=0) <= 1e-12
    assert np.array_equal(nz + (366 - hp.nsteps - len(hp.exc_doy)), np.isnan(oth).sum(axis=0))     # + absent labels
    if name == "30yr":
        assert hp.kp == 36 and list(hp.exc_doy) == [60] and hp.pool_rows <= 444 and len(hp.pat) <= 8      # 4 warps per SM


def test_topk_sweep_infinite_pentad_cube(em, oisst):
    tm = S.daily_time(2001, 2012)
    doy = S.doy366(tm)
    ts = S.synth_sst(len(tm), 6, S.season_table(tm))
    ts[500, 1] = np.inf
    ts[900, 2] = -np.inf
    ts[1300, 3] = np.inf
    ts[1302, 3] = -np.inf
    hp, thr, se, nz = sweep2(em, ts, doy, 366, 5, 0.9)
    with np.errstate(invalid="ignore"):
        oth, ose = O.threshold(ts, doy, 366, smoothPercentile=False, tstep=True)
    assert bit_equal(thr, oth)
    assert np.array_equal(np.isnan(se), np.isnan(ose)) and np.array_equal(np.isinf(se), np.isinf(ose))
    fin = np.isfinite(ose)
    assert np.abs(se[fin] - ose[fin]).max() <= 1e-12
    doy = np.tile(np.arange(1, 74), 30)                       # pentads, tstep calendar
    ts = S.synth_sst(len(doy), 16, S.season_table(len(doy)))
    hp, thr, se, nz = sweep2(em, ts, doy, 73, 5, 0.9)
    oth, ose = O.threshold(ts, doy, 73, smoothPercentile=False, tstep=True)
    assert hp is not None and len(hp.exc_doy) == 0
    assert bit_equal(thr, oth) and np.abs(se - ose).max() <= 1e-12
    d = O.add_doy(oisst["time"])                              # the reference's own test cube (2 years)
    cube = oisst["sst"].reshape(len(d), -1)
    hp, thr, se, nz = sweep2(em, cube, d, 366, 5, 0.9)
    oth, ose = O.threshold(cube, d, 366, smoothPercentile=False, tstep=True)
    assert hp is not None and bit_equal(thr, oth) and bit_equal(se, ose)


def test_sweep_infinite_samples(em):
    """+-inf samples poison the running window sum only while they are inside the window."""
    tm = S.daily_time(2001, 2012)
    doy = S.doy366(tm)
    ts = S.synth_sst(len(tm), 6, S.season_table(tm))
    ts[500, 1] = np.inf
    ts[900, 2] = -np.inf
    ts[1300, 3] = np.inf
    ts[1302, 3] = -np.inf
    hp, thr, se = sweep(em, ts, doy, 366, 5, 0.9)
    with np.errstate(invalid="ignore"):
        oth, ose = O.threshold(ts, doy, 366, smoothPercentile=False, tstep=True)
    assert bit_equal(thr, oth)
    assert np.array_equal(np.isnan(se), np.isnan(ose)) and np.array_equal(np.isinf(se), np.isinf(ose))
    fin = np.isfinite(ose)
    assert np.abs(se[fin] - ose[fin]).max() <= 1e-12 and (~fin).sum() >= 30


def test_sweep_pentad_and_reference_cube(em, oisst):
    doy = np.tile(np.arange(1, 74), 30)
    ts = S.synth_sst(len(doy), 16, S.season_table(len(doy)))
    hp, thr, se = sweep(em, ts, doy, 73, 5, 0.9)
    oth, ose = O.threshold(ts, doy, 73, smoothPercentile=False, tstep=True)
    assert bit_equal(thr, oth) and np.abs(se - ose).max() <= 1e-12
    d = O.add_doy(oisst["time"])
    cube = oisst["sst"].reshape(len(d), -1)
    hp, thr, se = sweep(em, cube, d, 366, 5, 0.9)
    oth, ose = O.threshold(cube, d, 366, smoothPercentile=False, tstep=True)
    assert bit_equal(thr, oth) and bit_equal(se, ose)


def test_plan_rejects_unsupported():
    with pytest.raises(NotImplementedError):
        P.build_clim_plan(np.tile(np.arange(1, 13), 3), 12, 6, 0.9)     # window wider than a year
    with pytest.raises(ValueError):
        P.build_clim_plan(np.array([0, 1, 2]), 3, 1, 0.9)
    lo, g = P.quantile_table(440, 0.9)
    olo, og = O.quantile_table(440, 0.9)
    assert np.array_equal(lo, olo) and np.array_equal(g, og)
    ptr, tidx = P.doy_csr(np.array([2, 1, 2, 3, 1]), 3)
    assert ptr.tolist() == [0, 2, 4, 5] and tidx.tolist() == [1, 4, 0, 2, 3]


def test_partition_sorted_and_front_helpers(em):
    """The branch-free building blocks of the sweep walk against plain numpy: placement of the cut
    in a sorted list (halving search through selects) and the sorted 4-entry front."""
    lib, _ = em
    rng = np.random.default_rng(11)
    out = np.zeros(3, np.int32)
    for trial in range(600):
        n = (8, 32, 40)[trial % 3]
        nvalid = int(rng.integers(0, n + 1))
        k = np.zeros(n, np.uint32)
        k[:nvalid] = np.sort(rng.integers(1, 50 if trial % 2 else 2 ** 32 - 1, nvalid, dtype=np.uint64))[::-1]
        piv = np.uint32(rng.choice([0, 0xffffffff, int(rng.integers(0, 2 ** 32 - 1)), int(k[rng.integers(0, n)])]))
        lib.emul_partition_sorted(_vp(k), n, C.c_uint32(int(piv)), _vp(out))
        ptr = int((k > piv).sum())
        assert out[0] == ptr
        assert np.uint32(out[1]) == (k[ptr - 1] if ptr > 0 else np.uint32(0xffffffff))
        assert np.uint32(out[2]) == (k[ptr] if ptr < n else np.uint32(0))
    f = np.zeros(4, np.uint32)
    g = np.zeros(4, np.int32)
    for trial in range(400):
        n = int(rng.integers(0, 14))
        keys = rng.integers(0, 40 if trial % 2 else 2 ** 32 - 2, n, dtype=np.uint64).astype(np.uint32)
        tags = np.arange(1, n + 1, dtype=np.int32)
        nrep = int(rng.integers(0, 6))
        rkeys = rng.integers(0, 60, nrep, dtype=np.uint64).astype(np.uint32)
        rkeys[rng.random(nrep) < 0.4] = 0xffffffff                      # "nothing enters"
        rtags = np.arange(100, 100 + nrep, dtype=np.int32)
        lib.emul_front(_vp(keys), _vp(tags), n, _vp(rkeys), _vp(rtags), nrep, _vp(f), _vp(g))
        # reference: stable sort by key (earlier entries first among ties), keep 4; replace = drop head, insert
        ref = sorted(zip(keys.tolist(), tags.tolist()), key=lambda kv: kv[0])[:4]
        for rk, rt in zip(rkeys.tolist(), rtags.tolist()):
            ref = ref[1:]
            if rk != 0xffffffff:
                pos = sum(1 for kv in ref if kv[0] <= rk)
                ref.insert(pos, (rk, rt))
            ref = ref[:4]
        for i, (kk, tt) in enumerate(ref):
            assert f[i] == kk and g[i] == tt, (trial, i)
        assert np.all(f[len(ref):] == 0xffffffff)


def test_run_finder_fuzz(em):
    lib, _ = em
    rng = np.random.default_rng(5)
    for trial in range(2500):
        T = int(rng.integers(1, 200)) if trial < 1500 else int(rng.integers(60, 700))
        b = (rng.random(T) < rng.uniform(0.05, 0.95)).astype(np.uint8)
        if trial % 2:
            b = np.repeat(b, rng.integers(1, 6))[:T].copy()
            T = len(b)
        minD = int(rng.integers(1, 8))
        maxG = int(rng.integers(0, max(1, minD)))
        join = int(rng.integers(0, 2))
        s = np.zeros(T + 1, np.int32)
        e = np.zeros(T + 1, np.int32)
        n = lib.emul_find_events(_vp(b), T, minD, join, maxG, _vp(s), _vp(e))
        so, eo = O.find_events(b.astype(bool), minD, bool(join), maxG)
        assert n == len(so) and np.array_equal(s[:n], so) and np.array_equal(e[:n], eo)
        # eager emission (fused detect kernel): identical events, each emitted no later than
        # maxGap + 32 steps after the word in which it became final
        s2 = np.zeros(T + 1, np.int32)
        e2 = np.zeros(T + 1, np.int32)
        at = np.zeros(T + 1, np.int32)
        n2 = lib.emul_find_events_eager(_vp(b), T, minD, join, maxG, _vp(s2), _vp(e2), _vp(at))
        assert n2 == n and np.array_equal(s2[:n], so) and np.array_equal(e2[:n], eo)
        assert np.all(at[:n] > eo) or T in at[:n]                  # never before the event has ended


def test_event_stats_vs_reference_tables(em, ref_cases):
    lib, cabi = em
    nev = 0
    for c in ref_cases:
        ts, th, se = c["ts"], c["th"], c["se"]
        minD, join, maxG = (int(v) for v in c["par"])
        T = len(ts)
        doy = np.arange(1, T + 1, dtype=np.int32)
        b = (ts.astype(np.float64) > th).astype(np.uint8)
        s = np.zeros(T + 1, np.int32)
        e = np.zeros(T + 1, np.int32)
        n = lib.emul_find_events(_vp(b), T, minD, join, maxG, _vp(s), _vp(e))
        assert n == len(c["index_start"])
        for k in range(n):
            oi = np.zeros(cabi.EI_COUNT, np.int32)
            of = np.zeros(cabi.EF_COUNT)
            lib.emul_event_stats(_vp(ts), _vp(th), _vp(se), _vp(doy), C.c_int64(1), T, int(s[k]), int(e[k]),
                                 _vp(oi), _vp(of))
            for j, f in enumerate(cabi.EI_FIELDS[1:], start=1):
                ref = c[f][k]
                assert oi[j] == (-1 if np.isnan(ref) else ref), f
            for j, f in enumerate(cabi.EF_FIELDS):
                tol = 5e-6 if f.endswith("_abs") else 1e-9
                np.testing.assert_allclose(of[j], c[f][k], rtol=tol, atol=tol, equal_nan=True, err_msg=f)
            nev += 1
    assert nev > 500
