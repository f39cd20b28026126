"""Shared helpers for the parity tests (comparison rules of SURVEY.md 8c)."""
import numpy as np

INT_FIELDS = ("cell", "index_start", "index_end", "index_peak", "duration", "category",
              "duration_moderate", "duration_strong", "duration_severe", "duration_extreme")
F32_FIELDS = ("intensity_max_abs", "intensity_mean_abs", "intensity_cumulative_abs", "intensity_var_abs")

# float tolerance from BASELINE.json north_star: 1e-5 degC absolute / 1 float32 ulp relative
ABS_TOL = 1e-5
REL_TOL = 2.0 ** -23


def bit_equal(a, b):
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    nan = np.isnan(a) & np.isnan(b)
    return a.shape == b.shape and bool(np.all((a.view(np.int64) == b.view(np.int64)) | nan))


def assert_events_match(got, exp, float_fields):
    """Integer/index outputs bit-exact, floats within max(1e-5 abs, 1 f32 ulp rel)."""
    assert len(got["cell"]) == len(exp["cell"]), "event count %d != %d" % (len(got["cell"]), len(exp["cell"]))
    for f in INT_FIELDS:
        assert np.array_equal(np.asarray(got[f], np.int64), np.asarray(exp[f], np.int64)), f
    for f in float_fields:
        g, e = np.asarray(got[f], np.float64), np.asarray(exp[f], np.float64)
        assert np.array_equal(np.isnan(g), np.isnan(e)), f
        inf = np.isinf(e)
        assert np.array_equal(g[inf], e[inf]), f
        ok = ~np.isnan(e) & ~inf
        err = np.abs(g[ok] - e[ok])
        lim = np.maximum(ABS_TOL, REL_TOL * np.abs(e[ok]))
        assert np.all(err <= lim), "%s: max err %.3e" % (f, err.max() if err.size else 0)
