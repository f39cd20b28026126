"""Names and small host-side pieces of the reference's xmhw/features.py.  The per-event
arithmetic (mhw_df, agg_df, properties, onset_decline: features.py:22-295) runs in the
`xmhw_event_stats_f32` kernel; this module holds the variable list the reference's
Dataset carries and `flip_cold`."""

# reference column order: features.py:115-151 then :181-189, :290-291
EVENT_VARIABLES = (
    "event", "index_start", "index_end", "time_start", "time_end", "time_peak",
    "intensity_max", "intensity_mean", "intensity_cumulative",
    "severity_max", "severity_mean", "severity_cumulative", "severity_var",
    "intensity_mean_relThresh", "intensity_cumulative_relThresh",
    "intensity_mean_abs", "intensity_cumulative_abs",
    "duration_moderate", "duration_strong", "duration_severe", "duration_extreme",
    "index_peak", "intensity_var", "intensity_max_relThresh", "intensity_max_abs",
    "intensity_var_relThresh", "intensity_var_abs", "category", "duration",
    "rate_onset", "rate_decline")
FLOAT32_VARIABLES = ("intensity_mean_abs", "intensity_cumulative_abs", "intensity_max_abs", "intensity_var_abs")


def flip_cold(columns):
    """features.py:298-315: negate variables whose name contains "intensity" but not "_var"."""
    for name in columns:
        if "intensity" in name and "_var" not in name:
            columns[name] = -1 * columns[name]
    return columns
