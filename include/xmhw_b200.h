/* xmhw_b200 -- C ABI of the B200 (sm_100a) marine-heatwave hot path.
 *
 * Drop-in boundary for the two per-cell loops of coecms/xmhw v0.9.3 (the
 * reference has no FFI of its own; these entry points are what a ctypes binding
 * inside the reference would call instead of its Python loops -- see
 * INTEGRATION.md for the stub):
 *
 *   xmhw/xmhw.py:184-197   for c in ts.cell: calc_clim(...)      + dask.compute
 *   xmhw/xmhw.py:440-454   for c in ts.cell: define_events(...)  + dask.compute
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller; the library never
 *     allocates or frees and keeps no global state (re-entrant);
 *   - all work is enqueued on the caller's cudaStream_t (`stream`, may be 0); no
 *     call synchronises;
 *   - return value: 0 = ok, negative = argument error (XMHW_E_*), positive =
 *     cudaError_t of the failed launch; no C++ exception crosses the boundary;
 *   - layouts are the reference's own: the series is (time, cell) row-major
 *     float32 exactly as xarray holds (time, lat, lon) after
 *     `stack(cell=sorted(dims))` (identify.py:520), climatologies are (doy, cell)
 *     row-major float64 (the reference's `thresh`/`seas` variables, xmhw.py:204-216).
 *     `cell` runs over the WHOLE grid, land included: land cells hold NaN
 *     (the reference drops them, identify.py:522-525; the host wrapper does the
 *     same from `nvalid`);
 *   - indices are 0-based positions along time like the reference's idxarr
 *     (xmhw.py:417); doy values are 1-based labels (identify.py:67-76).
 */
#ifndef XMHW_B200_H
#define XMHW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XMHW_E_ARG   (-1)   /* null pointer / non-positive size            */
#define XMHW_E_PLAN  (-2)   /* inconsistent climatology plan               */
#define XMHW_E_SMEM  (-3)   /* plan needs more shared memory than one SM has */

#define XMHW_ABI_VERSION 3

/* event table columns (struct-of-arrays, column c of event i at [c * cap + i]) */
enum xmhw_event_i32 {
  XMHW_EI_CELL = 0,        /* flat grid cell                                  */
  XMHW_EI_INDEX_START,     /* features.py:116                                  */
  XMHW_EI_INDEX_END,       /* features.py:117                                  */
  XMHW_EI_INDEX_PEAK,      /* features.py:120,181                              */
  XMHW_EI_DURATION,        /* features.py:189                                  */
  XMHW_EI_CATEGORY,        /* features.py:147,188 (-1 = undefined)             */
  XMHW_EI_DURATION_MODERATE, XMHW_EI_DURATION_STRONG,
  XMHW_EI_DURATION_SEVERE, XMHW_EI_DURATION_EXTREME,   /* features.py:148-151  */
  XMHW_EI_COUNT
};
enum xmhw_event_f64 {
  XMHW_EF_INTENSITY_MAX = 0, XMHW_EF_INTENSITY_MEAN, XMHW_EF_INTENSITY_CUMULATIVE, XMHW_EF_INTENSITY_VAR,
  XMHW_EF_SEVERITY_MAX, XMHW_EF_SEVERITY_MEAN, XMHW_EF_SEVERITY_CUMULATIVE, XMHW_EF_SEVERITY_VAR,
  XMHW_EF_INTENSITY_MAX_RELTHRESH, XMHW_EF_INTENSITY_MEAN_RELTHRESH,
  XMHW_EF_INTENSITY_CUMULATIVE_RELTHRESH, XMHW_EF_INTENSITY_VAR_RELTHRESH,
  XMHW_EF_INTENSITY_MAX_ABS, XMHW_EF_INTENSITY_MEAN_ABS,      /* float32-rounded, features.py:68 */
  XMHW_EF_INTENSITY_CUMULATIVE_ABS, XMHW_EF_INTENSITY_VAR_ABS,
  XMHW_EF_RATE_ONSET, XMHW_EF_RATE_DECLINE,                   /* features.py:290-291 */
  XMHW_EF_COUNT
};

/* Climatology sweep plan: which time rows form each sorted list, when lists
 * enter / leave the +-windowHalfWidth window as the day-of-year advances, and
 * numpy's linear-quantile index table.  Built on the host from the doy vector
 * (xmhw_b200/plan.py); all arrays are device-resident int32 unless noted.     */
typedef struct xmhw_clim_plan {
  int32_t nsteps;               /* = ndoy                                       */
  int32_t pool_rows;            /* shared-memory rows (128 B) per 32-cell warp  */
  int32_t nmax;                 /* q tables have nmax + 1 entries               */
  int32_t max_size;             /* largest list, <= 32                          */
  int32_t scratch_rows;         /* global scratch rows (128 B) per 32-cell warp  */
  int32_t reserved_;
  const int32_t* inst_base;     /* [ninst]                                      */
  const int32_t* inst_size;     /* [ninst] time rows of the list (1..32)         */
  const int32_t* inst_keep;     /* [ninst] key rows held in shared memory        */
  const int32_t* inst_sbase;    /* [ninst] first scratch row of the keys past keep */
  const int32_t* inst_row_off;  /* [ninst]                                      */
  const int32_t* rows;          /* time indices                                 */
  const int32_t* leave_off;     /* [nsteps+1]                                   */
  const int32_t* leave;
  const int32_t* enter_off;     /* [nsteps+1]                                   */
  const int32_t* enter;
  const int32_t* use_off;       /* [nsteps+1]                                   */
  const int32_t* use;
  const int32_t* step_rec;      /* [nsteps][32] fixed-size step records          */
  double q;                     /* quantile in [0,1] (numpy 'linear': (n-1) q)   */
} xmhw_clim_plan;

/* Plan of the two-stack top-K climatology sweep (xmhw_b200/plan2.py, csrc/xmhw_topk.h): the time
 * rows ordered into atoms (rows that enter and leave the doy windows together) such that every
 * doy's window is a contiguous range of atoms; per doy a fixed-size step record says which unit
 * slots leave, which atoms enter and where the front array of the query lives.  ONE plain HOST
 * struct (no pointers): the library copies it into the kernel's launch parameters, so the kernel
 * reads the plan with constant loads and there is no plan array in device memory.
 * Word layouts: csrc/xmhw_topk.h.                                                            */
#define XMHW_SC_MAX_STEPS 366
#define XMHW_SC_REC_WORDS 12
#define XMHW_SC_MAX_FLIP  768
#define XMHW_SC_MAX_PAT   16
#define XMHW_SC_PAT_LEN   48
#define XMHW_SC_MAX_INIT  32
typedef struct xmhw_clim_plan2 {
  int32_t nsteps;               /* sweep steps (doys computed by the sweep)        */
  int32_t kp;                   /* top-K capacity: 8, 16, 24, 36 or 48             */
  int32_t max_size;             /* rows of the largest atom, <= 48                 */
  int32_t slot_rows;            /* shared-memory rows (128 B) per unit slot = cap + 3 */
  int32_t nslots;               /* unit slots per 32-cell warp, <= 32              */
  int32_t n_init;               /* atoms pushed before the first step              */
  int32_t cap;                  /* key rows per slot                               */
  int32_t reserved_;
  double q;                     /* quantile in [0,1] (numpy 'linear': (n-1) q)     */
  uint32_t rec[XMHW_SC_MAX_STEPS][XMHW_SC_REC_WORDS];   /* step records            */
  uint32_t flip[XMHW_SC_MAX_FLIP];                      /* flip entries            */
  int32_t pat[XMHW_SC_MAX_PAT][XMHW_SC_PAT_LEN];        /* row patterns            */
  uint32_t init[XMHW_SC_MAX_INIT][2];                   /* atoms of the first window */
} xmhw_clim_plan2;

int xmhw_abi_version(void);
const char* xmhw_strerror(int code);

/* identify.py:184-270 window_roll + calculate_thresh + calculate_seas (before the
 * Feb-29 rule and smoothing) for every grid cell: the general sorted-list sweep (any calendar,
 * any quantile).  ts [T][ngrid] f32 -> thresh_raw, seas_raw [nsteps][ngrid] f64 (NaN = no
 * sample), nempty [ngrid] i32 = number of doys without any sample.
 * scratch: caller-owned workspace of ceil(ngrid/32) * plan->scratch_rows * 128 bytes
 * (sorted list tails; stays L2-resident while a warp needs it).                       */
int xmhw_clim_sweep_f32(const float* ts, int64_t T, int64_t ngrid, const xmhw_clim_plan* plan,
                        double* thresh_raw, double* seas_raw, int32_t* nempty, uint32_t* scratch, void* stream);

/* Same reference lines (identify.py:184-270) by the two-stack top-K sweep: every doy is the same
 * straight-line sorting / merging network code for all cells (no data-dependent walk).  Writes
 * the rows of thresh_raw / seas_raw [ndoy][ngrid] named by the plan's step records and
 * nempty [ngrid] i32 = number of those doys without any sample.  `plan` is a HOST pointer.
 * The few doys the plan excludes (doy 60 of the 366-day calendar: its window holds leap years
 * only) are computed by xmhw_clim_direct_f32 from their row list: rows [nrows] i32 time indices,
 * thresh_row / seas_row = that doy's row of the raw arrays, nempty += 1 where it has no sample.
 * group_order: NULL, or a device array [ceil(ngrid/32) + 1] i32: a permutation of the 32-cell groups = the
 * order the warps take them in (results do not depend on it), followed by one word the library uses as
 * a work ticket (zeroed by the library; only the persistent launch mode of the tensor-memory kernel, a
 * development option, draws groups from it).  The warps of a block advance in lockstep, so a block
 * whose groups have equal work wastes nothing: callers put the groups that look like land (all NaN in a
 * probe row) last.  When fewer than 8 warps' unit slots fit the shared memory of one SM (default window:
 * 4), the library launches the variant that keeps the remaining slots in tensor memory (tcgen05.ld / st),
 * 8 warps per SM; XMHW_B200_SWEEP2_TMEM=0 in the environment turns that off.                           */
int xmhw_clim_sweep2_f32(const float* ts, int64_t T, int64_t ngrid, const xmhw_clim_plan2* plan,
                         double* thresh_raw, double* seas_raw, int32_t* nempty, int32_t* group_order, void* stream);
/* A processing order for xmhw_clim_sweep2_f32: the 32-cell groups that hold data in at least one of three
 * probe rows (first, middle, last time step) first, the all-NaN ("land") groups last, each half in grid
 * order.  flags: workspace of ceil(ngrid/32) bytes; order: ceil(ngrid/32) + 1 i32 (the last word is the
 * sweep's work ticket, see xmhw_clim_sweep2_f32).                                                         */
int xmhw_group_order_f32(const float* ts, int64_t T, int64_t ngrid, uint8_t* flags, int32_t* order, void* stream);
int xmhw_clim_direct_f32(const float* ts, int64_t T, int64_t ngrid, const int32_t* rows, int32_t nrows, int32_t kp,
                         double q, double* thresh_row, double* seas_row, int32_t* nempty, void* stream);

/* identify.py:137-151 feb29 (if feb29 != 0: doy 60 <- mean of the doys 59,60,61 present) then
 * identify.py:154-181 runavg (circular centred mean, odd smooth_width; <= 1 = off).
 * raw, out [ndoy][ngrid] f64, out must not alias raw.  nempty [ngrid] (from the sweep): a cell
 * in which some doys have no sample is smoothed over its own compacted doy axis, as the
 * reference's groupby output lacks those doys (identify.py:175-180, :233-241); they stay NaN. */
int xmhw_clim_finish_f64(const double* raw, double* out, int32_t ndoy, int64_t ngrid,
                         int32_t feb29, int32_t smooth_width, const int32_t* nempty, void* stream);
/* same for thresh and seas in one call (one fused launch for the default width 31) */
int xmhw_clim_finish2_f64(const double* thresh_raw, double* thresh_out, const double* seas_raw,
                          double* seas_out, int32_t ndoy, int64_t ngrid, int32_t feb29,
                          int32_t smooth_width, const int32_t* nempty, void* stream);

/* identify.py:367-372: bthresh = ts > thresh[doy] (strict, float64 compare, NaN -> false).
 * doy_ptr [ndoy+1], doy_tidx [T]: CSR of time indices per doy label.
 * mask [ceil(ngrid/32)][T] u32: bit l of word (g, t) = exceedance of cell 32 g + l at t.
 * nvalid [ngrid] i32 (pre-zeroed): += number of non-NaN samples per cell
 * (land_check, identify.py:522-525).                                            */
int xmhw_exceed_mask_f32(const float* ts, int64_t T, int64_t ngrid, const int32_t* doy_ptr,
                         const int32_t* doy_tidx, int32_t ndoy, const double* thresh,
                         uint32_t* mask, int32_t* nvalid, void* stream);

/* identify.py:415-479 mhw_filter + :273-325 join_gaps: events per cell.
 * Phase 1 counts, caller scans, phase 2 fills rows CELL/INDEX_START/INDEX_END of ev_i32. */
int xmhw_events_count(const uint32_t* mask, int64_t T, int64_t ngrid, int32_t min_duration,
                      int32_t join_gaps, int32_t max_gap, int32_t* counts, void* stream);
/* offsets [n+1] i64 exclusive prefix sum of counts [n]; scratch [n/1024 + 2] i64. */
int xmhw_exclusive_scan_i32(const int32_t* counts, int64_t n, int64_t* offsets, int64_t* scratch,
                            void* stream);
int xmhw_events_fill(const uint32_t* mask, int64_t T, int64_t ngrid, int32_t min_duration,
                     int32_t join_gaps, int32_t max_gap, const int64_t* offsets, int64_t cap,
                     int32_t* ev_i32, void* stream);
/* One-pass variant of count + fill (same reference lines, identify.py:415-479, :273-325):
 * the count pass also parks the first cap_per_cell (start, end) pairs of every cell in
 * stage [ceil(ngrid/32)][2][cap_per_cell][32] i32 and ORs 1 into *overflow when a cell has
 * more; after the scan, xmhw_events_gather copies the staged pairs to ev_i32 at the
 * offsets.  If *overflow is set the caller runs xmhw_events_fill instead (exact path).   */
int xmhw_events_count_stage(const uint32_t* mask, int64_t T, int64_t ngrid, int32_t min_duration,
                            int32_t join_gaps, int32_t max_gap, int32_t* counts, int32_t* stage,
                            int32_t cap_per_cell, int32_t* overflow, void* stream);
int xmhw_events_gather(const int32_t* stage, int32_t cap_per_cell, const int32_t* counts,
                       const int64_t* offsets, int64_t ngrid, int64_t cap, int32_t* ev_i32,
                       void* stream);

/* Fused detect: ONE time-major pass over the series replaces xmhw_exceed_mask_f32 + xmhw_events_count_stage
 * + xmhw_events_gather + xmhw_event_stats_cm_f32 (identify.py:367-479, :273-325, features.py:22-295; the loop
 * xmhw.py:440-454).  A block owns 32 adjacent cells for the whole time axis: threshold compare against
 * float32 round-down thresholds held in shared memory, per-cell run rules, and the statistics of every
 * event while its rows are still L2-resident.  Records are appended in completion order to the staging
 * table stage_i32 [XMHW_EI_COUNT + 1][stage_cap] (extra last row: ordinal of the event in its cell) and
 * stage_f64 [XMHW_EF_COUNT][stage_cap]; counts [ngrid] and nvalid [ngrid] are written, counter[0] = number
 * of staged records, counter[1] != 0 when stage_cap was too small (the caller then uses the kernel chain).
 * After xmhw_exclusive_scan_i32 of counts, xmhw_events_scatter places the records at offsets[cell] + ordinal
 * in the (cell, start) ordered table.  clim_cm: xmhw_clim_cellmajor_f64.                                  */
int xmhw_detect_fused_f32(const float* ts, int64_t T, int64_t ngrid, const int32_t* doy, int32_t ndoy,
                          const double* thresh, const double* clim_cm, int32_t min_duration, int32_t join_gaps,
                          int32_t max_gap, int32_t* counts, int32_t* nvalid, int32_t* stage_i32, double* stage_f64,
                          int64_t stage_cap, int32_t* counter, void* stream);
int xmhw_events_scatter(const int32_t* stage_i32, const double* stage_f64, int64_t stage_cap, int64_t nstaged,
                        const int64_t* offsets, int64_t cap, int32_t* ev_i32, double* ev_f64, void* stream);

/* features.py:22-295 mhw_df + agg_df + properties + onset_decline for nev events.
 * doy [T] i32 (1-based labels); thresh, seas [ndoy][ngrid] f64;
 * ev_i32 [XMHW_EI_COUNT][cap], ev_f64 [XMHW_EF_COUNT][cap].                       */
int xmhw_event_stats_f32(const float* ts, int64_t T, int64_t ngrid, const int32_t* doy,
                         const double* thresh, const double* seas, int64_t nev, int64_t cap,
                         int32_t* ev_i32, double* ev_f64, void* stream);

/* Same statistics from a cell-major interleaved copy of the climatologies,
 * clim_cm[(cell * ndoy + d) * 2 + {0, 1}] = {thresh, seas}[d][cell] (16-byte aligned,
 * 2 * ndoy * ngrid doubles, built by xmhw_clim_cellmajor_f64): the consecutive days of an
 * event then read one contiguous run instead of two scattered 32-byte sectors per day.
 * Replaces the same reference lines as xmhw_event_stats_f32 (features.py:22-295; the
 * per-timestep label lookup th.sel(doy=ts.doy), identify.py:367-368). */
int xmhw_clim_cellmajor_f64(const double* thresh, const double* seas, int32_t ndoy, int64_t ngrid,
                            double* clim_cm, void* stream);
int xmhw_event_stats_cm_f32(const float* ts, int64_t T, int64_t ngrid, const int32_t* doy, int32_t ndoy,
                            const double* clim_cm, int64_t nev, int64_t cap,
                            int32_t* ev_i32, double* ev_f64, void* stream);

/* intermediate=True (identify.py:404-411): dense per-timestep fields of mhw_df
 * (features.py:22-69), every array [T][ngrid]; `events` must be pre-filled with NaN by the
 * caller (it receives the event label = start index on event days).  T <= 65535.           */
typedef struct xmhw_intermediate {
  double* events; double* seas; double* thresh; double* relSeas; double* relThresh;
  double* relThreshNorm; double* severity; double* cats; float* mabs;
  uint8_t* bthresh; uint8_t* duration_moderate; uint8_t* duration_strong;
  uint8_t* duration_severe; uint8_t* duration_extreme;
} xmhw_intermediate;
int xmhw_intermediate_f32(const float* ts, int64_t T, int64_t ngrid, const int32_t* doy,
                          const double* thresh, const double* seas, const int32_t* ev_i32,
                          int64_t nev, int64_t cap, const xmhw_intermediate* out, void* stream);

/* Downstream statistics of the reference's stats.py over the compact event table (ordered by cell,
 * then start; offsets [ngrid + 1] from the detect chain).  block_of_t [T] i32 maps every time step to
 * its block of years (0 .. nblocks-1, or -1 = outside the period), built on the host.
 *
 * xmhw_block_average (stats.py:27-183, agg_mhw :322-368): out [XMHW_BA_NCOL][nblocks][ngrid] f64 --
 * event count, NaN-skipping means of the 17 properties listed at XMHW_BA_MEAN0, the block maximum
 * of intensity_max and the block sum of intensity_cumulative; an event belongs to the block of its
 * start (use_peak = 0) or peak (1) day.
 * xmhw_block_ts_f32 (agg_ts :400-428): out [3][nblocks][ngrid] = mean, max, min of the series.
 * xmhw_block_cat_days_f32 (agg_cats :371-398): days [4][nblocks][ngrid] i32 = event days per category
 * (moderate, strong, severe, extreme), each day counted in the block of ITS OWN year.
 * xmhw_event_rank_f64 (mhw_rank :446-510): rank [nev] of every event within its cell for one property
 * column, 1 = largest, numpy semantics len - argsort(argsort(x)).                                */
#define XMHW_BA_COUNT_COL  0
#define XMHW_BA_MEAN0      1     /* duration, intensity_{max,mean,var,cumulative}, the same four _relThresh, */
#define XMHW_BA_NMEAN      17    /* the same four _abs, severity_{mean,cumulative}, rate_onset, rate_decline  */
#define XMHW_BA_IMAX_MAX   18
#define XMHW_BA_TOTAL_ICUM 19
#define XMHW_BA_NCOL       20
int xmhw_block_average(const int32_t* ev_i32, const double* ev_f64, int64_t cap, const int64_t* offsets,
                       int64_t ngrid, const int32_t* block_of_t, int32_t nblocks, int32_t use_peak,
                       double* out, void* stream);
int xmhw_block_ts_f32(const float* ts, int64_t T, int64_t ngrid, const int32_t* block_of_t, int32_t nblocks,
                      double* out, void* stream);
int xmhw_block_cat_days_f32(const float* ts, int64_t T, int64_t ngrid, const int32_t* doy, const double* thresh,
                            const double* seas, const int32_t* ev_i32, int64_t nev, int64_t cap,
                            const int32_t* block_of_t, int32_t nblocks, int32_t* days, void* stream);
int xmhw_event_rank_f64(const double* col, const int32_t* cells, const int64_t* offsets, int64_t nev,
                        double* rank, void* stream);

/* land_check census (identify.py:522-525): nvalid [ngrid] i32 = number of non-NaN samples of every
 * cell of ts [T][ngrid] (the array is zeroed by the call).                                     */
int xmhw_count_valid_f32(const float* ts, int64_t T, int64_t ngrid, int32_t* nvalid, void* stream);

/* Pre-step of both public functions (xmhw.py:159-160, :409-410, maxPadLength): in-place linear
 * interpolation along time of interior NaN runs of at most max_pad steps, per cell.        */
int xmhw_interp_gaps_f32(float* ts, int64_t T, int64_t ngrid, int32_t max_pad, void* stream);

/* Strided 2-D copy (cudaMemcpy2DAsync) used to move a column block of the (time, cell) host
 * array to / from the device: kind 0 = host->device, 1 = device->host, 2 = device->device. */
int xmhw_copy2d_async(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch,
                      int64_t width_bytes, int64_t height, int32_t kind, void* stream);

/* Deterministic synthetic SST: seasonal cycle + AR(1) noise rounded to
 * 0.01 degC, NaN on land; bit-identical to xmhw_b200/synth.py on the host.
 * season [T + 366] f64 host-computed sine table, land [ngrid] u8 (1 = land).
 * coherent >= 1: blocks of that many consecutive cells share mean/amplitude/phase
 * (1 = independent cells, the benchmark default; the noise is always per cell).   */
int xmhw_synth_sst_f32(float* ts, int64_t T, int64_t ngrid, int64_t cell0, const uint8_t* land,
                       const double* season, uint64_t seed, double rho, double sigma,
                       double noise_scale, uint32_t nan_per_million, uint32_t coherent, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XMHW_B200_H */
