// Per-lane algorithms of the xmhw_b200 kernels.
//
// Every kernel in xmhw_kernels.cu maps ONE GRID CELL TO ONE LANE (32 adjacent
// cells per warp), because the reference's arrays are (time, lat, lon) /
// (doy, lat, lon): with lane = cell every global access of a warp is one
// contiguous 128-byte (f32) or 256-byte (f64) row segment and no transpose of
// the 45 GB input is ever needed.  Lanes never exchange data (only ballots for
// warp-uniform early exits), so the per-lane logic lives here as plain
// __host__ __device__ functions.  The CUDA kernels are the only product users;
// tests/lane_emulator compiles the same header with g++ to exercise the logic
// one lane at a time on the GPU-less build box (test infrastructure only).
//
// Reference semantics (paths relative to the upstream checkout):
//   window pooling / quantile / mean   xmhw/identify.py:184-270
//   run-length encoding / gap joining  xmhw/identify.py:273-325, :415-479
//   per-event statistics               xmhw/features.py:22-295
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#ifdef __CUDACC__
#define XMHW_HD __host__ __device__ __forceinline__
#else
#define XMHW_HD inline
#endif

#include "sortnet_gen.h"

namespace xmhw {

// ---------------------------------------------------------------------------
// bit casts and order-preserving keys
// ---------------------------------------------------------------------------
XMHW_HD uint32_t f32_bits(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
XMHW_HD float bits_f32(uint32_t u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
XMHW_HD uint32_t f64_lo(double d) {
#ifdef __CUDA_ARCH__
  return (uint32_t)__double2loint(d);
#else
  uint64_t u; memcpy(&u, &d, 8); return (uint32_t)u;
#endif
}
XMHW_HD uint32_t f64_hi(double d) {
#ifdef __CUDA_ARCH__
  return (uint32_t)__double2hiint(d);
#else
  uint64_t u; memcpy(&u, &d, 8); return (uint32_t)(u >> 32);
#endif
}
XMHW_HD double f64_from(uint32_t lo, uint32_t hi) {
#ifdef __CUDA_ARCH__
  return __hiloint2double((int)hi, (int)lo);
#else
  uint64_t u = ((uint64_t)hi << 32) | lo; double d; memcpy(&d, &u, 8); return d;
#endif
}
XMHW_HD double qnan() { return f64_from(0u, 0x7ff80000u); }

// float -> uint32 whose unsigned order equals the float order; NaN -> 0 (below
// every real value: the smallest real key is key(-inf) = 0x007fffff).
XMHW_HD uint32_t f32_key(float f) {
  uint32_t b = f32_bits(f);
  bool valid = (b & 0x7fffffffu) <= 0x7f800000u;
  uint32_t k = b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
  return valid ? k : 0u;
}
XMHW_HD float key_f32(uint32_t k) {
  return bits_f32((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}
XMHW_HD int ctz32(uint32_t x) {
#ifdef __CUDA_ARCH__
  return __ffs((int)x) - 1;
#else
  return __builtin_ctz(x);
#endif
}

// ---------------------------------------------------------------------------
// climatology sweep plan (built on the host by xmhw_b200/plan.py)
// ---------------------------------------------------------------------------
// A "list instance" is a set of time rows that belong to exactly the same
// day-of-year windows (normally: all years of one calendar day).  It is loaded
// and sorted once, kept in the warp's shared-memory pool while the doy sweep
// needs it, and the window of doy d is the disjoint union of the instances in
// use[d].  This representation is derived from the actual doy vector, so leap
// days, series edges, pentad/monthly steps and any window half-width are all
// the same code path.
struct ClimPlan {
  int32_t nsteps;                // sweep steps (= ndoy), step s computes doy s+1
  int32_t pool_rows;             // shared-memory rows (32 words each) per warp
  int32_t nmax;                  // max samples per window (size of q tables - 1)
  int32_t max_size;              // largest instance (<= 32)
  const int32_t* inst_base;      // [ninst] first pool row of the instance block
  const int32_t* inst_size;      // [ninst] number of time rows (1..32)
  const int32_t* inst_row_off;   // [ninst] offset into rows[]
  const int32_t* rows;           // time indices
  const int32_t* leave_off;      // [nsteps+1]  -> leave[] (pool base rows)
  const int32_t* leave;
  const int32_t* enter_off;      // [nsteps+1]  -> enter[] (instance id | load flag << 30)
  const int32_t* enter;
  const int32_t* use_off;        // [nsteps+1]  -> use[] (pool base rows)
  const int32_t* use;
  const int32_t* q_lo;           // [nmax+1] floor((n-1) q)  (numpy 'linear')
  const double* q_gamma;         // [nmax+1] fractional part
};

// Instance block layout in the pool (row = 32 words, word index = lane):
//   row 0      meta: len | ptr << 8   (len = valid samples, ptr = #keys above the cut)
//   row 1, 2   f64 sum of the valid samples (lo, hi words)
//   row 3 + r  r-th largest key (r < size); invalid samples are key 0 at the end
enum { POOL_META = 0, POOL_SUM = 1, POOL_KEYS = 3 };

#ifdef __CUDA_ARCH__
#define XMHW_LDG(p) __ldg(p)
#else
#define XMHW_LDG(p) (*(p))
#endif

#define XMHW_CE(i, j) { uint32_t hi_ = k[i] > k[j] ? k[i] : k[j]; uint32_t lo_ = k[i] > k[j] ? k[j] : k[i]; k[i] = hi_; k[j] = lo_; }

template <int N> XMHW_HD void sort_desc(uint32_t (&k)[N]);
template <> XMHW_HD void sort_desc<4>(uint32_t (&k)[4]) { XMHW_SORTNET_4 }
template <> XMHW_HD void sort_desc<8>(uint32_t (&k)[8]) { XMHW_SORTNET_8 }
template <> XMHW_HD void sort_desc<12>(uint32_t (&k)[12]) { XMHW_SORTNET_12 }
template <> XMHW_HD void sort_desc<16>(uint32_t (&k)[16]) { XMHW_SORTNET_16 }
template <> XMHW_HD void sort_desc<20>(uint32_t (&k)[20]) { XMHW_SORTNET_20 }
template <> XMHW_HD void sort_desc<24>(uint32_t (&k)[24]) { XMHW_SORTNET_24 }
template <> XMHW_HD void sort_desc<28>(uint32_t (&k)[28]) { XMHW_SORTNET_28 }
template <> XMHW_HD void sort_desc<32>(uint32_t (&k)[32]) { XMHW_SORTNET_32 }

// Load one instance (size <= N rows) for this lane's cell, convert to keys,
// accumulate the f64 sum in row order, sort descending, store into the pool.
// Returns the number of valid (non-NaN) samples.  `any` = warp-uniform vote.
template <int N, class Env>
XMHW_HD int load_sort_store(const Env& env, uint32_t* pool, int lane, int base, int size,
                            const int32_t* rows, const float* col, int64_t ngrid, bool ok) {
  float v[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    v[i] = bits_f32(0x7fc00000u);
    if (i < size && ok) v[i] = XMHW_LDG(col + (int64_t)XMHW_LDG(rows + i) * ngrid);
  }
  uint32_t k[N];
  int len = 0;
  double sum = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    k[i] = f32_key(v[i]);
    if (k[i] != 0u) { ++len; sum = sum + (double)v[i]; }
  }
  if (env.any(len > 0)) {
    sort_desc<N>(k);
#pragma unroll
    for (int i = 0; i < N; ++i)
      if (i < size) pool[(base + POOL_KEYS + i) * 32 + lane] = k[i];
  }
  pool[(base + POOL_SUM) * 32 + lane] = f64_lo(sum);
  pool[(base + POOL_SUM + 1) * 32 + lane] = f64_hi(sum);
  return len;
}

// number of keys of the instance strictly above the pivot key (keys descending)
XMHW_HD int count_above(const uint32_t* pool, int lane, int base, int len, uint32_t pivot) {
  int lo = 0, hi = len;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (pool[(base + POOL_KEYS + mid) * 32 + lane] > pivot) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// numpy _lerp (lib/_function_base_impl.py): d = b - a in float32, result in
// float64 with two roundings, no FMA (the .cu is compiled with --fmad=false).
XMHW_HD double lerp_q(float a, float b, double g) {
  float d = b - a;
  double d64 = (double)d;
  double lo = (double)a + d64 * g;
  double hi = (double)b - d64 * (1.0 - g);
  return g >= 0.5 ? hi : lo;
}

// Persistent per-lane state of the doy sweep.
struct SweepState {
  int C;            // keys currently above the cut (sum of ptr over lists in use)
  int n;            // valid samples in the window (sum of len)
  uint32_t pivot;   // key of the smallest sample above the cut (cut value)
};

// One sweep step for one lane: update the window (leave / enter), move the cut
// to the rank numpy's linear quantile needs, return thresh and seas for this doy.
//
// Selection = k-th largest of a union of sorted lists.  Each list keeps ptr =
// number of its keys above the cut; the cut is "consistent" (every key above it
// >= every key below it).  A list entering the window gets ptr by binary search
// against the current cut value, which keeps the cut consistent; then single
// moves (drop the smallest key above / add the largest key below) restore
// C == target.  Consecutive doys share 10 of 11 lists, so only a few moves are
// needed (vs. sorting ~330 samples per doy).
template <class Env>
XMHW_HD void sweep_step(const Env& env, const ClimPlan& p, int s, SweepState& st, uint32_t* pool,
                        int lane, const float* col, int64_t ngrid, bool ok,
                        double& thresh, double& seas) {
  for (int j = XMHW_LDG(p.leave_off + s); j < XMHW_LDG(p.leave_off + s + 1); ++j) {
    uint32_t meta = pool[(XMHW_LDG(p.leave + j) + POOL_META) * 32 + lane];
    st.C -= (int)((meta >> 8) & 0xffu);
    st.n -= (int)(meta & 0xffu);
  }
  for (int j = XMHW_LDG(p.enter_off + s); j < XMHW_LDG(p.enter_off + s + 1); ++j) {
    int e = XMHW_LDG(p.enter + j);
    int id = e & 0x3fffffff;
    int base = XMHW_LDG(p.inst_base + id);
    int size = XMHW_LDG(p.inst_size + id);
    int len;
    if (e >> 30) {
      const int32_t* rows = p.rows + XMHW_LDG(p.inst_row_off + id);
      if (size <= 8) len = load_sort_store<8>(env, pool, lane, base, size, rows, col, ngrid, ok);
      else if (size <= 16) len = load_sort_store<16>(env, pool, lane, base, size, rows, col, ngrid, ok);
      else if (size <= 24) len = load_sort_store<24>(env, pool, lane, base, size, rows, col, ngrid, ok);
      else len = load_sort_store<32>(env, pool, lane, base, size, rows, col, ngrid, ok);
    } else {
      len = (int)(pool[(base + POOL_META) * 32 + lane] & 0xffu);
    }
    int ptr = count_above(pool, lane, base, len, st.pivot);
    pool[(base + POOL_META) * 32 + lane] = (uint32_t)len | ((uint32_t)ptr << 8);
    st.C += ptr;
    st.n += len;
  }
  const int u0 = XMHW_LDG(p.use_off + s), u1 = XMHW_LDG(p.use_off + s + 1);
  const bool live = st.n > 0;
  int target = 0;
  double gamma = 0.0;
  if (live) {
    target = st.n - XMHW_LDG(p.q_lo + st.n);   // rank (1-based, from the top) of s[floor v]
    gamma = XMHW_LDG(p.q_gamma + st.n);
  }
  if (!env.any(live)) { thresh = qnan(); seas = qnan(); return; }   // all-land warp
  // phase 1: lanes with too few keys above the cut add the largest key below it
  while (env.any(live && st.C < target)) {
    uint32_t best = 0u; int bbase = -1; uint32_t bmeta = 0u;
    for (int j = u0; j < u1; ++j) {
      int base = XMHW_LDG(p.use + j);
      uint32_t meta = pool[(base + POOL_META) * 32 + lane];
      int len = (int)(meta & 0xffu), ptr = (int)((meta >> 8) & 0xffu);
      uint32_t k = ptr < len ? pool[(base + POOL_KEYS + ptr) * 32 + lane] : 0u;
      if (k > best) { best = k; bbase = base; bmeta = meta; }
    }
    if (live && st.C < target && bbase >= 0) {
      pool[(bbase + POOL_META) * 32 + lane] = bmeta + 0x100u;
      ++st.C;
    }
  }
  // phase 2: lanes with too many drop the smallest key above the cut; lanes on
  // target read a = smallest key above the cut and b = next one up.
  uint32_t ka = 0u, kb = 0u;
  bool done = !live;
  while (true) {
    uint32_t m1 = 0xffffffffu, m2 = 0xffffffffu; int b1 = -1; uint32_t meta1 = 0u;
    for (int j = u0; j < u1; ++j) {
      int base = XMHW_LDG(p.use + j);
      uint32_t meta = pool[(base + POOL_META) * 32 + lane];
      int ptr = (int)((meta >> 8) & 0xffu);
      uint32_t k = ptr > 0 ? pool[(base + POOL_KEYS + ptr - 1) * 32 + lane] : 0xffffffffu;
      if (k < m1) { m2 = m1; m1 = k; b1 = base; meta1 = meta; }
      else if (k < m2) m2 = k;
    }
    if (!done) {
      if (st.C > target) {
        pool[(b1 + POOL_META) * 32 + lane] = meta1 - 0x100u;
        --st.C;
      } else {
        int ptr1 = (int)((meta1 >> 8) & 0xffu);
        uint32_t c2 = ptr1 >= 2 ? pool[(b1 + POOL_KEYS + ptr1 - 2) * 32 + lane] : 0xffffffffu;
        ka = m1;
        kb = target >= 2 ? (c2 < m2 ? c2 : m2) : m1;
        done = true;
      }
    }
    if (!env.any(!done)) break;
  }
  if (live) {
    st.pivot = ka;
    thresh = lerp_q(key_f32(ka), key_f32(kb), gamma);
    double sum = 0.0;
    for (int j = u0; j < u1; ++j) {
      int base = XMHW_LDG(p.use + j);
      sum = sum + f64_from(pool[(base + POOL_SUM) * 32 + lane], pool[(base + POOL_SUM + 1) * 32 + lane]);
    }
    seas = sum / (double)st.n;
  } else {
    thresh = qnan();
    seas = qnan();
  }
}

// ---------------------------------------------------------------------------
// event finding (identify.py:415-479 mhw_filter, :273-325 join_gaps)
// ---------------------------------------------------------------------------
// Plain rules (fuzz-verified against the reference, tests/test_oracle_vs_reference.py):
// index 0 is never in an event; maximal exceedance runs of length >= minDuration
// qualify; consecutive qualified events with start - prev_end - 1 <= maxGap merge.
struct RunFinder {
  int min_dur, join, max_gap;
  int run_start, ps, pe;
  XMHW_HD RunFinder(int md, int jn, int mg)
      : min_dur(md), join(jn), max_gap(mg), run_start(-1), ps(-1), pe(-1) {}

  template <class Emit> XMHW_HD void close(int s, int e, Emit& emit) {
    if (e - s + 1 < min_dur) return;
    if (join && ps >= 0 && s - pe - 1 <= max_gap) { pe = e; return; }
    if (ps >= 0) emit(ps, pe);
    ps = s; pe = e;
  }
  // bits: bit i = exceedance at time t0 + i (bits past the series end are 0)
  template <class Emit> XMHW_HD void feed(uint32_t bits, int t0, Emit& emit) {
    if (t0 == 0) bits &= ~1u;
    int pos = 0;
    while (pos < 32) {
      uint32_t rem = bits >> pos;
      if (run_start >= 0) {
        uint32_t z = (~rem) & (0xffffffffu >> pos);
        if (!z) break;
        int k = ctz32(z);
        close(run_start, t0 + pos + k - 1, emit);
        run_start = -1;
        pos += k + 1;
      } else {
        if (!rem) break;
        int k = ctz32(rem);
        run_start = t0 + pos + k;
        pos += k + 1;
      }
    }
  }
  template <class Emit> XMHW_HD void finish(int T, Emit& emit) {
    if (run_start >= 0) { close(run_start, T - 1, emit); run_start = -1; }
    if (ps >= 0) { emit(ps, pe); ps = -1; }
  }
};

// ---------------------------------------------------------------------------
// per-event statistics (features.py:22-69, :97-193, :225-295)
// ---------------------------------------------------------------------------
enum EvInt { EI_CELL = 0, EI_START, EI_END, EI_PEAK, EI_DURATION, EI_CATEGORY,
             EI_MODERATE, EI_STRONG, EI_SEVERE, EI_EXTREME, EI_COUNT };
enum EvF64 { EF_INT_MAX = 0, EF_INT_MEAN, EF_INT_CUM, EF_INT_VAR,
             EF_SEV_MAX, EF_SEV_MEAN, EF_SEV_CUM, EF_SEV_VAR,
             EF_RT_MAX, EF_RT_MEAN, EF_RT_CUM, EF_RT_VAR,
             EF_ABS_MAX, EF_ABS_MEAN, EF_ABS_CUM, EF_ABS_VAR,
             EF_RATE_ONSET, EF_RATE_DECLINE, EF_COUNT };

// NaN-skipping running moments (pandas groupby mean/sum/var(ddof=1) skip NaN;
// Welford update like pandas' group_var).
struct Moments {
  int n; double sum, mean, m2;
  XMHW_HD Moments() : n(0), sum(0.0), mean(0.0), m2(0.0) {}
  XMHW_HD void add(double x) {
    if (x != x) return;
    ++n; sum = sum + x;
    double d = x - mean;
    mean = mean + d / (double)n;
    m2 = m2 + d * (x - mean);
  }
  XMHW_HD double avg() const { return n ? sum / (double)n : qnan(); }
  XMHW_HD double sd() const { return n >= 2 ? sqrt(m2 / (double)(n - 1)) : qnan(); }
};

XMHW_HD double round_f32(double x) { return (double)(float)x; }

// One event [s, e] of the cell whose series starts at `col` (stride ngrid),
// thresholds/seasonal at th/se (doy-major, stride ngrid), doy[t] 1-based.
XMHW_HD void event_stats(const float* col, const double* th, const double* se, const int32_t* doy,
                         int64_t ngrid, int T, int s, int e, int32_t* oi, double* of, int64_t stride) {
  Moments mS, mV, mT, mA;
  double smax = -INFINITY, vmax = -INFINITY, catmax = -INFINITY;
  double t_at_peak = qnan(), x_at_peak = qnan();
  int peak = -1, nmod = 0, nstr = 0, nsev = 0, next = 0;
  bool have_cat = false, have_v = false;
  double relS_first = qnan(), relS_last = qnan();
  double anom_first = qnan(), anom_last = qnan();
  // anom[t] = ts - seas on the unmasked series (features.py:44); anom_plus is
  // anom[t-1], anom_minus anom[t+1] (features.py:45-46)
  double prev_anom = qnan();
  if (s >= 1) {
    int d = XMHW_LDG(doy + s - 1) - 1;
    prev_anom = (double)XMHW_LDG(col + (int64_t)(s - 1) * ngrid) - XMHW_LDG(se + (int64_t)d * ngrid);
  }
  for (int t = s; t <= e; ++t) {
    int d = XMHW_LDG(doy + t) - 1;
    double x = (double)XMHW_LDG(col + (int64_t)t * ngrid);
    double thr = XMHW_LDG(th + (int64_t)d * ngrid);
    double sea = XMHW_LDG(se + (int64_t)d * ngrid);
    double relS = x - sea;                 // features.py:52
    double relT = x - thr;                 // :53
    double ths = thr - sea;                // :54
    double norm = relT / ths;              // :57
    double sev = relS / -(ths);            // :59-61
    double cat = floor(1.0 + norm);        // :62
    if (anom_first != anom_first && prev_anom == prev_anom) anom_first = prev_anom;   // first non-null anom_plus
    if (t > s && relS == relS) anom_last = relS;     // anom_minus of day t-1 is anom[t]
    if (relS == relS) {
      if (relS_first != relS_first) relS_first = relS;
      relS_last = relS;
      if (relS > smax) { smax = relS; peak = t; t_at_peak = relT; x_at_peak = x; }   // first max (:120)
    }
    if (sev == sev) { have_v = true; if (sev > vmax) vmax = sev; }
    if (cat == cat) {
      have_cat = true;
      if (cat > catmax) catmax = cat;
      nmod += cat == 1.0; nstr += cat == 2.0; nsev += cat == 3.0; next += cat >= 4.0;   // :63-66
    }
    mS.add(relS); mV.add(sev); mT.add(relT); mA.add(x);
    prev_anom = relS;
  }
  if (e + 1 <= T - 1) {       // anom_minus of the last event day
    int d = XMHW_LDG(doy + e + 1) - 1;
    double a = (double)XMHW_LDG(col + (int64_t)(e + 1) * ngrid) - XMHW_LDG(se + (int64_t)d * ngrid);
    if (a == a) anom_last = a;
  }
  oi[EI_START * stride] = s;
  oi[EI_END * stride] = e;
  oi[EI_PEAK * stride] = peak;
  oi[EI_DURATION * stride] = e - s + 1;                       // :189
  double cm = catmax < 4.0 ? catmax : 4.0;                    // :188
  oi[EI_CATEGORY * stride] = have_cat ? (cm < -2147483000.0 ? -2147483647 : (int32_t)cm) : -1;
  oi[EI_MODERATE * stride] = nmod;
  oi[EI_STRONG * stride] = nstr;
  oi[EI_SEVERE * stride] = nsev;
  oi[EI_EXTREME * stride] = next;
  double imax = peak >= 0 ? smax : qnan();
  of[EF_INT_MAX * stride] = imax;
  of[EF_INT_MEAN * stride] = mS.avg();
  of[EF_INT_CUM * stride] = mS.sum;
  of[EF_INT_VAR * stride] = mS.sd();
  of[EF_SEV_MAX * stride] = have_v ? vmax : qnan();
  of[EF_SEV_MEAN * stride] = mV.avg();
  of[EF_SEV_CUM * stride] = mV.sum;
  of[EF_SEV_VAR * stride] = mV.sd();
  of[EF_RT_MAX * stride] = t_at_peak;                          // :184
  of[EF_RT_MEAN * stride] = mT.avg();
  of[EF_RT_CUM * stride] = mT.sum;
  of[EF_RT_VAR * stride] = mT.sd();
  of[EF_ABS_MAX * stride] = round_f32(x_at_peak);              // :185  (mabs is float32, :68)
  of[EF_ABS_MEAN * stride] = round_f32(mA.avg());
  of[EF_ABS_CUM * stride] = round_f32(mA.sum);
  of[EF_ABS_VAR * stride] = round_f32(mA.sd());
  // onset / decline (features.py:225-295)
  int pk = peak >= 0 ? peak - s : 0;
  double onset_period = (double)(pk != 0 ? pk : 1) + (s == 0 ? 0.0 : 0.5);            // :259-260
  double y = (pk != T - 1) ? (double)(e - s - pk) : 1.0;                              // :258,261
  double decline_period = y + (e == T - 1 ? 0.0 : 0.5);                               // :262
  double edge_s = 0.5 * (relS_first + (s == 0 ? relS_first : anom_first));            // :220-221,287
  double edge_e = 0.5 * (relS_last + (e == T - 1 ? relS_last : anom_last));           // :288
  of[EF_RATE_ONSET * stride] = (imax - edge_s) / onset_period;                        // :290
  of[EF_RATE_DECLINE * stride] = (imax - edge_e) / decline_period;                    // :291
}

}  // namespace xmhw
