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
//   row 3      cinc: smallest key above the cut (key[ptr-1]), 0xffffffff if ptr == 0
//   row 4      cexc: largest key below the cut (key[ptr]),    0 if ptr == len
//   row 5 + r  r-th largest key (r < size); invalid samples are key 0 at the end
// Block 0 of every pool is a "null list" (len 0) used to pad scans to multiples of 4;
// the two rows after plan.pool_rows hold the staged base rows of the lists in use.
enum { POOL_META = 0, POOL_SUM = 1, POOL_CINC = 3, POOL_CEXC = 4, POOL_KEYS = 5, POOL_NULL_ROWS = 5,
       POOL_STAGE_ROWS = 2, MAX_LISTS = 64 };

#ifdef __CUDA_ARCH__
#define XMHW_LDG(p) __ldg(p)
#else
#define XMHW_LDG(p) (*(p))
#endif

#define XMHW_CE(i, j) { uint32_t hi_ = k[i] > k[j] ? k[i] : k[j]; uint32_t lo_ = k[i] > k[j] ? k[j] : k[i]; k[i] = hi_; k[j] = lo_; }

template <int N> XMHW_HD void sort_desc(uint32_t* k);
template <> XMHW_HD void sort_desc<8>(uint32_t* k) { XMHW_SORTNET_8 }
template <> XMHW_HD void sort_desc<16>(uint32_t* k) { XMHW_SORTNET_16 }
template <> XMHW_HD void sort_desc<24>(uint32_t* k) { XMHW_SORTNET_24 }
template <> XMHW_HD void sort_desc<32>(uint32_t* k) { XMHW_SORTNET_32 }

// numpy _lerp (lib/_function_base_impl.py): d = b - a in float32, result in
// float64 with two roundings, no FMA (the .cu is compiled with --fmad=false).
XMHW_HD double lerp_q(float a, float b, double g) {
  float d = b - a;
  double d64 = (double)d;
  double lo = (double)a + d64 * g;
  double hi = (double)b - d64 * (1.0 - g);
  return g >= 0.5 ? hi : lo;
}

// number of keys of an instance strictly above the pivot key (keys descending)
XMHW_HD int count_above(const uint32_t* pool, int lane, int base, int len, uint32_t pivot) {
  int lo = 0, hi = len;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (pool[(base + POOL_KEYS + mid) * 32 + lane] > pivot) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Doy sweep of one lane (= one grid cell).  Selection = k-th largest of a union of
// sorted lists: each list keeps ptr = number of its keys above the cut, the cut is
// "consistent" (every key above it >= every key below it).  A list entering the
// window gets ptr by comparison with the current cut value, which keeps the cut
// consistent; then single moves (add the largest key below the cut / drop the
// smallest key above it) restore C == target rank.  Consecutive doys share all but
// one list, so a few moves replace a sort of ~330 samples per doy.  Each list
// caches its two keys adjacent to the cut (cinc/cexc rows), so one scan over the
// lists in use is one shared-memory load + compare/select per list.
template <class Env>
struct Sweeper {
  const Env& env;
  const ClimPlan& p;
  uint32_t* pool;
  const int lane;
  const float* col;
  const int64_t ngrid;
  const bool ok;
  int C, n;              // keys above the cut / valid samples, over the lists in use
  uint32_t pivot;        // cut value (key of the smallest sample above the cut)
  float pv[32];          // prefetched rows of the next instance to load
  int pf;                // index into plan.enter of that instance (or total)
  int total_enter;

  XMHW_HD Sweeper(const Env& e, const ClimPlan& pl, uint32_t* po, int ln, const float* c, int64_t ng, bool k)
      : env(e), p(pl), pool(po), lane(ln), col(c), ngrid(ng), ok(k), C(0), n(0), pivot(0xffffffffu) {}

  XMHW_HD uint32_t& at(int row) { return pool[row * 32 + lane]; }

  XMHW_HD void prefetch(int from) {
    int j = from;
    while (j < total_enter && !(XMHW_LDG(p.enter + j) >> 30)) ++j;
    pf = j;
    if (j >= total_enter) return;
    const int id = XMHW_LDG(p.enter + j) & 0x3fffffff;
    const int size = XMHW_LDG(p.inst_size + id);
    const int32_t* rows = p.rows + XMHW_LDG(p.inst_row_off + id);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      pv[i] = bits_f32(0x7fc00000u);
      if (i < size && ok) pv[i] = XMHW_LDG(col + (int64_t)XMHW_LDG(rows + i) * ngrid);
    }
  }

  // keys of the prefetched instance -> sorted block in the pool; returns (len, ptr)
  template <int N>
  XMHW_HD void consume(int base, int size, int& len, int& ptr) {
    uint32_t k[32];
    len = 0;
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      k[i] = f32_key(pv[i]);
      if (k[i] != 0u) { ++len; sum = sum + (double)pv[i]; }
    }
    ptr = 0;
    if (env.any(len > 0)) {
      sort_desc<N>(k);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        if (i < size) at(base + POOL_KEYS + i) = k[i];
        ptr += k[i] > pivot;
      }
    }
    at(base + POOL_SUM) = f64_lo(sum);
    at(base + POOL_SUM + 1) = f64_hi(sum);
  }

  XMHW_HD void set_block(int base, int len, int ptr) {
    at(base + POOL_META) = (uint32_t)len | ((uint32_t)ptr << 8);
    at(base + POOL_CINC) = ptr > 0 ? at(base + POOL_KEYS + ptr - 1) : 0xffffffffu;
    at(base + POOL_CEXC) = ptr < len ? at(base + POOL_KEYS + ptr) : 0u;
  }

  XMHW_HD void init() {
    at(POOL_META) = 0u; at(POOL_SUM) = 0u; at(POOL_SUM + 1) = 0u;
    at(POOL_CINC) = 0xffffffffu; at(POOL_CEXC) = 0u;
    total_enter = XMHW_LDG(p.enter_off + p.nsteps);
    prefetch(0);
  }

  XMHW_HD void step(int s, double& thresh, double& seas) {
    for (int j = XMHW_LDG(p.leave_off + s); j < XMHW_LDG(p.leave_off + s + 1); ++j) {
      uint32_t meta = at(XMHW_LDG(p.leave + j) + POOL_META);
      C -= (int)((meta >> 8) & 0xffu);
      n -= (int)(meta & 0xffu);
    }
    for (int j = XMHW_LDG(p.enter_off + s); j < XMHW_LDG(p.enter_off + s + 1); ++j) {
      const int e = XMHW_LDG(p.enter + j);
      const int id = e & 0x3fffffff;
      const int base = XMHW_LDG(p.inst_base + id);
      const int size = XMHW_LDG(p.inst_size + id);
      int len, ptr;
      if (e >> 30) {
        if (size <= 8) consume<8>(base, size, len, ptr);
        else if (size <= 16) consume<16>(base, size, len, ptr);
        else if (size <= 24) consume<24>(base, size, len, ptr);
        else consume<32>(base, size, len, ptr);
        prefetch(j + 1);
      } else {
        len = (int)(at(base + POOL_META) & 0xffu);
        ptr = count_above(pool, lane, base, len, pivot);
      }
      set_block(base, len, ptr);
      C += ptr;
      n += len;
    }
    // stage the base rows of the lists in use (padded to a multiple of 4 with the null list)
    const int u0 = XMHW_LDG(p.use_off + s);
    const int m = XMHW_LDG(p.use_off + s + 1) - u0;
    const int m4 = (m + 3) & ~3;
    uint32_t* ub = pool + p.pool_rows * 32;
    env.stage(ub, p.use + u0, m, m4, lane);

    const bool live = n > 0;
    if (!env.any(live)) { thresh = qnan(); seas = qnan(); return; }   // all-land warp
    int target = 0;
    double gamma = 0.0;
    if (live) {
      target = n - XMHW_LDG(p.q_lo + n);   // rank (1-based, from the top) of s[floor v]
      gamma = XMHW_LDG(p.q_gamma + n);
    }
    // phase 1: lanes with too few keys above the cut add the largest key below it
    while (env.any(live && C < target)) {
      uint32_t b0 = 0u, b1 = 0u, b2 = 0u, b3 = 0u;
      int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      for (int j = 0; j < m4; j += 4) {
        const int x0 = (int)ub[j], x1 = (int)ub[j + 1], x2 = (int)ub[j + 2], x3 = (int)ub[j + 3];
        const uint32_t k0 = at(x0 + POOL_CEXC), k1 = at(x1 + POOL_CEXC), k2 = at(x2 + POOL_CEXC), k3 = at(x3 + POOL_CEXC);
        if (k0 > b0) { b0 = k0; a0 = x0; }
        if (k1 > b1) { b1 = k1; a1 = x1; }
        if (k2 > b2) { b2 = k2; a2 = x2; }
        if (k3 > b3) { b3 = k3; a3 = x3; }
      }
      if (b1 > b0) { b0 = b1; a0 = a1; }
      if (b3 > b2) { b2 = b3; a2 = a3; }
      if (b2 > b0) { b0 = b2; a0 = a2; }
      if (live && C < target) {      // b0 > 0 is guaranteed: C < target <= n
        const uint32_t meta = at(a0 + POOL_META);
        const int len = (int)(meta & 0xffu), ptr = (int)((meta >> 8) & 0xffu) + 1;
        at(a0 + POOL_META) = meta + 0x100u;
        at(a0 + POOL_CINC) = b0;
        at(a0 + POOL_CEXC) = ptr < len ? at(a0 + POOL_KEYS + ptr) : 0u;
        ++C;
      }
    }
    // phase 2: lanes with too many drop the smallest key above the cut
    while (env.any(live && C > target)) {
      uint32_t b0 = 0xffffffffu, b1 = 0xffffffffu, b2 = 0xffffffffu, b3 = 0xffffffffu;
      int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      for (int j = 0; j < m4; j += 4) {
        const int x0 = (int)ub[j], x1 = (int)ub[j + 1], x2 = (int)ub[j + 2], x3 = (int)ub[j + 3];
        const uint32_t k0 = at(x0 + POOL_CINC), k1 = at(x1 + POOL_CINC), k2 = at(x2 + POOL_CINC), k3 = at(x3 + POOL_CINC);
        if (k0 < b0) { b0 = k0; a0 = x0; }
        if (k1 < b1) { b1 = k1; a1 = x1; }
        if (k2 < b2) { b2 = k2; a2 = x2; }
        if (k3 < b3) { b3 = k3; a3 = x3; }
      }
      if (b1 < b0) { b0 = b1; a0 = a1; }
      if (b3 < b2) { b2 = b3; a2 = a3; }
      if (b2 < b0) { b0 = b2; a0 = a2; }
      if (live && C > target) {
        const uint32_t meta = at(a0 + POOL_META);
        const int ptr = (int)((meta >> 8) & 0xffu) - 1;
        at(a0 + POOL_META) = meta - 0x100u;
        at(a0 + POOL_CEXC) = b0;
        at(a0 + POOL_CINC) = ptr > 0 ? at(a0 + POOL_KEYS + ptr - 1) : 0xffffffffu;
        --C;
      }
    }
    // final scan: a = smallest key above the cut, b = next one up; f64 sum of the window
    uint32_t m1 = 0xffffffffu, m2 = 0xffffffffu;
    int a1 = 0;
    double sum = 0.0;
    for (int j = 0; j < m4; ++j) {
      const int x = (int)ub[j];
      const uint32_t k = at(x + POOL_CINC);
      if (k < m1) { m2 = m1; m1 = k; a1 = x; }
      else if (k < m2) m2 = k;
      sum = sum + f64_from(at(x + POOL_SUM), at(x + POOL_SUM + 1));
    }
    if (live) {
      const int ptr1 = (int)((at(a1 + POOL_META) >> 8) & 0xffu);
      const uint32_t c2 = ptr1 >= 2 ? at(a1 + POOL_KEYS + ptr1 - 2) : 0xffffffffu;
      const uint32_t kb = target >= 2 ? (c2 < m2 ? c2 : m2) : m1;
      pivot = m1;
      thresh = lerp_q(key_f32(m1), key_f32(kb), gamma);
      seas = sum / (double)n;
    } else {
      thresh = qnan();
      seas = qnan();
    }
  }
};

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
