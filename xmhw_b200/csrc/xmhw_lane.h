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
#define XMHW_NOINLINE __host__ __device__ __noinline__
#else
#define XMHW_HD inline
#define XMHW_NOINLINE inline
#endif

#include "sortnet_gen.h"

// development statistics hooks (tools/sweep_stats.cpp defines them; no-ops in the product)
#ifndef XMHW_STAT_SCAN
#define XMHW_STAT_SCAN()
#define XMHW_STAT_POP(in_scratch)
#define XMHW_STAT_STEP(d0)
#endif
#ifndef XMHW_STAT_RANK
#define XMHW_STAT_RANK(r)
#endif
#ifndef XMHW_STAT_PTR0
#define XMHW_STAT_PTR0(base, ptr)
#define XMHW_STAT_OFF(base, r)
#endif

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
XMHW_HD int clz32(uint32_t x) {      // x != 0
#ifdef __CUDA_ARCH__
  return __clz((int)x);
#else
  return __builtin_clz(x);
#endif
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
  int32_t scratch_rows;          // global scratch rows (32 words each) per warp
  int32_t reserved_;
  const int32_t* inst_base;      // [ninst] first pool row of the instance block
  const int32_t* inst_size;      // [ninst] number of time rows (1..32)
  const int32_t* inst_keep;      // [ninst] key rows held in shared memory (1..size)
  const int32_t* inst_sbase;     // [ninst] first global scratch row of the keys past `keep`
  const int32_t* inst_row_off;   // [ninst] offset into rows[]
  const int32_t* rows;           // time indices
  const int32_t* leave_off;      // [nsteps+1]  -> leave[] (pool base rows)
  const int32_t* leave;
  const int32_t* enter_off;      // [nsteps+1]  -> enter[] (instance id | load flag << 30)
  const int32_t* enter;
  const int32_t* use_off;        // [nsteps+1]  -> use[] (pool base rows)
  const int32_t* use;
  const int32_t* step_rec;       // [nsteps][32] fixed-size step records (see STEP_* below)
  double q;                      // quantile in [0,1]; numpy 'linear': v = (n-1) * q
};

// Step record (32 int32 per sweep step, loaded with one coalesced warp load one step
// ahead so no plan lookup sits on the critical path).  Steps with more than 4 leaving or
// 4 entering lists (only the first step) set STEP_OVERFLOW and use the CSR arrays.
enum { STEP_COUNTS = 0,      // n_leave | n_enter << 8 | n_use << 16 | overflow << 31
       STEP_USE_OFF = 1,     // offset of this step's list bases in plan.use
       STEP_LEAVE = 2,       // 4 words: pool base rows of the leaving lists
       STEP_ENTER = 6,       // 4 x 3 words: (instance id | load flag << 30), (base | size << 16 | keep << 24), scratch base
       STEP_NEXT_USE_OFF = 18, STEP_NEXT_NUSE = 19,   // the same two numbers of the next step
       STEP_ENTER_OFF = 20,  // index of this step's first entry in plan.enter
       STEP_NEXT_LOAD = 21,  // 4 x 2 words: (offset into plan.rows, size) of the list to prefetch after entry j (size 0: none)
       STEP_WORDS = 32, STEP_MAX_INLINE = 4 };

// Instance block layout in the pool (row = 32 words, word index = lane):
//   row 0      meta: len | ptr << 6 | keep << 12 | scratch base row << 18
//              (len = valid samples, ptr = #keys above the cut, keep = key rows held here)
//   row 1      cinc: smallest key above the cut (key[ptr-1]), 0xffffffff if ptr == 0
//   row 2      cexc: largest key below the cut (key[ptr]),    0 if ptr == len
//   row 3 + r  r-th largest key, r < keep <= size.  Only the top `keep` keys of a list
//              live in shared memory; the sorted remainder is parked in a per-warp global
//              scratch (L2-resident, one load per access), so `keep` can be small and
//              many more warps stay resident per SM (the sweep is latency-chained).
// Scratch block of a list (rows of 32 words): row 0, 1 = f64 sum of its valid samples
// (lo, hi words), row 2 + i = key of rank keep + i.  Scratch rows 0, 1 of every warp are
// the null list's sum (0.0).
// Block 0 of every pool is a "null list" (len 0) used to pad scans to multiples of 4;
// the two rows after plan.pool_rows hold the staged base rows of the lists in use.
enum { POOL_META = 0, POOL_CINC = 1, POOL_CEXC = 2, POOL_KEYS = 3, POOL_NULL_ROWS = 3, SCR_SUM = 0, SCR_KEYS = 2,
       POOL_STAGE_ROWS = 2, MAX_LISTS = 64 };

XMHW_HD int meta_len(uint32_t m) { return (int)(m & 63u); }
XMHW_HD int meta_ptr(uint32_t m) { return (int)((m >> 6) & 63u); }
XMHW_HD int meta_keep(uint32_t m) { return (int)((m >> 12) & 63u); }
XMHW_HD int meta_sbase(uint32_t m) { return (int)(m >> 18); }
#define XMHW_META_PTR1 64u

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
template <> XMHW_HD void sort_desc<30>(uint32_t* k) { XMHW_SORTNET_30 }
template <> XMHW_HD void sort_desc<32>(uint32_t* k) { XMHW_SORTNET_32 }
template <> XMHW_HD void sort_desc<40>(uint32_t* k) { XMHW_SORTNET_40 }
template <> XMHW_HD void sort_desc<48>(uint32_t* k) { XMHW_SORTNET_48 }

XMHW_HD uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }
XMHW_HD uint32_t umax32(uint32_t a, uint32_t b) { return a > b ? a : b; }

// Branch-free insertion of (key t, tag x) into the ascending F-entry front (f[0] <= ... <= f[F-1])
// that keeps the F smallest; ties keep the earlier entry first.  All F compares and the 2F-1
// min/max are independent of each other (depth 2), unlike a chain of conditional swaps.
#ifndef XMHW_FRONT
#define XMHW_FRONT 4
#endif
template <int F>
XMHW_HD void front_insert(uint32_t t, int x, uint32_t (&f)[F], int (&g)[F]) {
  bool c[F];
#pragma unroll
  for (int i = 0; i < F; ++i) c[i] = t < f[i];
#pragma unroll
  for (int i = F - 1; i >= 1; --i) {
    g[i] = c[i - 1] ? g[i - 1] : (c[i] ? x : g[i]);
    f[i] = umin32(f[i], umax32(f[i - 1], t));
  }
  g[0] = c[0] ? x : g[0];
  f[0] = umin32(f[0], t);
}
// Drop the head of the front and insert (t, x) in one step (t = 0xffffffff: nothing enters, the
// last slot becomes empty): F-1 compares, 2F-2 min/max, 2F-1 selects instead of a shift + insert.
template <int F>
XMHW_HD void front_replace_head(uint32_t t, int x, uint32_t (&f)[F], int (&g)[F]) {
  bool c[F];
#pragma unroll
  for (int i = 1; i < F; ++i) c[i] = t < f[i];
  uint32_t nf[F];
  int ng[F];
  nf[0] = umin32(f[1], t);
  ng[0] = c[1] ? x : g[1];
#pragma unroll
  for (int i = 1; i + 1 < F; ++i) {
    nf[i] = umin32(f[i + 1], umax32(f[i], t));
    ng[i] = c[i] ? g[i] : (c[i + 1] ? x : g[i + 1]);
  }
  nf[F - 1] = umax32(f[F - 1], t);
  ng[F - 1] = c[F - 1] ? g[F - 1] : x;
#pragma unroll
  for (int i = 0; i < F; ++i) { f[i] = nf[i]; g[i] = ng[i]; }
}

// Position of `pivot` in a descending sorted register array: ptr = #{k[i] > pivot},
// cinc = k[ptr-1] (0xffffffff when ptr == 0), cexc = k[ptr] (0 when ptr == N).  For a power-of-two
// N a branch-free halving search that carries the window e[base-1 .. base+L] through selects
// (log2 N compares + ~N+10 selects) replaces N compares, N counts and 2N selects.
template <int N>
XMHW_HD void partition_sorted(const uint32_t (&k)[N], uint32_t pivot, int& ptr, uint32_t& cinc, uint32_t& cexc) {
  if ((N & (N - 1)) == 0) {
    uint32_t V[N + 2];
    V[0] = 0xffffffffu;
#pragma unroll
    for (int i = 0; i < N; ++i) V[i + 1] = k[i];
    V[N + 1] = 0u;
    int base = 0;
#pragma unroll
    for (int h = N / 2; h >= 1; h >>= 1) {        // window V[0 .. 2h+1] = e[base-1 .. base+2h]
      const bool t = V[h] > pivot;                // k[base + h - 1] > pivot: at least h more keys above
      base += t ? h : 0;
#pragma unroll
      for (int i = 0; i <= h + 1; ++i) V[i] = t ? V[h + i] : V[i];
    }
    const bool t = V[1] > pivot;                  // V = {k[base-1], k[base], k[base+1]}
    ptr = base + (t ? 1 : 0);
    cinc = t ? V[1] : V[0];
    cexc = t ? V[2] : V[1];
  } else {
    ptr = 0; cinc = 0xffffffffu; cexc = 0u;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const bool ab = k[i] > pivot;
      ptr += ab;
      if (ab) cinc = k[i]; else cexc = cexc > k[i] ? cexc : k[i];
    }
  }
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

// Doy sweep of one lane (= one grid cell).  Selection = k-th largest of a union of
// sorted lists: each list keeps ptr = number of its keys above the cut, the cut is
// "consistent" (every key above it >= every key below it).  A list entering the
// window gets ptr by comparison with the current cut value, which keeps the cut
// consistent; then single moves (add the largest key below the cut / drop the
// smallest key above it) restore C == target rank.  Consecutive doys share all but
// one list, so a few moves replace a sort of ~330 samples per doy.  Each list
// caches its two keys adjacent to the cut (cinc/cexc rows), so one scan over the
// lists in use is one shared-memory load + compare/select per list.
//
// Env supplies the warp-level pieces: any(pred) vote and Vec = a small int vector
// spread over the lanes (vload = one coalesced load, vget = broadcast of one entry).
template <class Env, int MAXN = 32>
struct Sweeper {
  typedef typename Env::Vec Vec;
  const Env& env;
  const ClimPlan& p;
  uint32_t* pool;
  uint32_t* scratch;     // this warp's global scratch rows (sorted keys past `keep`)
  const int lane;
  const float* col;
  const int64_t ngrid;
  const bool ok;
  int C, n;              // keys above the cut / valid samples, over the lists in use
  double wsum;           // running f64 sum of the samples of the lists in use (+ entering, - leaving list sums)
  uint32_t pivot;        // cut value (key of the smallest sample above the cut)
  float pv[MAXN];        // prefetched rows of the next instance to load (MAXN = 32 or 48 keys per list)
  int total_enter;
  int pend_off, pend_size;   // the list whose rows sit in (or are on their way to) the prefetch registers
  Vec pend_rows_v;           // its row indices
  Vec cur_use; int cur_m; uint32_t cur_plo, cur_phi;    // enter_phase(s) -> walk_phase(s)
  int nzero;             // steps without any sample (per-cell doy compaction of the smoothing)
  Vec rec_next, use_next;   // step record / list bases of the next step (prefetched)

  XMHW_HD Sweeper(const Env& e, const ClimPlan& pl, uint32_t* po, uint32_t* sc, int ln, const float* c, int64_t ng, bool k)
      : env(e), p(pl), pool(po), scratch(sc), lane(ln), col(c), ngrid(ng), ok(k), C(0), n(0), wsum(0.0), pivot(0xffffffffu), nzero(0) {}

  XMHW_HD uint32_t& at(int row) { return pool[row * 32 + lane]; }

  // (offset into plan.rows, size) of the next instance that has to be loaded, entry index >= from;
  // size 0 when there is none.  A chain of dependent plan loads: only used where the step record
  // does not carry the answer (start of the sweep, steps with more than 4 entering lists).
  XMHW_HD void next_load(int from, int& row_off, int& size) {
    int j = from;
    row_off = 0; size = 0;
    while (j < total_enter && !(XMHW_LDG(p.enter + j) >> 30)) ++j;
    if (j >= total_enter) return;
    const int id = XMHW_LDG(p.enter + j) & 0x3fffffff;
    row_off = XMHW_LDG(p.inst_row_off + id);
    size = XMHW_LDG(p.inst_size + id);
  }

  // issue the loads of the list whose time rows are the first `size` entries of rv
  XMHW_HD void prefetch_rows(const Vec& rv, int size) {
    // unconditional loads (entries past `size` read row 0 and are masked in consume), so the
    // compiler keeps all of them in flight instead of waiting on each predicated result.
    // Byte offsets: row * (4 ngrid) added to the column pointer is ONE 32x32+64 multiply-add per row.
    const uint32_t ng4 = (uint32_t)ngrid * 4u;      // ngrid < 2^30 (checked by the launcher)
    const char* const cb = reinterpret_cast<const char*>(col);
    if (MAXN == 32 || size <= 32) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        pv[i] = XMHW_LDG(reinterpret_cast<const float*>(cb + (uint64_t)(uint32_t)env.vget(rv, i) * ng4));
    } else {
#pragma unroll
      for (int i = 0; i < MAXN; ++i)
        pv[i] = XMHW_LDG(reinterpret_cast<const float*>(cb + (uint64_t)(uint32_t)env.vget(rv, i) * ng4));
    }
  }

  // key at rank r of the list at `base` (0 when !need)
  XMHW_HD uint32_t key_at(int base, uint32_t meta, int r, bool need) {
    const int keep = meta_keep(meta);
    const bool in_pool = r < keep;
    // shared-memory read is unconditional (row clamped into the block); the scratch read is the rare path
    uint32_t k = at(base + POOL_KEYS + ((in_pool && need) ? r : 0));
    if (need && !in_pool) k = scratch[(size_t)(meta_sbase(meta) + SCR_KEYS + r - keep) * 32 + lane];
    return need ? k : 0u;
  }

  // number of keys of the list strictly above `piv` (keys descending)
  XMHW_HD int count_above(int base, uint32_t meta, uint32_t piv) {
    int lo = 0, hi = meta_len(meta);
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (key_at(base, meta, mid, true) > piv) lo = mid + 1; else hi = mid;
    }
    return lo;
  }

  // keys of the prefetched instance -> sorted block in the pool (EXACT: the list has exactly N rows)
  template <int N, bool EXACT = false>
  XMHW_HD void consume(int base, int sbase, int size, int keep, int& len, int& ptr) {
    // all-land shortcut: a warp whose 32 cells have no valid sample in this list skips the key
    // conversion, sums and sort (ocean warps pay one compare + vote for the test)
    bool some = env.any(ok && pv[0] == pv[0]);
    if (!some) {
      bool anyv = false;
#pragma unroll
      for (int i = 1; i < N; ++i) anyv = anyv || (i < size && pv[i] == pv[i]);
      some = env.any(ok && anyv);
    }
    if (!some) {
      len = 0; ptr = 0;
      scratch[(size_t)(sbase + SCR_SUM) * 32 + lane] = 0u;
      scratch[(size_t)(sbase + SCR_SUM + 1) * 32 + lane] = 0u;
      at(base + POOL_META) = ((uint32_t)keep << 12) | ((uint32_t)sbase << 18);
      at(base + POOL_CINC) = 0xffffffffu;
      at(base + POOL_CEXC) = 0u;
      return;
    }
    uint32_t k[N];
    len = 0;
    double sum = 0.0;
    // Lanes past the grid edge read cell 0 and compute on it: nothing of theirs is ever stored
    // (lanes share no data), so validity is just "inside the list and not NaN".  An invalid sample
    // enters the f64 sum as +0.0f (exact) instead of being skipped by a select on the f64 halves.
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const float v = pv[i];
      const uint32_t b = f32_bits(v);
      const bool valid = (EXACT || i < size) && (v == v);
      k[i] = valid ? (b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u)) : 0u;
      len += valid ? 1 : 0;
      sum = sum + (double)(valid ? v : 0.0f);
    }
    ptr = 0;
    uint32_t cinc = 0xffffffffu, cexc = 0u;
    if (env.any(len > 0)) {
      sort_desc<N>(k);
      uint32_t* const srow = pool + (base + POOL_KEYS) * 32 + lane;
      uint32_t* const grow = scratch + ((size_t)(sbase + SCR_KEYS) - (size_t)keep) * 32 + lane;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        if (i < keep) srow[i * 32] = k[i];                  // top `keep` keys: shared memory
        if (i >= keep && i < size) grow[i * 32] = k[i];     // sorted remainder: global scratch
      }
      if (N == 30) {                 // the halving search wants a power of two: two 0 keys (below every pivot) appended
        uint32_t k32[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) k32[i] = i < N ? k[i] : 0u;
        partition_sorted<32>(k32, pivot, ptr, cinc, cexc);
      } else {
        partition_sorted<N>(k, pivot, ptr, cinc, cexc);
      }
    }
    scratch[(size_t)(sbase + SCR_SUM) * 32 + lane] = f64_lo(sum);
    scratch[(size_t)(sbase + SCR_SUM + 1) * 32 + lane] = f64_hi(sum);
    wsum = wsum + sum;
    at(base + POOL_META) = (uint32_t)len | ((uint32_t)ptr << 6) | ((uint32_t)keep << 12) | ((uint32_t)sbase << 18);
    at(base + POOL_CINC) = cinc;
    at(base + POOL_CEXC) = cexc;
  }

  // the leaving list's f64 sum comes back as two words so that the caller can defer the
  // subtraction until after the walk (the scratch load latency then hides behind it)
  XMHW_HD void leave_list(int base, uint32_t& slo, uint32_t& shi) {
    const uint32_t meta = at(base + POOL_META);
    C -= meta_ptr(meta);
    n -= meta_len(meta);
    const size_t sb = (size_t)meta_sbase(meta) * 32 + lane;
    slo = scratch[sb + SCR_SUM * 32];
    shi = scratch[sb + (SCR_SUM + 1) * 32];
  }

  // returns true when the list was loaded (the prefetched rows are consumed: the caller issues the
  // next prefetch -- at ONE site, the unrolled loads are the bulkiest code of the step after the sort)
  XMHW_HD bool enter_list(int e, int base, int size, int keep, int sbase) {
    int len, ptr;
    const bool loaded = (e >> 30) != 0;
    if (loaded) {
      if (size <= 8) consume<8>(base, sbase, size, keep, len, ptr);
      else if (size == 30) consume<30, true>(base, sbase, size, keep, len, ptr);      // 30-year series: the common list
      else if (MAXN == 32 || size <= 32) consume<32>(base, sbase, size, keep, len, ptr);
      else if (size <= 40) consume<(MAXN > 32 ? 40 : 32)>(base, sbase, size, keep, len, ptr);
      else consume<(MAXN > 32 ? 48 : 32)>(base, sbase, size, keep, len, ptr);
    } else {          // list re-enters after a hole (Feb 29): pointer against the current cut
      uint32_t meta = at(base + POOL_META);
      len = meta_len(meta);
      ptr = count_above(base, meta, pivot);
      meta = (meta & ~(63u << 6)) | ((uint32_t)ptr << 6);
      at(base + POOL_META) = meta;
      const uint32_t ci = key_at(base, meta, ptr - 1, ptr > 0);
      const uint32_t ce = key_at(base, meta, ptr, ptr < len);
      at(base + POOL_CINC) = ptr > 0 ? ci : 0xffffffffu;
      at(base + POOL_CEXC) = ce;
      const size_t sb = (size_t)meta_sbase(meta) * 32 + lane;
      wsum = wsum + f64_from(scratch[sb + SCR_SUM * 32], scratch[sb + (SCR_SUM + 1) * 32]);
    }
    C += ptr;
    n += len;
    return loaded;
  }

  XMHW_HD void init() {
    at(POOL_META) = 0u; at(POOL_CINC) = 0xffffffffu; at(POOL_CEXC) = 0u;
    scratch[SCR_SUM * 32 + lane] = 0u; scratch[(SCR_SUM + 1) * 32 + lane] = 0u;
    total_enter = XMHW_LDG(p.enter_off + p.nsteps);
    rec_next = env.vload(p.step_rec, STEP_WORDS, lane);
    use_next = env.vload(p.use + XMHW_LDG(p.step_rec + STEP_USE_OFF),
                         (XMHW_LDG(p.step_rec + STEP_COUNTS) >> 16) & 0x7f, lane);
    int off0, size0;
    next_load(0, off0, size0);
    pend_off = off0; pend_size = size0;
    pend_rows_v = env.vload(p.rows + off0, size0, lane);
    prefetch_rows(pend_rows_v, size0);
  }

  // One pass over the lists in use: the two smallest keys above the cut (i1 <= i2, lists
  // bi1 != bi2) and / or the two largest keys below it (e1 >= e2, lists be1 != be2).
  template <bool INC, bool EXC>
  XMHW_HD void scan(const uint32_t* ub, int m4, uint32_t& i1, uint32_t& i2, int& bi1, int& bi2,
                    uint32_t& e1, uint32_t& e2, int& be1, int& be2) {
    if (INC) { i1 = 0xffffffffu; i2 = 0xffffffffu; bi1 = 0; bi2 = 0; }
    if (EXC) { e1 = 0u; e2 = 0u; be1 = 0; be2 = 0; }
#pragma unroll 4
    for (int j = 0; j < m4; ++j) {
      const int x = (int)ub[j];
      if (INC) {
        const uint32_t ki = at(x + POOL_CINC);
        const bool c1 = ki < i1, c2 = ki < i2;            // branch-free 2-entry front (ties keep the earlier list)
        i2 = umin32(i2, umax32(i1, ki)); bi2 = c1 ? bi1 : (c2 ? x : bi2);
        i1 = umin32(i1, ki); bi1 = c1 ? x : bi1;
      }
      if (EXC) {
        const uint32_t ke = at(x + POOL_CEXC);
        const bool c1 = ke > e1, c2 = ke > e2;
        e2 = umax32(e2, umin32(e1, ke)); be2 = c1 ? be1 : (c2 ? x : be2);
        e1 = umax32(e1, ke); be1 = c1 ? x : be1;
      }
    }
  }

  // A sweep step = enter_phase(s) (lists leave / enter; consumes the prefetch registers), the ONE
  // unconditional prefetch of the list that is pending afterwards, then walk_phase(s) (selection +
  // output).  Measured variants of where the prefetch sits (B200, global grid, profiles/
  // kernel_ms_r02_general_sweep_variants.txt): inside the conditional entry loop 54.6 ms (round 1: ptxas
  // resolves the 32-register phi with copies right behind the loads, 14 % of the stall samples wait
  // there); one unconditional site after the entries 53.2 ms (this); at the end of the step 54.1 ms;
  // loop rotated so that the registers are written and read in one iteration 54.1 ms (code grows, the
  // instruction-fetch stalls eat the gain).
  XMHW_HD void step(int s, double& thresh, double& seas) {
    enter_phase(s);
    prefetch_pending();
    walk_phase(s, thresh, seas);
  }

  XMHW_HD void prefetch_pending() { prefetch_rows(pend_rows_v, pend_size); }

  XMHW_HD void enter_phase(int s) {
    const Vec rec = rec_next;
    const Vec usev = use_next;
    const uint32_t w0 = (uint32_t)env.vget(rec, STEP_COUNTS);
    const int m = (int)((w0 >> 16) & 0x7fu);
    if (s + 1 < p.nsteps) {     // everything the next step needs from the plan, one step ahead
      rec_next = env.vload(p.step_rec + (s + 1) * STEP_WORDS, STEP_WORDS, lane);
      use_next = env.vload(p.use + env.vget(rec, STEP_NEXT_USE_OFF), env.vget(rec, STEP_NEXT_NUSE), lane);
    }
    // leaving / entering lists: from the step record, or (first step) from the CSR arrays
    const bool ovf = (w0 >> 31) != 0u;
    const int l0 = ovf ? XMHW_LDG(p.leave_off + s) : 0;
    const int n_leave = ovf ? XMHW_LDG(p.leave_off + s + 1) - l0 : (int)(w0 & 0xffu);
    const int eoff = ovf ? XMHW_LDG(p.enter_off + s) : env.vget(rec, STEP_ENTER_OFF);
    const int n_enter = ovf ? XMHW_LDG(p.enter_off + s + 1) - eoff : (int)((w0 >> 8) & 0xffu);
    uint32_t plo = 0u, phi = 0u;          // sum of the first leaving list, subtracted after the walk
    if (n_leave > 0)                      // no use of the loaded words here: the load stays in flight
      leave_list(ovf ? XMHW_LDG(p.leave + l0) : env.vget(rec, STEP_LEAVE), plo, phi);
    for (int j = 1; j < n_leave; ++j) {
      uint32_t slo, shi;
      leave_list(ovf ? XMHW_LDG(p.leave + l0 + j) : env.vget(rec, (STEP_LEAVE + j) & 31), slo, shi);
      wsum = wsum - f64_from(slo, shi);
    }
    // The list that will sit in the prefetch registers AFTER this step: the one named by the last
    // loaded entry of the step, or (no load this step) the one already pending.  Its row indices
    // are requested now; its loads are issued by prefetch_pending(), unconditionally (a step without a
    // load simply re-reads the pending list: L2 hits).
    int jl = -1;
    if (!ovf) {
      for (int j = 0; j < n_enter; ++j)
        if (env.vget(rec, (STEP_ENTER + 3 * j) & 31) >> 30) jl = j;
      if (jl >= 0) {
        pend_off = env.vget(rec, (STEP_NEXT_LOAD + 2 * jl) & 31);
        pend_size = env.vget(rec, (STEP_NEXT_LOAD + 2 * jl + 1) & 31);
      }
    }
    if (!ovf) pend_rows_v = env.vload(p.rows + pend_off, pend_size, lane);
#pragma unroll 1
    for (int j = 0; j < n_enter; ++j) {
      int e, base, size, keep, sbase, next_off = 0, next_size = -1;
      if (ovf) {
        e = XMHW_LDG(p.enter + eoff + j);
        const int id = e & 0x3fffffff;
        base = XMHW_LDG(p.inst_base + id); size = XMHW_LDG(p.inst_size + id); keep = XMHW_LDG(p.inst_keep + id);
        sbase = XMHW_LDG(p.inst_sbase + id);
      } else {
        e = env.vget(rec, (STEP_ENTER + 3 * j) & 31);
        const uint32_t pk = (uint32_t)env.vget(rec, (STEP_ENTER + 3 * j + 1) & 31);
        sbase = env.vget(rec, (STEP_ENTER + 3 * j + 2) & 31);
        base = (int)(pk & 0xffffu); size = (int)((pk >> 16) & 0xffu); keep = (int)(pk >> 24);
        next_off = env.vget(rec, (STEP_NEXT_LOAD + 2 * j) & 31);
        next_size = env.vget(rec, (STEP_NEXT_LOAD + 2 * j + 1) & 31);
      }
      if (enter_list(e, base, size, keep, sbase) && (ovf || j != jl)) {
        // loaded, and another list of THIS step still has to be loaded (first step; the split
        // leap / non-leap lists around Feb 29): fetch it right away
        if (next_size < 0) next_load(eoff + j + 1, next_off, next_size);
        if (ovf) { pend_off = next_off; pend_size = next_size; }
        if (next_size > 0) prefetch_rows(env.vload(p.rows + next_off, next_size, lane), next_size);
      }
    }
    if (ovf) pend_rows_v = env.vload(p.rows + pend_off, pend_size, lane);
    cur_use = usev; cur_m = m; cur_plo = plo; cur_phi = phi;
  }

  XMHW_HD void walk_phase(int s, double& thresh, double& seas) {
    (void)s;
    const Vec usev = cur_use;
    const int m = cur_m;
    const uint32_t plo = cur_plo, phi = cur_phi;
    // stage the base rows of the lists in use (padded to a multiple of 4 with the null list)
    const int m4 = (m + 3) & ~3;
    uint32_t* ub = pool + p.pool_rows * 32;
    env.vstage(ub, usev, m, m4, lane);

    const bool live = n > 0;
    nzero += live ? 0 : 1;
    if (!env.any(live)) {                                             // all-land warp
      wsum = wsum - f64_from(plo, phi);
      thresh = qnan(); seas = qnan();
      return;
    }
    // numpy 'linear' quantile: v = (n-1) q, a = s[floor v], b = s[floor v + 1]; v >= n-1 -> max
    int target = 0;
    double gamma = 0.0;
    if (live) {
      const double nm1 = (double)(n - 1);
      const double v = nm1 * p.q;
      double fl = floor(v);
      gamma = v - fl;
      if (v >= nm1) { fl = nm1; gamma = 0.0; }
      target = n - (int)fl;       // rank (1-based, from the top) of s[floor v]
    }
    // Walk.  A lane is off target by d = C - target: it must DROP the d smallest keys above the
    // cut (d > 0) or ADD the -d largest keys below it (d < 0), in order.  With transformed
    // candidates (cinc for drops, ~cexc for adds) both are "pop the smallest head of the lists".
    // One scan certifies the FRONT smallest heads (all other heads are >= bound = the largest of
    // them); a pop replaces that list's head by its next key, which re-enters the front iff it
    // is <= bound, otherwise the certified front just shrinks.  A rescan happens only when some
    // lane exhausts its front -- typically one directional scan + the final scan per doy
    // instead of one scan per two moves.
    int d = live ? C - target : 0;
    XMHW_STAT_STEP(d);
    for (int j = 0; j < m; ++j) { XMHW_STAT_PTR0((int)ub[j], meta_ptr(at((int)ub[j] + POOL_META))); }
    while (env.any(d != 0)) {
      XMHW_STAT_SCAN();
      const bool drop = d >= 0;
      // direction constants of this lane (a walk never changes direction): transformed key = key ^ flip,
      // rowB = the cached head on the side being consumed (scanned here), rowA = the other side
      const uint32_t flip = drop ? 0u : 0xffffffffu;
      const int sgn = drop ? -1 : 1, nroff = drop ? -2 : 1;
      const int rowA = drop ? POOL_CEXC : POOL_CINC, rowB = drop ? POOL_CINC : POOL_CEXC;
      uint32_t f[XMHW_FRONT];
      int g[XMHW_FRONT];
#pragma unroll
      for (int i = 0; i < XMHW_FRONT; ++i) { f[i] = 0xffffffffu; g[i] = 0; }
#pragma unroll 4
      for (int j = 0; j < m4; ++j) {
        const int x = (int)ub[j];
        front_insert<XMHW_FRONT>(at(x + rowB) ^ flip, x, f, g);
      }
      const uint32_t bound = f[XMHW_FRONT - 1];
      int nf = XMHW_FRONT;
      // pop until every lane is on target or some lane has used up its certified front
      while (true) {
        const bool mv = d != 0 && nf > 0;
        if (!env.any(mv)) break;
        // ---- apply the move of key f[0] in list g[0]
        const int g0 = g[0];
        const uint32_t f0 = f[0];
        const uint32_t meta = at(g0 + POOL_META);
        const int pa = meta_ptr(meta), la = meta_len(meta);
        const int nr = pa + nroff;                                   // rank of the list's next head
        const bool has = (unsigned)nr < (unsigned)la;                // drop: ptr >= 2, add: ptr + 1 < len
        const uint32_t nk = key_at(g0, meta, nr, mv && has);
        if (mv) { XMHW_STAT_POP(has && nr >= meta_keep(meta)); if (has) { XMHW_STAT_RANK(nr); XMHW_STAT_OFF(g0, nr); } }
        if (mv) {
          const uint32_t ntk = has ? (nk ^ flip) : 0xffffffffu;      // transformed next head
          at(g0 + POOL_META) = meta + (uint32_t)(sgn * (int)XMHW_META_PTR1);
          at(g0 + rowA) = f0 ^ flip;                                 // the key that crossed the cut
          at(g0 + rowB) = ntk ^ flip;                                // next head, or the empty sentinel of that side
          C += sgn; d += sgn;
          // front: remove f0, insert the list's next head if it is certified (<= bound)
          const bool cert = ntk <= bound && ntk != 0xffffffffu;
          front_replace_head<XMHW_FRONT>(cert ? ntk : 0xffffffffu, g0, f, g);   // ~0 never enters
          nf += cert ? 0 : -1;
        }
      }
    }
    // final scan: the two smallest keys above the cut
    uint32_t i1, i2, e1 = 0u, e2 = 0u;
    int bi1, bi2, be1 = 0, be2 = 0;
    scan<true, false>(ub, m4, i1, i2, bi1, bi2, e1, e2, be1, be2);
    // f64 sum of the window = running sum of the list sums (exact whenever the list sums are:
    // f32 samples of similar exponent); a non-finite value (inf samples) is rebuilt from the lists
    wsum = wsum - f64_from(plo, phi);
    if (env.any(!(wsum - wsum == 0.0))) {
      double fresh = 0.0;
      for (int j = 0; j < m4; ++j) {
        const size_t sb = (size_t)meta_sbase(at((int)ub[j] + POOL_META)) * 32 + lane;
        fresh = fresh + f64_from(scratch[sb + SCR_SUM * 32], scratch[sb + (SCR_SUM + 1) * 32]);
      }
      if (!(wsum - wsum == 0.0)) wsum = fresh;
    }
    const double sum = wsum;
    const uint32_t m1 = i1, m2 = i2;
    const uint32_t meta1 = at(bi1 + POOL_META);
    const int ptr1 = meta_ptr(meta1);
    const uint32_t c2r = key_at(bi1, meta1, ptr1 - 2, live && ptr1 >= 2);
    if (live) {
      const uint32_t c2 = ptr1 >= 2 ? c2r : 0xffffffffu;
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
// Plain rules (fuzz-verified against the reference pandas code, tests/test_oracle_vs_reference.py):
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
  // Runs shorter than min_dur never qualify and never bridge a gap (the duration filter comes
  // before the join), so those lying strictly inside the word can be erased with a few
  // word-parallel operations before the sequential scan: most exceedance runs are short, and
  // the scan cost is per transition of the slowest lane of the warp.  Runs touching bit 0 or
  // bit 31 may continue in the neighbouring word and are kept as they are.
  XMHW_HD uint32_t erase_short_runs(uint32_t b) const {
    if (min_dur <= 1 || min_dur > 32) return b;
    uint32_t open = b;                                   // bit i: b[i .. i + min_dur - 1] all ones
    for (int rem = min_dur - 1, s = 1; rem > 0; s <<= 1) {
      const int step = s < rem ? s : rem;
      open &= open >> step;
      rem -= step;
    }
    uint32_t lng = open;                                 // dilate back: all bits of runs >= min_dur
    for (int rem = min_dur - 1, s = 1; rem > 0; s <<= 1) {
      const int step = s < rem ? s : rem;
      lng |= lng << step;
      rem -= step;
    }
    const uint32_t low_run = b & ~(b + 1u);              // trailing ones (run through bit 0)
    const uint32_t nb = ~b;
    const uint32_t high_run = nb ? ~(0xffffffffu >> clz32(nb)) : 0xffffffffu;   // leading ones (run through bit 31)
    return lng | low_run | (b & high_run);
  }

  // bits: bit i = exceedance at time t0 + i (bits past the series end are 0)
  template <class Emit> XMHW_HD void feed(uint32_t bits, int t0, Emit& emit) {
    if (t0 == 0) bits &= ~1u;
    bits = erase_short_runs(bits);
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
  // Eager emission (fused detect kernel): after everything before time t_end was fed, the pending
  // event is final as soon as no later qualified run can still join it -- the next run starts at
  // run_start (a run in progress) or at t_end at the earliest.  Same events, same order as with
  // finish() alone; only the moment of emission moves forward.
  template <class Emit> XMHW_HD void flush_pending(int t_end, Emit& emit) {
    if (ps < 0) return;
    const int nxt = run_start >= 0 ? run_start : t_end;
    if (!join || nxt - pe - 1 > max_gap) { emit(ps, pe); ps = -1; }
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

enum { EV_BATCH_DEFAULT = 4 };

// NaN-skipping running sums (pandas groupby mean/sum/var(ddof=1) skip NaN).  Plain f64
// sum and sum of squares: for this data (|x| < 1e3, an event is < 1e4 days) the cancellation
// error of (sq - sum^2/n) is < 1e-9 relative, far inside the 1e-5 / 1 f32 ulp contract, and
// it costs 3 FP64 operations per day instead of the division of a Welford update.
struct Moments {
  int n; double sum, sq;
  XMHW_HD Moments() : n(0), sum(0.0), sq(0.0) {}
  XMHW_HD void add(double x) {
    if (x != x) return;
    ++n; sum = sum + x; sq = sq + x * x;
  }
  XMHW_HD double avg() const { return n ? sum / (double)n : qnan(); }
  XMHW_HD double sd() const {
    if (n < 2) return qnan();
    const double v = (sq - sum * sum / (double)n) / (double)(n - 1);
    return v > 0.0 ? sqrt(v) : (v == v ? 0.0 : v);
  }
};

XMHW_HD double round_f32(double x) { return (double)(float)x; }

// One event [s, e] of the cell whose series starts at `col` (stride ngrid),
// thresholds/seasonal at th/se (doy-major, stride ngrid), doy[t] 1-based.
// Climatology access of one cell: get(d, thresh, seas) for the 0-based day-of-year d.
// Doy-major = the reference layout (doy, cell); the CUDA path also has a cell-major
// interleaved copy (xmhw_kernels.cu) so that the ~12 consecutive days of an event are one
// contiguous 200-byte run instead of 2 x 12 scattered 32-byte sectors.
struct ClimDoyMajor {
  const double* th; const double* se; int64_t ngrid;
  XMHW_HD void get(int d, double& t, double& s) const {
    t = XMHW_LDG(th + (int64_t)d * ngrid);
    s = XMHW_LDG(se + (int64_t)d * ngrid);
  }
};

template <int EV_BATCH, class Clim>
XMHW_HD void event_stats(const float* col, const Clim& clim, const int32_t* doy,
                         int64_t ngrid, int T, int s, int e, int32_t* oi, double* of, int64_t stride);

template <int EV_BATCH = EV_BATCH_DEFAULT>
XMHW_HD void event_stats(const float* col, const double* th, const double* se, const int32_t* doy,
                         int64_t ngrid, int T, int s, int e, int32_t* oi, double* of, int64_t stride) {
  ClimDoyMajor clim{th, se, ngrid};
  event_stats<EV_BATCH, ClimDoyMajor>(col, clim, doy, ngrid, T, s, e, oi, of, stride);
}

template <int EV_BATCH, class Clim>
XMHW_HD void event_stats(const float* col, const Clim& clim, const int32_t* doy,
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
    double t_, s_;
    clim.get(XMHW_LDG(doy + s - 1) - 1, t_, s_);
    prev_anom = (double)XMHW_LDG(col + (int64_t)(s - 1) * ngrid) - s_;
  }
  // days are processed in batches of EV_BATCH with all loads of a batch issued first
  // (memory-level parallelism; the per-day arithmetic is a serial f64 chain)
  for (int t0 = s; t0 <= e; t0 += EV_BATCH) {
    float xs[EV_BATCH];
    double thb[EV_BATCH], seb[EV_BATCH];
#pragma unroll
    for (int i = 0; i < EV_BATCH; ++i) {
      const int tt = t0 + i <= e ? t0 + i : e;           // clamped: loads stay unconditional
      const int dd = XMHW_LDG(doy + tt) - 1;
      xs[i] = XMHW_LDG(col + (int64_t)tt * ngrid);
      clim.get(dd, thb[i], seb[i]);
    }
#pragma unroll
    for (int i = 0; i < EV_BATCH; ++i) {
      const int t = t0 + i;
      if (t > e) break;
      const double x = (double)xs[i], thr = thb[i], sea = seb[i];
      double relS = x - sea;                 // features.py:52
      double relT = x - thr;                 // :53
      double ths = thr - sea;                // :54
      double norm = relT / ths;              // :57
      double one_norm = 1.0 + norm;
      // :59-61 severity = relS / -(ths); relS = relT + ths up to an ulp, so it is -(1 + norm)
      // to ~2 f64 ulp (inf / NaN cases agree) -- one division per day instead of two
      double sev = -one_norm;
      double cat = floor(one_norm);          // :62
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
  }
  if (e + 1 <= T - 1) {       // anom_minus of the last event day
    double t_, s_;
    clim.get(XMHW_LDG(doy + e + 1) - 1, t_, s_);
    double a = (double)XMHW_LDG(col + (int64_t)(e + 1) * ngrid) - s_;
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
