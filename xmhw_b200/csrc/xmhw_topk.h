// Two-stack top-K climatology sweep (xmhw_clim_sweep2_f32; the default sweep wherever its plan fits).
//
// Reference semantics: for every (cell, doy d) pool ts[t+k], doy[t] = d, |k| <= w, drop NaN,
// take numpy's 'linear' quantile and the mean (xmhw/identify.py:184-209 window_roll,
// :233 groupby("doy").quantile, :263 groupby("doy").mean).
//
// Design.  The quantile of a window of n samples only needs the K-th and (K-1)-th LARGEST
// sample, K = n - floor((n-1) q) (35 of 337 at q = 0.9).  The host plan (xmhw_b200/plan2.py)
// orders the time rows into ATOMS (rows that enter and leave the doy windows together,
// normally the ~30 years of one calendar day) such that the window of every doy is a
// contiguous range of atoms whose two ends only move forward.  A sliding range is a queue,
// and a queue of mergeable summaries is two stacks:
//   back stack   A = top-KP keys (sorted, registers) of everything pushed since the last flip;
//                each pushed atom is sorted once by a register network, merged into A by a
//                bitonic top-K merge and parked ("stash") in its unit's shared-memory slot;
//   front stack  one sorted top-KP array per UNIT (atoms that leave together) = top-KP of that
//                unit and all younger units of the front; built at a FLIP by walking the
//                stashed units from the youngest to the oldest, merging into the (then free)
//                accumulator and storing it over the unit's own stash;
//   query        K-th largest of (front array S) u (A):  max_i min(A[i-1], S[K-1-i]); the warp loops over
//                its DISTINCT ranks K (they differ between lanes only where samples are missing), so the
//                slot rows are addressed warp-uniformly -- which lets a slot live in shared memory OR in
//                tensor-memory columns (tcgen05.ld / st, see SplitPool in xmhw_kernels.cu).
// Every step is the same straight-line code for all lanes (sorting / merging networks and a
// fixed-length max-min scan): no data-dependent walk, no SIMT loss, no global key scratch.
// One lane = one grid cell like every other kernel of this library.
#pragma once
#include "xmhw_lane.h"

namespace xmhw {

// Plan of the two-stack sweep (host side: xmhw_b200/plan2.py; C mirror: xmhw_clim_plan2).  The
// whole plan is ONE plain struct of ~24 KB that travels in the kernel's parameter space
// (__grid_constant__): every field is read with warp-uniform constant loads that do not touch
// the load scoreboards the sample prefetch sits on, and there is no plan array in global memory.
enum { SC_MAX_STEPS = 366, SC_REC_WORDS = 12, SC_MAX_FLIP = 768, SC_MAX_PAT = 16, SC_PAT_LEN = 48, SC_MAX_INIT = 32 };
struct ClimPlan2 {
  int32_t nsteps;        // sweep steps (regular doys, in doy order)
  int32_t kp;            // capacity of the top-K arrays the plan needs (max rank over all sample counts)
  int32_t max_size;      // rows of the largest atom
  int32_t slot_rows;     // shared-memory rows per unit slot: 1 (len | guard) + cap + 2 (f64 sum)
  int32_t nslots;        // slots (= units alive at once)
  int32_t n_init;        // atoms pushed before the first step
  int32_t cap;           // key rows per slot = max(kp, rows of the largest unit)
  int32_t reserved_;
  double q;              // quantile in [0,1]; numpy 'linear': v = (n-1) q
  uint32_t rec[SC_MAX_STEPS][SC_REC_WORDS];   // step records (REC_*)
  uint32_t flip[SC_MAX_FLIP];                 // flip entries (FLIP_*)
  int32_t pat[SC_MAX_PAT][SC_PAT_LEN];        // row patterns: time index of row i of an atom = first row + pat[id][i]
  uint32_t init[SC_MAX_INIT][2];              // atoms of the first window, then the first atom pushed by the sweep
};

// atom descriptor, 2 words:
//   word 0  first row [0:24] | rows [24:30] | JOB_F_COPY, JOB_F_FIRST [30:32]
//   word 1  pattern [0:5] | slot [5:10] | key-row offset in the slot [10:17] | ragged [17] (rows != size class)
// step record, SC_REC_WORDS words:
//   word 0  n_pop [0:3] | n_push [3:5] | flip after the pushes [5] | n_flip [6:12] | front slot [12:17] | output row [17:32]
//   word 1  popped slots, 5 bits each        word 2  bit mask of the slots alive at the query
//   word 3  first flip entry                 words 4..9  up to 3 pushed atoms
//   words 10, 11  the atom pushed after this step's last push (prefetch target; rows 0 = none)
// flip entry, 1 word: source slot [0:5] | key-row offset [5:12] | rows [12:18] | JOB_F_* [18:24] | destination slot [24:29]
enum { REC_POPS = 1, REC_ALIVE = 2, REC_FLIP_OFF = 3, REC_PUSH = 4, REC_NEXT = 10 };

// job flags (pushes and flip entries)
enum { JOB_F_COPY = 1,      // the accumulator is empty: accumulator := this list (first push after a flip / first of a chain)
       JOB_F_FIRST = 2,     // push: first atom of its unit (initialises the slot's len row and f64 sum)
       JOB_F_STORE = 4,     // flip: store the accumulator over the unit's slot after merging this atom
       JOB_F_STOREP = 8,    // flip, no atom: store the accumulator as it is (the oldest unit's array = everything pushed)
       JOB_F_CLEAR = 16,    // flip, no atom: accumulator := empty
       JOB_F_RAGGED = 32 }; // the atom has fewer rows than its size class: predicated stash stores / loads

#define XMHW_GUARD 0xffffff00u          // len row = guard | len: above every real key (key(+inf) = 0xff800000)

template <int N> XMHW_HD void bitonic_valley_desc(uint32_t (&k)[N]);
template <> XMHW_HD void bitonic_valley_desc<8>(uint32_t (&k)[8]) { XMHW_BITONIC_8 }
template <> XMHW_HD void bitonic_valley_desc<16>(uint32_t (&k)[16]) { XMHW_BITONIC_16 }
template <> XMHW_HD void bitonic_valley_desc<24>(uint32_t (&k)[24]) { XMHW_BITONIC_24 }
template <> XMHW_HD void bitonic_valley_desc<36>(uint32_t (&k)[36]) { XMHW_BITONIC_36 }
template <> XMHW_HD void bitonic_valley_desc<48>(uint32_t (&k)[48]) { XMHW_BITONIC_48 }


// maximum of t[0..N): a balanced tree instead of a serial chain (depth log N; ptxas pairs the levels into
// 3-input VIMNMX3), t is clobbered
template <int W, int N> XMHW_HD uint32_t umax_tree_w(uint32_t (&t)[N]) {      // the first W entries of t are live
  if constexpr (W <= 1) {
    return t[0];
  } else {
    constexpr int W3 = (W + 2) / 3;
#pragma unroll
    for (int i = 0; i < W3; ++i) {
      uint32_t m = t[3 * i];
      if (3 * i + 1 < W) m = umax32(m, t[3 * i + 1]);
      if (3 * i + 2 < W) m = umax32(m, t[3 * i + 2]);
      t[i] = m;
    }
    return umax_tree_w<W3, N>(t);
  }
}
template <int N> XMHW_HD uint32_t umax_tree(uint32_t (&t)[N]) { return umax_tree_w<N, N>(t); }

// A (KP keys, descending) := the KP largest of A u L (N keys, descending), descending.
// max(A[i], L[KP-1-i]) is the top KP as a valley; the pruned bitonic merger sorts it.
template <int KP, int N>
XMHW_HD void merge_topk(uint32_t (&A)[KP], const uint32_t (&L)[N]) {
#pragma unroll
  for (int i = 0; i < KP; ++i) {
    const int j = KP - 1 - i;
    if (j < N) A[i] = umax32(A[i], L[j]);
  }
  bitonic_valley_desc<KP>(A);
}

// Where a warp's unit slots live.  Every access of the sweep names a WARP-UNIFORM row (the lane only selects
// the word of the row), so a pool may be any memory with "row of 32 lane-private words" addressing:
// PlainPool = shared memory on the device / plain memory in the host emulator; the CUDA file adds a pool whose
// upper rows are columns of tensor memory (tcgen05.ld / st, 32 lanes x 32 bit shape).
struct PlainPool {
  uint32_t* p;           // row 0, this lane's word
  int min_row;           // lowest row (<= 0) that is still readable memory: the rows of the block's earlier warps
  static constexpr bool kGather = false;      // query: load each front row right where it is used
  XMHW_HD PlainPool(uint32_t* base, int lane, int lowest = 0) : p(base + lane), min_row(lowest) {}
  XMHW_HD uint32_t ld(int row) const { return p[row * 32]; }
  XMHW_HD void st(int row, uint32_t v) const { p[row * 32] = v; }
  template <int N> XMHW_HD void ld_block(int row0, uint32_t (&k)[N]) const {
    const uint32_t* const r = p + row0 * 32;
#pragma unroll
    for (int i = 0; i < N; ++i) k[i] = r[i * 32];
  }
  template <int N> XMHW_HD void ld_block_n(int row0, int size, uint32_t (&k)[N]) const {     // rows past `size`: 0
    const uint32_t* const r = p + row0 * 32;
#pragma unroll
    for (int i = 0; i < N; ++i) k[i] = i < size ? r[i * 32] : 0u;
  }
  template <int N> XMHW_HD void st_block(int row0, const uint32_t (&k)[N]) const {
    uint32_t* const r = p + row0 * 32;
#pragma unroll
    for (int i = 0; i < N; ++i) r[i * 32] = k[i];
  }
  template <int N> XMHW_HD void st_block_n(int row0, int size, const uint32_t (&k)[N]) const { // only the first `size` rows
    uint32_t* const r = p + row0 * 32;
#pragma unroll
    for (int i = 0; i < N; ++i)
      if (i < size) r[i * 32] = k[i];
  }
  // Front rows of the query: sv[i] = row kk - i of the slot at `base` for i <= kk.  The entries past the
  // guard row (i > kk) only enter terms that are dominated whatever their value (min(A[i-1], x) <= A[i-1] <=
  // A[kk-1], the i = kk term), so when the rows below the slot are readable memory they are read as they
  // come (one address per row instead of compare + select + add); else they are clamped onto the guard row.
  template <int N> XMHW_HD void ld_front(int base, int kk, uint32_t (&sv)[N]) const {
    if (base + kk - (N - 1) >= min_row) {
      const uint32_t* const r = p + (base + kk) * 32;
#pragma unroll
      for (int i = 0; i < N; ++i) sv[i] = *(r - i * 32);
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) sv[i] = p[(base + (kk - i > 0 ? kk - i : 0)) * 32];
    }
  }
};

template <class Env, class Pool, int KP, int MAXN>
struct TopkSweeperP {
  const Env& env;
  const ClimPlan2& p;
  Pool pool;             // this warp's unit-slot rows (32 words each; word = lane)
  const float* col;
  const int64_t ngrid;
  const bool ok;
  uint32_t A[KP];        // back stack summary: top-KP keys pushed since the last flip
  float pv[MAXN];        // prefetched rows of the next atom
  int n;                 // valid samples in the window
  int nzero;             // steps without any sample (feeds the per-cell compaction of the smoothing)
  double wsum;           // f64 sum of the window (+ pushed unit sums, - popped)

  XMHW_HD TopkSweeperP(const Env& e, const ClimPlan2& pl, const Pool& po, const float* c, int64_t ng, bool k)
      : env(e), p(pl), pool(po), col(c), ngrid(ng), ok(k), n(0), nzero(0), wsum(0.0) {
#pragma unroll
    for (int i = 0; i < KP; ++i) A[i] = 0u;
  }

  XMHW_HD uint32_t at(int row) const { return pool.ld(row); }

  // issue the loads of the atom (d0, d1): row i = first row + pattern[i]
  XMHW_HD void prefetch(uint32_t d0, uint32_t d1) {
    const uint32_t ng4 = (uint32_t)ngrid * 4u;           // bytes per time row (ngrid < 2^30)
    const uint32_t first = d0 & 0xffffffu;
    const int32_t* const pt = p.pat[d1 & 31u];
    const char* const cb = reinterpret_cast<const char*>(col);
    // unconditional loads (rows past the atom repeat its first row and are masked in the job)
#pragma unroll
    for (int i = 0; i < MAXN; ++i)
      pv[i] = XMHW_LDG(reinterpret_cast<const float*>(cb + (uint64_t)(first + (uint32_t)pt[i]) * ng4));
  }

  // One list of at most N keys goes into the accumulator.  PUSH: the prefetched atom -- keys, f64
  // sum, register sort, stash in its unit's slot.  Otherwise (flip): a stashed atom read back from
  // its slot.  Both kinds share ONE merge site per size class (the merges are the bulk of the code).
  template <int N>
  XMHW_HD void job(bool push, int size, int flags, int slot_base, int off, bool alive) {
    uint32_t k[N];
    bool acc = alive;
    const int srow = slot_base + 1 + off;          // first key row of the atom in its slot
    if (push) {
      int len = 0;
      double sum = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const float v = pv[i];
        const uint32_t b = f32_bits(v);
        const bool valid = (i < size) && ok && (v == v);
        k[i] = valid ? (b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u)) : 0u;
        if (valid) { ++len; sum = sum + (double)v; }
      }
      acc = env.any(len > 0);
      if (acc) sort_desc<N>(k);
      if (flags & JOB_F_RAGGED) pool.template st_block_n<N>(srow, size, k);
      else pool.template st_block<N>(srow, k);
      const int lrow = slot_base, sum_row = slot_base + 1 + p.cap;
      if (flags & JOB_F_FIRST) {
        pool.st(lrow, XMHW_GUARD | (uint32_t)len);
        pool.st(sum_row, f64_lo(sum)); pool.st(sum_row + 1, f64_hi(sum));
      } else {
        pool.st(lrow, pool.ld(lrow) + (uint32_t)len);
        const double s2 = f64_from(pool.ld(sum_row), pool.ld(sum_row + 1)) + sum;
        pool.st(sum_row, f64_lo(s2)); pool.st(sum_row + 1, f64_hi(s2));
      }
      n += len;
      wsum = wsum + sum;
    } else if (flags & JOB_F_RAGGED) {
      pool.template ld_block_n<N>(srow, size, k);
    } else {
      pool.template ld_block<N>(srow, k);
    }
    if (flags & JOB_F_COPY) {
#pragma unroll
      for (int i = 0; i < KP; ++i) A[i] = (acc && i < N) ? k[i] : 0u;
    } else if (acc) {
      merge_topk<KP, N>(A, k);
    }
  }

  XMHW_HD void store_acc(int slot_base, bool alive) {
    if (alive) {
      pool.template st_block<KP>(slot_base + 1, A);
    } else {
      uint32_t z[KP];
#pragma unroll
      for (int i = 0; i < KP; ++i) z[i] = 0u;
      pool.template st_block<KP>(slot_base + 1, z);
    }
  }

  // ---- the same step with the pushes and the flip entries in two separate loops (step_phased): each loop
  // body updates the accumulator along ONE path, which keeps ptxas from copying the 36 accumulator
  // registers into a working set and back around every job (72 moves per job in step()).
  XMHW_HD void push_jobs(bool fill, const uint32_t* rec, int n_push) {
#pragma unroll 1
    for (int j = 0; j < n_push; ++j) {
      const uint32_t d0 = fill ? p.init[j][0] : rec[REC_PUSH + 2 * j];
      const uint32_t d1 = fill ? p.init[j][1] : rec[REC_PUSH + 2 * j + 1];
      const bool last = j + 1 == n_push;
      const uint32_t nx0 = fill ? p.init[j + 1][0] : (last ? rec[REC_NEXT] : rec[REC_PUSH + 2 * j + 2]);
      const uint32_t nx1 = fill ? p.init[j + 1][1] : (last ? rec[REC_NEXT + 1] : rec[REC_PUSH + 2 * j + 3]);
      const int size = (int)((d0 >> 24) & 63u);
      const int flags = (int)(d0 >> 30) | (int)(((d1 >> 17) & 1u) * JOB_F_RAGGED);
      const int slot_base = (int)((d1 >> 5) & 31u) * p.slot_rows;
      const int srow = slot_base + 1 + (int)((d1 >> 10) & 127u);
      uint32_t k[MAXN];
      // XMHW_PUSH_SUMS interleaved partial counts / f64 sums (1 = one sequential chain).  Measured on B200 (global
      // grid, same box): 1 chain 34.8 ms, 2 chains 38.1 ms, 3 chains 37.8 ms -- ptxas needs 10 more registers and
      // schedules the conversion worse; the sequential chain stays.
      int len3[3] = {0, 0, 0};
      double sum3[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int i = 0; i < MAXN; ++i) {
        const float v = pv[i];
        const uint32_t b = f32_bits(v);
        const bool valid = (i < size) && (v == v);      // lanes past the grid edge compute on cell 0: never stored
        k[i] = valid ? (b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u)) : 0u;
#ifndef XMHW_PUSH_SUMS
#define XMHW_PUSH_SUMS 1
#endif
        len3[i % XMHW_PUSH_SUMS] += valid ? 1 : 0;
        sum3[i % XMHW_PUSH_SUMS] = sum3[i % XMHW_PUSH_SUMS] + (double)(valid ? v : 0.0f);   // an invalid sample adds +0.0f (exact): one f32 select
      }
      const int len = (len3[0] + len3[1]) + len3[2];
      const double sum = (sum3[0] + sum3[1]) + sum3[2];
      if ((nx0 >> 24) & 63u) prefetch(nx0, nx1);              // pv is free again: the next atom's loads go out now
      const bool acc = env.any(len > 0);
      if (acc) sort_desc<MAXN>(k);
      if (flags & JOB_F_RAGGED) pool.template st_block_n<MAXN>(srow, size, k);
      else pool.template st_block<MAXN>(srow, k);
      const int lrow = slot_base, sum_row = slot_base + 1 + p.cap;
      if (flags & JOB_F_FIRST) {
        pool.st(lrow, XMHW_GUARD | (uint32_t)len);
        pool.st(sum_row, f64_lo(sum)); pool.st(sum_row + 1, f64_hi(sum));
      } else {
        pool.st(lrow, pool.ld(lrow) + (uint32_t)len);
        const double s2 = f64_from(pool.ld(sum_row), pool.ld(sum_row + 1)) + sum;
        pool.st(sum_row, f64_lo(s2)); pool.st(sum_row + 1, f64_hi(s2));
      }
      n += len;
      wsum = wsum + sum;
      if (flags & JOB_F_COPY) {                                // accumulator := this atom (= merge into an empty one)
#pragma unroll
        for (int i = 0; i < KP; ++i) A[i] = 0u;
      }
      if (acc) merge_topk<KP, MAXN>(A, k);
    }
  }

  XMHW_HD void flip_jobs(int n_flip, int flip_off) {
    const bool alive = env.any(n > 0);                         // the window as it is when the flip starts
#pragma unroll 1
    for (int e = 0; e < n_flip; ++e) {
      const uint32_t f0 = p.flip[flip_off + e];
      const int slot_base = (int)(f0 & 31u) * p.slot_rows;
      const int srow = slot_base + 1 + (int)((f0 >> 5) & 127u);
      const int size = (int)((f0 >> 12) & 63u), flags = (int)((f0 >> 18) & 63u);
      const int dst_base = (int)((f0 >> 24) & 31u) * p.slot_rows;
      if (flags & (JOB_F_CLEAR | JOB_F_COPY)) {
#pragma unroll
        for (int i = 0; i < KP; ++i) A[i] = 0u;
      }
      if (!(flags & (JOB_F_CLEAR | JOB_F_STOREP)) && alive) {
        uint32_t k[MAXN];
        if (flags & JOB_F_RAGGED) pool.template ld_block_n<MAXN>(srow, size, k);
        else pool.template ld_block<MAXN>(srow, k);
        merge_topk<KP, MAXN>(A, k);
      }
      if (flags & JOB_F_STOREP) store_acc(dst_base, alive);
      else if (flags & JOB_F_STORE) store_acc(slot_base, alive);
    }
  }

  XMHW_HD void step_phased(int s, double& thresh, double& seas, int& out_row) {
    const bool fill = s < 0;
    const uint32_t* const rec = p.rec[fill ? 0 : s];
    const uint32_t w0 = fill ? 0u : rec[0];
    const int n_pop = (int)(w0 & 7u), n_push = fill ? p.n_init : (int)((w0 >> 3) & 3u), n_flip = (int)((w0 >> 6) & 63u);
    const bool flip_late = ((w0 >> 5) & 1u) != 0u;
    const int front_base = (int)((w0 >> 12) & 31u) * p.slot_rows;
    out_row = (int)(w0 >> 17);
    double psum = 0.0;
    {
      const uint32_t pw = fill ? 0u : rec[REC_POPS];
      for (int j = 0; j < n_pop; ++j) {
        const int sb = (int)((pw >> (5 * j)) & 31u) * p.slot_rows;
        n -= (int)(at(sb) & 0xffu);
        psum = psum + f64_from(at(sb + 1 + p.cap), at(sb + 2 + p.cap));
      }
    }
    if (fill) prefetch(p.init[0][0], p.init[0][1]);
    const int flip_off = fill ? 0 : (int)rec[REC_FLIP_OFF];
#pragma unroll 1
    for (int ph = 0; ph < 2; ++ph) {                           // one site of each loop: flips before or after the pushes
      if ((ph == 0) == flip_late) push_jobs(fill, rec, n_push);
      else if (n_flip) flip_jobs(n_flip, flip_off);
    }
    if (fill) { thresh = qnan(); seas = qnan(); return; }
    query(rec, psum, front_base, thresh, seas);
  }

  // The first formulation of the step (one loop over pushes and flip entries through the shared job<>() body):
  // pops, then the step's jobs (a flip's entries before or after the pushes), then the query; s = -1 is the
  // initial fill (plan.init).  The kernels run step_phased() above; this one stays as the independent second
  // implementation the host emulator checks against the oracle (tests/test_lane_emulator.py runs both).
  XMHW_HD void step(int s, double& thresh, double& seas, int& out_row) {
    const bool fill = s < 0;
    const uint32_t* const rec = p.rec[fill ? 0 : s];
    const uint32_t w0 = fill ? 0u : rec[0];
    const int n_pop = (int)(w0 & 7u), n_push = fill ? p.n_init : (int)((w0 >> 3) & 3u), n_flip = (int)((w0 >> 6) & 63u);
    const bool flip_late = ((w0 >> 5) & 1u) != 0u;
    const int front_base = (int)((w0 >> 12) & 31u) * p.slot_rows;
    out_row = (int)(w0 >> 17);
    // pops: only the unit's sample count and sum leave the window (its keys live in no summary
    // that is still used: the front arrays are suffixes, the accumulator is younger)
    double psum = 0.0;
    {
      const uint32_t pw = fill ? 0u : rec[REC_POPS];
      for (int j = 0; j < n_pop; ++j) {
        const int sb = (int)((pw >> (5 * j)) & 31u) * p.slot_rows;
        n -= (int)(at(sb) & 0xffu);
        psum = psum + f64_from(at(sb + 1 + p.cap), at(sb + 2 + p.cap));
      }
    }
    if (fill) prefetch(p.init[0][0], p.init[0][1]);
    const int flip_off = fill ? 0 : (int)rec[REC_FLIP_OFF];
    bool alive = true;
    const int n_jobs = n_push + n_flip;
#pragma unroll 1
    for (int jb = 0; jb < n_jobs; ++jb) {
      const bool push = flip_late ? jb < n_push : jb >= n_flip;
      int size, flags, slot_base, off;
      uint32_t nx0 = 0u, nx1 = 0u;              // push: the atom to prefetch afterwards
      if (push) {
        const int j = flip_late ? jb : jb - n_flip;
        const uint32_t d0 = fill ? p.init[j][0] : rec[REC_PUSH + 2 * j];
        const uint32_t d1 = fill ? p.init[j][1] : rec[REC_PUSH + 2 * j + 1];
        const bool last = j + 1 == n_push;
        nx0 = fill ? p.init[j + 1][0] : (last ? rec[REC_NEXT] : rec[REC_PUSH + 2 * j + 2]);
        nx1 = fill ? p.init[j + 1][1] : (last ? rec[REC_NEXT + 1] : rec[REC_PUSH + 2 * j + 3]);
        size = (int)((d0 >> 24) & 63u);
        flags = (int)(d0 >> 30) | (int)(((d1 >> 17) & 1u) * JOB_F_RAGGED);
        slot_base = (int)((d1 >> 5) & 31u) * p.slot_rows;
        off = (int)((d1 >> 10) & 127u);
      } else {
        const int e = flip_late ? jb - n_push : jb;
        if (e == 0) alive = env.any(n > 0);         // the window as it is when the flip starts
        const uint32_t f0 = p.flip[flip_off + e];
        slot_base = (int)(f0 & 31u) * p.slot_rows;
        off = (int)((f0 >> 5) & 127u);
        size = (int)((f0 >> 12) & 63u);
        flags = (int)((f0 >> 18) & 63u);
        const int dst_base = (int)((f0 >> 24) & 31u) * p.slot_rows;
        if (flags & JOB_F_CLEAR) {
#pragma unroll
          for (int i = 0; i < KP; ++i) A[i] = 0u;
          continue;
        }
        if (flags & JOB_F_STOREP) { store_acc(dst_base, alive); continue; }
      }
      // ONE size class (MAXN keys, shorter atoms padded): a second instantiation of the sort / merge
      // code in this loop costs register shuffles at every call and doubles the instruction footprint
      job<MAXN>(push, size, flags, slot_base, off, alive);
      if (push) {
        if ((nx0 >> 24) & 63u) prefetch(nx0, nx1);
      } else if (flags & JOB_F_STORE) {
        store_acc(slot_base, alive);               // a unit's array lives in the unit's own slot
      }
    }
    if (fill) { thresh = qnan(); seas = qnan(); return; }
    query(rec, psum, front_base, thresh, seas);
  }

  // quantile + mean of the window as it stands after the step's pops / pushes / flips
  XMHW_HD void query(const uint32_t* rec, double psum, int front_base, double& thresh, double& seas) {
    wsum = wsum - psum;
    const bool live = n > 0;
    nzero += live ? 0 : 1;
    if (!env.any(live)) { thresh = qnan(); seas = qnan(); return; }
    // a non-finite running sum (inf samples) is rebuilt from the sums of the units alive
    if (env.any(!(wsum - wsum == 0.0))) {
      uint32_t alive_slots = rec[REC_ALIVE];
      double fresh = 0.0;
      while (alive_slots) {
        const int sb = ctz32(alive_slots) * p.slot_rows;
        alive_slots &= alive_slots - 1u;
        fresh = fresh + f64_from(at(sb + 1 + p.cap), at(sb + 2 + p.cap));
      }
      if (!(wsum - wsum == 0.0)) wsum = fresh;
    }
    // numpy 'linear': v = (n-1) q, a = s[floor v], b = s[floor v + 1]; v >= n-1 -> the maximum
    int target = 1;
    double gamma = 0.0;
    if (live) {
      const double nm1 = (double)(n - 1);
      const double v = nm1 * p.q;
      double fl = floor(v);
      gamma = v - fl;
      if (v >= nm1) { fl = nm1; gamma = 0.0; }
      target = n - (int)fl;               // rank from the top of s[floor v]; s[floor v + 1] is rank target - 1
    }
    // R(k) = k-th largest of S u A = max_i min(A[i-1], S[k-1-i]), A[-1] = S[-1] = +inf (the slot's
    // len | guard row sits at S[-1]; rows past it are clamped onto it, their terms are dominated).  s_i = S[k-1-i]
    // serves R(k) with A[i-1] and R(k-1) with A[i-2].  The rank k = target differs between lanes
    // only where samples are missing, so the warp loops over its DISTINCT ranks: inside the loop
    // every shared-memory row index is warp-uniform (no per-lane addressing).
    uint32_t r1 = 0u, r2 = 0u;
    int todo = live ? (target < KP ? target : KP) : 0;         // 0: nothing (left) to compute for this lane
    while (true) {
      const int kk = env.max_all(todo);
      if (kk == 0) break;
      // t1[i] = min(A[i-1], S[kk-i]) (A[-1] = +inf), t2[i] = min(A[i-2], S[kk-i]): R(kk) = max t1, R(kk-1) = max t2
      uint32_t t1[KP + 1], t2[KP + 1];
      {
        uint32_t sv[KP + 1];
        pool.template ld_front<KP + 1>(front_base, kk, sv);
#pragma unroll
        for (int i = 0; i <= KP; ++i) {
          t1[i] = i >= 1 ? umin32(A[i - 1], sv[i]) : sv[i];
          t2[i] = i >= 2 ? umin32(A[i - 2], sv[i]) : (i == 1 ? sv[i] : 0u);
        }
      }
      const uint32_t q1 = umax_tree<KP + 1>(t1), q2 = umax_tree<KP + 1>(t2);
      if (todo == kk) { r1 = q1; r2 = q2; todo = 0; }
    }
    if (live) {
      const uint32_t kb = target >= 2 ? r2 : r1;
      thresh = lerp_q(key_f32(r1), key_f32(kb), gamma);
      seas = wsum / (double)n;
    } else {
      thresh = qnan();
      seas = qnan();
    }
  }
};

// the sweep on a plain (shared / host) memory pool
template <class Env, int KP, int MAXN>
struct TopkSweeper : TopkSweeperP<Env, PlainPool, KP, MAXN> {
  XMHW_HD TopkSweeper(const Env& e, const ClimPlan2& pl, uint32_t* po, int ln, const float* c, int64_t ng, bool k,
                      int lowest_row = 0)
      : TopkSweeperP<Env, PlainPool, KP, MAXN>(e, pl, PlainPool(po, ln, lowest_row), c, ng, k) {}
};

// ---------------------------------------------------------------------------
// Direct selection for the few doys whose window is not a range of the atom order
// (doy 60: its window holds leap years only).  One lane streams the window's rows in
// chunks of 8 through the same sort / top-K merge and selects its ranks from registers.
// ---------------------------------------------------------------------------
template <int KP>
struct DirectSelect {
  uint32_t A[KP];
  int n;
  double sum;
  XMHW_HD DirectSelect() : n(0), sum(0.0) {
#pragma unroll
    for (int i = 0; i < KP; ++i) A[i] = 0u;
  }
  XMHW_HD void add8(const float (&v)[8], int cnt, bool ok) {
    uint32_t k[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t b = f32_bits(v[i]);
      const bool valid = i < cnt && ok && v[i] == v[i];
      k[i] = valid ? (b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u)) : 0u;
      if (valid) { ++n; sum = sum + (double)v[i]; }
    }
    sort_desc<8>(k);
    merge_topk<KP, 8>(A, k);
  }
  XMHW_HD uint32_t rank(int r) const {          // r-th largest, 1-based, r <= KP
    uint32_t x = 0u;
#pragma unroll
    for (int i = 0; i < KP; ++i) x = (i == r - 1) ? A[i] : x;
    return x;
  }
  XMHW_HD void result(double q, double& thresh, double& seas) const {
    if (n <= 0) { thresh = qnan(); seas = qnan(); return; }
    const double nm1 = (double)(n - 1);
    const double v = nm1 * q;
    double fl = floor(v), gamma = v - fl;
    if (v >= nm1) { fl = nm1; gamma = 0.0; }
    const int target = n - (int)fl;
    const uint32_t r1 = rank(target), r2 = target >= 2 ? rank(target - 1) : r1;
    thresh = lerp_q(key_f32(r1), key_f32(r2), gamma);
    seas = sum / (double)n;
  }
};

}  // namespace xmhw
