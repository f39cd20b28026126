// Two-stack top-K climatology sweep (the fast path of xmhw_clim_sweep2_f32).
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
//   query        K-th largest of (front array S) u (A):  max_i min(A[i-1], S[K-1-i]) -- the
//                array in registers is indexed statically, the one in shared memory per lane,
//                so every lane may have its own K (NaN data) without any divergence.
// Every step is the same straight-line code for all lanes (sorting / merging networks and a
// fixed-length max-min scan): no data-dependent walk, no SIMT loss, no global key scratch.
// One lane = one grid cell like every other kernel of this library.
#pragma once
#include "xmhw_lane.h"

namespace xmhw {

// plan of the two-stack sweep (host side: xmhw_b200/plan2.py; C mirror: xmhw_clim_plan2)
struct ClimPlan2 {
  int32_t nsteps;        // sweep steps (regular doys, in doy order)
  int32_t kp;            // capacity of the top-K arrays the plan needs (max rank + 1 over all n)
  int32_t max_size;      // rows of the largest atom
  int32_t slot_rows;     // shared-memory rows per unit slot: 1 (len | guard) + max(kp, unit rows)
  int32_t nslots;        // slots (= units alive at once)
  int32_t n_init;        // atoms pushed before the first step
  int32_t pool_rows;     // nslots * slot_rows
  int32_t reserved_;
  const int32_t* rows;       // time indices of all atoms, in push order
  const int32_t* atoms;      // [natoms][4] atom records in push order (ATOM_*)
  const int32_t* step_rec;   // [nsteps][32] step records (REC_*)
  const int32_t* flip;       // flip entries, 2 words each (FLIP_*)
  double q;                  // quantile in [0,1]; numpy 'linear': v = (n-1) q
};

// atom record: 4 words
enum { ATOM_ROWS_OFF = 0,   // offset of its time indices in plan.rows
       ATOM_SIZE = 1,       // rows | JOB_F_* flags << 8
       ATOM_DEST = 2,       // first pool row of its stash | slot base row << 16
       ATOM_SLOT = 3,       // slot index (scratch rows 2 slot, 2 slot + 1 hold the unit's f64 sum)
       ATOM_WORDS = 4 };

// step record: 32 words, loaded with one coalesced warp load one step ahead
enum { REC_COUNTS = 0,      // n_pop | n_push << 4 | flip_after_push << 8 | n_flip << 16
       REC_FLIP_OFF = 1,    // first flip entry (index into plan.flip / 2)
       REC_FRONT = 2,       // slot base row of the front array used by this step's query
       REC_OUT = 3,         // output row (doy - 1)
       REC_POP = 4,         // 4 words: slot base row | slot index << 16
       REC_PUSH = 8,        // 3 x ATOM_WORDS
       REC_NEXT = 20,       // 3 x 2 words: (rows offset, size) of the atom to prefetch after push j (size 0: none)
       REC_ALIVE = 26,      // bit mask of the slots alive at the query (sum rebuild for non-finite samples)
       REC_WORDS = 32, REC_MAX_POP = 4, REC_MAX_PUSH = 3 };

// flip entry: 2 words
enum { FLIP_SRC = 0,        // first pool row of the stashed atom | rows << 16 | JOB_F_* flags << 24
       FLIP_SLOT = 1 };     // slot base row the accumulator is stored to (JOB_F_STORE / JOB_F_STOREP)

// job flags (pushes and flip entries)
enum { JOB_F_COPY = 1,      // the accumulator is empty: accumulator := this list (first push after a flip / first of a chain)
       JOB_F_FIRST = 2,     // push: first atom of its unit (initialises the slot's len row and f64 sum)
       JOB_F_STORE = 4,     // flip: store the accumulator over the unit's slot after merging this atom
       JOB_F_STOREP = 8,    // flip, no atom: store the accumulator as it is (the oldest unit's array = everything pushed)
       JOB_F_CLEAR = 16 };  // flip, no atom: accumulator := empty

#define XMHW_GUARD 0xffffff00u          // len row = guard | len: above every real key (key(+inf) = 0xff800000)

template <int N> XMHW_HD void bitonic_valley_desc(uint32_t (&k)[N]);
#define XMHW_BITONIC_IMPL(N)                                              \
  template <> XMHW_HD void bitonic_valley_desc<N>(uint32_t (&k)[N]) {     \
    XMHW_BITONIC_##N                                                      \
    const int perm[N] = XMHW_BITONIC_PERM_##N;                            \
    uint32_t t[N];                                                        \
    _Pragma("unroll") for (int i = 0; i < N; ++i) t[i] = k[perm[i]];      \
    _Pragma("unroll") for (int i = 0; i < N; ++i) k[i] = t[i];            \
  }
XMHW_BITONIC_IMPL(8)
XMHW_BITONIC_IMPL(16)
XMHW_BITONIC_IMPL(24)
XMHW_BITONIC_IMPL(36)
XMHW_BITONIC_IMPL(48)
#undef XMHW_BITONIC_IMPL

template <> XMHW_HD void sort_desc<30>(uint32_t* k) { XMHW_SORTNET_30 }

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

template <class Env, int KP, int MAXN>
struct TopkSweeper {
  typedef typename Env::Vec Vec;
  const Env& env;
  const ClimPlan2& p;
  uint32_t* pool;        // this warp's shared-memory rows (32 words each; word = lane)
  uint32_t* scratch;     // this warp's global rows: f64 sum of every unit (2 rows per slot)
  const int lane;
  const float* col;
  const int64_t ngrid;
  const bool ok;
  uint32_t A[KP];        // back stack summary: top-KP keys pushed since the last flip
  float pv[MAXN];        // prefetched rows of the next atom
  int n;                 // valid samples in the window
  int nzero;             // steps without any sample (feeds the per-cell compaction of the smoothing)
  double wsum;           // f64 sum of the window (+ pushed unit sums, - popped)
  Vec rec_next;

  XMHW_HD TopkSweeper(const Env& e, const ClimPlan2& pl, uint32_t* po, uint32_t* sc, int ln, const float* c,
                      int64_t ng, bool k)
      : env(e), p(pl), pool(po), scratch(sc), lane(ln), col(c), ngrid(ng), ok(k), n(0), nzero(0), wsum(0.0) {}

  XMHW_HD uint32_t& at(int row) { return pool[row * 32 + lane]; }

  XMHW_HD void prefetch_rows(const Vec& rv, int size) {
    const uint32_t ng32 = (uint32_t)ngrid;
    if (MAXN == 32 || size <= 32) {
#pragma unroll
      for (int i = 0; i < 32; ++i) pv[i] = XMHW_LDG(col + (uint64_t)(uint32_t)env.vget(rv, i) * ng32);
    } else {
#pragma unroll
      for (int i = 0; i < MAXN; ++i) pv[i] = XMHW_LDG(col + (uint64_t)(uint32_t)env.vget(rv, i) * ng32);
    }
  }

  // One list of at most N keys goes into the accumulator.  PUSH: the prefetched atom -- keys, f64
  // sum, register sort, stash in its unit's slot.  Otherwise (flip): a stashed atom read back from
  // its slot; the accumulator is then stored over the slot when the unit is complete.  Both kinds
  // share ONE merge site per size class (the merges are the bulk of the code).
  template <int N>
  XMHW_HD void job(bool push, int size, int flags, int row, int slot_base, int slot, bool alive) {
    uint32_t k[N];
    bool acc = alive;
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
      uint32_t* const srow = pool + row * 32 + lane;
#pragma unroll
      for (int i = 0; i < N; ++i)
        if (i < size) srow[i * 32] = k[i];
      uint32_t* const sc = scratch + (size_t)slot * 64 + lane;
      if (flags & JOB_F_FIRST) {
        at(slot_base) = XMHW_GUARD | (uint32_t)len;
        sc[0] = f64_lo(sum); sc[32] = f64_hi(sum);
      } else {
        at(slot_base) = at(slot_base) + (uint32_t)len;
        const double s2 = f64_from(sc[0], sc[32]) + sum;
        sc[0] = f64_lo(s2); sc[32] = f64_hi(s2);
      }
      n += len;
      wsum = wsum + sum;
    } else {
      const uint32_t* const srow = pool + row * 32 + lane;
#pragma unroll
      for (int i = 0; i < N; ++i) k[i] = i < size ? srow[i * 32] : 0u;
    }
    if (flags & JOB_F_COPY) {
#pragma unroll
      for (int i = 0; i < KP; ++i) A[i] = (acc && i < N) ? k[i] : 0u;
    } else if (acc) {
      merge_topk<KP, N>(A, k);
    }
  }

  XMHW_HD void store_acc(int slot_base, bool alive) {
    uint32_t* const srow = pool + (slot_base + 1) * 32 + lane;
#pragma unroll
    for (int i = 0; i < KP; ++i) srow[i * 32] = alive ? A[i] : 0u;
  }

  // Pops, then the step's jobs (a flip's entries before or after the pushes), then the query.
  // s = -1 is the initial fill: the atoms of the first window, read from the plan's atom array.
  XMHW_HD void step(int s, double& thresh, double& seas, int& out_row) {
    const bool fill = s < 0;
    const Vec rec = rec_next;
    if (s + 1 < p.nsteps) rec_next = env.vload(p.step_rec + (size_t)(s + 1) * REC_WORDS, REC_WORDS, lane);
    const uint32_t w0 = fill ? 0u : (uint32_t)env.vget(rec, REC_COUNTS);
    const int n_pop = (int)(w0 & 15u), n_push = fill ? p.n_init : (int)((w0 >> 4) & 15u), n_flip = (int)(w0 >> 16);
    const bool flip_late = ((w0 >> 8) & 1u) != 0u;
    out_row = fill ? 0 : env.vget(rec, REC_OUT);
    // pops: only the unit's sample count and sum leave the window (its keys live in no summary
    // that is still used: the front arrays are suffixes, the accumulator is younger)
    double psum = 0.0;
    for (int j = 0; j < n_pop; ++j) {
      const uint32_t pw = (uint32_t)env.vget(rec, REC_POP + j);
      n -= (int)(at((int)(pw & 0xffffu)) & 0xffu);
      const uint32_t* sc = scratch + (size_t)(pw >> 16) * 64 + lane;
      psum = psum + f64_from(sc[0], sc[32]);
    }
    // row indices of the atom that is prefetched after this step's first push, and the flip
    // program: requested now so that they are here when needed (no dependent-load wait)
    int early_size = 0;
    Vec early_rows = rec;
    if (fill) {
      const int off0 = XMHW_LDG(p.atoms + ATOM_ROWS_OFF), size0 = XMHW_LDG(p.atoms + ATOM_SIZE) & 0xff;
      prefetch_rows(env.vload(p.rows + off0, size0, lane), size0);
    } else if (n_push > 0) {
      early_size = env.vget(rec, REC_NEXT + 1);
      if (early_size > 0) early_rows = env.vload(p.rows + env.vget(rec, REC_NEXT), early_size, lane);
    }
    Vec fv = rec;
    if (n_flip > 0) fv = env.vload(p.flip + 2 * env.vget(rec, REC_FLIP_OFF), 2 * n_flip, lane);
    bool alive = true;
    const int n_jobs = n_push + n_flip;
#pragma unroll 1
    for (int jb = 0; jb < n_jobs; ++jb) {
      const bool push = flip_late ? jb < n_push : jb >= n_flip;
      int size, flags, row, slot_base, slot = 0;
      if (push) {
        const int j = flip_late ? jb : jb - n_flip;
        int sz, dst;
        if (fill) {
          const int32_t* r = p.atoms + (size_t)j * ATOM_WORDS;
          sz = XMHW_LDG(r + ATOM_SIZE); dst = XMHW_LDG(r + ATOM_DEST); slot = XMHW_LDG(r + ATOM_SLOT);
        } else {
          sz = env.vget(rec, REC_PUSH + ATOM_WORDS * j + ATOM_SIZE);
          dst = env.vget(rec, REC_PUSH + ATOM_WORDS * j + ATOM_DEST);
          slot = env.vget(rec, REC_PUSH + ATOM_WORDS * j + ATOM_SLOT);
        }
        size = sz & 0xff; flags = sz >> 8; row = dst & 0xffff; slot_base = dst >> 16;
      } else {
        const int e = flip_late ? jb - n_push : jb;
        if (e == 0) alive = env.any(n > 0);         // the window as it is when the flip starts
        const uint32_t f0 = (uint32_t)env.vget(fv, 2 * e + FLIP_SRC);
        slot_base = env.vget(fv, 2 * e + FLIP_SLOT);
        row = (int)(f0 & 0xffffu); size = (int)((f0 >> 16) & 0xffu); flags = (int)(f0 >> 24);
        if (flags & JOB_F_CLEAR) {
#pragma unroll
          for (int i = 0; i < KP; ++i) A[i] = 0u;
          continue;
        }
        if (flags & JOB_F_STOREP) { store_acc(slot_base, alive); continue; }
      }
      if (size <= 8) job<8>(push, size, flags, row, slot_base, slot, alive);
      else if (MAXN == 32) {
        if (size <= 30) job<30>(push, size, flags, row, slot_base, slot, alive);
        else job<32>(push, size, flags, row, slot_base, slot, alive);
      }
      else if (size <= 32) job<(MAXN > 32 ? 32 : 8)>(push, size, flags, row, slot_base, slot, alive);
      else if (size <= 40) job<(MAXN > 32 ? 40 : 8)>(push, size, flags, row, slot_base, slot, alive);
      else job<(MAXN > 32 ? 48 : 8)>(push, size, flags, row, slot_base, slot, alive);
      if (push) {
        // prefetch the atom pushed next (push order = the plan's atom array)
        Vec rv = early_rows;
        int rsize = early_size;
        if (fill) {
          const int32_t* nx = p.atoms + (size_t)(jb + 1) * ATOM_WORDS;        // a zero record ends the array
          rsize = XMHW_LDG(nx + ATOM_SIZE) & 0xff;
          if (rsize > 0) rv = env.vload(p.rows + XMHW_LDG(nx + ATOM_ROWS_OFF), rsize, lane);
        } else if (jb != (flip_late ? 0 : n_flip)) {
          const int j = flip_late ? jb : jb - n_flip;
          rsize = env.vget(rec, REC_NEXT + 2 * j + 1);
          if (rsize > 0) rv = env.vload(p.rows + env.vget(rec, REC_NEXT + 2 * j), rsize, lane);
        }
        if (rsize > 0) prefetch_rows(rv, rsize);
      } else if (flags & JOB_F_STORE) {
        store_acc(slot_base, alive);
      }
    }
    if (fill) { thresh = qnan(); seas = qnan(); return; }
    wsum = wsum - psum;
    const bool live = n > 0;
    nzero += live ? 0 : 1;
    if (!env.any(live)) { thresh = qnan(); seas = qnan(); return; }
    // a non-finite running sum (inf samples) is rebuilt from the sums of the units alive
    if (env.any(!(wsum - wsum == 0.0))) {
      uint32_t alive_slots = (uint32_t)env.vget(rec, REC_ALIVE);
      double fresh = 0.0;
      while (alive_slots) {
        const int sl = ctz32(alive_slots);
        alive_slots &= alive_slots - 1u;
        const uint32_t* sc = scratch + (size_t)sl * 64 + lane;
        fresh = fresh + f64_from(sc[0], sc[32]);
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
    // len | guard row sits at S[-1]); rows past the guard are clamped onto it, their terms are
    // dominated.  s_i = S[target-1-i] serves R(target) with A[i-1] and R(target-1) with A[i-2].
    const int kk = target < KP ? target : KP;
    const uint32_t* const srow = pool + env.vget(rec, REC_FRONT) * 32 + lane;
    uint32_t r1 = 0u, r2 = 0u;
    uint32_t am1 = 0xffffffffu, am2 = 0u;         // A[i-1], A[i-2]
#pragma unroll
    for (int i = 0; i <= KP; ++i) {
      int row = kk - i;
      row = row > 0 ? row : 0;
      const uint32_t sv = srow[row * 32];
      r1 = umax32(r1, umin32(am1, sv));
      if (i >= 1) r2 = umax32(r2, umin32(am2, sv));
      am2 = am1;
      am1 = i < KP ? A[i] : 0u;
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

  XMHW_HD void init() {
#pragma unroll
    for (int i = 0; i < KP; ++i) A[i] = 0u;
    rec_next = env.vload(p.step_rec, REC_WORDS, lane);
  }
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
