// TEST INFRASTRUCTURE ONLY -- single-lane host emulation of xmhw_b200/csrc/xmhw_lane.h.
//
// The build container has no GPU.  Every xmhw_b200 kernel maps one grid cell to one
// lane and lanes never exchange data, so the per-lane device functions can be
// compiled for the host (this file, g++) and driven one cell at a time to debug
// the plan/selection/run-length/statistics logic before spending GPU time.
// The xmhw_b200 package never loads this library; product calls go to the CUDA
// library only (xmhw_b200/_cabi.py raises if it is missing).
#include <vector>
#include "../../xmhw_b200/csrc/xmhw_lane.h"
#include "../../xmhw_b200/csrc/xmhw_topk.h"

using namespace xmhw;

struct HostEnv {
  struct Vec { const int32_t* p; int n; };
  bool any(bool p) const { return p; }
  int max_all(int v) const { return v; }
  Vec vload(const int32_t* src, int count, int) const { return Vec{src, count}; }
  int32_t vget(const Vec& v, int i) const { return i < v.n ? v.p[i] : 0; }
  void vstage(uint32_t* ub, const Vec& v, int m, int m4, int) const {
    for (int i = 0; i < m4; ++i) ub[i] = i < m ? (uint32_t)v.p[i] : 0u;
  }
};

template <int MAXN>
static void sweep_cells(const float* ts, int64_t ngrid, const ClimPlan* plan, double* thr, double* seas) {
  std::vector<uint32_t> pool((size_t)(plan->pool_rows + POOL_STAGE_ROWS) * 32);
  std::vector<uint32_t> scratch((size_t)(plan->scratch_rows + 1) * 32);
  HostEnv env;
  for (int64_t cell = 0; cell < ngrid; ++cell) {
    const int lane = (int)(cell & 31);
    Sweeper<HostEnv, MAXN> sw(env, *plan, pool.data(), scratch.data(), lane, ts + cell, ngrid, true);
    sw.init();
    for (int s = 0; s < plan->nsteps; ++s) {
      double a, b;
      sw.step(s, a, b);
      thr[(int64_t)s * ngrid + cell] = a;
      seas[(int64_t)s * ngrid + cell] = b;
    }
  }
}

static int g_phased = 0;      // 1: TopkSweeperP::step_phased (the step of the tensor-memory kernel)

template <int KP, int MAXN>
static void sweep2_cells(const float* ts, int64_t ngrid, const ClimPlan2* plan, double* thr, double* seas,
                         int32_t* nzero) {
  std::vector<uint32_t> pool((size_t)plan->nslots * plan->slot_rows * 32);
  HostEnv env;
  for (int64_t cell = 0; cell < ngrid; ++cell) {
    const int lane = (int)(cell & 31);
    TopkSweeper<HostEnv, KP, MAXN> sw(env, *plan, pool.data(), lane, ts + cell, ngrid, true);
    for (int s = -1; s < plan->nsteps; ++s) {
      double a, b;
      int row;
      if (g_phased) sw.step_phased(s, a, b, row); else sw.step(s, a, b, row);
      if (s < 0) continue;
      thr[(int64_t)row * ngrid + cell] = a;
      seas[(int64_t)row * ngrid + cell] = b;
    }
    nzero[cell] = sw.nzero;
  }
}

template <int KP>
static void direct_cells(const float* ts, int64_t ngrid, const int32_t* rows, int nrows, double q, double* thr,
                         double* seas, int32_t* nzero) {
  for (int64_t cell = 0; cell < ngrid; ++cell) {
    DirectSelect<KP> ds;
    for (int r0 = 0; r0 < nrows; r0 += 8) {
      float v[8];
      for (int i = 0; i < 8; ++i) v[i] = r0 + i < nrows ? ts[(int64_t)rows[r0 + i] * ngrid + cell] : 0.0f;
      ds.add8(v, nrows - r0 < 8 ? nrows - r0 : 8, true);
    }
    ds.result(q, thr[cell], seas[cell]);
    if (ds.n == 0) nzero[cell] += 1;
  }
}

extern "C" {

void emul_set_phased(int on) { g_phased = on; }

int emul_clim_sweep2(const float* ts, int64_t T, int64_t ngrid, const ClimPlan2* plan, double* thr, double* seas,
                     int32_t* nzero) {
  (void)T;
  const bool big = plan->max_size > 32;
  const bool n30 = plan->max_size <= 30;
  switch (plan->kp) {
    case 8: if (big) sweep2_cells<8, 48>(ts, ngrid, plan, thr, seas, nzero); else if (n30) sweep2_cells<8, 30>(ts, ngrid, plan, thr, seas, nzero); else sweep2_cells<8, 32>(ts, ngrid, plan, thr, seas, nzero); break;
    case 16: if (big) sweep2_cells<16, 48>(ts, ngrid, plan, thr, seas, nzero); else if (n30) sweep2_cells<16, 30>(ts, ngrid, plan, thr, seas, nzero); else sweep2_cells<16, 32>(ts, ngrid, plan, thr, seas, nzero); break;
    case 24: if (big) sweep2_cells<24, 48>(ts, ngrid, plan, thr, seas, nzero); else if (n30) sweep2_cells<24, 30>(ts, ngrid, plan, thr, seas, nzero); else sweep2_cells<24, 32>(ts, ngrid, plan, thr, seas, nzero); break;
    case 36: if (big) sweep2_cells<36, 48>(ts, ngrid, plan, thr, seas, nzero); else if (n30) sweep2_cells<36, 30>(ts, ngrid, plan, thr, seas, nzero); else sweep2_cells<36, 32>(ts, ngrid, plan, thr, seas, nzero); break;
    case 48: if (big) sweep2_cells<48, 48>(ts, ngrid, plan, thr, seas, nzero); else if (n30) sweep2_cells<48, 30>(ts, ngrid, plan, thr, seas, nzero); else sweep2_cells<48, 32>(ts, ngrid, plan, thr, seas, nzero); break;
    default: return -1;
  }
  return 0;
}

// one exceptional doy: thr / seas point at that doy's output row
int emul_clim_direct(const float* ts, int64_t ngrid, const int32_t* rows, int nrows, int kp, double q, double* thr,
                     double* seas, int32_t* nzero) {
  switch (kp) {
    case 8: direct_cells<8>(ts, ngrid, rows, nrows, q, thr, seas, nzero); break;
    case 16: direct_cells<16>(ts, ngrid, rows, nrows, q, thr, seas, nzero); break;
    case 24: direct_cells<24>(ts, ngrid, rows, nrows, q, thr, seas, nzero); break;
    case 36: direct_cells<36>(ts, ngrid, rows, nrows, q, thr, seas, nzero); break;
    case 48: direct_cells<48>(ts, ngrid, rows, nrows, q, thr, seas, nzero); break;
    default: return -1;
  }
  return 0;
}

int emul_clim_sweep(const float* ts, int64_t T, int64_t ngrid, const ClimPlan* plan, double* thr, double* seas) {
  (void)T;
  if (plan->max_size <= 32) sweep_cells<32>(ts, ngrid, plan, thr, seas);
  else sweep_cells<48>(ts, ngrid, plan, thr, seas);
  return 0;
}

struct VecEmit { int32_t* s; int32_t* e; int n; void operator()(int a, int b) { s[n] = a; e[n] = b; ++n; } };

// b: exceedance booleans [T]; returns number of events, fills starts/ends
int emul_find_events(const uint8_t* b, int T, int min_dur, int join, int max_gap, int32_t* starts, int32_t* ends) {
  RunFinder rf(min_dur, join, max_gap);
  VecEmit em{starts, ends, 0};
  for (int t0 = 0; t0 < T; t0 += 32) {
    uint32_t bits = 0;
    for (int i = 0; i < 32 && t0 + i < T; ++i) bits |= (uint32_t)(b[t0 + i] != 0) << i;
    rf.feed(bits, t0, em);
  }
  rf.finish(T, em);
  return em.n;
}

// same with eager emission after every word (the fused detect kernel's use of RunFinder)
int emul_find_events_eager(const uint8_t* b, int T, int min_dur, int join, int max_gap, int32_t* starts, int32_t* ends,
                           int32_t* emitted_at) {
  RunFinder rf(min_dur, join, max_gap);
  VecEmit em{starts, ends, 0};
  int done = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    uint32_t bits = 0;
    for (int i = 0; i < 32 && t0 + i < T; ++i) bits |= (uint32_t)(b[t0 + i] != 0) << i;
    rf.feed(bits, t0, em);
    rf.flush_pending(t0 + 32 < T ? t0 + 32 : T, em);
    for (; done < em.n; ++done) emitted_at[done] = t0 + 32 < T ? t0 + 32 : T;
  }
  rf.finish(T, em);
  for (; done < em.n; ++done) emitted_at[done] = T;
  return em.n;
}

void emul_event_stats(const float* col, const double* th, const double* se, const int32_t* doy, int64_t ngrid,
                      int T, int s, int e, int32_t* oi, double* of) {
  event_stats(col, th, se, doy, ngrid, T, s, e, oi, of, 1);
}

// placement of a pivot in a descending sorted key array (halving search through selects)
void emul_partition_sorted(const uint32_t* k, int n, uint32_t pivot, int32_t* out) {
  int ptr = -1; uint32_t cinc = 0, cexc = 0;
  if (n == 8) { uint32_t a[8]; for (int i = 0; i < 8; ++i) a[i] = k[i]; partition_sorted<8>(a, pivot, ptr, cinc, cexc); }
  else if (n == 32) { uint32_t a[32]; for (int i = 0; i < 32; ++i) a[i] = k[i]; partition_sorted<32>(a, pivot, ptr, cinc, cexc); }
  else if (n == 40) { uint32_t a[40]; for (int i = 0; i < 40; ++i) a[i] = k[i]; partition_sorted<40>(a, pivot, ptr, cinc, cexc); }
  out[0] = ptr; out[1] = (int32_t)cinc; out[2] = (int32_t)cexc;
}

// the 4-entry front: insert a stream of (key, tag), then replace the head `nrep` times
void emul_front(const uint32_t* keys, const int32_t* tags, int n, const uint32_t* rkeys, const int32_t* rtags, int nrep,
                uint32_t* f_out, int32_t* g_out) {
  uint32_t f[4]; int g[4];
  for (int i = 0; i < 4; ++i) { f[i] = 0xffffffffu; g[i] = 0; }
  for (int i = 0; i < n; ++i) front_insert<4>(keys[i], tags[i], f, g);
  for (int i = 0; i < nrep; ++i) front_replace_head<4>(rkeys[i], rtags[i], f, g);
  for (int i = 0; i < 4; ++i) { f_out[i] = f[i]; g_out[i] = g[i]; }
}

uint32_t emul_f32_key(float f) { return f32_key(f); }
float emul_key_f32(uint32_t k) { return key_f32(k); }
}
