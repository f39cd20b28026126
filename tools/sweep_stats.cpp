// Development aid: per-lane walk statistics of the climatology sweep (host build of xmhw_lane.h)
// aggregated the way a warp executes them (max over 32 lanes).  g++ -O2 -o /tmp/ss tools/sweep_stats.cpp
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
static int g_step_d, g_nscan, g_pops_cur, g_scr;
static long g_hist[64], g_off[32];
static int g_ptr0[4096];
#define XMHW_STAT_PTR0(base, ptr) { g_ptr0[base] = (ptr); }
#define XMHW_STAT_OFF(base, r) { int o = (r) - g_ptr0[base] + 16; if (o >= 0 && o < 32) ++g_off[o]; }
static std::vector<int> g_pops;   // pops per scan of the current step
#define XMHW_STAT_STEP(d0) { g_step_d = (d0); g_nscan = 0; g_pops.clear(); g_scr = 0; }
#define XMHW_STAT_SCAN() { ++g_nscan; g_pops.push_back(0); }
#define XMHW_STAT_POP(s) { ++g_pops.back(); g_scr += (s) ? 1 : 0; }
#define XMHW_STAT_RANK(r) { if ((r) >= 0 && (r) < 64) ++g_hist[r]; }
#include "../xmhw_b200/csrc/xmhw_lane.h"
using namespace xmhw;
struct HostEnv {
  struct Vec { const int32_t* p; int n; };
  bool any(bool p) const { return p; }
  Vec vload(const int32_t* src, int count, int) const { return Vec{src, count}; }
  int32_t vget(const Vec& v, int i) const { return i < v.n ? v.p[i] : 0; }
  void vstage(uint32_t* ub, const Vec& v, int m, int m4, int) const { for (int i = 0; i < m4; ++i) ub[i] = i < m ? (uint32_t)v.p[i] : 0u; }
};
extern "C" int sweep_stats(const float* ts, int64_t ngrid, const ClimPlan* plan, double* out) {
  // out: [0] lane avg |d0|, [1] lane avg pops, [2] lane avg scans, [3] warp avg scans, [4] warp avg pop iterations,
  //      [5] lane avg scratch pops, [6] warp-steps with any scratch pop fraction, [7] warp avg max|d0|
  std::vector<uint32_t> pool((size_t)(plan->pool_rows + POOL_STAGE_ROWS) * 32), scratch((size_t)(plan->scratch_rows + 1) * 32);
  HostEnv env;
  const int ns = plan->nsteps;
  double sum_d = 0, sum_p = 0, sum_s = 0, sum_scr = 0, w_scans = 0, w_pops = 0, w_scr = 0, w_d = 0; long nl = 0, nw = 0;
  for (int64_t c0 = 0; c0 < ngrid; c0 += 32) {
    std::vector<std::vector<std::vector<int>>> pp(32, std::vector<std::vector<int>>(ns));
    std::vector<std::vector<int>> scr(32, std::vector<int>(ns)), dd(32, std::vector<int>(ns));
    for (int l = 0; l < 32 && c0 + l < ngrid; ++l) {
      Sweeper<HostEnv, 32> sw(env, *plan, pool.data(), scratch.data(), l, ts + c0 + l, ngrid, true);
      sw.init();
      for (int s = 0; s < ns; ++s) {
        double a, b; g_pops.clear(); g_nscan = 0; g_scr = 0; g_step_d = 0;
        sw.step(s, a, b);
        pp[l][s] = g_pops; scr[l][s] = g_scr; dd[l][s] = abs(g_step_d);
        int tp = 0; for (int x : g_pops) tp += x;
        sum_d += abs(g_step_d); sum_p += tp; sum_s += g_nscan; sum_scr += g_scr; ++nl;
      }
    }
    for (int s = 0; s < ns; ++s) {
      size_t ms = 0; int anyscr = 0, md = 0;
      for (int l = 0; l < 32; ++l) { ms = std::max(ms, pp[l][s].size()); anyscr |= scr[l][s] > 0; md = std::max(md, dd[l][s]); }
      int wp = 0;
      for (size_t i = 0; i < ms; ++i) { int m = 0; for (int l = 0; l < 32; ++l) if (i < pp[l][s].size()) m = std::max(m, pp[l][s][i]); wp += m; }
      w_scans += ms; w_pops += wp; w_scr += anyscr; w_d += md; ++nw;
    }
  }
  out[0] = sum_d / nl; out[1] = sum_p / nl; out[2] = sum_s / nl; out[3] = w_scans / nw; out[4] = w_pops / nw;
  for (int i = 0; i < 32; ++i) out[8 + i] = (double)g_hist[i];
  for (int i = 0; i < 32; ++i) out[40 + i] = (double)g_off[i];
  out[5] = sum_scr / nl; out[6] = w_scr / nw; out[7] = w_d / nw;
  return 0;
}
