// xmhw_b200 CUDA kernels (sm_100a) + C ABI.  See include/xmhw_b200.h for the
// boundary contract and xmhw_lane.h for the per-lane algorithms.
//
// All kernels map one grid cell to one lane (32 adjacent cells per warp): the
// reference layouts (time, cell) / (doy, cell) are then read and written as
// contiguous 128 B / 256 B row segments with no transpose pass.  None of this
// is a contraction, so no tensor-core instruction is issued; tensor MEMORY is
// used as lane-private storage by the dominant kernel (clim_sweep2_tm_kernel).
// The kernels are HBM-, LSU- and ALU-bound.  Compiled with --fmad=false so that the float64 expressions round
// exactly like numpy/pandas (no FMA contraction).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/xmhw_b200.h"
#include "xmhw_lane.h"
#include "xmhw_topk.h"

namespace {

using namespace xmhw;

struct WarpEnv {
  // small int vector spread over the lanes: entry i lives in lane (i & 31), word (i >> 5)
  struct Vec { int32_t a, b; };
  __device__ __forceinline__ bool any(bool p) const { return __any_sync(0xffffffffu, p); }
  __device__ __forceinline__ int max_all(int v) const { return (int)__reduce_max_sync(0xffffffffu, (unsigned)v); }
  __device__ __forceinline__ Vec vload(const int32_t* src, int count, int lane) const {
    Vec v;
    v.a = lane < count ? __ldg(src + lane) : 0;
    v.b = lane + 32 < count ? __ldg(src + lane + 32) : 0;
    return v;
  }
  __device__ __forceinline__ int32_t vget(const Vec& v, int i) const {      // i is warp-uniform
    return __shfl_sync(0xffffffffu, i < 32 ? v.a : v.b, i & 31);
  }
  // write the first m entries (padded with the null list 0 up to m4) to shared memory
  __device__ __forceinline__ void vstage(uint32_t* ub, const Vec& v, int m, int m4, int lane) const {
    __syncwarp();
    if (lane < m4) ub[lane] = lane < m ? (uint32_t)v.a : 0u;
    if (lane + 32 < m4) ub[lane + 32] = lane + 32 < m ? (uint32_t)v.b : 0u;
    __syncwarp();
  }
};

static_assert(sizeof(xmhw_clim_plan) == sizeof(ClimPlan), "plan ABI mismatch");
static_assert(sizeof(xmhw_clim_plan2) == sizeof(ClimPlan2), "plan2 ABI mismatch");
static_assert((int)XMHW_EI_COUNT == (int)EI_COUNT && (int)XMHW_EF_COUNT == (int)EF_COUNT, "event ABI mismatch");

// ---------------------------------------------------------------------------
// K1  climatology sweep: one warp = 32 cells, sequential over day-of-year.
// Shared memory per warp: plan.pool_rows rows of 32 words (sorted lists of the
// current +-w window, their f64 sums and per-lane cut pointers).
// ---------------------------------------------------------------------------
// MAXN = keys per sorted list: 32 (series of <= 32 years: one list per calendar day, ~96
// registers, 20 warps/SM) or 48 (longer series; larger register sorting networks, fewer warps).
template <int MAXN, int MINB>
__global__ void __launch_bounds__(32, MINB) clim_sweep_kernel(
    ClimPlan p, const float* __restrict__ ts, int64_t ngrid, double* __restrict__ thr, double* __restrict__ seas,
    int32_t* __restrict__ nempty, uint32_t* __restrict__ scratch) {
  extern __shared__ uint32_t pool[];
  const int lane = threadIdx.x;
  const int64_t cell = (int64_t)blockIdx.x * 32 + lane;
  const bool ok = cell < ngrid;
  const float* col = ts + (ok ? cell : 0);
  WarpEnv env;
  Sweeper<WarpEnv, MAXN> sw(env, p, pool, scratch + (size_t)blockIdx.x * p.scratch_rows * 32, lane, col, ngrid, ok);
  sw.init();
  for (int s = 0; s < p.nsteps; ++s) {
    double a, b;
    sw.step(s, a, b);
    if (ok) {
      thr[(int64_t)s * ngrid + cell] = a;
      seas[(int64_t)s * ngrid + cell] = b;
    }
  }
  if (ok) nempty[cell] = sw.nzero;
}

// ---------------------------------------------------------------------------
// K1'  two-stack top-K climatology sweep (xmhw_topk.h): one warp = 32 cells, straight-line
// sorting / merging networks per doy, the unit slots of the window in shared memory.
// ---------------------------------------------------------------------------
// WPB warps per block; `sync_every` > 0 makes them advance in lockstep (a barrier every so many doys) so that
// they stream the same straight-line code through the instruction cache -- a gain while the step carried 70
// register copies per job, a loss since (launcher: default 0).
// 32-cell group of warp `w` of the grid: order[w] when the caller passed a processing order (land-looking groups
// last, so that the warps of one lockstep block have the same amount of work), w otherwise; past the grid: none
__device__ __forceinline__ int64_t sweep2_group(const int32_t* order, int64_t w, int64_t ngrid) {
  const int64_t ncg = (ngrid + 31) / 32;
  if (w >= ncg) return ncg;
  return order ? (int64_t)__ldg(order + w) : w;
}

template <int KP, int MAXN, int WPB, int MINB>
__global__ void __launch_bounds__(32 * WPB, MINB) clim_sweep2_kernel(
    const __grid_constant__ ClimPlan2 p, const float* __restrict__ ts, int64_t ngrid, double* __restrict__ thr,
    double* __restrict__ seas, int32_t* __restrict__ nempty, const int32_t* __restrict__ order, int sync_every) {
  extern __shared__ uint32_t pool[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t cell = sweep2_group(order, (int64_t)blockIdx.x * WPB + wib, ngrid) * 32 + lane;
  const bool ok = cell < ngrid;
  const float* col = ts + (ok ? cell : 0);
  WarpEnv env;
  TopkSweeper<WarpEnv, KP, MAXN> sw(env, p, pool + (size_t)wib * p.nslots * p.slot_rows * 32, lane, col, ngrid, ok,
                                    -wib * p.nslots * p.slot_rows);       // rows below this warp's pool: the earlier warps'
  for (int s = -1; s < p.nsteps; ++s) {          // s = -1: initial fill of the first window
    double a, b;
    int row;
    sw.step_phased(s, a, b, row);
    if (ok && s >= 0) {
      thr[(int64_t)row * ngrid + cell] = a;
      seas[(int64_t)row * ngrid + cell] = b;
    }
    if (WPB > 1 && sync_every > 0 && (s + 1) % sync_every == 0) __syncthreads();
  }
  if (ok) nempty[cell] = sw.nzero;
}

// ---------------------------------------------------------------------------
// K1t  the top-K sweep with half of every warp's unit slots in TENSOR MEMORY.
// The sweep is bound by the warps its slots leave room for (55 KB per warp at the default window: four
// warps fill the 227 KB of shared memory, one per scheduler).  The slot rows are only ever addressed
// warp-uniformly (row = same for all lanes, word = lane), which is exactly the 32-lane x 32-bit shape of
// tcgen05.ld / st: a TMEM column is a pool row.  A block = 8 warps = the whole SM: warps w and w + 4 own
// the TMEM lane quarter w % 4, 256 columns each; the first `smem_slots` slots of a warp stay in shared
// memory, the others are TMEM columns.  No tensor-core instruction is involved -- TMEM is used as 256 KB
// of extra lane-private storage, doubling the resident warps of this kernel.
// ---------------------------------------------------------------------------
#include "tmem_gen.h"

template <int N> __device__ __forceinline__ void tm_ld_block(uint32_t taddr, uint32_t* k) {
  if constexpr (N >= 32) { tm_ld32(taddr, k); tm_ld_block<N - 32>(taddr + 32, k + 32); }
  else if constexpr (N >= 16) { tm_ld16(taddr, k); tm_ld_block<N - 16>(taddr + 16, k + 16); }
  else if constexpr (N >= 8) { tm_ld8(taddr, k); tm_ld_block<N - 8>(taddr + 8, k + 8); }
  else if constexpr (N >= 4) { tm_ld4(taddr, k); tm_ld_block<N - 4>(taddr + 4, k + 4); }
  else if constexpr (N >= 2) { tm_ld2(taddr, k); tm_ld_block<N - 2>(taddr + 2, k + 2); }
  else if constexpr (N >= 1) { tm_ld1(taddr, k); }
}
template <int N> __device__ __forceinline__ void tm_st_block(uint32_t taddr, const uint32_t* k) {
  if constexpr (N >= 32) { tm_st32(taddr, k); tm_st_block<N - 32>(taddr + 32, k + 32); }
  else if constexpr (N >= 16) { tm_st16(taddr, k); tm_st_block<N - 16>(taddr + 16, k + 16); }
  else if constexpr (N >= 8) { tm_st8(taddr, k); tm_st_block<N - 8>(taddr + 8, k + 8); }
  else if constexpr (N >= 4) { tm_st4(taddr, k); tm_st_block<N - 4>(taddr + 4, k + 4); }
  else if constexpr (N >= 2) { tm_st2(taddr, k); tm_st_block<N - 2>(taddr + 2, k + 2); }
  else if constexpr (N >= 1) { tm_st1(taddr, k); }
}
template <int N> __device__ __forceinline__ void tm_gather(const uint32_t* a, uint32_t* k);
template <> __device__ __forceinline__ void tm_gather<9>(const uint32_t* a, uint32_t* k) { tm_gather9(a, k); }
template <> __device__ __forceinline__ void tm_gather<17>(const uint32_t* a, uint32_t* k) { tm_gather17(a, k); }
template <> __device__ __forceinline__ void tm_gather<25>(const uint32_t* a, uint32_t* k) { tm_gather25(a, k); }
template <> __device__ __forceinline__ void tm_gather<37>(const uint32_t* a, uint32_t* k) { tm_gather37(a, k); }
template <> __device__ __forceinline__ void tm_gather<49>(const uint32_t* a, uint32_t* k) { tm_gather49(a, k); }

// rows [0, split) of the warp's pool: shared memory; rows >= split: TMEM column tcol + (row - split).
// A slot never straddles the split (it is a multiple of the slot size), so a block access is one or the other.
struct SplitPool {
  uint32_t* p;           // shared-memory row 0, this lane's word
  int split;
  uint32_t tcol;         // TMEM address (lane quarter << 16 | column) of row `split`
  int min_srow, min_tcol;     // lowest readable shared-memory row / TMEM column relative to this warp's (<= 0)
  static constexpr bool kGather = true;
  __device__ __forceinline__ uint32_t ld(int row) const {
    if (row < split) return p[row * 32];
    uint32_t v;
    tm_ld1(tcol + (uint32_t)(row - split), &v);
    return v;
  }
  __device__ __forceinline__ void st(int row, uint32_t v) const {
    if (row < split) { p[row * 32] = v; return; }
    tm_st1(tcol + (uint32_t)(row - split), &v);
  }
  template <int N> __device__ __forceinline__ void ld_block(int row0, uint32_t (&k)[N]) const {
    if (row0 < split) {
      const uint32_t* const r = p + row0 * 32;
#pragma unroll
      for (int i = 0; i < N; ++i) k[i] = r[i * 32];
    } else {
      tm_ld_block<N>(tcol + (uint32_t)(row0 - split), k);
    }
  }
  template <int N> __device__ __forceinline__ void ld_block_n(int row0, int size, uint32_t (&k)[N]) const {
    if (row0 < split) {
      const uint32_t* const r = p + row0 * 32;
#pragma unroll
      for (int i = 0; i < N; ++i) k[i] = i < size ? r[i * 32] : 0u;
    } else {                                   // rare (atoms that share a slot): column by column
#pragma unroll
      for (int i = 0; i < N; ++i) {
        k[i] = 0u;
        if (i < size) tm_ld1(tcol + (uint32_t)(row0 - split + i), &k[i]);
      }
    }
  }
  template <int N> __device__ __forceinline__ void st_block(int row0, const uint32_t (&k)[N]) const {
    if (row0 < split) {
      uint32_t* const r = p + row0 * 32;
#pragma unroll
      for (int i = 0; i < N; ++i) r[i * 32] = k[i];
    } else {
      tm_st_block<N>(tcol + (uint32_t)(row0 - split), k);
    }
  }
  template <int N> __device__ __forceinline__ void st_block_n(int row0, int size, const uint32_t (&k)[N]) const {
    if (row0 < split) {
      uint32_t* const r = p + row0 * 32;
#pragma unroll
      for (int i = 0; i < N; ++i)
        if (i < size) r[i * 32] = k[i];
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i)
        if (i < size) tm_st1(tcol + (uint32_t)(row0 - split + i), &k[i]);
    }
  }
  // see PlainPool::ld_front: rows past the guard are read unclamped when that memory exists
  template <int N> __device__ __forceinline__ void ld_front(int base, int kk, uint32_t (&sv)[N]) const {
    if (base < split) {
      if (base + kk - (N - 1) >= min_srow) {
        const uint32_t* const r = p + (base + kk) * 32;
#pragma unroll
        for (int i = 0; i < N; ++i) sv[i] = *(r - i * 32);
      } else {
#pragma unroll
        for (int i = 0; i < N; ++i) sv[i] = p[(base + (kk - i > 0 ? kk - i : 0)) * 32];
      }
    } else {
      uint32_t a[N];
      const int c0 = base - split + kk;
      if (c0 - (N - 1) >= min_tcol) {
#pragma unroll
        for (int i = 0; i < N; ++i) a[i] = tcol + (uint32_t)(c0 - i);
      } else {
#pragma unroll
        for (int i = 0; i < N; ++i) a[i] = tcol + (uint32_t)(base - split + (kk - i > 0 ? kk - i : 0));
      }
      tm_gather<N>(a, sv);
    }
  }
};

enum { TM_WARPS = 8, TM_COLS_PER_WARP = 256 };

template <int KP, int MAXN>
__global__ void __launch_bounds__(32 * TM_WARPS, 1) clim_sweep2_tm_kernel(
    const __grid_constant__ ClimPlan2 p, const float* __restrict__ ts, int64_t ngrid, double* __restrict__ thr,
    double* __restrict__ seas, int32_t* __restrict__ nempty, const int32_t* __restrict__ order,
    unsigned* __restrict__ ticket, int smem_slots) {
  extern __shared__ uint32_t pool[];
  __shared__ uint32_t tm_base;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  if (wib == 0) {               // the block owns all 512 columns of its SM (one block per SM: launch bounds + shared memory)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;"
                 :: "r"((uint32_t)__cvta_generic_to_shared(&tm_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tm_base;
  const int split = smem_slots * p.slot_rows;
  SplitPool pl;
  pl.p = pool + (size_t)wib * split * 32 + lane;
  pl.split = split;
  pl.tcol = tbase + ((uint32_t)((wib & 3) * 32) << 16) + (uint32_t)((wib >> 2) * TM_COLS_PER_WARP);
  pl.min_srow = -wib * split;
  pl.min_tcol = -(wib >> 2) * TM_COLS_PER_WARP;
  WarpEnv env;
  // Independent warps: a warp sweeps the 32-cell group its position names; in the persistent launch mode (one
  // block per SM, see the launcher) it goes on with groups drawn from a ticket counter in the caller's
  // processing order.
  const int64_t ncg = (ngrid + 31) / 32;
  const int64_t slots = (int64_t)gridDim.x * TM_WARPS;
  int64_t w = (int64_t)blockIdx.x * TM_WARPS + wib;
  while (w < ncg) {
    const int64_t cell = (order ? (int64_t)__ldg(order + w) : w) * 32 + lane;
    const bool ok = cell < ngrid;
    const float* col = ts + (ok ? cell : 0);
    {
      TopkSweeperP<WarpEnv, SplitPool, KP, MAXN> sw(env, p, pl, col, ngrid, ok);
      for (int s = -1; s < p.nsteps; ++s) {          // s = -1: initial fill of the first window
        double a, b;
        int row;
        sw.step_phased(s, a, b, row);
        if (ok && s >= 0) {
          thr[(int64_t)row * ngrid + cell] = a;
          seas[(int64_t)row * ngrid + cell] = b;
        }
      }
      if (ok) nempty[cell] = sw.nzero;
    }
    if (ticket) {
      unsigned t = 0;
      if (lane == 0) t = atomicAdd(ticket, 1u);
      w = slots + (int64_t)__shfl_sync(0xffffffffu, t, 0);
    } else {
      w += slots;
    }
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (wib == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tbase) : "memory");
}

// doys whose window is not a range of the atom order (doy 60): direct selection, one thread = one cell
template <int KP>
__global__ void __launch_bounds__(128) clim_direct_kernel(const float* __restrict__ ts, int64_t ngrid,
                                                          const int32_t* __restrict__ rows, int nrows, double q,
                                                          double* __restrict__ thr_row, double* __restrict__ seas_row,
                                                          int32_t* __restrict__ nempty) {
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = cell < ngrid;
  const float* col = ts + (ok ? cell : 0);
  DirectSelect<KP> ds;
  for (int r0 = 0; r0 < nrows; r0 += 8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldg(col + (int64_t)__ldg(rows + min(r0 + i, nrows - 1)) * ngrid);
    ds.add8(v, nrows - r0 < 8 ? nrows - r0 : 8, ok);
  }
  if (!ok) return;
  double a, b;
  ds.result(q, a, b);
  thr_row[cell] = a;
  seas_row[cell] = b;
  if (ds.n == 0) nempty[cell] += 1;
}

// ---------------------------------------------------------------------------
// K1b  Feb-29 rule + circular running mean over doy.  One thread = one cell.
//
// Summation order (shared with the oracle, oracle/xmhw_oracle.py:runavg, so results are
// bit-equal): with the wrapped sequence e[j] = x[(j - h) mod ndoy], j = 0 .. ndoy + W - 2,
// cut into blocks of W, output d = b W + p is
//     suf_b[p]                      (p == 0: the whole block, summed right to left)
//     suf_b[p] + pre_{b+1}[p - 1]   (p >= 1)
// where suf_b[p] = e[bW+p] + (e[bW+p+1] + (... + e[bW+W-1])) and pre_b[i] = ((e[bW] + e[bW+1]) + ...) + e[bW+i].
// Every window is one suffix of a block plus one prefix of the next, so an output costs
// 3 FP64 additions instead of W - 1 (the FP64 pipe, not HBM, bounded the direct sum).
// ---------------------------------------------------------------------------
struct FinishSrc {
  const double* __restrict__ raw;     // this cell's column (stride ngrid)
  int64_t ngrid;
  int ndoy;
  bool feb;                           // substitute doy 60 (index 59)
  double v59;                         // mean of the non-NaN values at doy 59, 60, 61 (identify.py:137-151)
  __device__ __forceinline__ void init(const double* r, int64_t ng, int nd, int feb29) {
    raw = r; ngrid = ng; ndoy = nd; feb = feb29 && nd >= 61; v59 = 0.0;
    if (feb) {
      double acc = 0.0; int n = 0;
#pragma unroll
      for (int d = 58; d <= 60; ++d) {
        const double v = raw[(int64_t)d * ngrid];
        if (v == v) { acc = acc + v; ++n; }
      }
      v59 = n ? acc / (double)n : qnan();
    }
  }
  __device__ __forceinline__ double at(int idx) const {          // idx in [0, ndoy)
    const double v = raw[(int64_t)idx * ngrid];
    return (feb && idx == 59) ? v59 : v;
  }
};

// x / W, correctly rounded, for a small odd constant W: q = x * RN(1/W) refined by one
// exact-residual step (Markstein); non-finite or tiny values take the true division.
template <int W>
__device__ __forceinline__ double div_const(double x) {
  const double w = (double)W, rc = 1.0 / (double)W;
  const double q = x * rc;
  const double rem = __fma_rn(-q, w, x);
  const double q2 = __fma_rn(rem, rc, q);
  const double ax = fabs(x);
  return (ax > 1e-280 && ax < 1e300) ? q2 : x / w;
}

// generic odd width: W slots per thread in shared memory (column = thread: conflict free)
__global__ void clim_finish_kernel(const double* __restrict__ raw, double* __restrict__ out, int ndoy,
                                   int64_t ngrid, int feb29, int W, const int32_t* __restrict__ nempty) {
  extern __shared__ double slots[];
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ngrid) return;
  const int ne = nempty[cell];
  if (ne > 0 && ne < ndoy) return;          // some doys without samples: clim_finish_compact_kernel
  FinishSrc src;
  src.init(raw + cell, ngrid, ndoy, feb29);
  double* __restrict__ o = out + cell;
  if (W <= 1) {
    for (int d = 0; d < ndoy; ++d) o[(int64_t)d * ngrid] = src.at(d);
    return;
  }
  const int h = (W - 1) / 2;
  const int nt = blockDim.x, tid = threadIdx.x;
  int idx = ((-h) % ndoy + ndoy) % ndoy;                        // index of e[j], advanced with wrap
  for (int i = 0; i < W; ++i) { slots[i * nt + tid] = src.at(idx); idx = idx + 1 == ndoy ? 0 : idx + 1; }
  for (int i = W - 2; i >= 0; --i) slots[i * nt + tid] = slots[i * nt + tid] + slots[(i + 1) * nt + tid];
  const double w = (double)W;
  for (int d0 = 0; d0 < ndoy; d0 += W) {
    double pre = 0.0;
    for (int p = 0; p < W && d0 + p < ndoy; ++p) {
      double S = slots[p * nt + tid];
      if (p >= 1) {
        const double en = src.at(idx); idx = idx + 1 == ndoy ? 0 : idx + 1;
        pre = p == 1 ? en : pre + en;
        S = S + pre;
        slots[(p - 1) * nt + tid] = en;
      }
      o[(int64_t)(d0 + p) * ngrid] = S / w;
    }
    if (d0 + W < ndoy) {
      slots[(W - 1) * nt + tid] = src.at(idx); idx = idx + 1 == ndoy ? 0 : idx + 1;
      for (int i = W - 2; i >= 0; --i) slots[i * nt + tid] = slots[i * nt + tid] + slots[(i + 1) * nt + tid];
    }
  }
}

// compile-time width (the default 31): the W slots are statically indexed registers
template <int W>
__global__ void __launch_bounds__(128) clim_finish_reg_kernel(const double* __restrict__ raw0, double* __restrict__ out0,
                                                              const double* __restrict__ raw1, double* __restrict__ out1,
                                                              int ndoy, int64_t ngrid, int feb29,
                                                              const int32_t* __restrict__ nempty) {
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ngrid) return;
  const int ne = nempty[cell];
  if (ne > 0 && ne < ndoy) return;          // some doys without samples: clim_finish_compact_kernel
  FinishSrc src;
  src.init((blockIdx.y ? raw1 : raw0) + cell, ngrid, ndoy, feb29);
  double* __restrict__ o = (blockIdx.y ? out1 : out0) + cell;
  constexpr int h = (W - 1) / 2;
  double slot[W];
  int idx = ((-h) % ndoy + ndoy) % ndoy;
#pragma unroll
  for (int i = 0; i < W; ++i) { slot[i] = src.at(idx); idx = idx + 1 == ndoy ? 0 : idx + 1; }
#pragma unroll
  for (int i = W - 2; i >= 0; --i) slot[i] = slot[i] + slot[i + 1];
  for (int d0 = 0; d0 < ndoy; d0 += W) {
    double pre = 0.0;
    if (d0 + W < ndoy) {                        // full block of outputs, another block follows
      constexpr int G = 8;                      // loads in flight per thread
#pragma unroll
      for (int p0 = 0; p0 < W; p0 += G) {
        double en[G];
#pragma unroll
        for (int i = 0; i < G; ++i)
          if (p0 + i < W) { en[i] = src.at(idx); idx = idx + 1 == ndoy ? 0 : idx + 1; }
        // en[i] = e[(b+1)W + p0 + i]: feeds output p0 + i + 1 and becomes slot[p0 + i] of the next block
#pragma unroll
        for (int i = 0; i < G; ++i) {
          const int p = p0 + i;                 // output p (needs en of p - 1, loaded in the previous group)
          if (p < W) {
            if (p == 0) o[(int64_t)d0 * ngrid] = div_const<W>(slot[0]);
            if (p + 1 < W) {
              pre = p == 0 ? en[i] : pre + en[i];
              o[(int64_t)(d0 + p + 1) * ngrid] = div_const<W>(slot[p + 1] + pre);
            }
            slot[p] = en[i];
          }
        }
      }
    } else {
#pragma unroll
      for (int p = 0; p < W; ++p) {
        if (d0 + p < ndoy) {
          double S = slot[p];
          if (p >= 1) {
            const double e1 = src.at(idx); idx = idx + 1 == ndoy ? 0 : idx + 1;
            pre = p == 1 ? e1 : pre + e1;
            S = S + pre;
            slot[p - 1] = e1;
          }
          o[(int64_t)(d0 + p) * ngrid] = div_const<W>(S);
        }
      }
    }
    if (d0 + W < ndoy) {
#pragma unroll
      for (int i = W - 2; i >= 0; --i) slot[i] = slot[i] + slot[i + 1];
    }
  }
}

// Cells in which SOME doys have no sample (e.g. a sea-ice season that is NaN every year): the
// reference's groupby output simply lacks those doys for that cell, so feb29 acts on the labels
// that are present and runavg pads / rolls over the cell's own compacted doy axis
// (identify.py:137-151, :175-180, :233-241).  One warp = one such cell: the column is compacted
// into shared memory, every lane then evaluates outputs in the summation order of the block
// scheme above (oracle/xmhw_oracle.py:runavg on the compacted axis), absent doys become NaN.
constexpr int FINC_WARPS = 4;
__global__ void __launch_bounds__(FINC_WARPS * 32) clim_finish_compact_kernel(
    const double* __restrict__ raw0, double* __restrict__ out0, const double* __restrict__ raw1,
    double* __restrict__ out1, int ndoy, int64_t ngrid, int feb29, int W, const int32_t* __restrict__ nempty) {
  extern __shared__ double fc_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t cell = (int64_t)blockIdx.x * FINC_WARPS + wib;
  if (cell >= ngrid) return;
  const int ne = nempty[cell];
  if (ne <= 0 || ne >= ndoy) return;
  double* c = fc_smem + (size_t)wib * ndoy * 2;              // compacted values
  int* lab = (int*)(c + ndoy);                               // their doy index (0-based)
  for (int var = 0; var < 2; ++var) {
    const double* raw = (var ? raw1 : raw0) + cell;
    double* o = (var ? out1 : out0) + cell;
    if (!raw || !o) continue;
    int m = 0;
    for (int d0 = 0; d0 < ndoy; d0 += 32) {
      const int d = d0 + lane;
      double v = d < ndoy ? raw[(int64_t)d * ngrid] : qnan();
      const bool pres = v == v;
      if (pres && feb29 && ndoy >= 61 && d == 59) {          // label 60 <- mean of the labels 59, 60, 61 present
        double acc = 0.0; int n = 0;
        for (int dd = 58; dd <= 60; ++dd) {
          const double x = raw[(int64_t)dd * ngrid];
          if (x == x) { acc = acc + x; ++n; }
        }
        v = acc / (double)n;
      }
      const uint32_t bal = __ballot_sync(0xffffffffu, pres);
      const int pos = m + __popc(bal & ((1u << lane) - 1u));
      if (pres) { c[pos] = v; lab[pos] = d; }
      else if (d < ndoy) o[(int64_t)d * ngrid] = qnan();
      m += __popc(bal);
    }
    __syncwarp();
    const int h = W > 1 ? (W - 1) / 2 : 0;
    const int L = m + W - 1;
    const double w = (double)W;
    for (int i = lane; i < m; i += 32) {
      double S;
      if (W <= 1) {
        S = c[i];
      } else {
        // e[j] = c[(j - h) mod m]; block b = i / W: suffix of block b from i (right to left)
        // plus prefix of block b + 1 up to i + W - 1 (left to right)
        const int b = i / W, p = i - b * W;
        const int hi = min(L, (b + 1) * W);
        int k = (((hi - 1 - h) % m) + m) % m;                 // index of e[hi - 1]
        double acc = c[k];
        for (int j = hi - 2; j >= i; --j) { k = k == 0 ? m - 1 : k - 1; acc = c[k] + acc; }
        S = acc;
        if (p > 0) {
          k = ((((b + 1) * W - h) % m) + m) % m;              // index of e[(b + 1) W]
          double pre = c[k];
          for (int j = (b + 1) * W + 1; j <= i + W - 1; ++j) { k = k + 1 == m ? 0 : k + 1; pre = pre + c[k]; }
          S = S + pre;
        }
        S = S / w;
      }
      o[(int64_t)lab[i] * ngrid] = S;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// K2  exceedance sweep, doy-major: the threshold of (cell, doy) stays in a
// register while every year's sample of that doy streams through, so no
// per-timestep threshold fetch is needed.  ts > thresh in float64 is evaluated
// exactly as ts > round_down_f32(thresh).
// ---------------------------------------------------------------------------
constexpr int EXC_WARPS = 8;
constexpr int EXC_BATCH = 16;   // rows in flight per lane (memory-level parallelism)

__global__ void __launch_bounds__(EXC_WARPS * 32) exceed_kernel(
    const float* __restrict__ ts, int64_t T, int64_t ngrid, const int32_t* __restrict__ doy_ptr,
    const int32_t* __restrict__ doy_tidx, int ndoy, int nchunk, const double* __restrict__ thresh,
    uint32_t* __restrict__ mask, int32_t* __restrict__ nvalid) {
  // the warps of a block take ADJACENT cell groups of the same doy chunk, so a block reads
  // EXC_WARPS x 128 contiguous bytes of every row (DRAM page locality)
  const int lane = threadIdx.x & 31;
  const int64_t ncg = (ngrid + 31) / 32;
  const int64_t cg = (int64_t)(blockIdx.x / nchunk) * EXC_WARPS + (threadIdx.x >> 5);
  const int chunk = (int)(blockIdx.x % nchunk);
  if (cg >= ncg) return;
  const int64_t cell = cg * 32 + lane;
  const bool ok = cell < ngrid;
  const float* col = ts + (ok ? cell : 0);
  uint32_t* mrow = mask + cg * T;
  const int d0 = (int)((int64_t)ndoy * chunk / nchunk), d1 = (int)((int64_t)ndoy * (chunk + 1) / nchunk);
  int cnt = 0;
  const uint32_t ng32 = (uint32_t)ngrid;          // row offset t * ngrid as one 32x32->64 multiply
  for (int d = d0; d < d1; ++d) {
    const double th = ok ? thresh[(int64_t)d * ngrid + cell] : qnan();
    // NaN threshold or out-of-grid lane -> +inf: nothing exceeds; ts > th (f64) == ts > rd_f32(th)
    const float thr = (th == th) ? __double2float_rd(th) : __int_as_float(0x7f800000);
    const int j1 = doy_ptr[d + 1];
    for (int j = doy_ptr[d]; j < j1; j += EXC_BATCH) {
      float v[EXC_BATCH];
      uint32_t t[EXC_BATCH];
#pragma unroll
      for (int i = 0; i < EXC_BATCH; ++i) {
        t[i] = (uint32_t)__ldg(doy_tidx + min(j + i, j1 - 1));        // tail rows repeat the last one
        v[i] = __ldg(col + (uint64_t)t[i] * ng32);
      }
#pragma unroll
      for (int i = 0; i < EXC_BATCH; ++i) {
        const bool act = j + i < j1;                                  // warp-uniform
        cnt += act && v[i] == v[i];
        const uint32_t bits = __ballot_sync(0xffffffffu, v[i] > thr);
        if (lane == 0 && act) mrow[t[i]] = bits;
      }
    }
  }
  if (ok && cnt) atomicAdd(nvalid + cell, cnt);
}

// Wide variant (grid a multiple of 4 cells, 16-byte aligned rows): one lane = 4 adjacent cells
// (float4), one warp = 128 cells = 512 contiguous bytes per row -- 4x fewer DRAM row
// activations and load instructions per byte than the 128-byte variant.  The 4 exceedance
// bits of a lane are merged over 8-lane groups into the 4 mask words of the 4 cell groups.
__global__ void __launch_bounds__(EXC_WARPS * 32) exceed4_kernel(
    const float* __restrict__ ts, int64_t T, int64_t ngrid, const int32_t* __restrict__ doy_ptr,
    const int32_t* __restrict__ doy_tidx, int ndoy, int nchunk, const double* __restrict__ thresh,
    uint32_t* __restrict__ mask, int32_t* __restrict__ nvalid) {
  const int lane = threadIdx.x & 31;
  const int64_t nsg = (ngrid + 127) / 128;                  // super-groups of 128 cells
  // adjacent super-groups per block, same doy chunk: EXC_WARPS x 512 contiguous bytes per row
  const int64_t sg = (int64_t)(blockIdx.x / nchunk) * EXC_WARPS + (threadIdx.x >> 5);
  const int chunk = (int)(blockIdx.x % nchunk);
  if (sg >= nsg) return;
  const int64_t cell = sg * 128 + 4 * lane;                 // first of this lane's 4 cells
  const bool ok = cell < ngrid;                             // ngrid % 4 == 0: all 4 or none
  const float* col = ts + (ok ? cell : 0);
  const int64_t cg = sg * 4 + (lane >> 3);                  // 32-cell group of this lane's cells
  uint32_t* mrow = mask + cg * T;
  const bool writer = (lane & 7) == 0 && cg < (ngrid + 31) / 32;
  const int sh = 4 * (lane & 7);
  const int d0 = (int)((int64_t)ndoy * chunk / nchunk), d1 = (int)((int64_t)ndoy * (chunk + 1) / nchunk);
  int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  const uint32_t ng32 = (uint32_t)ngrid;
  const float inf = __int_as_float(0x7f800000);
  for (int d = d0; d < d1; ++d) {
    float thr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double th = ok ? thresh[(int64_t)d * ngrid + cell + k] : qnan();
      thr[k] = (th == th) ? __double2float_rd(th) : inf;
    }
    const int j1 = doy_ptr[d + 1];
    for (int j = doy_ptr[d]; j < j1; j += EXC_BATCH) {
      float4 v[EXC_BATCH];
      uint32_t t[EXC_BATCH];
#pragma unroll
      for (int i = 0; i < EXC_BATCH; ++i) {
        t[i] = (uint32_t)__ldg(doy_tidx + min(j + i, j1 - 1));
        v[i] = __ldg(reinterpret_cast<const float4*>(col + (uint64_t)t[i] * ng32));
      }
#pragma unroll
      for (int i = 0; i < EXC_BATCH; ++i) {
        const bool act = j + i < j1;
        c0 += act && v[i].x == v[i].x; c1 += act && v[i].y == v[i].y;
        c2 += act && v[i].z == v[i].z; c3 += act && v[i].w == v[i].w;
        uint32_t x = ((uint32_t)(v[i].x > thr[0]) | ((uint32_t)(v[i].y > thr[1]) << 1) |
                      ((uint32_t)(v[i].z > thr[2]) << 2) | ((uint32_t)(v[i].w > thr[3]) << 3)) << sh;
        if (!ok) x = 0u;
        x |= __shfl_xor_sync(0xffffffffu, x, 1);
        x |= __shfl_xor_sync(0xffffffffu, x, 2);
        x |= __shfl_xor_sync(0xffffffffu, x, 4);
        if (writer && act) mrow[t[i]] = x;
      }
    }
  }
  if (ok) {
    if (c0) atomicAdd(nvalid + cell, c0);
    if (c1) atomicAdd(nvalid + cell + 1, c1);
    if (c2) atomicAdd(nvalid + cell + 2, c2);
    if (c3) atomicAdd(nvalid + cell + 3, c3);
  }
}

// ---------------------------------------------------------------------------
// K3  event finding: warp = 32 cells; 32 mask words (32 consecutive times) are
// bit-transposed in the warp so each lane holds 32 time steps of its own cell,
// then the lane runs the run-length / min-duration / gap-join rules.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t transpose32(uint32_t x, int lane) {
  // lane i holds row i (bit c = column c); returns column `lane` (bit r = row r)
#pragma unroll
  for (int j = 16; j >= 1; j >>= 1) {
    const uint32_t m = j == 16 ? 0x0000ffffu : j == 8 ? 0x00ff00ffu : j == 4 ? 0x0f0f0f0fu
                     : j == 2 ? 0x33333333u : 0x55555555u;
    uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
    x = (lane & j) ? ((x & ~m) | ((y >> j) & m)) : ((x & m) | ((y & m) << j));
  }
  return x;
}

constexpr int EVT_WARPS = 4;

struct CountEmit {
  int n;
  __device__ __forceinline__ void operator()(int, int) { ++n; }
};
// counts and also parks the first `cap` events of the cell in a staging table
// (slot-major: [cell group][slot][lane]) so that no second pass over the mask is needed
struct StageEmit {
  int n, cap;
  int32_t* s_row; int32_t* e_row;      // this lane's column of the group's staging block
  __device__ __forceinline__ void operator()(int s, int e) {
    if (n < cap) { s_row[(size_t)n * 32] = s; e_row[(size_t)n * 32] = e; }
    ++n;
  }
};
struct FillEmit {
  int32_t* ev; int64_t cap, pos; int32_t cell;
  __device__ __forceinline__ void operator()(int s, int e) {
    ev[EI_CELL * cap + pos] = cell;
    ev[EI_START * cap + pos] = s;
    ev[EI_END * cap + pos] = e;
    ++pos;
  }
};

template <bool FILL>
__global__ void __launch_bounds__(EVT_WARPS * 32) events_kernel(
    const uint32_t* __restrict__ mask, int64_t T, int64_t ngrid, int min_dur, int join, int max_gap,
    int32_t* __restrict__ counts, const int64_t* __restrict__ offsets, int64_t cap, int32_t* __restrict__ ev) {
  const int lane = threadIdx.x & 31;
  const int64_t cg = (int64_t)blockIdx.x * EVT_WARPS + (threadIdx.x >> 5);
  if (cg >= (ngrid + 31) / 32) return;
  const int64_t cell = cg * 32 + lane;
  const bool ok = cell < ngrid;
  const uint32_t* mrow = mask + cg * T;
  RunFinder rf(min_dur, join, max_gap);
  CountEmit ce; ce.n = 0;
  FillEmit fe; fe.ev = ev; fe.cap = cap; fe.cell = (int32_t)cell; fe.pos = (FILL && ok) ? offsets[cell] : 0;
  for (int64_t t0 = 0; t0 < T; t0 += 32) {
    uint32_t x = (t0 + lane < T) ? __ldg(mrow + t0 + lane) : 0u;
    uint32_t bits = transpose32(x, lane);
    if (FILL) { if (ok) rf.feed(bits, (int)t0, fe); } else rf.feed(bits, (int)t0, ce);
  }
  if (FILL) { if (ok) rf.finish((int)T, fe); }
  else { rf.finish((int)T, ce); if (ok) counts[cell] = ce.n; }
}

// count pass that also stages the events: stage [ncg][2][cap][32] int32 (starts, ends)
__global__ void __launch_bounds__(EVT_WARPS * 32) events_stage_kernel(
    const uint32_t* __restrict__ mask, int64_t T, int64_t ngrid, int min_dur, int join, int max_gap,
    int32_t* __restrict__ counts, int32_t* __restrict__ stage, int cap, int32_t* __restrict__ overflow) {
  const int lane = threadIdx.x & 31;
  const int64_t cg = (int64_t)blockIdx.x * EVT_WARPS + (threadIdx.x >> 5);
  if (cg >= (ngrid + 31) / 32) return;
  const int64_t cell = cg * 32 + lane;
  const bool ok = cell < ngrid;
  const uint32_t* mrow = mask + cg * T;
  RunFinder rf(min_dur, join, max_gap);
  StageEmit se;
  se.n = 0; se.cap = ok ? cap : 0;
  se.s_row = stage + (size_t)cg * 2 * cap * 32 + lane;
  se.e_row = se.s_row + (size_t)cap * 32;
  // 4 mask words in flight: the run scan of one word is a data-dependent loop the compiler cannot
  // hoist the next load across, and a warp's words are consumed strictly in order
  for (int64_t t0 = 0; t0 < T; t0 += 128) {
    uint32_t x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) x[u] = (t0 + 32 * u + lane < T) ? __ldg(mrow + t0 + 32 * u + lane) : 0u;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (t0 + 32 * u < T) rf.feed(transpose32(x[u], lane), (int)(t0 + 32 * u), se);
  }
  rf.finish((int)T, se);
  if (ok) {
    counts[cell] = se.n;
    if (se.n > cap) atomicOr(overflow, 1);
  }
}

// staged events -> event table columns (cell, start, end) at the scanned offsets
__global__ void __launch_bounds__(EVT_WARPS * 32) events_gather_kernel(
    const int32_t* __restrict__ stage, int cap_per_cell, const int32_t* __restrict__ counts,
    const int64_t* __restrict__ offsets, int64_t ngrid, int64_t cap, int32_t* __restrict__ ev) {
  const int lane = threadIdx.x & 31;
  const int64_t cg = (int64_t)blockIdx.x * EVT_WARPS + (threadIdx.x >> 5);
  if (cg >= (ngrid + 31) / 32) return;
  const int64_t cell = cg * 32 + lane;
  const bool ok = cell < ngrid;
  const int n = ok ? counts[cell] : 0;
  const int64_t pos = ok ? offsets[cell] : 0;
  const int32_t* s_row = stage + (size_t)cg * 2 * cap_per_cell * 32 + lane;
  const int32_t* e_row = s_row + (size_t)cap_per_cell * 32;
  for (int k = 0; k < n; ++k) {
    ev[EI_CELL * cap + pos + k] = (int32_t)cell;
    ev[EI_START * cap + pos + k] = s_row[(size_t)k * 32];
    ev[EI_END * cap + pos + k] = e_row[(size_t)k * 32];
  }
}

// ---------------------------------------------------------------------------
// K3f  fused detect: ONE time-major pass over the series does the threshold compare, the run
// rules and the event statistics (identify.py:367-479, :273-325, features.py:22-295).
//
// A block owns 32 adjacent cells (lane = cell) for the whole time axis; its DF_WARPS warps take
// consecutive 32-step words of a chunk, so every lane builds the exceedance bits of ITS OWN cell
// from coalesced 128-byte row loads -- no mask array, no bit transpose.  The float32 round-down
// thresholds of the 32 cells (ts > thresh in float64 == ts > round_down_f32(thresh)) sit in shared
// memory for all doys.  Warp 0 then feeds the words of the chunk, in time order, to the per-cell
// run finder, which emits an event as soon as no later run can join it; events queue up in shared
// memory and, a block-full at a time, every thread computes one event's statistics while the rows
// it needs are still L2-resident (read a few hundred steps ago by this very block), so the sparse
// 4-bytes-per-32-byte-sector reads of the statistics hit L2 instead of DRAM.  Records go to a
// staging table in completion order together with (cell, ordinal in cell); after the scan of the
// per-cell counts events_scatter_kernel moves them to their place in the (cell, start) ordered table.
// ---------------------------------------------------------------------------
struct ClimCellMajor {
  const double2* cm;       // this cell's [ndoy] {thresh, seas} pairs
  __device__ __forceinline__ void get(int d, double& t, double& s) const {
    const double2 v = __ldg(cm + d);
    t = v.x; s = v.y;
  }
};

constexpr int DF_WARPS = 8;
constexpr int DF_QFLUSH = 32 * DF_WARPS;          // run the statistics when this many events wait

struct QueueEmit {
  int4* q; int* q_n; int lane; int k;
  __device__ __forceinline__ void operator()(int s, int e) {
    const int i = atomicAdd(q_n, 1);
    q[i] = make_int4(lane, s, e, k);
    ++k;
  }
};

__global__ void __launch_bounds__(DF_WARPS * 32) detect_fused_kernel(
    const float* __restrict__ ts, int64_t T, int64_t ngrid, const int32_t* __restrict__ doy, int ndoy,
    const double* __restrict__ thresh, const double2* __restrict__ cm, int min_dur, int join, int max_gap,
    int32_t* __restrict__ counts, int32_t* __restrict__ nvalid, int32_t* __restrict__ stage_i,
    double* __restrict__ stage_f, int64_t scap, int32_t* __restrict__ counter /* [0] staged, [1] overflow */) {
  extern __shared__ float df_smem[];
  float* thrs = df_smem;                                            // [ndoy][32]
  uint32_t* words = reinterpret_cast<uint32_t*>(thrs + (size_t)ndoy * 32);   // [DF_WARPS][32]
  int* sh = reinterpret_cast<int*>(words + DF_WARPS * 32);          // [0] queue length, [1] staging base, [2..33] valid counts
  int4* queue = reinterpret_cast<int4*>(sh + 36);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t cell = (int64_t)blockIdx.x * 32 + lane;
  const bool ok = cell < ngrid;
  const float* col = ts + (ok ? cell : 0);
  const float inf = __int_as_float(0x7f800000);
  for (int d = warp; d < ndoy; d += DF_WARPS) {
    const double th = ok ? thresh[(int64_t)d * ngrid + cell] : qnan();
    thrs[d * 32 + lane] = (th == th) ? __double2float_rd(th) : inf;   // NaN threshold / out-of-grid lane: nothing exceeds
  }
  if (threadIdx.x < 36) sh[threadIdx.x] = 0;
  __syncthreads();
  RunFinder rf(min_dur, join, max_gap);
  QueueEmit em{queue, sh, lane, 0};
  int cnt = 0;
  const uint32_t ng32 = (uint32_t)ngrid;

  auto stats_phase = [&]() {
    // every thread one queued event; records go to the staging table at a block-wide atomic base
    const int nq = sh[0];
    if (threadIdx.x == 0) sh[1] = atomicAdd(counter, nq);
    __syncthreads();
    const int64_t base = sh[1];
    for (int i = threadIdx.x; i < nq; i += DF_WARPS * 32) {
      const int4 ev = queue[i];
      const int64_t pos = base + i;
      if (pos >= scap) { atomicOr(counter + 1, 1); continue; }
      const int64_t c = (int64_t)blockIdx.x * 32 + ev.x;
      ClimCellMajor clim{cm + c * ndoy};
      event_stats<8, ClimCellMajor>(ts + c, clim, doy, ngrid, (int)T, ev.y, ev.z, stage_i + pos, stage_f + pos, scap);
      stage_i[(size_t)EI_CELL * scap + pos] = (int32_t)c;
      stage_i[(size_t)EI_COUNT * scap + pos] = ev.w;                  // ordinal of the event within its cell
    }
    __syncthreads();
    if (threadIdx.x == 0) sh[0] = 0;
    __syncthreads();
  };

  for (int64_t t0 = 0; t0 < T; t0 += 32 * DF_WARPS) {
    const int64_t tw = t0 + 32 * warp;                               // this warp's word
    uint32_t bits = 0u;
    if (tw < T) {
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int64_t t = tw + i < T ? tw + i : T - 1;               // clamped: loads stay unconditional
        v[i] = __ldg(col + (uint64_t)t * ng32);
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const bool in = tw + i < T;
        const int d = __ldg(doy + (in ? tw + i : T - 1)) - 1;
        const float thr = thrs[d * 32 + lane];
        bits |= (uint32_t)(in && ok && v[i] > thr) << i;
        cnt += in && ok && v[i] == v[i];
      }
    }
    words[warp * 32 + lane] = bits;
    __syncthreads();
    if (warp == 0) {
#pragma unroll 1
      for (int w = 0; w < DF_WARPS; ++w)
        if (t0 + 32 * w < T) rf.feed(words[w * 32 + lane], (int)(t0 + 32 * w), em);
      const int64_t t_end = t0 + 32 * DF_WARPS < T ? t0 + 32 * DF_WARPS : T;
      if (t_end < T) rf.flush_pending((int)t_end, em);
      else rf.finish((int)T, em);
    }
    __syncthreads();
    if (sh[0] >= DF_QFLUSH || t0 + 32 * DF_WARPS >= T) stats_phase();
  }
  if (cnt) atomicAdd(sh + 2 + lane, cnt);
  __syncthreads();
  if (warp == 0 && ok) {
    counts[cell] = em.k;
    nvalid[cell] = sh[2 + lane];
  }
}

// staged records -> the (cell, start) ordered event table: position = offsets[cell] + ordinal
__global__ void __launch_bounds__(256) events_scatter_kernel(const int32_t* __restrict__ stage_i,
                                                             const double* __restrict__ stage_f, int64_t scap,
                                                             int64_t nstaged, const int64_t* __restrict__ offsets,
                                                             int64_t cap, int32_t* __restrict__ ei, double* __restrict__ ef) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nstaged) return;
  const int64_t pos = offsets[stage_i[(size_t)EI_CELL * scap + j]] + stage_i[(size_t)EI_COUNT * scap + j];
#pragma unroll
  for (int k = 0; k < EI_COUNT; ++k) ei[(size_t)k * cap + pos] = stage_i[(size_t)k * scap + j];
#pragma unroll
  for (int k = 0; k < EF_COUNT; ++k) ef[(size_t)k * cap + pos] = stage_f[(size_t)k * scap + j];
}

// ---------------------------------------------------------------------------
// exclusive scan int32 -> int64 (three small kernels; n <= a few million)
// ---------------------------------------------------------------------------
constexpr int SCAN_ITEMS = 1024;

__global__ void scan_block_sums(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ bsum) {
  __shared__ long long red[32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_ITEMS;
  long long s = 0;
  for (int i = threadIdx.x; i < SCAN_ITEMS; i += blockDim.x)
    if (base + i < n) s += in[base + i];
  for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long tot = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
    bsum[blockIdx.x] = tot;
  }
}
__global__ void scan_of_sums(int64_t* bsum, int64_t nb) {
  // single warp: chunks of 32 block sums, exclusive scan in place, total at bsum[nb]
  const int lane = threadIdx.x;
  long long carry = 0;
  for (int64_t b0 = 0; b0 < nb; b0 += 32) {
    long long v = (b0 + lane < nb) ? bsum[b0 + lane] : 0, x = v;
    for (int o = 1; o < 32; o <<= 1) {
      long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (b0 + lane < nb) bsum[b0 + lane] = carry + x - v;
    carry += __shfl_sync(0xffffffffu, x, 31);
  }
  if (lane == 0) bsum[nb] = carry;
}
__global__ void scan_apply(const int32_t* __restrict__ in, int64_t n, const int64_t* __restrict__ bsum,
                           int64_t nb, int64_t* __restrict__ out) {
  // one warp per 1024-item block, sequential over 32 chunks of 32
  const int lane = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * SCAN_ITEMS;
  long long carry = bsum[blockIdx.x];
  for (int c = 0; c < SCAN_ITEMS / 32; ++c) {
    int64_t i = base + c * 32 + lane;
    long long v = i < n ? in[i] : 0, x = v;
    for (int o = 1; o < 32; o <<= 1) {
      long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (i < n) out[i] = carry + x - v;
    carry += __shfl_sync(0xffffffffu, x, 31);
  }
  if (blockIdx.x == nb - 1 && lane == 0) out[n] = bsum[nb];
}

// ---------------------------------------------------------------------------
// K4  per-event statistics: one thread = one event.
// ---------------------------------------------------------------------------
template <int BATCH, int MINB>
__global__ void __launch_bounds__(128, MINB) event_stats_kernel(const float* __restrict__ ts, int64_t T, int64_t ngrid,
                                   const int32_t* __restrict__ doy, const double* __restrict__ thr,
                                   const double* __restrict__ seas, int64_t nev, int64_t cap,
                                   int32_t* __restrict__ ei, double* __restrict__ ef) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nev) return;
  const int64_t cell = ei[EI_CELL * cap + i];
  const int s = ei[EI_START * cap + i], e = ei[EI_END * cap + i];
  event_stats<BATCH>(ts + cell, thr + cell, seas + cell, doy, ngrid, (int)T, s, e, ei + i, ef + i, cap);
}

// Cell-major interleaved climatology copy: cm[cell * ndoy + d] = {thresh[d][cell], seas[d][cell]}.
// 32 x 32 (doy x cell) tiles through shared memory: coalesced 256 B reads, 512 B writes.
__global__ void __launch_bounds__(256) clim_cellmajor_kernel(const double* __restrict__ thr, const double* __restrict__ seas,
                                                             int ndoy, int64_t ngrid, double2* __restrict__ cm) {
  __shared__ double tt[32][33], tsn[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t c0 = (int64_t)blockIdx.x * 32;
  const int d0 = blockIdx.y * 32;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int dl = ty + 8 * j, d = d0 + dl;
    const int64_t c = c0 + tx;
    if (d < ndoy && c < ngrid) {
      tt[dl][tx] = __ldg(thr + (int64_t)d * ngrid + c);
      tsn[dl][tx] = __ldg(seas + (int64_t)d * ngrid + c);
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int cl = ty + 8 * j, d = d0 + tx;
    const int64_t c = c0 + cl;
    if (d < ndoy && c < ngrid) cm[c * ndoy + d] = make_double2(tt[tx][cl], tsn[tx][cl]);
  }
}

template <int BATCH, int MINB>
__global__ void __launch_bounds__(128, MINB) event_stats_cm_kernel(const float* __restrict__ ts, int64_t T, int64_t ngrid,
                                   const int32_t* __restrict__ doy, int ndoy, const double2* __restrict__ cm,
                                   int64_t nev, int64_t cap, int32_t* __restrict__ ei, double* __restrict__ ef) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nev) return;
  const int64_t cell = ei[EI_CELL * cap + i];
  const int s = ei[EI_START * cap + i], e = ei[EI_END * cap + i];
  ClimCellMajor clim{cm + cell * ndoy};
  event_stats<BATCH, ClimCellMajor>(ts + cell, clim, doy, ngrid, (int)T, s, e, ei + i, ef + i, cap);
}

// ---------------------------------------------------------------------------
// synthetic SST generator (bit-identical to xmhw_b200/synth.py)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

__global__ void synth_kernel(float* __restrict__ ts, int64_t T, int64_t ngrid, int64_t cell0,
                             const uint8_t* __restrict__ land, const double* __restrict__ season,
                             uint64_t seed, double rho, double sigma, double noise_scale,
                             uint32_t nan_ppm, uint32_t coherent) {
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ngrid) return;
  const uint64_t gid = (uint64_t)(cell0 + cell);
  const bool is_land = land && land[cell];
  const uint64_t h0 = splitmix64(seed ^ (gid * 0xD1B54A32D192ED03ull));
  // climatology parameters (mean, amplitude, phase) are shared by blocks of `coherent` consecutive
  // cells (1 = every cell its own, the benchmark default); the AR(1) noise is always per cell
  const uint64_t pgid = (gid / coherent) * coherent;
  const uint64_t hp = splitmix64(seed ^ (pgid * 0xD1B54A32D192ED03ull));
  const double m = 28.0 * u01(splitmix64(hp + 1));
  const double A = 1.0 + 5.0 * u01(splitmix64(hp + 2));
  const int phi = (int)(365.0 * u01(splitmix64(hp + 3)));
  double x = 0.0;
  for (int64_t t = 0; t < T; ++t) {
    const uint64_t r = splitmix64(h0 ^ ((uint64_t)t * 0x9E3779B97F4A7C15ull + 0x1234567ull));
    const int sum16 = (int)(r & 0xffff) + (int)((r >> 16) & 0xffff) + (int)((r >> 32) & 0xffff) + (int)(r >> 48);
    const double eps = (double)(sum16 - 131070) * noise_scale;
    x = rho * x + sigma * eps;
    const double v = (m + A * season[t + phi]) + x;
    float out = (float)(rint(v * 100.0) / 100.0);
    if (nan_ppm) {
      const uint64_t q = splitmix64(r ^ 0xA5A5A5A5A5A5A5A5ull);
      if ((uint32_t)(q % 1000000ull) < nan_ppm) out = __uint_as_float(0x7fc00000u);
    }
    if (is_land) out = __uint_as_float(0x7fc00000u);
    ts[t * ngrid + cell] = out;
  }
}

// ---------------------------------------------------------------------------
// pre-step: linear interpolation along time of NaN runs no longer than max_pad
// (xmhw.py:159-160, :409-410 `interpolate_na(dim=tdim, max_gap=maxPadLength)`), in place.
// One thread = one cell (coalesced rows); np.interp arithmetic: slope * (x - x0) + y0 in f64.
// ---------------------------------------------------------------------------
__global__ void interp_gaps_kernel(float* __restrict__ ts, int64_t T, int64_t ngrid, int max_pad) {
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ngrid) return;
  float* col = ts + cell;
  int64_t last_t = -1;
  float last_v = 0.0f;
  for (int64_t t = 0; t < T; ++t) {
    const float v = col[t * ngrid];
    if (v == v) {
      const int64_t gap = t - last_t - 1;
      if (last_t >= 0 && gap >= 1 && gap <= max_pad) {
        const double slope = ((double)v - (double)last_v) / (double)(gap + 1);
        for (int64_t k = 1; k <= gap; ++k)
          col[(last_t + k) * ngrid] = (float)(slope * (double)k + (double)last_v);
      }
      last_t = t;
      last_v = v;
    }
  }
}

// ---------------------------------------------------------------------------
// land_check census (identify.py:522-525): non-NaN samples per cell.  One lane = 4 adjacent cells
// (float4 rows) when the grid allows, 8 rows in flight; the time axis is split over blockIdx.y.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) count_valid_kernel(const float* __restrict__ ts, int64_t T, int64_t ngrid,
                                                          int tchunk, int32_t* __restrict__ nvalid) {
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ngrid) return;
  const int64_t t0 = (int64_t)blockIdx.y * tchunk, t1 = min(T, t0 + tchunk);
  const float* col = ts + cell;
  int cnt = 0;
  for (int64_t t = t0; t < t1; t += 8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldg(col + min(t + i, t1 - 1) * ngrid);
#pragma unroll
    for (int i = 0; i < 8; ++i) cnt += (t + i < t1) && v[i] == v[i];
  }
  if (cnt) atomicAdd(nvalid + cell, cnt);
}

// ---------------------------------------------------------------------------
// processing order of the 32-cell groups for the top-K sweep (xmhw_clim_sweep2_f32 group_order): groups whose 32
// cells are all NaN in three probe rows (first, middle, last time step: land, as far as a probe can tell) go
// last, both halves in grid order -- a stable partition.  Any permutation is correct; this one makes the
// blocks of the sweep homogeneous at the cost of three row reads.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) group_flags_kernel(const float* __restrict__ ts, int64_t T, int64_t ngrid,
                                                          uint8_t* __restrict__ flags) {
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;       // one warp = one group
  const int lane = threadIdx.x & 31;
  const int64_t ncg = (ngrid + 31) / 32;
  if (g >= ncg) return;
  const int64_t cell = g * 32 + lane;
  bool valid = false;
  if (cell < ngrid) {
    const float a = __ldg(ts + cell), b = __ldg(ts + (T / 2) * ngrid + cell), c = __ldg(ts + (T - 1) * ngrid + cell);
    valid = a == a || b == b || c == c;
  }
  const bool any = __any_sync(0xffffffffu, valid);
  if (lane == 0) flags[g] = any ? 0 : 1;
}

__global__ void __launch_bounds__(1024) group_order_kernel(const uint8_t* __restrict__ flags, int64_t ncg,
                                                           int32_t* __restrict__ order) {
  // ONE block: count the groups with data, then assign positions chunk by chunk (ballot ranks + warp totals)
  __shared__ int wtot[2][32];
  __shared__ int base[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int cnt = 0;
  for (int64_t g = tid; g < ncg; g += 1024) cnt += flags[g] == 0;
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if (lane == 0) wtot[0][warp] = cnt;
  __syncthreads();
  if (tid == 0) {
    int n = 0;
    for (int w = 0; w < 32; ++w) n += wtot[0][w];
    base[0] = 0; base[1] = n;
  }
  __syncthreads();
  for (int64_t g0 = 0; g0 < ncg; g0 += 1024) {
    const int64_t g = g0 + tid;
    const int f = g < ncg ? (int)flags[g] : 2;                 // 0 data, 1 land-looking, 2 past the end
    const unsigned b0 = __ballot_sync(0xffffffffu, f == 0), b1 = __ballot_sync(0xffffffffu, f == 1);
    const unsigned lt = (1u << lane) - 1u;
    if (lane == 0) { wtot[0][warp] = __popc(b0); wtot[1][warp] = __popc(b1); }
    __syncthreads();
    if (f < 2) {
      int off = 0;
      for (int w = 0; w < warp; ++w) off += wtot[f][w];
      order[base[f] + off + __popc((f == 0 ? b0 : b1) & lt)] = (int32_t)g;
    }
    __syncthreads();
    if (tid == 0) {
      int n0 = 0, n1 = 0;
      for (int w = 0; w < 32; ++w) { n0 += wtot[0][w]; n1 += wtot[1][w]; }
      base[0] += n0; base[1] += n1;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// intermediate=True (identify.py:404-411): per-timestep fields of mhw_df (features.py:22-69)
// ---------------------------------------------------------------------------
__global__ void event_labels_kernel(const int32_t* __restrict__ ei, int64_t nev, int64_t cap, int64_t ngrid,
                                    double* __restrict__ events) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nev) return;
  const int64_t cell = ei[EI_CELL * cap + i];
  const int s = ei[EI_START * cap + i], e = ei[EI_END * cap + i];
  for (int t = s; t <= e; ++t) events[(int64_t)t * ngrid + cell] = (double)s;      // label = start index
}

__global__ void intermediate_kernel(const float* __restrict__ ts, int64_t T, int64_t ngrid,
                                    const int32_t* __restrict__ doy, const double* __restrict__ thresh,
                                    const double* __restrict__ seas, const double* __restrict__ events,
                                    xmhw_intermediate out) {
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t t = blockIdx.y;
  if (cell >= ngrid) return;
  const int64_t i = t * ngrid + cell;
  const int d = doy[t] - 1;
  const float xf = ts[i];
  const double x = (double)xf, th = thresh[(int64_t)d * ngrid + cell], se = seas[(int64_t)d * ngrid + cell];
  out.bthresh[i] = (uint8_t)(x > th);                       // identify.py:372
  const bool in_ev = events[i] == events[i];                // features.py:38
  const double nan = qnan();
  const double relS = x - se, relT = x - th, ths = th - se; // features.py:52-54
  const double norm = relT / ths, sev = relS / -(ths), cat = floor(1.0 + norm);     // :57-62
  out.seas[i] = in_ev ? se : nan;
  out.thresh[i] = in_ev ? th : nan;
  out.relSeas[i] = in_ev ? relS : nan;
  out.relThresh[i] = in_ev ? relT : nan;
  out.relThreshNorm[i] = in_ev ? norm : nan;
  out.severity[i] = in_ev ? sev : nan;
  out.cats[i] = in_ev ? cat : nan;
  out.mabs[i] = in_ev ? xf : __uint_as_float(0x7fc00000u);  // float32, features.py:68
  out.duration_moderate[i] = (uint8_t)(in_ev && cat == 1.0);                        // :63-66
  out.duration_strong[i] = (uint8_t)(in_ev && cat == 2.0);
  out.duration_severe[i] = (uint8_t)(in_ev && cat == 3.0);
  out.duration_extreme[i] = (uint8_t)(in_ev && cat >= 4.0);
}

// ---------------------------------------------------------------------------
// Downstream statistics over the compact event table (reference xmhw/stats.py; semantics: Eric
// Oliver's blockAverage / rank, see xmhw_b200/stats.py).  Events are ordered by cell then start,
// so the events of one (cell, block of years) are one contiguous run of that cell's range.
// ---------------------------------------------------------------------------
// one thread = one cell: count / mean / max / sum of the event properties per block of years
__global__ void __launch_bounds__(128) block_average_kernel(
    const int32_t* __restrict__ ei, const double* __restrict__ ef, int64_t cap, const int64_t* __restrict__ offsets,
    int64_t ngrid, const int32_t* __restrict__ block_of_t, int nblocks, int time_col, double* __restrict__ out) {
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ngrid) return;
  const int64_t e0 = offsets[cell], e1 = offsets[cell + 1];
  const size_t plane = (size_t)nblocks * ngrid;
  int64_t e = e0;
  for (int b = 0; b < nblocks; ++b) {
    int cnt = 0;
    double sum[XMHW_BA_NMEAN];
#pragma unroll
    for (int k = 0; k < XMHW_BA_NMEAN; ++k) sum[k] = 0.0;
    int nn[XMHW_BA_NMEAN];
#pragma unroll
    for (int k = 0; k < XMHW_BA_NMEAN; ++k) nn[k] = 0;
    double imax = -INFINITY, icum = 0.0;
    bool any_max = false;
    // events whose block is < b were consumed (or lie outside every block: skipped)
    while (e < e1) {
      const int be = block_of_t[ei[(size_t)time_col * cap + e]];
      if (be > b) break;
      if (be == b) {
        ++cnt;
        const double dur = (double)ei[(size_t)EI_DURATION * cap + e];
        const double v[XMHW_BA_NMEAN] = {
            dur, ef[(size_t)EF_INT_MAX * cap + e], ef[(size_t)EF_INT_MEAN * cap + e], ef[(size_t)EF_INT_VAR * cap + e],
            ef[(size_t)EF_INT_CUM * cap + e], ef[(size_t)EF_RT_MAX * cap + e], ef[(size_t)EF_RT_MEAN * cap + e],
            ef[(size_t)EF_RT_VAR * cap + e], ef[(size_t)EF_RT_CUM * cap + e], ef[(size_t)EF_ABS_MAX * cap + e],
            ef[(size_t)EF_ABS_MEAN * cap + e], ef[(size_t)EF_ABS_VAR * cap + e], ef[(size_t)EF_ABS_CUM * cap + e],
            ef[(size_t)EF_SEV_MEAN * cap + e], ef[(size_t)EF_SEV_CUM * cap + e], ef[(size_t)EF_RATE_ONSET * cap + e],
            ef[(size_t)EF_RATE_DECLINE * cap + e]};
#pragma unroll
        for (int k = 0; k < XMHW_BA_NMEAN; ++k)
          if (v[k] == v[k]) { sum[k] = sum[k] + v[k]; ++nn[k]; }          // pandas mean skips NaN
        if (v[1] == v[1]) { any_max = true; imax = v[1] > imax ? v[1] : imax; }
        if (v[4] == v[4]) icum = icum + v[4];
      }
      ++e;
    }
    const size_t o = (size_t)b * ngrid + cell;
    out[(size_t)XMHW_BA_COUNT_COL * plane + o] = (double)cnt;
#pragma unroll
    for (int k = 0; k < XMHW_BA_NMEAN; ++k) out[(size_t)(XMHW_BA_MEAN0 + k) * plane + o] = nn[k] ? sum[k] / (double)nn[k] : qnan();
    out[(size_t)XMHW_BA_IMAX_MAX * plane + o] = any_max ? imax : qnan();
    out[(size_t)XMHW_BA_TOTAL_ICUM * plane + o] = icum;                     // pandas sum of nothing = 0
  }
}

// one thread = one cell: mean / max / min of the series per block of years (NaN skipped)
__global__ void __launch_bounds__(128) block_ts_kernel(const float* __restrict__ ts, int64_t T, int64_t ngrid,
                                                       const int32_t* __restrict__ block_of_t, int nblocks,
                                                       double* __restrict__ out) {
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ngrid) return;
  const size_t plane = (size_t)nblocks * ngrid;
  for (int b = 0; b < nblocks; ++b) {
    out[0 * plane + (size_t)b * ngrid + cell] = qnan();
    out[1 * plane + (size_t)b * ngrid + cell] = qnan();
    out[2 * plane + (size_t)b * ngrid + cell] = qnan();
  }
  int cur = -1, n = 0;
  double sum = 0.0;
  float mx = -INFINITY, mn = INFINITY;
  for (int64_t t = 0; t <= T; ++t) {
    const int b = t < T ? block_of_t[t] : -2;
    if (b != cur) {
      if (cur >= 0 && n > 0) {
        out[0 * plane + (size_t)cur * ngrid + cell] = sum / (double)n;
        out[1 * plane + (size_t)cur * ngrid + cell] = (double)mx;
        out[2 * plane + (size_t)cur * ngrid + cell] = (double)mn;
      }
      cur = b; n = 0; sum = 0.0; mx = -INFINITY; mn = INFINITY;
    }
    if (t < T && b >= 0) {
      const float v = ts[t * ngrid + cell];
      if (v == v) { ++n; sum = sum + (double)v; mx = v > mx ? v : mx; mn = v < mn ? v : mn; }
    }
  }
}

// one thread = one event: its days counted per category into the block of years of EACH DAY
__global__ void __launch_bounds__(128) block_cat_days_kernel(
    const float* __restrict__ ts, int64_t ngrid, const int32_t* __restrict__ doy, const double* __restrict__ thr,
    const double* __restrict__ seas, const int32_t* __restrict__ ei, int64_t nev, int64_t cap,
    const int32_t* __restrict__ block_of_t, int nblocks, int32_t* __restrict__ days) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nev) return;
  const int64_t cell = ei[(size_t)EI_CELL * cap + i];
  const int s = ei[(size_t)EI_START * cap + i], e = ei[(size_t)EI_END * cap + i];
  const size_t plane = (size_t)nblocks * ngrid;
  for (int t = s; t <= e; ++t) {
    const int b = block_of_t[t];
    if (b < 0) continue;
    const int d = doy[t] - 1;
    const double x = (double)ts[(int64_t)t * ngrid + cell];
    const double th = thr[(int64_t)d * ngrid + cell], se = seas[(int64_t)d * ngrid + cell];
    const double cat = floor(1.0 + (x - th) / (th - se));                   // features.py:57-62
    if (cat >= 1.0) {
      const int c = cat >= 4.0 ? 3 : (int)cat - 1;
      atomicAdd(days + (size_t)c * plane + (size_t)b * ngrid + cell, 1);
    }
  }
}

// one thread = one event: rank of its value among the events of its cell, 1 = largest
// (numpy: len - argsort(argsort(values)); NaN sorts last, ties keep table order)
__global__ void __launch_bounds__(128) event_rank_kernel(const double* __restrict__ col, const int32_t* __restrict__ cells,
                                                         const int64_t* __restrict__ offsets, int64_t nev,
                                                         double* __restrict__ rank) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nev) return;
  const int64_t c = cells[i];
  const int64_t e0 = offsets[c], e1 = offsets[c + 1];
  const double v = col[i];
  const bool vnan = v != v;
  int64_t below = 0;                       // position in the ascending stable sort
  for (int64_t j = e0; j < e1; ++j) {
    const double w = col[j];
    const bool wnan = w != w;
    const bool lt = vnan ? (!wnan || j < i) : (!wnan && (w < v || (w == v && j < i)));
    below += lt;
  }
  rank[i] = (double)((e1 - e0) - below);
}

inline int cuda_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace

extern "C" {

int xmhw_abi_version(void) { return XMHW_ABI_VERSION; }

const char* xmhw_strerror(int code) {
  switch (code) {
    case 0: return "ok";
    case XMHW_E_ARG: return "invalid argument (null pointer or non-positive size)";
    case XMHW_E_PLAN: return "inconsistent climatology plan";
    case XMHW_E_SMEM: return "climatology plan needs more shared memory than one SM provides";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown xmhw error";
  }
}

int xmhw_clim_sweep_f32(const float* ts, int64_t T, int64_t ngrid, const xmhw_clim_plan* plan,
                        double* thresh_raw, double* seas_raw, int32_t* nempty, uint32_t* scratch, void* stream) {
  if (!ts || !plan || !thresh_raw || !seas_raw || !nempty || T <= 0 || ngrid <= 0) return XMHW_E_ARG;
  if (plan->scratch_rows < 0 || (plan->scratch_rows > 0 && !scratch)) return XMHW_E_ARG;
  if (ngrid > 0xffffffffll || T > 0x7fffffffll) return XMHW_E_ARG;
  if (plan->nsteps <= 0 || plan->pool_rows <= 0 || plan->max_size > 48 || plan->nmax <= 0) return XMHW_E_PLAN;
  size_t smem = (size_t)(plan->pool_rows + POOL_STAGE_ROWS) * 128;
  static const int smem_pad = getenv("XMHW_B200_SWEEP_SMEM_PAD") ? atoi(getenv("XMHW_B200_SWEEP_SMEM_PAD")) : 0;
  smem += (size_t)smem_pad;       // development knob: lowers the resident warps per SM without touching the code
  if (smem > 227 * 1024) return XMHW_E_SMEM;
  ClimPlan p;
  memcpy(&p, plan, sizeof(p));
  const int64_t ncg = (ngrid + 31) / 32;
  cudaError_t e;
  static const int minb = getenv("XMHW_B200_SWEEP_MINB") ? atoi(getenv("XMHW_B200_SWEEP_MINB")) : 16;   // development knob
#define XMHW_SWEEP(N, B)                                                                                            \
  {                                                                                                                 \
    e = cudaFuncSetAttribute(clim_sweep_kernel<N, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
    if (e != cudaSuccess) return (int)e;                                                                            \
    clim_sweep_kernel<N, B><<<(unsigned)ncg, 32, smem, (cudaStream_t)stream>>>(p, ts, ngrid, thresh_raw, seas_raw,  \
                                                                               nempty, scratch);                   \
  }
  if (plan->max_size <= 32) {
    if (minb >= 24) XMHW_SWEEP(32, 24) else if (minb >= 20) XMHW_SWEEP(32, 20) else if (minb >= 16) XMHW_SWEEP(32, 16)
    else if (minb >= 14) XMHW_SWEEP(32, 14) else XMHW_SWEEP(32, 12)
  } else {
    XMHW_SWEEP(48, 10)
  }
#undef XMHW_SWEEP
  return cuda_status();
}

int xmhw_clim_sweep2_f32(const float* ts, int64_t T, int64_t ngrid, const xmhw_clim_plan2* plan,
                         double* thresh_raw, double* seas_raw, int32_t* nempty, int32_t* group_order, void* stream) {
  if (!ts || !plan || !thresh_raw || !seas_raw || !nempty || T <= 0 || ngrid <= 0) return XMHW_E_ARG;
  if (ngrid > 0xffffffffll || T > 0x7fffffffll) return XMHW_E_ARG;
  if (plan->nsteps <= 0 || plan->nsteps > SC_MAX_STEPS || plan->nslots <= 0 || plan->nslots > 32 ||
      plan->cap < plan->kp || plan->slot_rows != plan->cap + 3 || plan->max_size <= 0 || plan->max_size > 48 ||
      plan->n_init <= 0 || plan->n_init >= SC_MAX_INIT)
    return XMHW_E_PLAN;
  const size_t smem1 = (size_t)plan->nslots * plan->slot_rows * 128;      // per warp
  if (smem1 > 227 * 1024) return XMHW_E_SMEM;
  if ((uint64_t)ngrid * 4u > 0xffffffffull) return XMHW_E_ARG;              // byte offset of a time row in 32 bits
  const ClimPlan2& p = *reinterpret_cast<const ClimPlan2*>(plan);      // copied into the launch parameters
  const int64_t ncg = (ngrid + 31) / 32;
  // development knob: warps per block (lockstep group).  Default: as many as fit one SM, at most 4.
  static const int wpb_env = getenv("XMHW_B200_SWEEP2_WPB") ? atoi(getenv("XMHW_B200_SWEEP2_WPB")) : 0;
  int fit = (int)((227 * 1024) / (smem1 + 256));
  int wpb = wpb_env > 0 ? wpb_env : (fit >= 4 ? 4 : (fit >= 2 ? 2 : 1));
  if (wpb != 1 && wpb != 2 && wpb != 4) wpb = 1;
  if (wpb > fit) wpb = fit >= 2 ? 2 : 1;
  const size_t smem = smem1 * wpb;
  cudaError_t e;
#define XMHW_SWEEP2_W(K, N, W)                                                                                       \
  {                                                                                                                  \
    e = cudaFuncSetAttribute(clim_sweep2_kernel<K, N, W, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return (int)e;                                                                             \
    clim_sweep2_kernel<K, N, W, 1><<<(unsigned)((ncg + W - 1) / W), 32 * W, smem, (cudaStream_t)stream>>>(           \
        p, ts, ngrid, thresh_raw, seas_raw, nempty, group_order, sync_every);                                        \
  }
#define XMHW_SWEEP2(K, N)                                                                                            \
  { if (wpb == 4) XMHW_SWEEP2_W(K, N, 4) else if (wpb == 2) XMHW_SWEEP2_W(K, N, 2) else XMHW_SWEEP2_W(K, N, 1) }
  const bool big = plan->max_size > 32;
  const bool n30 = plan->max_size <= 30;      // 30-year daily series: the 30-input sorting network, 30-row atoms
  // Tensor-memory kernel (8 warps per SM, the slots that do not fit shared memory live in TMEM): taken when
  // fewer than 8 warps of the shared-memory kernel fit one SM and the slots split.  XMHW_B200_SWEEP2_TMEM = 0 / 1
  // forces it off / on (development knob, read per call).
  // The warps of a block may advance in lockstep (a barrier every `sync_every` doys) so that they stream the same
  // code through the instruction cache.  Measured (profiles/kernel_ms_r02r_sweep2_tmem.txt): since the step
  // lost its register copies the kernels run best WITHOUT the barrier (0); XMHW_B200_SWEEP2_SYNC = n sets it.
  const int sync_every = getenv("XMHW_B200_SWEEP2_SYNC") ? atoi(getenv("XMHW_B200_SWEEP2_SYNC")) : 0;
  const int tm_env = getenv("XMHW_B200_SWEEP2_TMEM") ? atoi(getenv("XMHW_B200_SWEEP2_TMEM")) : -1;
  const bool tm_on = tm_env == 1 || (tm_env < 0 && fit < TM_WARPS);
  if (tm_on) {
    const int max_smem_slots = (int)((227 * 1024 - 1024) / ((size_t)TM_WARPS * plan->slot_rows * 128));
    const int min_smem_slots = plan->nslots - TM_COLS_PER_WARP / plan->slot_rows;
    if (min_smem_slots <= max_smem_slots) {
      const int smem_slots = min_smem_slots > 0 ? min_smem_slots : 0;
      const size_t tsmem = (size_t)TM_WARPS * smem_slots * plan->slot_rows * 128;
      // persistent blocks, one per SM; the ticket counter lives behind the caller's processing order
      int dev = 0, sms = 0;
      if ((e = cudaGetDevice(&dev)) != cudaSuccess) return (int)e;
      if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
      const int64_t nblk = (ncg + TM_WARPS - 1) / TM_WARPS;
      // One block per 8 groups by default.  XMHW_B200_SWEEP2_PERSIST=1 (development knob) launches one PERSISTENT
      // block per SM whose warps draw their next group from the ticket word instead.  Measured on B200: 36.5 vs
      // 35.0 ms on the global grid, 9.05 vs 9.0 ms on the quarter grid, 4.51 vs 4.47 ms on an eighth of it -- the
      // block scheduler already fills the tail with the light land blocks that the processing order puts last.
      const bool persist = getenv("XMHW_B200_SWEEP2_PERSIST") && atoi(getenv("XMHW_B200_SWEEP2_PERSIST")) != 0;
      const unsigned tm_grid = (unsigned)(persist && nblk > sms ? sms : nblk);
      unsigned* const ticket = persist && group_order ? reinterpret_cast<unsigned*>(group_order + ncg) : nullptr;
      if (ticket && (e = cudaMemsetAsync(ticket, 0, sizeof(unsigned), (cudaStream_t)stream)) != cudaSuccess) return (int)e;
#define XMHW_TM(K, N)                                                                                                 \
  {                                                                                                                   \
    e = cudaFuncSetAttribute(clim_sweep2_tm_kernel<K, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem);    \
    if (e != cudaSuccess) return (int)e;                                                                              \
    clim_sweep2_tm_kernel<K, N><<<tm_grid, 32 * TM_WARPS, tsmem, (cudaStream_t)stream>>>(                             \
        p, ts, ngrid, thresh_raw, seas_raw, nempty, group_order, ticket, smem_slots);                                 \
  }
      switch (plan->kp) {
        case 8: if (big) XMHW_TM(8, 48) else if (n30) XMHW_TM(8, 30) else XMHW_TM(8, 32) break;
        case 16: if (big) XMHW_TM(16, 48) else if (n30) XMHW_TM(16, 30) else XMHW_TM(16, 32) break;
        case 24: if (big) XMHW_TM(24, 48) else if (n30) XMHW_TM(24, 30) else XMHW_TM(24, 32) break;
        case 36: if (big) XMHW_TM(36, 48) else if (n30) XMHW_TM(36, 30) else XMHW_TM(36, 32) break;
        case 48: if (big) XMHW_TM(48, 48) else if (n30) XMHW_TM(48, 30) else XMHW_TM(48, 32) break;
        default: return XMHW_E_PLAN;
      }
#undef XMHW_TM
      return cuda_status();
    }
  }
  switch (plan->kp) {
    case 8: if (big) XMHW_SWEEP2(8, 48) else if (n30) XMHW_SWEEP2(8, 30) else XMHW_SWEEP2(8, 32) break;
    case 16: if (big) XMHW_SWEEP2(16, 48) else if (n30) XMHW_SWEEP2(16, 30) else XMHW_SWEEP2(16, 32) break;
    case 24: if (big) XMHW_SWEEP2(24, 48) else if (n30) XMHW_SWEEP2(24, 30) else XMHW_SWEEP2(24, 32) break;
    case 36: if (big) XMHW_SWEEP2(36, 48) else if (n30) XMHW_SWEEP2(36, 30) else XMHW_SWEEP2(36, 32) break;
    case 48: if (big) XMHW_SWEEP2(48, 48) else if (n30) XMHW_SWEEP2(48, 30) else XMHW_SWEEP2(48, 32) break;
    default: return XMHW_E_PLAN;
  }
#undef XMHW_SWEEP2
#undef XMHW_SWEEP2_W
  return cuda_status();
}

int xmhw_clim_direct_f32(const float* ts, int64_t T, int64_t ngrid, const int32_t* rows, int32_t nrows, int32_t kp,
                         double q, double* thresh_row, double* seas_row, int32_t* nempty, void* stream) {
  if (!ts || !rows || !thresh_row || !seas_row || !nempty || T <= 0 || ngrid <= 0 || nrows <= 0) return XMHW_E_ARG;
  const int nt = 128;
  const unsigned nb = (unsigned)((ngrid + nt - 1) / nt);
  cudaStream_t st = (cudaStream_t)stream;
  switch (kp) {
    case 8: clim_direct_kernel<8><<<nb, nt, 0, st>>>(ts, ngrid, rows, nrows, q, thresh_row, seas_row, nempty); break;
    case 16: clim_direct_kernel<16><<<nb, nt, 0, st>>>(ts, ngrid, rows, nrows, q, thresh_row, seas_row, nempty); break;
    case 24: clim_direct_kernel<24><<<nb, nt, 0, st>>>(ts, ngrid, rows, nrows, q, thresh_row, seas_row, nempty); break;
    case 36: clim_direct_kernel<36><<<nb, nt, 0, st>>>(ts, ngrid, rows, nrows, q, thresh_row, seas_row, nempty); break;
    case 48: clim_direct_kernel<48><<<nb, nt, 0, st>>>(ts, ngrid, rows, nrows, q, thresh_row, seas_row, nempty); break;
    default: return XMHW_E_PLAN;
  }
  return cuda_status();
}

// cells with some (not all) doys empty: per-cell compacted doy axis
static int finish_compact(const double* raw0, double* out0, const double* raw1, double* out1, int32_t ndoy,
                          int64_t ngrid, int32_t feb29, int32_t W, const int32_t* nempty, cudaStream_t st) {
  const size_t smem = (size_t)FINC_WARPS * ndoy * 2 * sizeof(double);
  if (smem > 227 * 1024) return XMHW_E_SMEM;
  cudaError_t e = cudaFuncSetAttribute(clim_finish_compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  clim_finish_compact_kernel<<<(unsigned)((ngrid + FINC_WARPS - 1) / FINC_WARPS), FINC_WARPS * 32, smem, st>>>(
      raw0, out0, raw1, out1, ndoy, ngrid, feb29, W, nempty);
  return cuda_status();
}

static int finish_one(const double* raw, double* out, int32_t ndoy, int64_t ngrid, int32_t feb29,
                      int32_t smooth_width, const int32_t* nempty, cudaStream_t st) {
  const int nt = 128;
  const size_t smem = smooth_width > 1 ? (size_t)smooth_width * nt * sizeof(double) : 0;
  if (smem > 227 * 1024) return XMHW_E_SMEM;
  cudaError_t e = cudaFuncSetAttribute(clim_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  clim_finish_kernel<<<(unsigned)((ngrid + nt - 1) / nt), nt, smem, st>>>(raw, out, ndoy, ngrid, feb29, smooth_width,
                                                                           nempty);
  return cuda_status();
}

int xmhw_clim_finish_f64(const double* raw, double* out, int32_t ndoy, int64_t ngrid, int32_t feb29,
                         int32_t smooth_width, const int32_t* nempty, void* stream) {
  if (!raw || !out || !nempty || raw == out || ndoy <= 0 || ngrid <= 0) return XMHW_E_ARG;
  if (smooth_width > 1 && smooth_width % 2 == 0) return XMHW_E_ARG;
  int rc = finish_one(raw, out, ndoy, ngrid, feb29, smooth_width, nempty, (cudaStream_t)stream);
  if (rc) return rc;
  return finish_compact(raw, out, nullptr, nullptr, ndoy, ngrid, feb29, smooth_width, nempty, (cudaStream_t)stream);
}

int xmhw_clim_finish2_f64(const double* thresh_raw, double* thresh_out, const double* seas_raw, double* seas_out,
                          int32_t ndoy, int64_t ngrid, int32_t feb29, int32_t smooth_width, const int32_t* nempty,
                          void* stream) {
  if (!thresh_raw || !thresh_out || !seas_raw || !seas_out || !nempty || thresh_raw == thresh_out ||
      seas_raw == seas_out || ndoy <= 0 || ngrid <= 0) return XMHW_E_ARG;
  if (smooth_width > 1 && smooth_width % 2 == 0) return XMHW_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (smooth_width == 31) {
    const int nt = 128;
    dim3 grid((unsigned)((ngrid + nt - 1) / nt), 2);
    clim_finish_reg_kernel<31><<<grid, nt, 0, st>>>(thresh_raw, thresh_out, seas_raw, seas_out, ndoy, ngrid, feb29,
                                                    nempty);
    int rc = cuda_status();
    if (rc) return rc;
  } else {
    int rc = finish_one(thresh_raw, thresh_out, ndoy, ngrid, feb29, smooth_width, nempty, st);
    if (rc) return rc;
    rc = finish_one(seas_raw, seas_out, ndoy, ngrid, feb29, smooth_width, nempty, st);
    if (rc) return rc;
  }
  return finish_compact(thresh_raw, thresh_out, seas_raw, seas_out, ndoy, ngrid, feb29, smooth_width, nempty, st);
}

int xmhw_exceed_mask_f32(const float* ts, int64_t T, int64_t ngrid, const int32_t* doy_ptr,
                         const int32_t* doy_tidx, int32_t ndoy, const double* thresh, uint32_t* mask,
                         int32_t* nvalid, void* stream) {
  if (!ts || !doy_ptr || !doy_tidx || !thresh || !mask || !nvalid || T <= 0 || ngrid <= 0 || ndoy <= 0 ||
      ngrid > 0xffffffffll || T > 0x7fffffffll)
    return XMHW_E_ARG;
  const int64_t ncg = (ngrid + 31) / 32;
  // doy chunks: enough warps to fill the machine (148 SMs x 64 warps x 2) on small grids, and never
  // more than ~5 doys per warp -- measured on B200: 9.1 ms at 5 doys/chunk vs 10.6 ms at 120 (1440x720x30yr);
  // consecutive blocks are the chunks of the same cell groups, so the resident blocks read the whole
  // time extent of a narrow column band
  const int by_doy = (ndoy + 4) / 5;
  int nchunk = (int)((148 * 64 * 2 + ncg - 1) / ncg);
  if (nchunk < by_doy) nchunk = by_doy;
  if (nchunk > ndoy) nchunk = ndoy;
  if (ngrid % 4 == 0 && ((uintptr_t)ts & 15) == 0) {
    const int64_t nsg = (ngrid + 127) / 128;
    int nc4 = (int)((148 * 64 * 2 + nsg - 1) / nsg);
    nc4 = nc4 < by_doy ? by_doy : nc4;
    nc4 = nc4 > ndoy ? ndoy : nc4;
    const int64_t nb4 = ((nsg + EXC_WARPS - 1) / EXC_WARPS) * nc4;
    exceed4_kernel<<<(unsigned)nb4, EXC_WARPS * 32, 0, (cudaStream_t)stream>>>(
        ts, T, ngrid, doy_ptr, doy_tidx, ndoy, nc4, thresh, mask, nvalid);
    return cuda_status();
  }
  const int64_t nblk = ((ncg + EXC_WARPS - 1) / EXC_WARPS) * nchunk;
  exceed_kernel<<<(unsigned)nblk, EXC_WARPS * 32, 0, (cudaStream_t)stream>>>(
      ts, T, ngrid, doy_ptr, doy_tidx, ndoy, nchunk, thresh, mask, nvalid);
  return cuda_status();
}

int xmhw_events_count(const uint32_t* mask, int64_t T, int64_t ngrid, int32_t min_duration,
                      int32_t join_gaps, int32_t max_gap, int32_t* counts, void* stream) {
  if (!mask || !counts || T <= 0 || ngrid <= 0 || min_duration < 1 || max_gap < 0) return XMHW_E_ARG;
  const int64_t ncg = (ngrid + 31) / 32;
  events_kernel<false><<<(unsigned)((ncg + EVT_WARPS - 1) / EVT_WARPS), EVT_WARPS * 32, 0, (cudaStream_t)stream>>>(
      mask, T, ngrid, min_duration, join_gaps, max_gap, counts, nullptr, 0, nullptr);
  return cuda_status();
}

int xmhw_events_count_stage(const uint32_t* mask, int64_t T, int64_t ngrid, int32_t min_duration,
                            int32_t join_gaps, int32_t max_gap, int32_t* counts, int32_t* stage,
                            int32_t cap_per_cell, int32_t* overflow, void* stream) {
  if (!mask || !counts || !stage || !overflow || T <= 0 || ngrid <= 0 || min_duration < 1 || max_gap < 0 ||
      cap_per_cell < 1)
    return XMHW_E_ARG;
  const int64_t ncg = (ngrid + 31) / 32;
  events_stage_kernel<<<(unsigned)((ncg + EVT_WARPS - 1) / EVT_WARPS), EVT_WARPS * 32, 0, (cudaStream_t)stream>>>(
      mask, T, ngrid, min_duration, join_gaps, max_gap, counts, stage, cap_per_cell, overflow);
  return cuda_status();
}

int xmhw_events_gather(const int32_t* stage, int32_t cap_per_cell, const int32_t* counts, const int64_t* offsets,
                       int64_t ngrid, int64_t cap, int32_t* ev_i32, void* stream) {
  if (!stage || !counts || !offsets || !ev_i32 || ngrid <= 0 || cap <= 0 || cap_per_cell < 1) return XMHW_E_ARG;
  const int64_t ncg = (ngrid + 31) / 32;
  events_gather_kernel<<<(unsigned)((ncg + EVT_WARPS - 1) / EVT_WARPS), EVT_WARPS * 32, 0, (cudaStream_t)stream>>>(
      stage, cap_per_cell, counts, offsets, ngrid, cap, ev_i32);
  return cuda_status();
}

int xmhw_detect_fused_f32(const float* ts, int64_t T, int64_t ngrid, const int32_t* doy, int32_t ndoy,
                          const double* thresh, const double* clim_cm, int32_t min_duration, int32_t join_gaps,
                          int32_t max_gap, int32_t* counts, int32_t* nvalid, int32_t* stage_i32, double* stage_f64,
                          int64_t stage_cap, int32_t* counter, void* stream) {
  if (!ts || !doy || !thresh || !clim_cm || !counts || !nvalid || !stage_i32 || !stage_f64 || !counter || T <= 0 ||
      ngrid <= 0 || ndoy <= 0 || min_duration < 1 || max_gap < 0 || stage_cap <= 0 || ((uintptr_t)clim_cm & 15) ||
      ngrid > 0xffffffffll || T > 0x7fffffffll)
    return XMHW_E_ARG;
  // queue: what waits below the flush level plus what one chunk can emit (per lane at most one event
  // per min_duration + 1 steps, + 2 for the chunk edges)
  const int per_lane = (32 * DF_WARPS) / (min_duration + 1) + 2;
  const size_t qcap = (size_t)DF_QFLUSH + 32u * per_lane;
  const size_t smem = (size_t)ndoy * 32 * 4 + DF_WARPS * 32 * 4 + 36 * 4 + qcap * sizeof(int4) + 16;
  if (smem > 227 * 1024) return XMHW_E_SMEM;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(counter, 0, 2 * sizeof(int32_t), st);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(detect_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int64_t ncg = (ngrid + 31) / 32;
  detect_fused_kernel<<<(unsigned)ncg, DF_WARPS * 32, smem, st>>>(ts, T, ngrid, doy, ndoy, thresh,
                                                                  (const double2*)clim_cm, min_duration, join_gaps, max_gap,
                                                                  counts, nvalid, stage_i32, stage_f64, stage_cap, counter);
  return cuda_status();
}

int xmhw_events_scatter(const int32_t* stage_i32, const double* stage_f64, int64_t stage_cap, int64_t nstaged,
                        const int64_t* offsets, int64_t cap, int32_t* ev_i32, double* ev_f64, void* stream) {
  if (!stage_i32 || !stage_f64 || !offsets || !ev_i32 || !ev_f64 || stage_cap <= 0 || nstaged < 0 || nstaged > stage_cap ||
      cap < nstaged)
    return XMHW_E_ARG;
  if (nstaged == 0) return 0;
  events_scatter_kernel<<<(unsigned)((nstaged + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      stage_i32, stage_f64, stage_cap, nstaged, offsets, cap, ev_i32, ev_f64);
  return cuda_status();
}

int xmhw_exclusive_scan_i32(const int32_t* counts, int64_t n, int64_t* offsets, int64_t* scratch, void* stream) {
  if (!counts || !offsets || !scratch || n <= 0) return XMHW_E_ARG;
  const int64_t nb = (n + SCAN_ITEMS - 1) / SCAN_ITEMS;
  cudaStream_t st = (cudaStream_t)stream;
  scan_block_sums<<<(unsigned)nb, 256, 0, st>>>(counts, n, scratch);
  scan_of_sums<<<1, 32, 0, st>>>(scratch, nb);
  scan_apply<<<(unsigned)nb, 32, 0, st>>>(counts, n, scratch, nb, offsets);
  return cuda_status();
}

int xmhw_events_fill(const uint32_t* mask, int64_t T, int64_t ngrid, int32_t min_duration, int32_t join_gaps,
                     int32_t max_gap, const int64_t* offsets, int64_t cap, int32_t* ev_i32, void* stream) {
  if (!mask || !offsets || !ev_i32 || T <= 0 || ngrid <= 0 || cap <= 0 || min_duration < 1 || max_gap < 0)
    return XMHW_E_ARG;
  const int64_t ncg = (ngrid + 31) / 32;
  events_kernel<true><<<(unsigned)((ncg + EVT_WARPS - 1) / EVT_WARPS), EVT_WARPS * 32, 0, (cudaStream_t)stream>>>(
      mask, T, ngrid, min_duration, join_gaps, max_gap, nullptr, offsets, cap, ev_i32);
  return cuda_status();
}

int xmhw_event_stats_f32(const float* ts, int64_t T, int64_t ngrid, const int32_t* doy, const double* thresh,
                         const double* seas, int64_t nev, int64_t cap, int32_t* ev_i32, double* ev_f64,
                         void* stream) {
  if (!ts || !doy || !thresh || !seas || !ev_i32 || !ev_f64 || T <= 0 || ngrid <= 0 || nev < 0 || cap < nev)
    return XMHW_E_ARG;
  if (nev == 0) return 0;
  const int nt = 128;
  event_stats_kernel<4, 4><<<(unsigned)((nev + nt - 1) / nt), nt, 0, (cudaStream_t)stream>>>(
      ts, T, ngrid, doy, thresh, seas, nev, cap, ev_i32, ev_f64);
  return cuda_status();
}

int xmhw_clim_cellmajor_f64(const double* thresh, const double* seas, int32_t ndoy, int64_t ngrid, double* clim_cm,
                            void* stream) {
  if (!thresh || !seas || !clim_cm || ndoy <= 0 || ngrid <= 0 || ((uintptr_t)clim_cm & 15)) return XMHW_E_ARG;
  dim3 grid((unsigned)((ngrid + 31) / 32), (unsigned)((ndoy + 31) / 32));
  clim_cellmajor_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(thresh, seas, ndoy, ngrid, (double2*)clim_cm);
  return cuda_status();
}

int xmhw_event_stats_cm_f32(const float* ts, int64_t T, int64_t ngrid, const int32_t* doy, int32_t ndoy,
                            const double* clim_cm, int64_t nev, int64_t cap, int32_t* ev_i32, double* ev_f64,
                            void* stream) {
  if (!ts || !doy || !clim_cm || !ev_i32 || !ev_f64 || T <= 0 || ngrid <= 0 || ndoy <= 0 || nev < 0 || cap < nev ||
      ((uintptr_t)clim_cm & 15))
    return XMHW_E_ARG;
  if (nev == 0) return 0;
  const int nt = 128;
  // 8 days of loads in flight per thread, 4 blocks/SM: measured best of {4,6,8,12} x {3..6 blocks}
  event_stats_cm_kernel<8, 4><<<(unsigned)((nev + nt - 1) / nt), nt, 0, (cudaStream_t)stream>>>(
      ts, T, ngrid, doy, ndoy, (const double2*)clim_cm, nev, cap, ev_i32, ev_f64);
  return cuda_status();
}

int xmhw_intermediate_f32(const float* ts, int64_t T, int64_t ngrid, const int32_t* doy, const double* thresh,
                          const double* seas, const int32_t* ev_i32, int64_t nev, int64_t cap,
                          const xmhw_intermediate* out, void* stream) {
  if (!ts || !doy || !thresh || !seas || !out || !out->events || T <= 0 || ngrid <= 0 || nev < 0 ||
      (nev > 0 && !ev_i32) || T > 65535)
    return XMHW_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  // events is pre-filled with NaN by the caller; labels first, then the elementwise fields
  if (nev) event_labels_kernel<<<(unsigned)((nev + 127) / 128), 128, 0, st>>>(ev_i32, nev, cap, ngrid, out->events);
  dim3 grid((unsigned)((ngrid + 127) / 128), (unsigned)T);
  intermediate_kernel<<<grid, 128, 0, st>>>(ts, T, ngrid, doy, thresh, seas, out->events, *out);
  return cuda_status();
}

int xmhw_block_average(const int32_t* ev_i32, const double* ev_f64, int64_t cap, const int64_t* offsets, int64_t ngrid,
                       const int32_t* block_of_t, int32_t nblocks, int32_t use_peak, double* out, void* stream) {
  if (!ev_i32 || !ev_f64 || !offsets || !block_of_t || !out || cap <= 0 || ngrid <= 0 || nblocks <= 0) return XMHW_E_ARG;
  const int nt = 128;
  block_average_kernel<<<(unsigned)((ngrid + nt - 1) / nt), nt, 0, (cudaStream_t)stream>>>(
      ev_i32, ev_f64, cap, offsets, ngrid, block_of_t, nblocks, use_peak ? EI_PEAK : EI_START, out);
  return cuda_status();
}

int xmhw_block_ts_f32(const float* ts, int64_t T, int64_t ngrid, const int32_t* block_of_t, int32_t nblocks, double* out,
                      void* stream) {
  if (!ts || !block_of_t || !out || T <= 0 || ngrid <= 0 || nblocks <= 0) return XMHW_E_ARG;
  const int nt = 128;
  block_ts_kernel<<<(unsigned)((ngrid + nt - 1) / nt), nt, 0, (cudaStream_t)stream>>>(ts, T, ngrid, block_of_t, nblocks, out);
  return cuda_status();
}

int xmhw_block_cat_days_f32(const float* ts, int64_t T, int64_t ngrid, const int32_t* doy, const double* thresh,
                            const double* seas, const int32_t* ev_i32, int64_t nev, int64_t cap,
                            const int32_t* block_of_t, int32_t nblocks, int32_t* days, void* stream) {
  if (!ts || !doy || !thresh || !seas || !block_of_t || !days || T <= 0 || ngrid <= 0 || nblocks <= 0 || nev < 0 ||
      (nev > 0 && !ev_i32) || cap < nev)
    return XMHW_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(days, 0, (size_t)4 * nblocks * ngrid * sizeof(int32_t), st);
  if (e != cudaSuccess) return (int)e;
  if (nev == 0) return 0;
  block_cat_days_kernel<<<(unsigned)((nev + 127) / 128), 128, 0, st>>>(ts, ngrid, doy, thresh, seas, ev_i32, nev, cap,
                                                                        block_of_t, nblocks, days);
  return cuda_status();
}

int xmhw_event_rank_f64(const double* col, const int32_t* cells, const int64_t* offsets, int64_t nev, double* rank,
                        void* stream) {
  if (!col || !cells || !offsets || !rank || nev < 0) return XMHW_E_ARG;
  if (nev == 0) return 0;
  event_rank_kernel<<<(unsigned)((nev + 127) / 128), 128, 0, (cudaStream_t)stream>>>(col, cells, offsets, nev, rank);
  return cuda_status();
}

int xmhw_count_valid_f32(const float* ts, int64_t T, int64_t ngrid, int32_t* nvalid, void* stream) {
  if (!ts || !nvalid || T <= 0 || ngrid <= 0) return XMHW_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(nvalid, 0, (size_t)ngrid * sizeof(int32_t), st);
  if (e != cudaSuccess) return (int)e;
  const int nt = 256;
  const int64_t nbx = (ngrid + nt - 1) / nt;
  int ny = (int)((148 * 8 + nbx - 1) / nbx);                 // enough blocks to fill the machine on narrow grids
  ny = ny < 1 ? 1 : (ny > 64 ? 64 : ny);
  const int tchunk = (int)((T + ny - 1) / ny);
  count_valid_kernel<<<dim3((unsigned)nbx, (unsigned)ny), nt, 0, st>>>(ts, T, ngrid, tchunk, nvalid);
  return cuda_status();
}

int xmhw_group_order_f32(const float* ts, int64_t T, int64_t ngrid, uint8_t* flags, int32_t* order, void* stream) {
  if (!ts || !flags || !order || T <= 0 || ngrid <= 0 || ngrid > 0x3fffffffll * 32) return XMHW_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t ncg = (ngrid + 31) / 32;
  group_flags_kernel<<<(unsigned)((ncg * 32 + 255) / 256), 256, 0, st>>>(ts, T, ngrid, flags);
  group_order_kernel<<<1, 1024, 0, st>>>(flags, ncg, order);
  return cuda_status();
}

int xmhw_interp_gaps_f32(float* ts, int64_t T, int64_t ngrid, int32_t max_pad, void* stream) {
  if (!ts || T <= 0 || ngrid <= 0 || max_pad < 0) return XMHW_E_ARG;
  if (max_pad == 0) return 0;
  const int nt = 128;
  interp_gaps_kernel<<<(unsigned)((ngrid + nt - 1) / nt), nt, 0, (cudaStream_t)stream>>>(ts, T, ngrid, max_pad);
  return cuda_status();
}

int xmhw_copy2d_async(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t width_bytes,
                      int64_t height, int32_t kind, void* stream) {
  if (!dst || !src || width_bytes <= 0 || height <= 0 || dst_pitch < width_bytes || src_pitch < width_bytes)
    return XMHW_E_ARG;
  const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost
                                                                           : cudaMemcpyDeviceToDevice;
  cudaError_t e = cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)width_bytes,
                                    (size_t)height, k, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : (int)e;
}

int xmhw_synth_sst_f32(float* ts, int64_t T, int64_t ngrid, int64_t cell0, const uint8_t* land,
                       const double* season, uint64_t seed, double rho, double sigma, double noise_scale,
                       uint32_t nan_per_million, uint32_t coherent, void* stream) {
  if (!ts || !season || T <= 0 || ngrid <= 0 || coherent < 1) return XMHW_E_ARG;
  const int nt = 128;
  synth_kernel<<<(unsigned)((ngrid + nt - 1) / nt), nt, 0, (cudaStream_t)stream>>>(
      ts, T, ngrid, cell0, land, season, seed, rho, sigma, noise_scale, nan_per_million, coherent);
  return cuda_status();
}

}  // extern "C"
