"""Host-side planning for the climatology sweep kernel.

The reference pools, for every day-of-year d, the samples ts[t+k] with
doy[t] == d and |k| <= windowHalfWidth (xmhw/identify.py:184-209 window_roll,
then groupby("doy") at identify.py:233/263).  Windows of consecutive doys share
all but one "column" of samples, so the GPU kernel sorts each column once and
slides over doy.  This module derives that structure from the ACTUAL doy
vector, which makes leap days (doy 60 only exists in leap years,
identify.py:73-76), truncated windows at the series edges, `tstep=True`
calendars (identify.py:58-71) and any half-width one uniform case:

* the *signature* of time row t' is the set of doys whose window contains it,
  {doy[t] : |t - t'| <= w};
* rows with equal signature form a *class*; a class is cut into *instances*
  of at most 32 rows (one register-resident sorting network each) and into
  separate visits when its doys are not contiguous in sweep order (the year
  wrap-around);
* the window of doy d is the disjoint union of the instances whose signature
  contains d.

The plan also carries numpy's 'linear' quantile index table
(floor((n-1) q), fractional part) for every possible sample count n, computed
here in float64 exactly as numpy does, and a static shared-memory allocation
for the instances (the kernel never allocates).
"""
from dataclasses import dataclass

import numpy as np

MAX_LIST = 32          # rows per instance = keys per register sorting network (short series)
MAX_LIST_LONG = 48     # ... when a calendar day has more than 32 samples (series > 32 years)
META_ROWS = 3          # meta word, cinc, cexc per instance and lane
NULL_ROWS = 3          # block 0 of the pool is the kernel's null list
SCRATCH_HEAD = 2       # scratch rows per instance before its tail keys: f64 sum (lo, hi)
STAGE_ROWS = 2         # two rows after the pool: staged bases of the lists in use
MAX_LISTS = 64
LOAD_FLAG = 1 << 30
MAX_GAP_STEPS = 2      # a class stays resident across holes of up to this many sweep steps


@dataclass
class ClimPlanHost:
    nsteps: int
    pool_rows: int
    nmax: int
    max_size: int
    scratch_rows: int
    inst_base: np.ndarray
    inst_size: np.ndarray
    inst_keep: np.ndarray
    inst_sbase: np.ndarray
    inst_row_off: np.ndarray
    rows: np.ndarray
    leave_off: np.ndarray
    leave: np.ndarray
    enter_off: np.ndarray
    enter: np.ndarray
    use_off: np.ndarray
    use: np.ndarray
    step_rec: np.ndarray
    q: float
    # diagnostics
    n_instances: int = 0
    n_loads: int = 0
    rows_loaded: int = 0
    max_lists: int = 0

    def smem_bytes(self):
        return (self.pool_rows + STAGE_ROWS) * 128


def quantile_table(nmax, q):
    """numpy 2.x `method="linear"`: virtual index v = (n-1)*q evaluated in float64
    (numpy/lib/_function_base_impl.py `_QuantileMethods['linear']`), previous index
    floor(v), gamma = v - floor(v); v >= n-1 selects the maximum (`_get_indexes`)."""
    n = np.arange(nmax + 1, dtype=np.int64)
    v = (n - 1) * np.float64(q)
    lo = np.floor(v)
    gamma = v - lo
    above = v >= (n - 1)
    lo = np.where(above, n - 1, lo)
    gamma = np.where(above, 0.0, gamma)
    lo = np.where(n == 0, 0, lo)
    return lo.astype(np.int32), gamma.astype(np.float64)


def default_keep():
    """Key rows per list kept in shared memory (the rest is re-derived on demand)."""
    import os
    return int(os.environ.get("XMHW_B200_KEEP", "6"))


# pool rows that let N = 32, 31, ... 1 single-warp blocks share one SM's 227 KB of shared
# memory (1 KB per block is reserved by the system, 2 rows per pool are staging rows)
POOL_ROW_STEPS = tuple((227 * 1024 // nw - 1024) // 128 - STAGE_ROWS for nw in range(32, 0, -1))


def default_pool_rows():
    """Row budget of the per-warp pool from the environment (0 = choose automatically)."""
    import os
    return int(os.environ.get("XMHW_B200_POOL_ROWS", "0"))


def build_clim_plan(doy, ndoy, w, q, keep=None, max_rows=None):
    """doy: int array [T] of 1-based labels in 1..ndoy; w: window half width; q in [0,1]."""
    keep = default_keep() if keep is None else int(keep)
    if not 1 <= keep <= MAX_LIST:
        raise ValueError("keep must be in 1..32")
    doy = np.asarray(doy, dtype=np.int64)
    T = len(doy)
    if T == 0:
        raise ValueError("empty time axis")
    if doy.min() < 1 or doy.max() > ndoy:
        raise ValueError("doy labels must lie in 1..ndoy")
    if w < 0:
        raise ValueError("windowHalfWidth must be >= 0")
    # signature of every row
    classes = {}
    for tp in range(T):
        lo, hi = max(0, tp - w), min(T, tp + w + 1)
        seg = doy[lo:hi]
        sig = tuple(sorted(set(seg.tolist())))
        if len(sig) != hi - lo:
            raise NotImplementedError(
                "window (2*windowHalfWidth+1) spans repeated day-of-year labels; "
                "windows wider than one year are not supported")
        classes.setdefault(sig, []).append(tp)

    # instances: (rows, sorted step list) split into visits and into <= max_list pieces; one
    # 48-key list per calendar day (wider register sorting network) beats two 32-key lists
    max_list = MAX_LIST if max(len(r) for r in classes.values()) <= MAX_LIST else MAX_LIST_LONG
    insts = []   # dict(rows=array, steps=list)
    for sig, rws in classes.items():
        steps = [d - 1 for d in sig]
        visits = [[steps[0]]]
        for s in steps[1:]:
            if s - visits[-1][-1] <= MAX_GAP_STEPS:
                visits[-1].append(s)
            else:
                visits.append([s])
        rws = np.asarray(rws, np.int32)
        npieces = -(-len(rws) // max_list)
        pieces = np.array_split(rws, npieces)
        for v in visits:
            for pc in pieces:
                insts.append({"rows": pc, "steps": v})
    insts.sort(key=lambda i: (i["steps"][0], int(i["rows"][0])))
    ninst = len(insts)

    by_first = [[] for _ in range(ndoy)]
    in_use = [[] for _ in range(ndoy)]
    for i, it in enumerate(insts):
        by_first[it["steps"][0]].append(i)
        for s in it["steps"]:
            in_use[s].append(i)

    sizes_list = [len(it["rows"]) for it in insts]
    keeps = [min(sz, keep) for sz in sizes_list]
    budget = default_pool_rows() if max_rows is None else int(max_rows)
    if budget <= 0:
        # occupancy step that holds a REGULAR step with full `keep`; rare peaks (the split
        # leap / non-leap lists around Feb 29) are squeezed into it by shrinking their keeps
        alive = np.zeros(ndoy, np.int64)
        for i, it in enumerate(insts):
            alive[it["steps"][0]:it["steps"][-1] + 1] += keeps[i] + META_ROWS
        regular = int(np.median(alive)) + NULL_ROWS
        budget = next((b for b in POOL_ROW_STEPS if b >= regular), POOL_ROW_STEPS[-1])

    def allocate(keeps):
        """Static first-fit allocation of the instance blocks over the sweep."""
        free = [(NULL_ROWS, 1 << 30)]
        base = np.zeros(ninst, np.int32)
        state = {"rows": 0, "peak_step": 0}
        release_at = [[] for _ in range(ndoy + 1)]

        def alloc(n, s):
            for k, (a, sz) in enumerate(free):
                if sz >= n:
                    if sz == n:
                        free.pop(k)
                    else:
                        free[k] = (a + n, sz - n)
                    if a + n > state["rows"]:
                        state["rows"], state["peak_step"] = a + n, s
                    return a
            raise RuntimeError("pool exhausted")

        def release(a, n):
            free.append((a, n))
            free.sort()
            merged = []
            for seg in free:
                if merged and merged[-1][0] + merged[-1][1] == seg[0]:
                    merged[-1] = (merged[-1][0], merged[-1][1] + seg[1])
                else:
                    merged.append(seg)
            free[:] = merged

        leave_off, leave, enter_off, enter, use_off, use = [0], [], [0], [], [0], []
        prev_use = set()
        loaded = set()
        n_loads = rows_loaded = max_lists = 0
        for s in range(ndoy):
            for (a, n) in release_at[s]:
                release(a, n)
            cur = in_use[s]
            cur_set = set(cur)
            for i in sorted(prev_use - cur_set):
                leave.append(int(base[i]))
            for i in cur:
                if i in prev_use:
                    continue
                if i not in loaded:
                    n = keeps[i] + META_ROWS
                    base[i] = alloc(n, s)
                    release_at[insts[i]["steps"][-1] + 1].append((int(base[i]), n))
                    loaded.add(i)
                    enter.append(i | LOAD_FLAG)
                    n_loads += 1
                    rows_loaded += sizes_list[i]
                else:
                    enter.append(i)
            for i in cur:
                use.append(int(base[i]))
            max_lists = max(max_lists, len(cur))
            leave_off.append(len(leave))
            enter_off.append(len(enter))
            use_off.append(len(use))
            prev_use = cur_set
        return (base, state["rows"], state["peak_step"], leave_off, leave, enter_off, enter, use_off, use,
                n_loads, rows_loaded, max_lists)

    # shrink the lists alive at the peak step until the pool fits the row budget
    for _ in range(4096):
        (base, pool_rows, peak_step, leave_off, leave, enter_off, enter, use_off, use,
         n_loads, rows_loaded, max_lists) = allocate(keeps)
        if pool_rows <= budget:
            break
        lo_s, hi_s = max(0, peak_step - 1), min(ndoy - 1, peak_step + 1)
        live = sorted({i for s in range(lo_s, hi_s + 1) for i in in_use[s]}, key=lambda i: -keeps[i])
        shrunk = False
        for i in live[:max(1, len(live) // 2)]:
            if keeps[i] > 1:
                keeps[i] -= 1
                shrunk = True
        if not shrunk:
            break

    if max_lists > MAX_LISTS:
        raise NotImplementedError("more than %d sorted lists per window (windowHalfWidth too large)" % MAX_LISTS)
    # global scratch rows for the sorted keys past `keep` (same lifetime as the pool block)
    sfree = [(SCRATCH_HEAD, 1 << 30)]       # rows 0,1 = sum of the null list
    sbase = np.zeros(ninst, np.int32)
    scratch_rows = SCRATCH_HEAD
    srel = [[] for _ in range(ndoy + 1)]
    for s in range(ndoy):
        for (a, n) in srel[s]:
            sfree.append((a, n))
            sfree.sort()
            merged = []
            for seg in sfree:
                if merged and merged[-1][0] + merged[-1][1] == seg[0]:
                    merged[-1] = (merged[-1][0], merged[-1][1] + seg[1])
                else:
                    merged.append(seg)
            sfree[:] = merged
        for e in enter[enter_off[s]:enter_off[s + 1]]:
            if not e & LOAD_FLAG:
                continue
            i = e & (LOAD_FLAG - 1)
            n = SCRATCH_HEAD + max(0, sizes_list[i] - keeps[i])
            for k, (a, sz) in enumerate(sfree):
                if sz >= n:
                    sfree[k] = (a + n, sz - n)
                    sbase[i] = a
                    scratch_rows = max(scratch_rows, a + n)
                    break
            srel[insts[i]["steps"][-1] + 1].append((int(sbase[i]), n))
    if scratch_rows >= (1 << 14):
        raise NotImplementedError("scratch too large for the pool meta word")
    sizes = np.array([len(it["rows"]) for it in insts], np.int32)
    row_off = np.concatenate(([0], np.cumsum(sizes)[:-1])).astype(np.int32)
    rows = np.concatenate([it["rows"] for it in insts]).astype(np.int32)
    nmax = max(1, max(sum(int(sizes[i]) for i in in_use[s]) for s in range(ndoy)))
    # fixed-size step records (csrc/xmhw_lane.h STEP_*)
    rec = np.zeros((ndoy, 32), np.int64)
    keeps_arr = np.asarray(keeps, np.int64)
    for s in range(ndoy):
        nl = leave_off[s + 1] - leave_off[s]
        ne = enter_off[s + 1] - enter_off[s]
        nu = use_off[s + 1] - use_off[s]
        ovf = nl > 4 or ne > 4
        rec[s, 0] = (0 if ovf else nl) | ((0 if ovf else ne) << 8) | (nu << 16) | ((1 << 31) if ovf else 0)
        rec[s, 1] = use_off[s]
        if not ovf:
            for j in range(nl):
                rec[s, 2 + j] = leave[leave_off[s] + j]
            for j in range(ne):
                e = enter[enter_off[s] + j]
                i = e & (LOAD_FLAG - 1)
                rec[s, 6 + 3 * j] = e
                rec[s, 7 + 3 * j] = int(base[i]) | (int(sizes[i]) << 16) | (int(keeps_arr[i]) << 24)
                rec[s, 8 + 3 * j] = int(sbase[i])
        if s + 1 < ndoy:
            rec[s, 18] = use_off[s + 1]
            rec[s, 19] = use_off[s + 2] - use_off[s + 1]
        rec[s, 20] = enter_off[s]
        if not ovf:
            # the list to prefetch after each entering load: (row offset, size) of the next load entry
            for j in range(ne):
                g = enter_off[s] + j + 1
                while g < len(enter) and not enter[g] & LOAD_FLAG:
                    g += 1
                if g < len(enter):
                    i = enter[g] & (LOAD_FLAG - 1)
                    rec[s, 21 + 2 * j] = int(row_off[i])
                    rec[s, 22 + 2 * j] = int(sizes[i])
    step_rec = rec.astype(np.uint32).view(np.int32).reshape(-1)

    def arr(x):
        a = np.asarray(x, np.int32)
        return a if a.size else np.zeros(1, np.int32)

    return ClimPlanHost(
        nsteps=ndoy, pool_rows=int(pool_rows), nmax=int(nmax), max_size=int(sizes.max()), scratch_rows=int(scratch_rows),
        inst_base=base, inst_size=sizes, inst_keep=np.asarray(keeps, np.int32), inst_sbase=sbase,
        inst_row_off=row_off, rows=rows,
        leave_off=arr(leave_off), leave=arr(leave), enter_off=arr(enter_off), enter=arr(enter),
        use_off=arr(use_off), use=arr(use), step_rec=step_rec, q=float(q),
        n_instances=ninst, n_loads=n_loads, rows_loaded=rows_loaded, max_lists=max_lists)


def doy_csr(doy, ndoy):
    """CSR of time indices per doy label: (ptr [ndoy+1], tidx [T]) int32."""
    doy = np.asarray(doy, np.int64)
    order = np.argsort(doy, kind="stable").astype(np.int32)
    counts = np.bincount(doy - 1, minlength=ndoy)
    ptr = np.concatenate(([0], np.cumsum(counts))).astype(np.int32)
    return ptr, order
