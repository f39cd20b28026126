#!/usr/bin/env python
"""Headline benchmark: cell-years/s of threshold + detect on synthetic global 0.25-degree
30-year daily SST (BASELINE.json metric), one process per GPU.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference          # the reference's own pandas detect on the host cores

A "step" is one pass of the hot path (xmhw_b200.core.threshold_arrays + detect_arrays, i.e. what
replaces xmhw.threshold + xmhw.detect) over the rank's cells with the series resident in HBM.
`value` = ocean cells x calendar years / device time (max over ranks).

Scaling.  Default `--scaling strong`: ONE grid is cut into N contiguous ocean-balanced cell ranges
(xmhw_b200.shard / multi), each rank generates and processes only its column block, events carry
global cell ids, and partition-independent checksums of the event table and the climatologies are
summed over the ranks (`result_checksum`: identical for every N, so the driver's N = 1, 2, 4, 8
lines prove the N-rank table equals the 1-rank table).  `--scaling weak` gives every rank its own
realisation of the full grid (secondary mode, labelled in the line).  There is no data-path
collective; the process group is gloo (barriers, timing max, checksum sums) -- NCCL is not used.

`e2e` is the same metric through the host-buffer entry point (pinned host series -> device ->
results back to host, copies inside the timed region); `api_e2e` (1 GPU) is the drop-in public API
(xmhw_b200.xmhw.threshold + detect(compact=True)) on host arrays.  The input of a rank is far
larger than L2 (126 MB), so no L2 flush is needed between iterations.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: dict(grid (nlat, nlon), years, land fraction, calendar, options)
    "global025_30yr": dict(grid=(720, 1440), years=(1982, 2011), land=0.33),      # BASELINE configs[2]: the metric's config
    "regional_40yr": dict(grid=(160, 240), years=(1982, 2021), land=0.0),         # BASELINE configs[1]
    "global025_skipna99": dict(grid=(720, 1440), years=(1982, 2011), land=0.33, nan_ppm=10000, winter_blocks=0.02,
                               threshold=dict(pctile=99)),                       # BASELINE configs[3] (i)
    "global025_pentad": dict(grid=(720, 1440), years=(1982, 2011), land=0.33, pentad=True,
                             threshold=dict(windowHalfWidth=5, smoothPercentileWidth=5, feb29=False),
                             detect=dict(minDuration=3, maxGap=1)),              # BASELINE configs[3] (ii)
    "global010_30yr": dict(grid=(1800, 3600), years=(1982, 2011), land=0.33),     # BASELINE configs[4]: 284 GB, chunked
    "global025_quarter": dict(grid=(180, 1440), years=(1982, 2011), land=0.33),   # development timing
    "quarter_w1": dict(grid=(180, 1440), years=(1982, 2011), land=0.33, threshold=dict(windowHalfWidth=1)),   # development:
    "quarter_w2": dict(grid=(180, 1440), years=(1982, 2011), land=0.33, threshold=dict(windowHalfWidth=2)),   # narrower windows =
    "quarter_w3": dict(grid=(180, 1440), years=(1982, 2011), land=0.33, threshold=dict(windowHalfWidth=3)),   # smaller pools, more warps/SM
    "small": dict(grid=(32, 64), years=(2001, 2010), land=0.2),
}
METRIC = "cell-years/s, threshold+detect, global 0.25deg 30-yr SST"


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_axes(wl):
    """(doy labels, ndoy, T, years) of a workload's time axis."""
    from xmhw_b200 import synth
    y0, y1 = wl["years"]
    if wl.get("pentad"):
        nyr = y1 - y0 + 1
        return np.tile(np.arange(1, 74), nyr), 73, 73 * nyr, nyr, None
    tm = synth.daily_time(y0, y1)
    return synth.doy366(tm), 366, len(tm), y1 - y0 + 1, tm


def static_config(name, wl, T, nocean, ngrid):
    """The part of `config` that both arms print identically (workload identity)."""
    return {"workload": name, "grid": list(wl["grid"]), "years": list(wl["years"]), "T": int(T),
            "grid_cells": int(ngrid), "ocean_cells": int(nocean),
            "options": {k: wl[k] for k in ("nan_ppm", "winter_blocks", "pentad", "threshold", "detect") if k in wl}}


class ClockSampler(threading.Thread):
    """Samples the SM clock and the throttle reasons while the timed region runs (NVML every 20 ms, else
    nvidia-smi every 200 ms)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        if self._run_nvml():
            return
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def _run_nvml(self):
        """The same fields through NVML (nvidia_ml_py), every 20 ms: an nvidia-smi process per sample takes
        ~150 ms, too coarse for a timed region of a few hundred ms.  False when NVML is not usable."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
            get_reasons(h)
        except Exception:
            return False
        bits = ((0x8, 3), (0x40, 4), (0x20, 5), (0x4, 6))      # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap
        while not self.stop_flag:
            try:
                r = int(get_reasons(h))
                row = [str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx), "0", "", "", "", ""]
                for bit, col in bits:
                    row[col] = "Active" if r & bit else "Not Active"
                self.rows.append(row)
            except Exception:
                pass
            time.sleep(0.02)
        return True

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU arm: the reference on the host cores
def _cpu_worker(args):
    """threshold + detect of a slab of cells in one host process.  Detect is the UNMODIFIED reference
    pandas code (identify.mhw_filter / join_gaps, features.mhw_df / mhw_features, one call chain per
    cell exactly like the reference's loop xmhw.py:440-454) when oracle/_ref or /root/reference is
    present; the climatology (xarray glue, not importable here) is the oracle's numpy port."""
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import ref_harness as rh
    from oracle import xmhw_oracle as O
    from xmhw_b200 import synth
    (name, cell0, ncell, use_ref) = args
    wl = WORKLOADS[name]
    doy, ndoy, T, years, tm = workload_axes(wl)
    ts = synth.synth_sst(T, ncell, synth.season_table(tm if tm is not None else T), cell0=cell0,
                         nan_ppm=wl.get("nan_ppm", 0))
    tkw = dict(wl.get("threshold", {}))
    tkw["tstep"] = not tkw.pop("feb29", True)
    dkw = wl.get("detect", {})
    t0 = time.perf_counter()
    th, se = O.threshold(ts, doy, ndoy, **tkw)
    nev = 0
    if use_ref:
        for c in range(ncell):
            df = rh.ref_define_events(ts[:, c], th[doy - 1, c], se[doy - 1, c], dkw.get("minDuration", 5), True,
                                      dkw.get("maxGap", 2))
            nev += 0 if df is None else len(df)
    else:
        nev = len(O.detect(ts, doy, th, se, dkw.get("minDuration", 5), True, dkw.get("maxGap", 2))["cell"])
    return time.perf_counter() - t0, nev


def cpu_reference_rate(name, cells_per_proc, nproc):
    """cell-years/s of the CPU arm using `nproc` host processes on a bounded sample of ocean cells of
    the workload.  Returns (rate, slowest worker seconds, kind)."""
    import multiprocessing as mp
    from oracle import ref_harness as rh
    use_ref = rh.available()
    years = WORKLOADS[name]["years"]
    jobs = [(name, 1000 + i * cells_per_proc, cells_per_proc, use_ref) for i in range(nproc)]
    with mp.get_context("spawn").Pool(nproc) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = max(r[0] for r in res)            # workers run concurrently; the slowest bounds the rate
    kind = "reference-pandas+glue-port" if use_ref else "port"
    return cells_per_proc * nproc * (years[1] - years[0] + 1) / wall, wall, kind


def cpu_sample_text(kind, ncell, years, per_proc, wall=None):
    what = ("reference pandas detect (unmodified xmhw identify.py / features.py, one call chain per cell) + numpy "
            "port of the xarray climatology glue" if kind.startswith("reference") else
            "oracle numpy port of xmhw threshold+detect")
    t = "" if wall is None else " (%.1f s)" % wall
    return "%d ocean cells x %d yr, %d per host process%s: %s" % (ncell, years, per_proc, t, what)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    doy, ndoy, T, years, _ = workload_axes(wl)
    from xmhw_b200 import synth
    nlat, nlon = wl["grid"]
    ngrid = nlat * nlon
    nocean = ngrid - (int(synth.land_mask(nlat, nlon, wl["land"]).sum()) if wl["land"] else 0)
    cores = os.cpu_count() or 1
    cpp = args.cpu_cells
    rates, walls, kind = [], [], "port"
    for _ in range(args.warmup):
        cpu_reference_rate(args.workload, 1, min(cores, 2))
    for _ in range(args.steps):
        r, w, kind = cpu_reference_rate(args.workload, cpp, cores)
        rates.append(r)
        walls.append(w)
    value = float(np.mean(rates))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "cell-years/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(walls)) * 1e3,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": static_config(args.workload, wl, T, nocean, ngrid),
            "cpu_baseline": {"value": value, "unit": "cell-years/s", "cores": cores, "kind": kind,
                             "sample": cpu_sample_text(kind, cpp * cores, years, cpp)},
            "e2e": {"value": value, "unit": "cell-years/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------- GPU arm
def winter_blocks_(ts, doy, cell0, frac, seed=44):
    """BASELINE config 4(i): `frac` of the cells get a 60-120-day NaN block every winter (in place)."""
    import torch
    T, n = ts.shape
    g = np.random.default_rng(seed + cell0)
    cells = np.flatnonzero(g.random(n) < frac)
    year0 = np.flatnonzero(np.asarray(doy) == 1)
    rows, cols = [], []
    for c in cells:
        start, length = int(g.integers(330, 366)), int(g.integers(60, 121))
        for y0 in np.concatenate((year0, [year0[-1] + 365])):
            a = max(0, y0 + start - 365)
            b = min(T, max(0, y0 + start - 365 + length))
            if b > a:
                rows.append(np.arange(a, b))
                cols.append(np.full(b - a, c))
    if rows:
        r = torch.from_numpy(np.concatenate(rows)).to(ts.device)
        c = torch.from_numpy(np.concatenate(cols)).to(ts.device)
        ts[r, c] = float("nan")
    return ts


def run_ours(args):
    import torch
    import torch.distributed as dist
    from xmhw_b200 import core, multi, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("gloo")          # barriers / timing max / checksum sums only: no data-path collective

    wl = WORKLOADS[args.workload]
    nlat, nlon = wl["grid"]
    doy, ndoy, T, years, tm = workload_axes(wl)
    ngrid = nlat * nlon
    land = synth.land_mask(nlat, nlon, wl["land"]).ravel() if wl["land"] else np.zeros(ngrid, np.uint8)
    ocean = land == 0
    nocean_total = int(ocean.sum())
    season = synth.season_table(tm if tm is not None else T)
    tkw, dkw = wl.get("threshold", {}), wl.get("detect", {})
    strong = args.scaling == "strong"
    a, b = multi.rank_range(ocean, rank, world) if strong else (0, ngrid)
    cell0 = a if strong else rank * ngrid             # seeds follow the GLOBAL cell id
    nloc = b - a
    nocean_loc = int(ocean[a:b].sum())
    # cells per device-resident chunk (config 5 at N = 1 does not fit: 284 GB)
    free_b = torch.cuda.mem_get_info(dev)[0]
    per_cell = T * 4 + 4 * ndoy * 8 + 3 * years * 200 + 2048        # series + raw/smoothed climatologies + events + transients
    chunk = int(min(nloc, max(32 * 1024, (int(free_b * 0.80) // per_cell) // 32 * 32)))
    chunks = [(c0, min(nloc, c0 + chunk)) for c0 in range(0, nloc, chunk)]
    resident = len(chunks) == 1

    def make_block(c0, c1):
        ts = core.synth_sst_device(T, c1 - c0, season, land=land[a + c0:a + c1], cell0=cell0 + c0, device=dev,
                                   nan_ppm=wl.get("nan_ppm", 0))
        if wl.get("winter_blocks"):
            winter_blocks_(ts, doy, cell0 + c0, wl["winter_blocks"])
        return ts

    ts_res = make_block(0, nloc) if resident else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    state = {}

    def step(timers=None):
        nev, sums = 0, {"events": 0, "table": 0, "clim": 0}
        gen_ms = 0.0
        for (c0, c1) in chunks:
            if resident:
                ts = ts_res
            else:                                   # streamed chunk: its generation is NOT part of the hot path
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                ts = make_block(c0, c1)
                g1.record()
                if timers is not None:
                    timers.append((g0, g1))
            th, se = core.threshold_arrays(ts, doy, ndoy, **tkw)
            ev = core.detect_arrays(ts, doy, ndoy, th, se, **dkw)
            if args.checksum or "sums" not in state:
                s = {"events": ev.n, "table": multi.event_checksum(ev, cell0 + c0), "clim": multi.clim_checksum(th, se)}
                sums = {k: (sums[k] + s[k]) & ((1 << 64) - 1) for k in sums}
            nev += ev.n
            del th, se, ev
            if not resident:
                del ts
        if "sums" not in state:
            state["sums"] = sums
        return nev

    for _ in range(max(1, args.warmup)):
        nev = step()
    # clock ramp: an idle B200 sits at ~0.7 GHz and needs sustained load to reach its 1.97 GHz; with short steps
    # (N = 8: 9 ms) the W warm-up steps -- the first one mostly host-side plan building -- are over before that,
    # so ~0.25 s of further untimed steps follow (their number from the duration of one synchronised step)
    torch.cuda.synchronize()
    t_one = time.perf_counter()
    nev = step()
    torch.cuda.synchronize()
    t_one = max(1e-4, time.perf_counter() - t_one)
    extra = 1 + min(64, int(0.25 / t_one))
    for _ in range(extra - 1):
        nev = step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    core.TRACE = []
    core.LAUNCHES["n"] = 0
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    gen_timers = []
    barrier()
    e0.record()
    for _ in range(args.steps):
        nev = step(gen_timers)
    e1.record()
    barrier()
    sampler.stop_flag = True
    ms = e0.elapsed_time(e1) - sum(g0.elapsed_time(g1) for g0, g1 in gen_timers)
    trace = core.TRACE
    core.TRACE = None
    launches = core.LAUNCHES["n"]
    per_kernel = {}
    for name, x0, x1 in trace:
        per_kernel.setdefault(name, []).append(x0.elapsed_time(x1))
    kernel_ms = {k: float(np.sum(v)) / args.steps for k, v in per_kernel.items()}      # per step (all chunks)

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([x], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ms_step = allmax(ms) / args.steps
    nocean_all = nocean_total if strong else world * nocean_total
    value = nocean_all * years / (ms_step * 1e-3)
    nev_all = int(allsum(nev))
    sums = multi.combine_checksums(state["sums"])

    # roofline of the dominant kernel (climatology sweep): algorithmic bytes / measured duration, this rank
    peak, peak_src = hbm_peak()
    sweep_name = "xmhw_clim_sweep2_f32" if "xmhw_clim_sweep2_f32" in kernel_ms else "xmhw_clim_sweep_f32"
    sweep_ms = kernel_ms[sweep_name]
    sweep_kernel = "clim_sweep_kernel"
    if sweep_name.endswith("sweep2_f32"):
        dp2 = core.device_plan2(doy, ndoy, tkw.get("windowHalfWidth", 5), tkw.get("pctile", 90) / 100.0, dev)
        sweep_kernel = core.sweep2_kernel_name(dp2.host) if dp2 is not None else "clim_sweep2_kernel"
    sweep_bytes = nloc * T * 4 + nocean_loc * 2 * ndoy * 8
    ach = sweep_bytes / (sweep_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.isfile(tp):
        traffic = json.load(open(tp)).get(args.workload, {}).get(sweep_name, {}).get("dram_bytes")
    b_alg = nloc * T * 4 + nocean_loc * 2 * ndoy * 8 + nocean_loc * 4 + nev * 180
    ms_rank = ms / args.steps

    # end-to-end through the host-buffer entry point (rank-local block, inputs in pinned host memory)
    e2e = None
    need_host = T * nloc * 4 + 2 * ndoy * nloc * 8 + int(nev * 1.05) * (core.EI_COUNT * 4 + core.EF_COUNT * 8)
    try:
        import psutil
        avail = psutil.virtual_memory().available / max(1, world)
    except Exception:
        avail = float("inf")
    if args.no_e2e or not resident:
        e2e = {"value": None, "unit": "cell-years/s", "skipped": "disabled" if args.no_e2e else "chunked workload"}
    elif avail < need_host * 1.15:
        e2e = {"value": None, "unit": "cell-years/s", "skipped": "host memory: %.0f GB available per rank, "
               "%.0f GB of pinned buffers needed" % (avail / 1e9, need_host / 1e9)}
    else:
        host = torch.empty((T, nloc), dtype=torch.float32, pin_memory=True)
        host.copy_(ts_res)
        torch.cuda.synchronize()
        ts_res = None
        torch.cuda.empty_cache()
        cap = int(nev * 1.05) + 1024
        out = {"thresh": torch.empty((ndoy, nloc), dtype=torch.float64, pin_memory=True),
               "seas": torch.empty((ndoy, nloc), dtype=torch.float64, pin_memory=True),
               "nvalid": torch.empty(nloc, dtype=torch.int32, pin_memory=True),
               "ev_i32": torch.empty((core.EI_COUNT, cap), dtype=torch.int32, pin_memory=True),
               "ev_f64": torch.empty((core.EF_COUNT, cap), dtype=torch.float64, pin_memory=True)}
        kw = dict(tkw)
        kw.update(dkw)
        core.threshold_detect_host(host, doy, ndoy, device=dev, out=out, slabs=args.slabs, **kw)      # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            res = core.threshold_detect_host(host, doy, ndoy, device=dev, out=out, slabs=args.slabs, **kw)
        barrier()
        dt = allmax((time.perf_counter() - t0) / args.e2e_steps)
        dts = torch.empty((T, nloc), dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        th0 = time.perf_counter()
        dts.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        h2d_ms = (time.perf_counter() - th0) * 1e3
        del dts
        e2e = {"value": nocean_all * years / dt, "unit": "cell-years/s", "h2d_only_ms": h2d_ms,
               "h2d_gbs_this_rank": T * nloc * 4 / (h2d_ms * 1e-3) / 1e9,
               "h2d_bytes_per_step": int(res["h2d_bytes"]), "d2h_bytes_per_step": int(res["d2h_bytes"]),
               "ms_per_step": dt * 1e3, "note": "host<->device copies of this rank's block; the aggregate H2D rate of "
               "the host (all ranks share its memory bandwidth) is the e2e limiter at N > 1"}
        api = None
        if world == 1 and not args.no_api:
            api = api_e2e(host, doy, tm, wl, nlat, nlon, nocean_total, years, profile=args.api_profile)
        del host, out, res
        if api is not None:
            e2e["api"] = api

    cpu = None
    if rank == 0 and not args.no_cpu:
        cores = os.cpu_count() or 1
        r, w, kind = cpu_reference_rate(args.workload, args.cpu_cells, cores)
        cpu = {"value": r, "unit": "cell-years/s", "cores": cores, "kind": kind,
               "sample": cpu_sample_text(kind, args.cpu_cells * cores, years, args.cpu_cells, w)}
    if rank == 0:
        cfg = static_config(args.workload, wl, T, nocean_total, ngrid)
        line = {"metric": METRIC, "value": value, "unit": "cell-years/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "warmup_extra_steps": extra, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f32 keys / f64 statistics", "data": "synthetic",
                "config": cfg,
                "partition": {"mode": "one grid cut into ocean-balanced contiguous cell ranges, one per rank" if strong
                              else "every rank its own realisation of the grid (weak scaling)",
                              "rank0_cells": [int(a), int(b)], "rank0_ocean_cells": nocean_loc,
                              "chunks_per_rank": len(chunks), "collective": "none (gloo: barrier, timing max, checksum sums)",
                              "l2": "input (%.1f GB per rank) >> L2, no flush needed" % (nloc * T * 4 / 1e9)},
                "events": nev_all,
                "result_checksum": {"events": int(sums["events"]), "table": "%016x" % sums["table"],
                                    "clim": "%016x" % sums["clim"]},
                "e2e": e2e, "gpu_launches": launches,
                "roofline": {"bound": "hbm", "kernel": sweep_kernel,
                             "achieved": ach, "peak": peak,
                             "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": sweep_bytes // len(chunks), "ms_per_launch": sweep_ms / len(chunks),
                             "whole_step": {"algorithmic_bytes": b_alg, "achieved": b_alg / (ms_rank * 1e-3) / 1e9,
                                            "frac": b_alg / (ms_rank * 1e-3) / 1e9 / peak}},
                "kernel_ms": kernel_ms,
                "cpu_baseline": cpu, "clocks": sampler.summary()}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def api_e2e(host, doy, tm, wl, nlat, nlon, nocean, years, profile=None):
    """The drop-in public API on host arrays: xmhw.threshold + xmhw.detect(compact=True), wall clock.
    profile: path of a cProfile listing of one extra (untimed) pass (development aid)."""
    if tm is None or wl.get("winter_blocks") or wl.get("threshold") or wl.get("detect"):
        return None
    import torch
    from xmhw_b200 import labeled
    from xmhw_b200 import xmhw as api
    T = host.shape[0]
    lat = np.linspace(-89.875, 89.875, nlat)
    lon = np.linspace(0.125, 359.875, nlon)
    da = labeled.DataArray(host.numpy().reshape(T, nlat, nlon), ("time", "lat", "lon"),
                           coords={"time": np.asarray(tm).astype("datetime64[ns]"), "lat": lat, "lon": lon})
    out = {}
    for it in range(2):                              # first pass warms plans / allocator
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        clim = api.threshold(da)
        t1 = time.perf_counter()
        ev = api.detect(da, clim["thresh"], clim["seas"], compact=True)
        t2 = time.perf_counter()
        out = {"value": nocean * years / (t2 - t0), "unit": "cell-years/s", "threshold_s": t1 - t0, "detect_s": t2 - t1,
               "events": int(len(ev["index_start"].values)),
               "what": "xmhw_b200.xmhw.threshold + detect(compact=True) on host arrays (labeled.DataArray), wall clock"}
        del clim, ev
    if profile:
        import cProfile
        import io
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        clim = api.threshold(da)
        ev = api.detect(da, clim["thresh"], clim["seas"], compact=True)
        pr.disable()
        buf = io.StringIO()
        pstats.Stats(pr, stream=buf).sort_stats("tottime").print_stats(45)
        open(profile, "w").write(buf.getvalue())
        del clim, ev
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="global025_30yr", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--cpu-cells", type=int, default=24, help="cells per host process in the CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--slabs", type=int, default=24, help="column blocks of the host-buffer (e2e) pipeline")
    ap.add_argument("--checksum", action="store_true", help="recompute the result checksums in every step (default: first step)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-api", action="store_true")
    ap.add_argument("--api-profile", default=None, help="write a cProfile listing of the public-API pass to this file")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: everything libraries write to file descriptor 1 goes to stderr
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(saved, "w", buffering=1)
    if args.impl == "reference":
        run_reference(args)
    else:
        import __graft_entry__ as g
        g.build()
        run_ours(args)


if __name__ == "__main__":
    main()
