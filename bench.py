#!/usr/bin/env python
"""Headline benchmark: cell-years/s of threshold + detect on synthetic global 0.25-degree
30-year daily SST (BASELINE.json metric), one process per GPU.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference          # the CPU stand-in for the reference (oracle port)

A "step" is one pass of the hot path (xmhw_b200.core.threshold_arrays + detect_arrays, i.e.
what replaces xmhw.threshold + xmhw.detect) over the whole grid with the series resident in
HBM.  `value` = ocean cells x calendar years / device time (max over ranks); `e2e` is the same
metric through the host-buffer entry point (pinned host series -> device -> results back to
host, copies inside the timed region).  The input (45 GB) is far larger than L2 (126 MB), so no
L2 flush is needed between iterations.  Scaling is weak: every rank processes its own
realisation of the full grid (cells are independent; no data-path collective).
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
    # name: (nlat, nlon, first year, last year, land fraction)
    "global025_30yr": (720, 1440, 1982, 2011, 0.33),     # BASELINE configs[2], the config the metric names
    "regional_40yr": (160, 240, 1982, 2021, 0.0),        # BASELINE configs[1]
    "global025_quarter": (180, 1440, 1982, 2011, 0.33),  # a quarter of the global grid (development timing)
    # diagnostic: blocks of 64 cells share mean/amplitude/phase (spatially coherent climatology, as in
    # real SST); NOT the headline -- it shows how much of the sweep is SIMT loss on independent cells
    "global025_30yr_coherent": (720, 1440, 1982, 2011, 0.33),
    "small": (32, 64, 2001, 2010, 0.2),
}
METRIC = "cell-years/s, threshold+detect, global 0.25deg 30-yr SST"


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

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


# --------------------------------------------------------------------------- CPU stand-in for the reference
def _cpu_worker(args):
    """threshold + detect of the oracle on a slab of cells (one host process)."""
    from oracle import xmhw_oracle as O
    from xmhw_b200 import synth
    (y0, y1, cell0, ncell) = args
    tm = synth.daily_time(y0, y1)
    doy = synth.doy366(tm)
    ts = synth.synth_sst(len(tm), ncell, synth.season_table(tm), cell0=cell0)
    t0 = time.perf_counter()
    th, se = O.threshold(ts, doy, 366)
    ev = O.detect(ts, doy, th, se)
    return time.perf_counter() - t0, len(ev["cell"])


def cpu_reference_rate(y0, y1, cells_per_proc, nproc):
    """cell-years/s of the CPU oracle (numpy port of the reference's algorithm) using `nproc`
    host processes, on a bounded sample of the same synthetic workload (ocean cells)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    jobs = [(y0, y1, 1000 + i * cells_per_proc, cells_per_proc) for i in range(nproc)]
    t0 = time.perf_counter()
    with ctx.Pool(nproc) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = max(r[0] for r in res)            # workers run concurrently; slowest one bounds the rate
    ncell = cells_per_proc * nproc
    return ncell * (y1 - y0 + 1) / wall, wall, time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nlat, nlon, y0, y1, _ = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    cpp = args.cpu_cells
    rates, walls = [], []
    for _ in range(args.warmup):
        cpu_reference_rate(y0, y1, 1, min(cores, 2))
    for _ in range(args.steps):
        r, w, _ = cpu_reference_rate(y0, y1, cpp, cores)
        rates.append(r)
        walls.append(w)
    value = float(np.mean(rates))
    sample = "%d ocean cells x %d yr per step (%d per process), oracle numpy port of xmhw threshold+detect" % (
        cpp * cores, y1 - y0 + 1, cpp)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "cell-years/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(walls)) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "grid": [nlat, nlon], "years": [y0, y1]},
            "cpu_baseline": {"value": value, "unit": "cell-years/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "cell-years/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from xmhw_b200 import core, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    nlat, nlon, y0, y1, land_frac = WORKLOADS[args.workload]
    years = y1 - y0 + 1
    tm = synth.daily_time(y0, y1)
    doy = synth.doy366(tm)
    T, ngrid = len(tm), nlat * nlon
    land = synth.land_mask(nlat, nlon, land_frac).ravel() if land_frac else None
    nocean = ngrid - (int(land.sum()) if land is not None else 0)
    season = synth.season_table(tm)
    # every rank generates its own realisation of the grid (weak scaling), seeded by global cell id
    ts = core.synth_sst_device(T, ngrid, season, land=land, cell0=rank * ngrid, device=dev,
                               coherent=64 if args.workload.endswith("_coherent") else 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        th, se = core.threshold_arrays(ts, doy, 366)
        ev = core.detect_arrays(ts, doy, 366, th, se)
        return th, se, ev

    for _ in range(args.warmup):
        th, se, ev = step()
        nev = ev.n
        del th, se, ev
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    core.TRACE = []
    core.LAUNCHES["n"] = 0
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        th, se, ev = step()
        nev = ev.n
        del th, se, ev
    e1.record()
    barrier()
    sampler.stop_flag = True
    ms = e0.elapsed_time(e1)
    trace = core.TRACE
    core.TRACE = None
    launches = core.LAUNCHES["n"]
    per_kernel = {}
    for name, a, b in trace:
        per_kernel.setdefault(name, []).append(a.elapsed_time(b))
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax.item()) / args.steps
    value = world * nocean * years / (ms_step * 1e-3)

    # roofline of the dominant kernel (climatology sweep): algorithmic bytes / measured duration
    peak, peak_src = hbm_peak()
    sweep_name = "xmhw_clim_sweep2_f32" if "xmhw_clim_sweep2_f32" in per_kernel else "xmhw_clim_sweep_f32"
    sweep_ms = float(np.mean(per_kernel[sweep_name]))
    sweep_bytes = ngrid * T * 4 + nocean * 2 * 366 * 8
    ach = sweep_bytes / (sweep_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.isfile(tp):
        traffic = json.load(open(tp)).get(args.workload, {}).get("clim_sweep_dram_bytes")
    b_alg = ngrid * T * 4 + nocean * 2 * 366 * 8 + nocean * 4 + nev * 180

    # end-to-end through the host-buffer entry point (rank-local, inputs in pinned host memory)
    e2e = None
    need_host = T * ngrid * 4 + 2 * 366 * ngrid * 8 + int(nev * 1.05) * (core.EI_COUNT * 4 + core.EF_COUNT * 8)
    try:
        import psutil
        avail = psutil.virtual_memory().available / max(1, world)
    except Exception:
        avail = float("inf")
    if not args.no_e2e and avail < need_host * 1.15:
        e2e = {"value": None, "unit": "cell-years/s", "skipped": "host memory: %.0f GB available per rank, "
               "%.0f GB of pinned buffers needed" % (avail / 1e9, need_host / 1e9)}
    elif not args.no_e2e:
        host = torch.empty((T, ngrid), dtype=torch.float32, pin_memory=True)
        host.copy_(ts)
        torch.cuda.synchronize()
        del ts
        torch.cuda.empty_cache()
        cap = int(nev * 1.05) + 1024
        out = {"thresh": torch.empty((366, ngrid), dtype=torch.float64, pin_memory=True),
               "seas": torch.empty((366, ngrid), dtype=torch.float64, pin_memory=True),
               "nvalid": torch.empty(ngrid, dtype=torch.int32, pin_memory=True),
               "ev_i32": torch.empty((core.EI_COUNT, cap), dtype=torch.int32, pin_memory=True),
               "ev_f64": torch.empty((core.EF_COUNT, cap), dtype=torch.float64, pin_memory=True)}
        core.threshold_detect_host(host, doy, 366, device=dev, out=out, slabs=args.slabs)      # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            res = core.threshold_detect_host(host, doy, 366, device=dev, out=out, slabs=args.slabs)
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / args.e2e_steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        # where the end-to-end time goes: the H2D copy alone, timed once more on its own
        dts = torch.empty((T, ngrid), dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        th0 = time.perf_counter()
        dts.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        h2d_ms = (time.perf_counter() - th0) * 1e3
        del dts
        e2e = {"value": world * nocean * years / float(dt.item()), "unit": "cell-years/s", "h2d_only_ms": h2d_ms,
               "h2d_bytes_per_step": int(res["h2d_bytes"]), "d2h_bytes_per_step": int(res["d2h_bytes"]),
               "ms_per_step": float(dt.item()) * 1e3}
        del host, out, res

    cpu = None
    if rank == 0 and not args.no_cpu:
        cores = os.cpu_count() or 1
        r, w, _ = cpu_reference_rate(y0, y1, args.cpu_cells, cores)
        cpu = {"value": r, "unit": "cell-years/s", "cores": cores, "kind": "port",
               "sample": "%d ocean cells x %d yr, one slab per host process (%.1f s), oracle numpy port of xmhw "
                         "threshold+detect" % (args.cpu_cells * cores, years, w)}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "cell-years/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32 keys / f64 statistics", "data": "synthetic",
                "config": {"workload": args.workload, "grid": [nlat, nlon], "years": [y0, y1], "T": T,
                           "ocean_cells_per_gpu": nocean, "events_per_gpu": nev,
                           "l2": "input (%.1f GB per GPU) >> L2, no flush needed" % (ngrid * T * 4 / 1e9),
                           "parallelism": "cells sharded, no collective"},
                "e2e": e2e, "gpu_launches": launches,
                "roofline": {"bound": "hbm", "kernel": "clim_sweep2_kernel" if sweep_name.endswith("sweep2_f32") else "clim_sweep_kernel", "achieved": ach, "peak": peak,
                             "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": sweep_bytes, "ms_per_launch": sweep_ms,
                             "whole_step": {"algorithmic_bytes": b_alg, "achieved": b_alg / (ms_step * 1e-3) / 1e9,
                                            "frac": b_alg / (ms_step * 1e-3) / 1e9 / peak}},
                "kernel_ms": {k: float(np.mean(v)) for k, v in per_kernel.items()},
                "cpu_baseline": cpu, "clocks": sampler.summary()}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="global025_30yr", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-cells", type=int, default=400, help="cells per host process in the CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--slabs", type=int, default=24, help="column blocks of the host-buffer (e2e) pipeline")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: everything libraries write to file descriptor 1 (NCCL's
    # version banner, nvcc during build()) goes to stderr; the line is written to the saved descriptor
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
