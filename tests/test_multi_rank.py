"""world_size-2 test (gloo, CPU) of the N>1 host logic: ocean-balanced contiguous cell
ranges, per-shard processing with global cell ids, final gather on rank 0.  The per-shard
compute is the CPU oracle here (checker standing in for the GPU so the test runs on the
GPU-less box); the sharding / gather code under test is the product's (xmhw_b200.shard)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _table(ev, ngrid):
    """oracle event dict -> a core.EventTable on CPU tensors (what multi.event_checksum reads)"""
    import torch
    from xmhw_b200 import core
    n = len(ev["cell"])
    i32 = torch.from_numpy(np.stack([np.asarray(ev[f], np.int32) for f in core.EI_FIELDS]).reshape(core.EI_COUNT, n))
    f64 = torch.from_numpy(np.stack([np.asarray(ev[f], np.float64) for f in core.EF_FIELDS]).reshape(core.EF_COUNT, n))
    return core.EventTable(i32, f64, n, None, None, 0, ngrid)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import xmhw_oracle as O
    from xmhw_b200 import shard, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tm = synth.daily_time(2001, 2004)
    doy = synth.doy366(tm)
    land = synth.land_mask(4, 48).ravel()
    ranges = shard.balanced_ranges(~land.astype(bool), world)
    a, b = ranges[rank]
    ts = synth.synth_sst(len(tm), b - a, synth.season_table(tm), land=land[a:b], cell0=a)   # own shard only
    th, se = O.threshold(ts, doy, 366)
    loc = O.detect(ts, doy, th, se)
    ev = shard.globalize(loc, a)
    # the product's partition-independent checksums (xmhw_b200.multi), summed over the ranks with gloo
    from xmhw_b200 import multi
    sums = {"events": len(loc["cell"]), "table": multi.event_checksum(_table(loc, b - a), a),
            "clim": multi.clim_checksum(torch.from_numpy(th), torch.from_numpy(se))}
    sums = multi.combine_checksums(sums)
    got = shard.gather_results({"range": (a, b), "thresh": th, "events": ev, "sums": sums,
                                "rank_range": multi.rank_range(~land.astype(bool), rank, world)}, dst=0)
    if rank == 0:
        q.put(got)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
    import torch.multiprocessing as mp
    from oracle import xmhw_oracle as O
    from xmhw_b200 import shard, synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process answer on the whole grid
    tm = synth.daily_time(2001, 2004)
    doy = synth.doy366(tm)
    land = synth.land_mask(4, 48).ravel()
    ts = synth.synth_sst(len(tm), land.size, synth.season_table(tm), land=land)
    th, se = O.threshold(ts, doy, 366)
    ev = O.detect(ts, doy, th, se)
    assert [g["range"] for g in got] == shard.balanced_ranges(~land.astype(bool), 2)
    assert got[0]["range"][1] == got[1]["range"][0] and got[1]["range"][1] == land.size
    th_all = np.concatenate([g["thresh"] for g in got], axis=1)
    assert np.array_equal(th_all, th, equal_nan=True)
    allev = shard.concat_tables([g["events"] for g in got])
    for k in ev:
        assert np.array_equal(allev[k], ev[k], equal_nan=True), k
    # checksums: both ranks hold the same sums, equal to the one-process checksums of the whole grid
    import torch
    from xmhw_b200 import multi
    whole = {"events": len(ev["cell"]), "table": multi.event_checksum(_table(ev, land.size), 0),
             "clim": multi.clim_checksum(torch.from_numpy(th), torch.from_numpy(se))}
    assert got[0]["sums"] == got[1]["sums"] == whole
    assert [g["rank_range"] for g in got] == [g["range"] for g in got]


def test_balanced_ranges_properties():
    from xmhw_b200 import shard, synth
    land = synth.land_mask(72, 144).ravel().astype(bool)
    for n in (1, 2, 4, 8):
        r = shard.balanced_ranges(~land, n)
        assert r[0][0] == 0 and r[-1][1] == land.size and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert all(a % 32 == 0 for a, _ in r)
        counts = [int((~land[a:b]).sum()) for a, b in r]
        assert max(counts) - min(counts) <= 64 + 0.02 * sum(counts) / n
    assert shard.balanced_ranges(np.zeros(10, bool), 3)[-1][1] == 10      # all land: still a valid partition
