"""Cell sharding across GPUs (one process per GPU) -- SURVEY.md 8e.

Every statistic of the hot path is per cell, so the grid is cut into contiguous ranges of
the flattened (sorted-dim) cell axis, balanced by OCEAN-cell count, one range per rank.
Each rank reads its column block of the (time, cell) array and produces thresh/seas for
its block and an event table with GLOBAL cell ids.  There is no data-path collective: the
only exchange is the final gather of results to rank 0 over the process group (host
objects; NCCL is deliberately not used because nothing is reduced across shards).
"""
import numpy as np


def balanced_ranges(ocean_mask, nparts, align=32):
    """Split cells [0, n) into `nparts` contiguous ranges with ~equal ocean-cell counts.
    Boundaries are multiples of `align` (a warp handles 32 adjacent cells), so that row
    segments of every shard stay 128-byte aligned.  Returns [(start, stop), ...]."""
    ocean = np.asarray(ocean_mask, bool).ravel()
    n = len(ocean)
    if nparts < 1:
        raise ValueError("nparts must be >= 1")
    csum = np.concatenate(([0], np.cumsum(ocean)))
    total = csum[-1]
    bounds = [0]
    for k in range(1, nparts):
        target = total * k / nparts
        b = int(np.searchsorted(csum, target, side="left"))
        b = min(n, max(bounds[-1], int(round(b / align)) * align))
        bounds.append(b)
    bounds.append(n)
    return [(bounds[i], bounds[i + 1]) for i in range(nparts)]


def column_block(ts_time_major, start, stop):
    """The (time, cell[start:stop]) block of a host array as a contiguous copy."""
    return np.ascontiguousarray(ts_time_major[:, start:stop])


def globalize(table, start):
    """Shift the `cell` column of a per-shard event table (dict of arrays) to global ids."""
    out = dict(table)
    out["cell"] = np.asarray(table["cell"]) + start
    return out


def concat_tables(tables):
    """Concatenate per-shard event tables (already ordered by cell) in shard order."""
    keys = tables[0].keys()
    return {k: np.concatenate([np.asarray(t[k]) for t in tables]) for k in keys}


def gather_results(local, group=None, dst=0):
    """Gather per-rank results (any picklable object) on rank `dst`; returns the list there
    and None elsewhere.  Works with any torch.distributed backend (gloo on CPU, nccl on GPU
    boxes -- objects travel through host memory either way)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    out = [None] * world if rank == dst else None
    dist.gather_object(local, out, dst=dst, group=group)
    return out
