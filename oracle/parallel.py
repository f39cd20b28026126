"""TEST INFRASTRUCTURE: the numpy oracle (oracle/xmhw_oracle.py) over column chunks on all host
cores.  Cells are independent, so the chunks are too; the workers are spawned (never forked: the
test process usually holds a CUDA context) and import nothing but numpy and the oracle."""
import multiprocessing as mp
import os

import numpy as np


def _work(args):
    from oracle import xmhw_oracle as O
    ts, doy, ndoy, tkw, dkw, with_detect = args
    th, se = O.threshold(ts, doy, ndoy, **tkw)
    ev = O.detect(ts, doy, th, se, **dkw) if with_detect else None
    return th, se, ev


def threshold_detect(ts, doy, ndoy, chunk=256, procs=None, detect=True, tkw=None, dkw=None):
    """(thresh, seas [ndoy, ncell] f64, event dict with cell ids local to ts) of the oracle."""
    ts = np.ascontiguousarray(ts, np.float32)
    ncell = ts.shape[1]
    jobs = [(np.ascontiguousarray(ts[:, a:a + chunk]), np.asarray(doy), ndoy, tkw or {}, dkw or {}, detect)
            for a in range(0, ncell, chunk)]
    procs = procs or min(len(jobs), os.cpu_count() or 1)
    if procs <= 1:
        res = [_work(j) for j in jobs]
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(_work, jobs)
    th = np.concatenate([r[0] for r in res], axis=1)
    se = np.concatenate([r[1] for r in res], axis=1)
    ev = None
    if detect:
        ev = {}
        for k in res[0][2]:
            parts = []
            for i, r in enumerate(res):
                v = r[2][k]
                parts.append(v + i * chunk if k == "cell" else v)
            ev[k] = np.concatenate(parts)
    return th, se, ev
