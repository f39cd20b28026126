"""Host-side plan of the two-stack top-K climatology sweep (csrc/xmhw_topk.h).

The reference pools, for every day-of-year d, the samples ts[t+k] with doy[t] == d and
|k| <= windowHalfWidth (xmhw/identify.py:184-209, then groupby("doy") at :233 / :263).
Everything here is derived from the ACTUAL doy vector, like plan.py:

* the *signature* of time row t' is the set of doys whose window contains it;
* rows with equal signature form an ATOM (normally: all years of one calendar day);
* if the atoms can be ordered so that every doy's window is a contiguous range of atoms
  whose two ends only move forward (true for daily 366-day calendars except doy 60, for
  `tstep` calendars and at the series edges), the sweep is a queue: each doy pops the atoms
  that left and pushes the atoms that entered.  Atoms that are popped together form a UNIT
  and share one shared-memory slot;
* doys that break the order (doy 60: its window holds leap years only) are computed by the
  direct selection kernel from their row list ("exceptional" doys);
* the queue is simulated here once (two stacks: pushes go to the back, a FLIP turns the back
  into front arrays when the front runs empty), and every step becomes a fixed-size record:
  the kernel never takes a data- or plan-dependent decision that is not warp-uniform.

`build_clim_plan2` returns None when the calendar does not fit (the caller then uses the
general sorted-list sweep of plan.py).
"""
from dataclasses import dataclass, field

import numpy as np

KP_CLASSES = (8, 16, 24, 36, 48)       # top-K capacities the CUDA library instantiates
MAX_ATOM = 48                          # rows per atom (register sorting network)
JOB_F_COPY, JOB_F_FIRST, JOB_F_STORE, JOB_F_STOREP, JOB_F_CLEAR, JOB_F_RAGGED = 1, 2, 4, 8, 16, 32   # csrc/xmhw_topk.h
MAX_POP, MAX_PUSH = 4, 3
SMEM_LIMIT = 227 * 1024
# static sizes of the plan block that travels in the kernel's parameter space (xmhw_clim_plan2)
SC_MAX_STEPS, SC_REC_WORDS, SC_MAX_FLIP, SC_MAX_PAT, SC_PAT_LEN, SC_MAX_INIT = 366, 12, 768, 16, 48, 32


@dataclass
class ClimPlan2Host:
    nsteps: int
    kp: int
    max_size: int
    slot_rows: int
    nslots: int
    n_init: int
    cap: int                               # key rows per slot
    pool_rows: int
    rec: np.ndarray                        # [nsteps][SC_REC_WORDS] uint32 step records
    flip: np.ndarray                       # [n] uint32 flip entries
    pat: np.ndarray                        # [npat][SC_PAT_LEN] int32 row patterns (rows of an atom - its first row)
    init: np.ndarray                       # [n_init + 1][2] uint32 atoms of the first window (+ the first push)
    q: float
    nmax: int
    step_doy: np.ndarray                   # [nsteps] doy label of every sweep step
    exc_doy: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))      # exceptional doy labels
    exc_off: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int32))      # CSR into exc_rows
    exc_rows: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    n_merges: int = 0                      # diagnostics: top-K merges over the whole sweep
    n_flips: int = 0

    def smem_bytes(self):
        return self.pool_rows * 128


LAST_FAIL = {"code": 0}


def _fail(code):
    """build_clim_plan2 gives up (the caller uses the general sweep); the code tells tests / tools where."""
    LAST_FAIL["code"] = code
    return None


def max_rank(nmax, q):
    """Largest rank from the top, n - floor((n-1) q), that the quantile of 1..nmax samples needs."""
    n = np.arange(1, nmax + 1, dtype=np.int64)
    v = (n - 1) * np.float64(q)
    fl = np.floor(v)
    fl = np.where(v >= n - 1, n - 1, fl)
    return int((n - fl.astype(np.int64)).max())


def _signatures(doy, w):
    T = len(doy)
    sigs = []
    for tp in range(T):
        lo, hi = max(0, tp - w), min(T, tp + w + 1)
        seg = doy[lo:hi]
        s = set(seg.tolist())
        if len(s) != hi - lo:
            return None
        sigs.append(s)
    return sigs


def _try_order(doy, ndoy, sigs, exc):
    """Atoms (rows, a, b) over the positions of the regular doys, ordered so that windows are ranges.
    Returns (atoms, regular doys) or None."""
    regular = [d for d in range(1, ndoy + 1) if d not in exc]
    pos = {d: i for i, d in enumerate(regular)}
    groups = {}
    for tp, s in enumerate(sigs):
        ps = sorted(pos[d] for d in s if d in pos)
        if not ps:
            continue
        run = [ps[0]]
        runs = []
        for x in ps[1:]:
            if x == run[-1] + 1:
                run.append(x)
            else:
                runs.append(run)
                run = [x]
        runs.append(run)
        for r in runs:
            groups.setdefault((r[0], r[-1]), []).append(tp)
    atoms = sorted(((a, b, rows) for (a, b), rows in groups.items()), key=lambda x: (x[0], x[1]))
    last_b = -1
    for a, b, _ in atoms:
        if b < last_b:
            return None                    # an atom nested inside an older one: windows are not ranges
        last_b = b
    # every regular position must be covered consistently: window(p) = atoms with a <= p <= b
    return atoms, regular


def build_clim_plan2(doy, ndoy, w, q):
    """doy: int array [T] of 1-based labels in 1..ndoy; w: window half width; q in [0,1].
    Returns a ClimPlan2Host, or None when this calendar / quantile needs the general sweep."""
    doy = np.asarray(doy, dtype=np.int64)
    T = len(doy)
    if T == 0 or doy.min() < 1 or doy.max() > ndoy or w < 0:
        return _fail(1)
    sigs = _signatures(doy, w)
    if sigs is None:
        return _fail(2)
    counts = np.bincount(doy - 1, minlength=ndoy)
    present = [d for d in range(1, ndoy + 1) if counts[d - 1] > 0]
    absent = set(range(1, ndoy + 1)) - set(present)
    # exceptional candidates: none, then the doys with far fewer time steps than the typical one
    typical = np.median(counts[counts > 0])
    rare = {d for d in present if counts[d - 1] * 2 < typical}
    fewer = {d for d in present if counts[d - 1] < typical}
    ordered = None
    for exc in (set(), rare, fewer):
        if len(exc) > 4:
            continue
        got = _try_order(doy, ndoy, sigs, exc | absent)
        if got is not None:
            ordered, exc_used = got, exc
            break
    if ordered is None:
        return _fail(3)
    atoms0, regular = ordered
    nsteps = len(regular)
    if nsteps == 0:
        return _fail(4)
    # split oversized atoms (series longer than MAX_ATOM years)
    atoms = []
    for a, b, rows in atoms0:
        rows = np.asarray(rows, np.int32)
        for pc in np.array_split(rows, -(-len(rows) // MAX_ATOM)):
            atoms.append((a, b, pc))
    natoms = len(atoms)
    a_arr = np.array([x[0] for x in atoms])
    b_arr = np.array([x[1] for x in atoms])
    sizes = np.array([len(x[2]) for x in atoms])
    max_size = int(sizes.max())
    maxn = 32 if max_size <= 32 else 48
    # window sizes -> needed capacity
    win = np.zeros(nsteps, np.int64)
    for a, b, rows in atoms:
        win[a:b + 1] += len(rows)
    nmax = int(win.max())
    kneed = max_rank(nmax, q)
    kp = next((k for k in KP_CLASSES if k >= kneed), None)
    if kp is None:
        return _fail(5)
    # units: atoms popped together (same b), consecutive in the order
    unit_of = np.zeros(natoms, np.int64)
    units = []          # (first atom, last atom + 1)
    i = 0
    while i < natoms:
        j = i
        # (atoms that stay to the last step are never popped: one unit each, or the units of the
        # final window would lump together)
        while j + 1 < natoms and b_arr[j + 1] == b_arr[i] and b_arr[i] < nsteps - 1:
            j += 1
        units.append((i, j + 1))
        unit_of[i:j + 1] = len(units) - 1
        i = j + 1
    unit_rows = [int(sizes[u0:u1].sum()) for u0, u1 in units]
    cap = max(kp, max(unit_rows))                 # key rows of a slot
    slot_rows = 1 + cap + 2                       # len | guard row, keys, f64 sum of the unit (lo, hi)
    if cap > 127:
        return _fail(6)
    # row patterns: rows of an atom relative to its first row (a handful of distinct ones)
    pat_id = {}
    pats = []
    atom_pat = np.zeros(natoms, np.int64)
    for i, (_, _, rws) in enumerate(atoms):
        rel = tuple(int(x) for x in (rws - rws[0]))
        if rel not in pat_id:
            pat_id[rel] = len(pats)
            pats.append(rel)
        atom_pat[i] = pat_id[rel]
    if len(pats) > SC_MAX_PAT or max_size > SC_PAT_LEN or nsteps > SC_MAX_STEPS or int(rows_max := max(int(r[2][0]) for r in atoms)) >= (1 << 24):
        return _fail(17)

    # ---- simulate the queue
    def ragged(i_or_size, off, alone):
        """The kernel stores / reloads an atom as `maxn` rows (its one size class, zero padded).  That is
        only right when the atom is alone in its unit's slot and the padded rows fit: otherwise the
        kernel takes the predicated (exact row count) path."""
        return 0 if (alone and off + maxn <= cap) else 1

    unit_len = [u1 - u0 for u0, u1 in units]

    def atom_words(i, flags, slot, off):
        """two-word atom descriptor: first row | size << 24 | flags << 30, pattern | slot << 5 | offset << 10 | ragged << 17"""
        sz = int(sizes[i])
        return (int(atoms[i][2][0]) | (sz << 24) | (flags << 30),
                int(atom_pat[i]) | (slot << 5) | (off << 10) | (ragged(sz, off, unit_len[int(unit_of[i])] == 1) << 17))

    free_slots = []
    nslots = 0
    slot_of_unit = {}
    stash_fill = {}                      # unit -> rows already stashed in its slot
    atom_desc = np.zeros((natoms + 1, 2), np.int64)     # trailing zero descriptor: nothing left to prefetch
    rec = np.zeros((nsteps, SC_REC_WORDS), np.int64)
    flip_entries = []
    front = []                           # units with a front array, oldest first
    back = []                            # units stashed since the last flip, oldest first
    back_atoms = {}                      # unit -> list of (dest row, size) stashed
    accumulator_empty = True
    n_merges = n_flips = 0
    next_push = 0                        # next atom to push (atoms are pushed in order)
    next_pop_unit = 0

    def slot_base(u):
        return slot_of_unit[u]           # flip entries / records name slots; the kernel multiplies by slot_rows

    def is_partial(u):
        return len(back_atoms[u]) != units[u][1] - units[u][0]

    def do_flip(allow_storep):
        """Turn the fully pushed units of the back into front arrays.  A partly pushed unit (always the
        youngest) stays in the back and the accumulator is reloaded from its stash."""
        nonlocal accumulator_empty, n_merges, n_flips
        off = len(flip_entries)
        partial = back[-1] if back and is_partial(back[-1]) else None
        full = [u for u in back if u != partial]
        if not full:
            return off, 0
        n_flips += 1
        chain = list(reversed(full))                     # youngest first
        if allow_storep and partial is None:
            # the oldest unit's array = everything pushed since the last flip = the accumulator as it is
            flip_entries.append((0, 0, 0, JOB_F_STOREP, slot_base(full[0])))
            chain = chain[:-1]
        first = True
        for u in chain:
            ats = back_atoms[u]
            for m, (dest, size) in enumerate(reversed(ats)):
                fl = JOB_F_COPY if first else 0
                if not first:
                    n_merges += 1
                first = False
                if m == len(ats) - 1:
                    fl |= JOB_F_STORE
                if ragged(size, dest, len(ats) == 1):
                    fl |= JOB_F_RAGGED
                flip_entries.append((slot_base(u), dest, size, fl, slot_base(u)))
        front.extend(full)
        back.clear()
        if partial is None:
            flip_entries.append((0, 0, 0, JOB_F_CLEAR, 0))
            accumulator_empty = True
        else:
            back.append(partial)
            for m, (dest, size) in enumerate(back_atoms[partial]):
                flip_entries.append((slot_base(partial), dest, size,
                                     (JOB_F_COPY if m == 0 else 0) | JOB_F_RAGGED, 0))
                if m:
                    n_merges += 1
            accumulator_empty = False
        return off, len(flip_entries) - off

    def do_push(s, j_list):
        """push atom next_push; returns its record"""
        nonlocal next_push, nslots, accumulator_empty, n_merges
        i = next_push
        next_push += 1
        u = int(unit_of[i])
        flags = 0
        if u not in slot_of_unit:
            if free_slots:
                slot_of_unit[u] = free_slots.pop(0)
            else:
                slot_of_unit[u] = nslots
                nslots += 1
            stash_fill[u] = 0
            back.append(u)
            back_atoms[u] = []
            flags |= JOB_F_FIRST
        elif u not in back:
            raise RuntimeError("unit split across a flip")        # guarded below: handled by returning None
        if accumulator_empty:
            flags |= JOB_F_COPY
            accumulator_empty = False
        else:
            n_merges += 1
        dest = stash_fill[u]             # key-row offset inside the slot
        stash_fill[u] += int(sizes[i])
        back_atoms[u].append((dest, int(sizes[i])))
        r = atom_words(i, flags, slot_of_unit[u], dest)
        atom_desc[i] = r
        return r

    def do_pop():
        nonlocal next_pop_unit
        u = next_pop_unit
        next_pop_unit += 1
        if front and front[0] == u:
            front.pop(0)
        elif back and back[0] == u:
            back.pop(0)                  # popped straight from the back: the accumulator is stale, flip follows
        else:
            raise RuntimeError("pop order")
        sl = slot_of_unit.pop(u)
        free_slots.append(sl)
        free_slots.sort()
        return sl

    push_idx = []
    try:
        # initial fill: every atom of the first window
        n_init = int(np.searchsorted(a_arr, 0, side="right"))
        for _ in range(n_init):
            do_push(0, None)
        for s in range(nsteps):
            pops = []
            stale = False
            while next_pop_unit < len(units) and b_arr[units[next_pop_unit][0]] < s:
                u = next_pop_unit
                if not (front and front[0] == u):
                    stale = True         # the unit is still in the back: its keys sit in the accumulator
                pops.append(do_pop())
            if len(pops) > MAX_POP:
                return _fail(7)
            flip_off, n_flip, flip_late = 0, 0, 0
            if stale:
                if front:
                    return _fail(8)          # cannot happen: pops are oldest first, the front is older than the back
                # the accumulator held the popped unit too: rebuild the front from the stashes (no shortcut)
                flip_off, n_flip = do_flip(False)
            pushes = []
            while next_push < natoms and a_arr[next_push] <= s:
                if s == 0:
                    break                # all of step 0's atoms were pushed by the initial fill
                pushes.append(do_push(s, None))
            if len(pushes) > MAX_PUSH:
                return _fail(9)
            if not front:
                if n_flip:
                    return _fail(10)          # two flips in one step: not representable
                flip_off, n_flip = do_flip(True)
                flip_late = 1
            if not front or n_flip > 63:
                return _fail(11)
            rec[s, 0] = (len(pops) | (len(pushes) << 3) | (flip_late << 5) | (n_flip << 6) | (slot_base(front[0]) << 12)
                         | ((regular[s] - 1) << 17))
            rec[s, 1] = sum(sl << (5 * j) for j, sl in enumerate(pops))
            rec[s, 3] = flip_off
            for j, r in enumerate(pushes):
                rec[s, 4 + 2 * j], rec[s, 5 + 2 * j] = r
            push_idx.append((s, next_push))      # the atom pushed after this step's last push: filled in below
            alive = 0
            for u in front + back:
                alive |= 1 << slot_of_unit[u]
            rec[s, 2] = alive
            # consistency: the units alive are exactly the window of position s
            expect = {int(unit_of[i]) for i in range(natoms) if a_arr[i] <= s <= b_arr[i]}
            if set(front + back) != expect:
                return _fail(12)
    except RuntimeError:
        return _fail(13)
    if nslots > 32 or regular[-1] >= (1 << 15):
        return _fail(14)
    pool_rows = nslots * slot_rows
    if pool_rows * 128 > SMEM_LIMIT or n_init > SC_MAX_INIT - 1 or len(flip_entries) > SC_MAX_FLIP:
        return _fail(15)
    # the descriptor of the atom to prefetch after a step's last push (push order = atom order)
    for s_, nx in push_idx:
        rec[s_, 10], rec[s_, 11] = atom_desc[nx]
    fl = np.zeros(max(1, len(flip_entries)), np.int64)
    for k, (ssl, soff, size, flags, dsl) in enumerate(flip_entries):
        fl[k] = ssl | (soff << 5) | (size << 12) | (flags << 18) | (dsl << 24)
    pat = np.zeros((len(pats), SC_PAT_LEN), np.int32)
    for k, rel in enumerate(pats):
        pat[k, :len(rel)] = rel
    # exceptional doys: their window rows, straight from the signatures
    exc_list = sorted(exc_used)
    exc_off = [0]
    exc_rows = []
    for d in exc_list:
        r = [tp for tp, sg in enumerate(sigs) if d in sg]
        exc_rows.extend(r)
        exc_off.append(len(exc_rows))
        nmax_d = len(r)
        if max_rank(max(1, nmax_d), q) > kp:
            return _fail(16)
    return ClimPlan2Host(
        nsteps=nsteps, kp=kp, max_size=max_size, slot_rows=slot_rows, nslots=nslots, n_init=n_init, cap=cap,
        pool_rows=pool_rows, rec=rec.astype(np.uint32), flip=fl.astype(np.uint32), pat=pat,
        init=atom_desc[:n_init + 1].astype(np.uint32), q=float(q), nmax=nmax,
        step_doy=np.asarray(regular, np.int32),
        exc_doy=np.asarray(exc_list, np.int32), exc_off=np.asarray(exc_off, np.int32),
        exc_rows=np.asarray(exc_rows if exc_rows else [0], np.int32),
        n_merges=n_merges, n_flips=n_flips)
