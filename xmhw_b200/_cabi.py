"""ctypes binding of the C ABI in include/xmhw_b200.h.

The CUDA library is the ONLY compute path of this package: if the shared
object is missing or does not load, importing this module raises -- there is no
CPU or PyTorch fallback (build it with `python -c "import __graft_entry__ as g; g.build()"`).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("XMHW_B200_LIB") or os.path.join(_HERE, "_xmhw_b200.so")   # env override: development builds

EI_FIELDS = ("cell", "index_start", "index_end", "index_peak", "duration", "category",
             "duration_moderate", "duration_strong", "duration_severe", "duration_extreme")
EF_FIELDS = ("intensity_max", "intensity_mean", "intensity_cumulative", "intensity_var",
             "severity_max", "severity_mean", "severity_cumulative", "severity_var",
             "intensity_max_relThresh", "intensity_mean_relThresh",
             "intensity_cumulative_relThresh", "intensity_var_relThresh",
             "intensity_max_abs", "intensity_mean_abs", "intensity_cumulative_abs",
             "intensity_var_abs", "rate_onset", "rate_decline")
EI_COUNT = len(EI_FIELDS)
EF_COUNT = len(EF_FIELDS)

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


class ClimPlanStruct(C.Structure):
    """Mirror of `xmhw_clim_plan` (include/xmhw_b200.h)."""
    _fields_ = [("nsteps", C.c_int32), ("pool_rows", C.c_int32), ("nmax", C.c_int32),
                ("max_size", C.c_int32), ("scratch_rows", C.c_int32), ("reserved_", C.c_int32),
                ("inst_base", C.c_void_p), ("inst_size", C.c_void_p), ("inst_keep", C.c_void_p),
                ("inst_sbase", C.c_void_p),
                ("inst_row_off", C.c_void_p),
                ("rows", C.c_void_p),
                ("leave_off", C.c_void_p), ("leave", C.c_void_p),
                ("enter_off", C.c_void_p), ("enter", C.c_void_p),
                ("use_off", C.c_void_p), ("use", C.c_void_p),
                ("step_rec", C.c_void_p), ("q", C.c_double)]


SC_MAX_STEPS, SC_REC_WORDS, SC_MAX_FLIP, SC_MAX_PAT, SC_PAT_LEN, SC_MAX_INIT = 366, 12, 768, 16, 48, 32


class ClimPlan2Struct(C.Structure):
    """Mirror of `xmhw_clim_plan2` (include/xmhw_b200.h): the two-stack top-K sweep, one plain host
    struct that the library copies into the kernel's launch parameters."""
    _fields_ = [("nsteps", C.c_int32), ("kp", C.c_int32), ("max_size", C.c_int32), ("slot_rows", C.c_int32),
                ("nslots", C.c_int32), ("n_init", C.c_int32), ("cap", C.c_int32), ("reserved_", C.c_int32),
                ("q", C.c_double),
                ("rec", C.c_uint32 * SC_REC_WORDS * SC_MAX_STEPS), ("flip", C.c_uint32 * SC_MAX_FLIP),
                ("pat", C.c_int32 * SC_PAT_LEN * SC_MAX_PAT), ("init", C.c_uint32 * 2 * SC_MAX_INIT)]


INTERMEDIATE_FIELDS = (("events", "f8"), ("seas", "f8"), ("thresh", "f8"), ("relSeas", "f8"),
                       ("relThresh", "f8"), ("relThreshNorm", "f8"), ("severity", "f8"), ("cats", "f8"),
                       ("mabs", "f4"), ("bthresh", "u1"), ("duration_moderate", "u1"),
                       ("duration_strong", "u1"), ("duration_severe", "u1"), ("duration_extreme", "u1"))


class IntermediateStruct(C.Structure):
    """Mirror of `xmhw_intermediate` (include/xmhw_b200.h)."""
    _fields_ = [(name, C.c_void_p) for name, _ in INTERMEDIATE_FIELDS]


PLAN_ARRAYS = ("inst_base", "inst_size", "inst_keep", "inst_sbase", "inst_row_off", "rows", "leave_off", "leave",
               "enter_off", "enter", "use_off", "use", "step_rec")

_SIGNATURES = {
    "xmhw_abi_version": (C.c_int, []),
    "xmhw_strerror": (C.c_char_p, [C.c_int]),
    "xmhw_clim_sweep_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(ClimPlanStruct),
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "xmhw_group_order_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "xmhw_clim_sweep2_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(ClimPlan2Struct),
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "xmhw_clim_direct_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int32, C.c_int32,
                                       C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "xmhw_clim_finish_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32,
                                       C.c_int32, C.c_void_p, C.c_void_p]),
    "xmhw_clim_finish2_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64,
                                        C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "xmhw_exceed_mask_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                       C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "xmhw_events_count": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_void_p, C.c_void_p]),
    "xmhw_events_count_stage": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "xmhw_events_gather": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                     C.c_void_p, C.c_void_p]),
    "xmhw_detect_fused_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int64, C.c_void_p, C.c_void_p]),
    "xmhw_events_scatter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "xmhw_exclusive_scan_i32": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "xmhw_events_fill": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "xmhw_event_stats_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "xmhw_clim_cellmajor_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "xmhw_event_stats_cm_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p,
                                          C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "xmhw_intermediate_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int64, C.c_int64, C.POINTER(IntermediateStruct),
                                        C.c_void_p]),
    "xmhw_interp_gaps_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p]),
    "xmhw_block_average": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                     C.c_int32, C.c_void_p, C.c_void_p]),
    "xmhw_block_ts_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "xmhw_block_cat_days_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p,
                                          C.c_void_p]),
    "xmhw_event_rank_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "xmhw_count_valid_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "xmhw_copy2d_async": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                    C.c_int32, C.c_void_p]),
    "xmhw_synth_sst_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                     C.c_void_p, C.c_uint64, C.c_double, C.c_double, C.c_double,
                                     C.c_uint32, C.c_uint32, C.c_void_p]),
}

EXPORTS = tuple(_SIGNATURES)


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "xmhw_b200: CUDA library %s not found. Build it first "
            "(python -c \"import __graft_entry__ as g; g.build()\"). "
            "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()
ABI_VERSION = 3
if lib.xmhw_abi_version() != ABI_VERSION:
    raise ImportError("xmhw_b200: ABI version mismatch in %s" % LIB_PATH)


class XmhwCudaError(RuntimeError):
    pass


def check(code, what):
    if code != 0:
        msg = lib.xmhw_strerror(code).decode()
        raise XmhwCudaError("%s failed: %s (code %d)" % (what, msg, code))


def plan_struct(host_plan, pointers):
    """Build a ClimPlanStruct from a plan.ClimPlanHost and {array name: address}."""
    s = ClimPlanStruct()
    s.nsteps, s.pool_rows = host_plan.nsteps, host_plan.pool_rows
    s.nmax, s.max_size = host_plan.nmax, host_plan.max_size
    s.q = float(host_plan.q)
    s.scratch_rows = int(host_plan.scratch_rows)
    for name in PLAN_ARRAYS:
        setattr(s, name, pointers[name])
    return s


def plan2_struct(host_plan):
    """Build a ClimPlan2Struct (host memory) from a plan2.ClimPlan2Host."""
    s = ClimPlan2Struct()
    for f in ("nsteps", "kp", "max_size", "slot_rows", "nslots", "n_init", "cap"):
        setattr(s, f, int(getattr(host_plan, f)))
    s.q = float(host_plan.q)

    def fill(dst, src):
        src = np.ascontiguousarray(src)
        if src.nbytes > C.sizeof(dst):
            raise ValueError("plan array larger than its slot in xmhw_clim_plan2")
        C.memmove(C.addressof(dst), src.ctypes.data, src.nbytes)

    fill(s.rec, host_plan.rec.astype(np.uint32))
    fill(s.flip, host_plan.flip.astype(np.uint32))
    fill(s.pat, host_plan.pat.astype(np.int32))
    fill(s.init, host_plan.init.astype(np.uint32))
    return s


def numpy_plan_struct(host_plan):
    """Plan struct over HOST arrays (used only by the test-side lane emulator)."""
    keep = {n: np.ascontiguousarray(getattr(host_plan, n)) for n in PLAN_ARRAYS}
    s = plan_struct(host_plan, {n: a.ctypes.data for n, a in keep.items()})
    return s, keep
