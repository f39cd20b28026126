"""The C-ABI shared library loads on a GPU-less box and exports every symbol that
include/xmhw_b200.h declares (no compute call is made here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header():
    from xmhw_b200 import _cabi
    hdr = open(os.path.join(ROOT, "include", "xmhw_b200.h")).read()
    declared = set(re.findall(r"\b(xmhw_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_cabi.EXPORTS), declared ^ set(_cabi.EXPORTS)
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _cabi.lib.xmhw_abi_version() == _cabi.ABI_VERSION == int(re.search(r"#define XMHW_ABI_VERSION (\d+)", hdr).group(1))
    assert b"plan" in _cabi.lib.xmhw_strerror(-2)


def test_event_field_tables_match_header():
    from xmhw_b200 import _cabi
    hdr = open(os.path.join(ROOT, "include", "xmhw_b200.h")).read()
    ei = re.search(r"enum xmhw_event_i32 \{(.*?)\};", hdr, re.S).group(1)
    ef = re.search(r"enum xmhw_event_f64 \{(.*?)\};", hdr, re.S).group(1)
    names_i = [n for n in re.findall(r"XMHW_EI_([A-Z0-9_]+)", ei) if n != "COUNT"]
    names_f = [n for n in re.findall(r"XMHW_EF_([A-Z0-9_]+)", ef) if n != "COUNT"]
    assert [n.lower() for n in names_i] == [f.lower() for f in _cabi.EI_FIELDS]
    assert [n.lower() for n in names_f] == [f.lower() for f in _cabi.EF_FIELDS]


def test_argument_errors_without_gpu():
    """Null pointers / bad sizes are rejected before any CUDA call."""
    from xmhw_b200 import _cabi
    assert _cabi.lib.xmhw_clim_finish_f64(None, None, 366, 10, 1, 31, None, None) == -1
    assert _cabi.lib.xmhw_clim_sweep2_f32(None, 10, 10, None, None, None, None, None, None) == -1
    assert _cabi.lib.xmhw_events_count(None, 10, 10, 5, 1, 2, None, None) == -1
    assert _cabi.lib.xmhw_exclusive_scan_i32(None, 0, None, None, None) == -1
