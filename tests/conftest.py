import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the CUDA library and the host lane emulator exist (nvcc cross-compiles
    here without a GPU; on the GPU box the prebuilt .so files travel with the snapshot)."""
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def oisst():
    z = np.load(os.path.join(GOLD, "oisst_2003_2004.npz"))
    time = np.datetime64("2003-01-01T12:00:00") + z["time"].astype("timedelta64[D]")
    return {"sst": z["sst"], "time": time, "lat": z["lat"], "lon": z["lon"]}


@pytest.fixture(scope="session")
def clim_gold():
    return (dict(np.load(os.path.join(GOLD, "clim_oisst.npz"))),
            dict(np.load(os.path.join(GOLD, "clim_oisst_nosmooth.npz"))))


@pytest.fixture(scope="session")
def ref_cases():
    z = np.load(os.path.join(GOLD, "ref_detect_cases.npz"))
    n = int(z["ncase"])
    cases = []
    for i in range(n):
        p = "c%03d_" % i
        cases.append({k[len(p):]: z[k] for k in z.files if k.startswith(p)})
    return cases
