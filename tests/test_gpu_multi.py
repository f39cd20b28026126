"""GPU test of the N > 1 product path (xmhw_b200.multi) on one device: the ranks of a 3-way partition of
ONE grid are run one after the other, each on its own column block with global cell ids; their results
concatenate to the single-rank result bit for bit and their checksums add up to the single-rank ones."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_partitioned_grid_equals_whole_grid():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from xmhw_b200 import core, multi, synth
    tm = synth.daily_time(1993, 2004)
    doy = synth.doy366(tm)
    nlat, nlon = 12, 80
    land = synth.land_mask(nlat, nlon, 0.33).ravel()
    ocean = land == 0
    season = synth.season_table(tm)
    T = len(tm)

    def loader(a, b):
        return core.synth_sst_device(T, b - a, season, land=land[a:b], cell0=a, nan_ppm=2000)

    whole = multi.run_rank(loader, ocean, doy, 366, 0, 1)
    assert whole["range"] == (0, nlat * nlon)
    world = 3
    parts = [multi.run_rank(loader, ocean, doy, 366, r, world) for r in range(world)]
    assert [p["range"][0] for p in parts][0] == 0 and parts[-1]["range"][1] == nlat * nlon
    assert all(p["range"][0] % 32 == 0 for p in parts)
    th = torch.cat([p["thresh"] for p in parts], dim=1)
    assert torch.equal(torch.nan_to_num(th, nan=-1.0), torch.nan_to_num(whole["thresh"], nan=-1.0))
    n = whole["events"].n
    assert sum(p["events"].n for p in parts) == n
    i32 = torch.cat([p["events"].i32[:, :p["events"].n] for p in parts], dim=1)
    f64 = torch.cat([p["events"].f64[:, :p["events"].n] for p in parts], dim=1)
    assert torch.equal(i32, whole["events"].i32[:, :n])              # global cell ids, cell order
    assert np.array_equal(f64.cpu().numpy().view(np.int64), whole["events"].f64[:, :n].cpu().numpy().view(np.int64))
    mask64 = (1 << 64) - 1
    for k in ("events", "table", "clim"):
        tot = sum(p["checksums"][k] for p in parts)
        assert (tot if k == "events" else tot & mask64) == whole["checksums"][k], k
    # ocean-balanced: the ranks' ocean counts differ by little
    counts = [int(ocean[a:b].sum()) for a, b in (p["range"] for p in parts)]
    assert max(counts) - min(counts) <= 64
