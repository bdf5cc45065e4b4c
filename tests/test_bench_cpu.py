"""bench.py host logic that needs no GPU: presets, workload naming, shard arithmetic, the committed ncu CSV reader."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _parse(argv):
    import bench
    old = sys.argv
    sys.argv = ["bench.py"] + argv
    try:
        return bench, bench.parse()
    finally:
        sys.argv = old


def test_presets_match_baseline_configs():
    bench, a = _parse([])
    assert (a.config, a.n, a.d, a.nq, a.M, a.efc, a.k, a.ef, a.metric, a.custom) == ("c2", 1000000, 128, 10000, 16, 200, 10, 64, "DistL2", False)
    assert a.steps == 100 and a.warmup >= 3 and not a.strong
    assert "C2" in bench.workload_name(a) and "10000 queries/step/GPU" in bench.workload_name(a)
    _, c5 = _parse(["--config", "c5", "--gpus", "8"])
    assert c5.strong and c5.nq == 1000000 and "over 8 GPU(s) (125000/GPU)" in bench.workload_name(c5, 8)
    _, c3 = _parse(["--config", "c3"])
    assert (c3.n, c3.d, c3.M, c3.ef, c3.metric, c3.data) == (1183514, 25, 24, 128, "DistCosine", "unit")
    _, c4 = _parse(["--config", "c4"])
    assert (c4.n, c4.d, c4.M, c4.ef) == (60000, 784, 32, 200)
    _, cu = _parse(["--config", "c1", "--ef", "48"])
    assert cu.custom and "custom" in bench.workload_name(cu)


def test_shard_bounds_cover_the_batch():
    import bench
    for n in (0, 1, 9, 1000001):
        for w in (1, 2, 3, 8):
            b = [bench.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1


def test_traffic_comes_from_the_committed_ncu_summary():
    bench, a = _parse([])
    t, src = bench.committed_traffic(a)
    assert src == os.path.join("profiles", "r2_search_lean_kernel_ncu_full_selected.csv")
    assert 5.4e9 < t < 7.0e9          # dram read + write of one launch, a little above the algorithmic 5.45 GB
    _, c1 = _parse(["--config", "c1"])
    assert bench.committed_traffic(c1) == (None, None)
