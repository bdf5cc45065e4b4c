"""CPU tests of the oracle (the parity checker itself).

The reference holds no golden vectors and cannot be built here: what its own tests
DO assert is re-expressed below against the oracle, plus unit checks of the Rust-std BinaryHeap
restatement and of the distance definitions.
"""
import numpy as np
import pytest

from util import recall_ids


def test_rheap_matches_rust_std_semantics(po):
    # max-heap order, pop order = descending keys; into_sorted_vec ascending
    keys = [5.0, 1.0, 9.0, 3.0, 7.0, 2.0, 8.0]
    ops = list(range(len(keys))) + [-1, -1]
    vals = keys + [0, 0]
    out = po.rheap_script(ops, vals)
    assert out[:2].tolist() == [2, 6]                      # 9.0 then 8.0
    assert [keys[i] for i in out[2:]] == [1.0, 2.0, 3.0, 5.0, 7.0]
    # tie behaviour of std BinaryHeap: push does not sift past an EQUAL parent (sift_up stops on <=), so
    # with all-equal keys the root stays the first pushed; pop moves the LAST element to the root and
    # sift_down_to_bottom prefers the right child on ties.  Known answer worked by hand:
    # push a,b,c (equal keys): vec [a,b,c]; pop -> returns a; last (c) goes to root: [c,b], child b <= hole? the
    # sift goes to bottom then up: end state [b,c]?  hole=c at 0, child=1==end-1 -> move b up: [b,_], pos=1,
    # sift_up(0,1): c <= parent b -> stays: [b,c].  pop -> b.  then c.
    out = po.rheap_script([0, 1, 2, -1, -1, -1], [4.0] * 6)
    assert out.tolist() == [0, 1, 2]
    # 5 equal keys a..e (ids 0..4), worked by hand with the std rules (push never sifts past an equal parent):
    # [a,b,c,d,e] pop: last e swapped into root, returns a; bottom walk picks the RIGHT child on ties (c up),
    #   e lands at 2 -> [c,b,e,d];  pop: returns c, d into root, right child e up -> [e,b,d];
    #   pop: returns e, d into root, single child b up -> [b,d];  pop: returns b;  then d.
    out = po.rheap_script([0, 1, 2, 3, 4, -1, -1, -1, -1, -1], [1.0] * 10)
    assert out.tolist() == [0, 2, 4, 1, 3]


def test_rheap_det_mode_is_total_order(po):
    rng = np.random.default_rng(0)
    vals = rng.integers(0, 4, 200).astype(np.float32)  # many ties
    ops = list(range(200))
    out = po.rheap_script(ops, vals, mode=po.MODE_DET)
    keys = [(vals[i], i) for i in out]
    assert keys == sorted(keys)
    out = po.rheap_script(ops, -vals, mode=po.MODE_DET, neg=True)  # negative heap: ascending stored value
    keys = [(-vals[i], -i) for i in out]                           # == descending (d, id)
    assert keys == sorted(keys)


def test_distance_definitions(po):
    rng = np.random.default_rng(3)
    for d in (1, 7, 8, 25, 128, 131):
        a, b = rng.random(d, dtype=np.float32), rng.random(d, dtype=np.float32)
        a64, b64 = a.astype(np.float64), b.astype(np.float64)
        want = {
            "DistL1": np.abs(a64 - b64).sum(),
            "DistL2": np.sqrt(((a64 - b64) ** 2).sum()),          # un-squared (ann-sift1m example :172-178)
            "DistCosine": max(0.0, 1 - (a64 @ b64) / np.sqrt((a64 @ a64) * (b64 @ b64))),
            "DistHellinger": np.sqrt(max(0.0, 1 - np.sqrt(a64 * b64).sum())),
        }
        for name, w in want.items():
            for order in (po.ORDER_REF, po.ORDER_GPU):
                got = po.dist(a, b, name, order)
                assert abs(got - w) <= 2e-5 * max(1.0, abs(w)), (name, d, order, got, w)
        au, bu = a / np.linalg.norm(a), b / np.linalg.norm(b)
        w = max(0.0, 1 - float(au.astype(np.float64) @ bu.astype(np.float64)))
        assert abs(po.dist(au, bu, "DistDot") - w) < 1e-5
        # the two summation orders agree within the 1e-5 relative tolerance of north_star
        for name in ("DistL1", "DistL2"):
            r, g = po.dist(a, b, name, po.ORDER_REF), po.dist(a, b, name, po.ORDER_GPU)
            assert abs(r - g) <= 1e-5 * abs(r)
    x = rng.random(40, dtype=np.float32)
    for name in ("DistL1", "DistL2", "DistCosine"):
        assert po.dist(x, x, name) == 0.0                    # d(x,x) == 0 exactly (hnsw.rs:1878-1879)
    u = rng.integers(0, 5, 64).astype(np.uint16)
    v = rng.integers(0, 5, 64).astype(np.uint16)
    assert po.dist(u, v, "DistHamming") == np.float32((u != v).sum() / 64)
    assert abs(po.dist(u, v, "DistJaccard") - (1 - np.minimum(u, v).sum() / np.maximum(u, v).sum())) < 1e-6


@pytest.fixture(scope="module")
def c1(pkg, po):
    X = pkg.datagen.uniform(10000, 25, 1)
    o = po.Oracle(16, 10000, 9, 200, "DistL2", 25)
    o.insert_batch(X)
    return X, o


def test_c1_random_config_recall(pkg, po, c1):
    """BASELINE.json configs[0] (examples/random.rs shape): recall vs brute force, and the traversal
    counters of SURVEY.md App. E as sanity anchors (~26 expansions, ~660 evals at ef=24)."""
    X, o = c1
    Q = pkg.datagen.uniform(1000, 25, 2)
    ti, td = po.bruteforce(X, Q, 10, "DistL2")
    o.counters()
    oo, od, oi, pid, oc = o.search_batch(Q, 10, 24, nthreads=4)
    c = o.counters()
    r = recall_ids(oi, oc, ti)
    assert 0.85 < r < 0.97
    assert 600 < c["evals"] / 1000 < 720 and 24 < c["expansions"] / 1000 < 30
    assert np.all(np.diff(od, axis=1) >= 0)                  # ascending (hnsw.rs:1544)
    assert np.all(oc == 10)
    # PointId(level, rank) identifies the point: rank is unique within its level
    lv, rk, og = o.export_points()
    assert len(set(zip(lv.tolist(), rk.tolist()))) == len(lv)


def test_parallel_search_keeps_input_order(pkg, po, c1):
    X, o = c1
    Q = pkg.datagen.uniform(257, 25, 5)
    a = o.search_batch(Q, 5, 32, nthreads=1)
    b = o.search_batch(Q, 5, 32, nthreads=7)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_std_and_det_modes_agree_without_ties(pkg, po, c1):
    X, o = c1
    Q = pkg.datagen.uniform(500, 25, 6)
    a = o.search_batch(Q, 10, 64)
    o.set_mode(po.MODE_DET)
    b = o.search_batch(Q, 10, 64)
    o.set_mode(po.MODE_STD)
    assert np.array_equal(a[2], b[2])


def test_self_query_any_level(po):
    """hnsw.rs:1871-1879 test_sparse_search: a single inserted point is found whatever level it drew."""
    for lvl in (0, 3, 15):
        o = po.Oracle(16, 10, 16, 50, "DistL1", 8)
        v = np.arange(8, dtype=np.float32)
        o.insert_batch(v[None, :], ids=[77], levels=[lvl])
        og, d, it, pid, cnt = o.search_batch(v[None, :], 3, 10)
        assert cnt[0] == 1 and og[0, 0] == 77 and d[0, 0] == 0.0 and pid[0, 0, 0] == lvl


def test_filter_semantics(pkg, po):
    """tests/filtertest.rs: always-false filter => 0 hits (263-269); single-admit filter => <=1 hit (258);
    ids found by a filtered search carry the same distances as in an unfiltered search (211)."""
    X = pkg.datagen.uniform(3000, 8, 7)
    o = po.Oracle(8, 3000, 16, 100, "DistL2", 8)
    o.insert_batch(X)
    Q = pkg.datagen.uniform(50, 8, 8)
    og, d, it, pid, cnt = o.search_batch(Q, 10, 64, filter_fn=lambda i: False)
    assert np.all(cnt == 0)
    og, d, it, pid, cnt = o.search_batch(Q, 10, 4, filter_ids=[1234])
    assert np.all(cnt <= 1)
    allow = np.arange(0, 3000, 3)
    fo, fd, fi, _, fc = o.search_batch(Q, 10, 64, filter_ids=allow)
    assert fc.min() >= 1
    for i in range(len(Q)):
        assert np.all(fo[i, :fc[i]] % 3 == 0)
        for j in range(fc[i]):
            ref = po.dist(Q[i], X[int(fo[i, j])], "DistL2")
            assert abs(fd[i, j] - ref) <= 1e-5 * max(ref, 1e-30)
    # callback and sorted-list forms of FilterT agree (filter.rs:11-24)
    go, gd, gi, _, gc = o.search_batch(Q, 10, 64, filter_fn=lambda i: i % 3 == 0)
    assert np.array_equal(fo, go) and np.array_equal(fc, gc)


def test_parallel_insert_quality(pkg, po):
    """racy parallel insert (hnsw.rs:1224-1238) gives a graph of the same quality as the serial one."""
    X = pkg.datagen.uniform(6000, 16, 9)
    Q = pkg.datagen.uniform(300, 16, 10)
    ti, _ = po.bruteforce(X, Q, 10, "DistL2")
    rec = []
    for nth in (1, 6):
        o = po.Oracle(16, 6000, 16, 100, "DistL2", 16)
        o.insert_batch(X, nthreads=nth)
        assert len(o) == 6000
        r = o.search_batch(Q, 10, 64)
        rec.append(recall_ids(r[0], r[4], ti.astype(np.uint64)))   # origin ids: racy inserts number points in arrival order
    assert abs(rec[0] - rec[1]) < 0.02 and rec[0] > 0.9


def test_export_import_roundtrip(pkg, po):
    X = pkg.datagen.uniform(2000, 12, 3)
    o = po.Oracle(8, 2000, 16, 60, "DistL2", 12)
    o.insert_batch(X)
    lv, rk, og = o.export_points()
    o2 = po.Oracle(8, 2000, 16, 60, "DistL2", 12)
    o2.import_graph(o.export_vectors(), og, lv, o.entry, {l: o.export_layer(l) for l in range(16)})
    Q = pkg.datagen.uniform(100, 12, 4)
    a, b = o.search_batch(Q, 10, 32), o2.search_batch(Q, 10, 32)
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[1], b[1])


def test_level_law(po):
    """LayerGenerator (hnsw.rs:363-374): P(level >= l) = M^-l for scale 1/ln(M)."""
    o = po.Oracle(16, 10, 16, 50, "DistL2", 4)
    lv = o.draw_levels(400000)
    for l in (1, 2):
        p = (lv >= l).mean()
        assert abs(p - 16.0 ** -l) < 3 * np.sqrt(16.0 ** -l / 400000) + 1e-4
    o = po.Oracle(16, 10, 16, 50, "DistL2", 4)
    o.modify_level_scale(0.5)
    lv = o.draw_levels(400000)
    assert abs((lv >= 1).mean() - 16.0 ** -2) < 1e-3
