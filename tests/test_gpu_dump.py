"""GPU tests of SURVEY §8 row f2: dump / reload in the reference's file format, cross-checked against the independent
Python restatement of the format (oracle/dumpfmt.py), and the reference's own dump/reload assertions
(hnswio.rs:1413-1460: reloaded graph equal, self-query distance 0; :1689-1700: empty dump is an error)."""
import os

import numpy as np
import pytest

from util import oracle_layers

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,metric", [(np.float32, "DistL2"), (np.uint8, "DistHamming"), (np.uint16, "DistJaccard")])
def test_engine_dump_is_readable_and_reloads(tmp_path, pkg, po, dtype, metric):
    import dumpfmt
    n, d, M, efc = 1200, 20, 8, 48
    rng = np.random.default_rng(3)
    X = pkg.datagen.uniform(n, d, 3) if dtype == np.float32 else rng.integers(0, 5, (n, d)).astype(dtype)
    Q = X[:50]
    h = pkg.Hnsw(M, n, 16, efc, metric, dtype=dtype)
    h.insert_flat(X, ids=np.arange(5000, 5000 + n))
    used = h.file_dump(tmp_path, "dumpA")
    assert used == "dumpA" and os.path.exists(tmp_path / "dumpA.hnsw.graph") and os.path.exists(tmp_path / "dumpA.hnsw.data")
    assert h.file_dump(tmp_path, "dumpA", overwrite=False) != "dumpA"      # DumpInit: unique basename, hnswio.rs:153-185
    # (1) the independent reader parses the engine's files and finds the exported graph
    b = dumpfmt.read_dump(str(tmp_path / "dumpA"), dtype)
    lv, rk, og, entry = h.export_points()
    order = np.lexsort((rk, lv))
    inv = np.empty(n, np.int64); inv[order] = np.arange(n)
    assert b["M"] == M and b["ef"] == efc and b["n"] == n and b["d"] == d and b["distname"].endswith("::" + metric)
    assert np.array_equal(b["origin"], og[order]) and np.array_equal(b["vecs"], X[order]) and b["entry"] == inv[entry]
    off, ids, ds = h.export_layer(0)
    for p in range(0, n, 97):
        want = [(int(inv[ids[j]]), float(ds[j])) for j in range(int(off[p]), int(off[p + 1]))]
        assert b["lists"][0][int(inv[p])] == want
    # (2) reload through the C ABI: same answers as before the dump, self-query distance 0
    h2 = pkg.Hnsw.load(tmp_path, "dumpA", metric, dtype=dtype)
    assert h2.get_nb_point() == n
    a1, a2 = h.search_flat(Q, 5, 32), h2.search_flat(Q, 5, 32)
    if dtype == np.float32:
        assert np.array_equal(a1[0], a2[0]) and np.array_equal(a1[1].view(np.uint32), a2[1].view(np.uint32))
    else:
        # a reload renumbers internal ids in file order (layer by layer), and ties are ordered by internal id:
        # equal-distance neighbours may swap; the distance lists must agree for (almost) every query
        assert np.mean(np.all(a1[1] == a2[1], axis=1)) > 0.9
    assert np.all(a2[1][:, 0] == 0.0)
    # (3) inserting after a reload is supported (hnswio.rs:1611-1632)
    extra = X[:10].copy()
    h2.insert_flat(extra, ids=np.arange(9000, 9010))
    assert h2.get_nb_point() == n + 10
    # wrong type / distance => loud failure
    with pytest.raises(pkg.HnswError):
        pkg.Hnsw.load(tmp_path, "dumpA", "DistL1" if metric != "DistL1" else "DistL2", dtype=dtype)


def test_engine_loads_dump_written_by_independent_writer(tmp_path, pkg, po):
    """a dump produced from an ORACLE-built graph by the Python writer (i.e. not by the engine) loads and searches
    exactly like the oracle"""
    import dumpfmt
    n, d, M, efc = 2000, 16, 8, 60
    X = pkg.datagen.uniform(n, d, 5)
    o = po.Oracle(M, n, 16, efc, "DistL2", d, mode=po.MODE_DET, order=po.ORDER_GPU)
    o.insert_batch(X, ids=np.arange(100, 100 + n))
    lv, rk, og = o.export_points()
    layers = oracle_layers(o, int(lv.max()) + 1)
    dumpfmt.write_dump(str(tmp_path / "orc"), X, og, lv, o.entry, layers, M, efc, 1 / np.log(M), "DistL2")
    h = pkg.Hnsw.load(tmp_path, "orc", "DistL2")
    Q = pkg.datagen.uniform(100, d, 6)
    oo, od, oi, _, oc = o.search_batch(Q, 10, 48)
    go, gd, gi, _, gc = h.search_flat(Q, 10, 48)
    # internal ids are renumbered in file order on reload; origin ids and distances must match exactly
    assert np.array_equal(gc, oc) and np.array_equal(go, oo)
    assert np.array_equal(gd.view(np.uint32), od.view(np.uint32))
    L = pkg.load_library()
    import ctypes as C
    path = str(tmp_path / "orc.hnsw.graph").encode()
    dptr = L.load_hnsw_description(len(path), path)
    assert dptr
    L.hnsw_b200_free_description(dptr)


def test_empty_index_dump_is_an_error(tmp_path, pkg):
    h = pkg.Hnsw(8, 10, 16, 20, "DistL2")
    with pytest.raises(pkg.HnswError):
        h.file_dump(tmp_path, "empty")


def test_flat_neighborhood_matches_oracle_lists(pkg, po):
    """FlatNeighborhood (flatten.rs:50-126): all layers merged, ascending distance; flatten.rs:146-197 checks it is
    identical before and after a dump/reload."""
    import tempfile
    n, d = 1500, 10
    X = pkg.datagen.uniform(n, d, 12)
    o = po.Oracle(6, n, 16, 40, "DistL2", d, mode=po.MODE_DET, order=po.ORDER_GPU)
    lv = o.draw_levels(n)
    o.insert_batch(X, ids=np.arange(300, 300 + n), levels=lv)
    h = pkg.Hnsw(6, n, 16, 40, "DistL2")
    h.set_insert_batching(1 << 30, 1)
    h.insert_flat(X, ids=np.arange(300, 300 + n), levels=lv)
    flat = h.flat_neighborhood()
    olv, ork, oog = o.export_points()
    layers = [o.export_layer(l) for l in range(int(olv.max()) + 1)]
    for p in (0, 5, 77, n - 1):
        own, full = [], []   # oracle lists of layers <= the point's level / of all 16 layers
        for l, (off, ids, ds) in enumerate(layers):
            item = [(np.float32(ds[j]), int(oog[ids[j]])) for j in range(int(off[p]), int(off[p + 1]))]
            full += item
            if l == 0 or olv[p] >= l:
                own += item
        got = [(np.float32(dd), i) for i, dd in flat[int(oog[p])]]
        assert got == sorted(got, key=lambda t: t[0])            # ascending distance (flatten.rs:82)
        # the engine materialises every list a search can reach: at least the layers up to the point's level, plus
        # (for former entry points) the layers it was promoted to; never a list the oracle does not have
        assert set(own) <= set(got) <= set(full)
    with tempfile.TemporaryDirectory() as td:
        h.file_dump(td, "fl")
        h2 = pkg.Hnsw.load(td, "fl", "DistL2")
        flat2 = h2.flat_neighborhood()
    assert all(flat[k] == flat2[k] for k in flat)
