"""CPU checks of the dump-format restatement (oracle/dumpfmt.py) against the byte layout the reference source
defines (hnswio.rs:878-919, 1063-1115, 1303-1340, 1382-1383): hand-packed known answer for a 2-point index."""
import struct

import numpy as np


def test_two_point_dump_bytes(tmp_path, po):
    import dumpfmt
    vecs = np.array([[1.0, 2.0], [3.0, 4.0]], np.float32)
    origin = [70, 71]
    levels = [0, 1]
    off = np.array([0, 1, 2], np.uint64)
    layers = [(off, np.array([1, 0], np.uint32), np.array([2.5, 2.5], np.float32))]
    base = str(tmp_path / "tiny")
    dumpfmt.write_dump(base, vecs, origin, levels, 1, layers, 16, 200, 0.36, "DistL2")
    g = open(base + ".hnsw.graph", "rb").read()
    dn = b"anndists::dist::distances::DistL2"
    want = struct.pack("=I", 0x002A6779) + bytes([1, 16]) + struct.pack("=d", 0.36) + bytes([16])
    want += struct.pack("=QQQ", 200, 2, 2) + struct.pack("=Q", len(dn)) + dn + struct.pack("=Q", 3) + b"f32"
    want += bytes([16])
    # layer 0: point (70, level 0, rank 0) with one layer-0 neighbour = origin 71, PointId(1,0), dist 2.5
    want += struct.pack("=IQ", 0x000A676F, 1) + struct.pack("=IQBi", 0x000A678F, 70, 0, 0)
    want += struct.pack("=Q", 1) + struct.pack("=QBif", 71, 1, 0, 2.5) + struct.pack("=Q", 0) * 15
    # layer 1: point (71, level 1, rank 0)
    want += struct.pack("=IQ", 0x000A676F, 1) + struct.pack("=IQBi", 0x000A678F, 71, 1, 0)
    want += struct.pack("=Q", 1) + struct.pack("=QBif", 70, 0, 0, 2.5) + struct.pack("=Q", 0) * 15
    for _ in range(14):
        want += struct.pack("=IQ", 0x000A676F, 0)
    want += struct.pack("=QBi", 71, 1, 0)     # entry point
    assert g == want
    d = open(base + ".hnsw.data", "rb").read()
    wd = struct.pack("=IQ", 0xA67F0000, 2)
    wd += struct.pack("=IQQ", 0xA67F0000, 70, 8) + vecs[0].tobytes()
    wd += struct.pack("=IQQ", 0xA67F0000, 71, 8) + vecs[1].tobytes()
    assert d == wd
    back = dumpfmt.read_dump(base, np.float32)
    assert back["entry"] == 1 and back["lists"][0][0] == [(1, 2.5)] and np.array_equal(back["vecs"], vecs)


def test_oracle_graph_roundtrip(tmp_path, pkg, po):
    import dumpfmt
    X = pkg.datagen.uniform(500, 7, 3)
    o = po.Oracle(6, 500, 16, 40, "DistL1", 7)
    o.insert_batch(X, ids=np.arange(1000, 1500))
    lv, rk, og = o.export_points()
    layers = [o.export_layer(l) for l in range(int(lv.max()) + 1)]
    base = str(tmp_path / "orc")
    dumpfmt.write_dump(base, X, og, lv, o.entry, layers, 6, 40, 1 / np.log(6), "DistL1")
    b = dumpfmt.read_dump(base, np.float32)
    # file order = layer by layer; map back through (level, rank)
    order = np.lexsort((rk, lv))
    assert np.array_equal(b["origin"], og[order]) and np.array_equal(b["vecs"], X[order])
    inv = np.empty(500, np.int64); inv[order] = np.arange(500)
    off, ids, ds = layers[0]
    for p in (0, 17, 499):
        want = [(int(inv[ids[j]]), float(ds[j])) for j in range(int(off[p]), int(off[p + 1]))]
        assert b["lists"][0][int(inv[p])] == want


class _DescriptionFFI(__import__("ctypes").Structure):
    """#[repr(C)] DescriptionFFI, libext.rs:1121-1141 (include/hnsw_b200.h)."""
    import ctypes as _C
    _fields_ = [("dumpmode", _C.c_uint8), ("max_nb_connection", _C.c_uint8), ("nb_layer", _C.c_uint8),
                ("ef", _C.c_size_t), ("nb_point", _C.c_size_t), ("data_dimension", _C.c_size_t),
                ("distname_len", _C.c_size_t), ("distname", _C.c_void_p),
                ("t_name_len", _C.c_size_t), ("t_name", _C.c_void_p)]


def test_library_reads_description_of_independent_dump(tmp_path, pkg, po):
    """load_hnsw_description (libext.rs:1170-1232) is host-only: the library's reader against a dump written by the
    independent writer of oracle/dumpfmt.py, and its refusal of a file that is not a dump."""
    import ctypes as C
    import dumpfmt
    X = pkg.datagen.uniform(300, 9, 4)
    o = po.Oracle(12, 300, 16, 50, "DistL1", 9)
    o.insert_batch(X, ids=np.arange(300))
    lv, rk, og = o.export_points()
    layers = [o.export_layer(l) for l in range(int(lv.max()) + 1)]
    base = str(tmp_path / "desc")
    dumpfmt.write_dump(base, X, og, lv, o.entry, layers, 12, 50, 1 / np.log(12), "DistL1")
    L = pkg.load_library()
    path = (base + ".hnsw.graph").encode()
    dptr = L.load_hnsw_description(len(path), path)
    assert dptr
    d = C.cast(dptr, C.POINTER(_DescriptionFFI)).contents
    assert (d.dumpmode, d.max_nb_connection, d.nb_layer) == (1, 12, 16)
    assert (d.ef, d.nb_point, d.data_dimension) == (50, 300, 9)
    assert C.string_at(d.distname, d.distname_len).endswith(b"DistL1")
    assert C.string_at(d.t_name, d.t_name_len) == b"f32"
    L.hnsw_b200_free_description(dptr)
    bad = tmp_path / "bad.hnsw.graph"
    bad.write_bytes(b"\x00" * 64)
    p2 = str(bad).encode()
    assert not L.load_hnsw_description(len(p2), p2)
    assert L.hnsw_b200_last_error()
    missing = str(tmp_path / "nope.hnsw.graph").encode()
    assert not L.load_hnsw_description(len(missing), missing)
