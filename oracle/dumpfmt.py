"""TEST INFRASTRUCTURE ONLY.  Independent restatement (struct packing) of the reference's dump format, used to
cross-check the engine's writer/reader (hnswlib-rs_b200/csrc/hnswio.cu).  PARITY UNPINNED: no dump written by the real
crate is available here (it cannot be built), so the format is pinned by this second implementation written
directly from the reference source:
  Description::dump        /root/reference/src/hnswio.rs:878-919   (v4: magic 0x002a6779)
  Hnsw::dump               /root/reference/src/hnswio.rs:1355-1387 (data header: 0xa67f0000 + dimension)
  PointIndexation::dump    /root/reference/src/hnswio.rs:1303-1340 (nb_layer u8; per layer 0x000a676f + count; entry point)
  dump_point               /root/reference/src/hnswio.rs:1063-1115 (0x000a678f, origin, level u8, rank i32, 16 lists)
Native endian, usize = 8 bytes.
"""
import struct

import numpy as np

MAGICPOINT, MAGICDESCR_4, MAGICLAYER, MAGICDATAP = 0x000A678F, 0x002A6779, 0x000A676F, 0xA67F0000
TNAME = {np.dtype(np.float32): "f32", np.dtype(np.uint8): "u8", np.dtype(np.uint16): "u16",
         np.dtype(np.uint32): "u32", np.dtype(np.int32): "i32"}


def write_dump(path_base, vecs, origin, levels, entry, layers, M, ef_c, level_scale, distname, nb_layer=16):
    """layers: list (index = layer) of (offsets, ids, dists) CSR over internal ids in insertion order."""
    vecs = np.ascontiguousarray(vecs)
    n, d = vecs.shape
    levels = np.asarray(levels)
    ranks = np.zeros(n, np.int32)
    by_level = [[] for _ in range(16)]
    for p in range(n):
        ranks[p] = len(by_level[levels[p]])
        by_level[levels[p]].append(p)
    g = bytearray()
    g += struct.pack("=IBBdBQQQ", MAGICDESCR_4, 1, M & 0xFF, level_scale, nb_layer, ef_c, n, d)
    dn = ("anndists::dist::distances::" + distname).encode()
    tn = TNAME[vecs.dtype].encode()
    g += struct.pack("=Q", len(dn)) + dn + struct.pack("=Q", len(tn)) + tn
    dat = bytearray(struct.pack("=IQ", MAGICDATAP, d))
    g += struct.pack("=B", nb_layer)
    for lay in range(nb_layer):
        g += struct.pack("=IQ", MAGICLAYER, len(by_level[lay]))
        for p in by_level[lay]:
            g += struct.pack("=IQBi", MAGICPOINT, int(origin[p]), int(levels[p]), int(ranks[p]))
            for l in range(16):
                if l < len(layers):
                    off, ids, ds = layers[l]
                    b, e = int(off[p]), int(off[p + 1])
                else:
                    b = e = 0
                g += struct.pack("=Q", e - b)
                for j in range(b, e):
                    q = int(ids[j])
                    g += struct.pack("=QBif", int(origin[q]), int(levels[q]), int(ranks[q]), float(ds[j]))
            raw = vecs[p].tobytes()
            dat += struct.pack("=IQQ", MAGICDATAP, int(origin[p]), len(raw)) + raw
    g += struct.pack("=QBi", int(origin[entry]), int(levels[entry]), int(ranks[entry]))
    open(path_base + ".hnsw.graph", "wb").write(bytes(g))
    open(path_base + ".hnsw.data", "wb").write(bytes(dat))


def read_dump(path_base, dtype):
    """-> dict(description..., origin, levels, ranks, vecs, entry (file-order index), lists[layer][point] = [(idx, dist)])
    Points are indexed in FILE order (layer by layer, rank order)."""
    g = open(path_base + ".hnsw.graph", "rb").read()
    dat = open(path_base + ".hnsw.data", "rb").read()
    pos = 0

    def take(fmt):
        nonlocal pos
        v = struct.unpack_from("=" + fmt, g, pos)
        pos += struct.calcsize("=" + fmt)
        return v if len(v) > 1 else v[0]
    magic, mode, M, scale, nb_layer, ef, n, d = take("IBBdBQQQ")
    assert magic == MAGICDESCR_4 and mode == 1
    ln = take("Q"); distname = g[pos:pos + ln].decode(); pos += ln
    ln = take("Q"); tname = g[pos:pos + ln].decode(); pos += ln
    assert tname == TNAME[np.dtype(dtype)]
    dpos = 0
    dm, dd = struct.unpack_from("=IQ", dat, dpos); dpos += 12
    assert dm == MAGICDATAP and dd == d
    nl = take("B")
    origin, levels, ranks, raw_lists, vecs = [], [], [], [], []
    start = []
    es = np.dtype(dtype).itemsize
    for lay in range(nl):
        m, cnt = take("IQ")
        assert m == MAGICLAYER
        start.append(len(origin))
        for j in range(cnt):
            m, oid, lv, rk = take("IQBi")
            assert m == MAGICPOINT and lv == lay and rk == j
            origin.append(oid); levels.append(lv); ranks.append(rk)
            pl = []
            for l in range(nb_layer):
                nn = take("Q")
                pl.append([take("QBif") for _ in range(nn)])
            raw_lists.append(pl)
            m2, oid2, blen = struct.unpack_from("=IQQ", dat, dpos); dpos += 20
            assert m2 == MAGICDATAP and oid2 == oid and blen == d * es
            vecs.append(np.frombuffer(dat, dtype, d, dpos).copy()); dpos += blen
    start.append(len(origin))
    e_oid, e_lv, e_rk = take("QBi")
    assert pos == len(g) and dpos == len(dat), "trailing bytes"
    idx = lambda lv, rk: start[lv] + rk
    lists = [[[(idx(lv, rk), ds) for (_o, lv, rk, ds) in raw_lists[p][l]] for p in range(len(origin))] for l in range(nb_layer)]
    return {"M": M, "level_scale": scale, "nb_layer": nb_layer, "ef": ef, "n": n, "d": d, "distname": distname,
            "t_name": tname, "origin": np.array(origin, np.uint64), "levels": np.array(levels, np.uint8),
            "ranks": np.array(ranks, np.int32), "vecs": np.array(vecs).reshape(len(origin), d), "entry": idx(e_lv, e_rk),
            "lists": lists}
