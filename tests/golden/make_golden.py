"""Generates tests/golden/oracle_small.npz.

NOT reference-generated: the reference (Rust) cannot be built or imported in this environment, so these vectors are
produced by the CPU oracle in its LITERAL reference mode (MODE_STD heaps, ORDER_REF sums).  They pin the oracle (and
through it the engine) against accidental drift between rounds; they do not pin the oracle against the real crate.
Run:  python tests/golden/make_golden.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle as po  # noqa: E402

pkg = importlib.import_module("hnswlib-rs_b200")


def build():
    n, d, M, efc = 2000, 25, 16, 200            # BASELINE configs[0] shape (examples/random.rs), scaled down
    X = pkg.datagen.uniform(n, d, 1)
    Q = pkg.datagen.uniform(64, d, 2)
    o = po.Oracle(M, n, 9, efc, "DistL2", d, mode=po.MODE_STD, order=po.ORDER_REF)
    levels = o.draw_levels(n)
    o.insert_batch(X, levels=levels)
    og, ds, it, pid, cnt = o.search_batch(Q, 10, 24)
    off, ids, eds = o.export_layer(0)
    fo, fd, fi, _, fc = o.search_batch(Q, 10, 24, filter_ids=np.arange(0, n, 3))
    return dict(n=n, d=d, M=M, efc=efc, levels=levels, entry=o.entry, ids=it, dists=ds, counts=cnt, pid=pid,
                l0_off=off, l0_ids=ids, l0_dists=eds, f_ids=fi, f_dists=fd, f_counts=fc)


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_small.npz")
    np.savez_compressed(out, **build())
    print("wrote", out, os.path.getsize(out), "bytes")
