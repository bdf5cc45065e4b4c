"""Scratch measurements for profiles/: insert-path counters + throughput, stand-alone distance kernel (K1), scale."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("hnswlib-rs_b200")
what = sys.argv[1]
if what == "insert":
    n, d = 1000000, 128
    X = pkg.datagen.clustered(n, d, 1)
    h = pkg.Hnsw(16, n, 16, 200, "DistL2")
    h.enable_stats(True)
    t = time.time(); h.insert_flat(X); dt = time.time() - t
    st = h.get_stats()
    E, A = st["evals"] / n, st["adj_read"] / n
    print(f"INSERT n={n}: {dt:.2f}s wall {n/dt:.0f} inserts/s  evals/insert {E:.0f} expansions/insert {st['expansions']/n:.1f} adj/insert {A:.0f}  "
          f"algorithmic bytes/insert {(E*d*4 + A*4)/1e6:.2f} MB", flush=True)
elif what == "dist":
    n, d, nq, m = 1000000, 128, 10000, 512
    X = pkg.datagen.clustered(n, d, 1)
    h = pkg.Hnsw(16, n, 16, 200, "DistL2")
    lv = np.zeros(n, np.uint8)
    off = np.zeros(n + 1, np.uint64)
    h.import_graph(X, np.arange(n, dtype=np.uint64), lv, 0, [(off, np.zeros(0, np.uint32), None)])   # point store only
    Q = pkg.datagen.clustered(nq, d, 2)
    cand = np.random.default_rng(3).integers(0, n, (nq, m)).astype(np.uint32)
    for _ in range(3):
        t = time.time(); out = h.dist_batch(Q, cand); dt = time.time() - t
    print(f"DIST_BATCH nq={nq} m={m}: {dt*1e3:.1f} ms wall incl. copies; rows {nq*m*d*4/1e9:.2f} GB", flush=True)
elif what == "scale":
    n, d = 10000000, 128
    t = time.time(); X = pkg.datagen.clustered(n, d, 1); print("gen", time.time() - t, flush=True)
    h = pkg.Hnsw(16, n, 16, 200, "DistL2")
    t = time.time(); h.insert_flat(X); dt = time.time() - t
    print(f"SCALE build n={n}: {dt:.1f}s  {n/dt:.0f} inserts/s", flush=True)
    Q = pkg.datagen.clustered(10000, d, 2)
    for _ in range(3):
        t = time.time(); r = h.search_flat(Q, 10, 64, with_internal=True, with_pid=False); ts = time.time() - t
    bi, bd = h.bruteforce(Q[:200], 10)
    rec = np.mean([len(set(r[2][i, :r[4][i]].tolist()) & set(bi[i].tolist())) / 10 for i in range(200)])
    print(f"SCALE search 10k q: {ts*1e3:.2f} ms e2e {10000/ts:.0f} qps recall@10 {rec:.3f}", flush=True)
