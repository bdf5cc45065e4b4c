"""Scratch measurement: GPU build time and search rate at a given size (not the bench)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
pkg = importlib.import_module("hnswlib-rs_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
kind = sys.argv[3] if len(sys.argv) > 3 else "clustered"
M, efc, k, ef, nq = 16, 200, 10, 64, 10000
t = time.time(); X = pkg.datagen.make(kind, n, d, 1); Q = pkg.datagen.make(kind, nq, d, 2); print("gen", time.time() - t, flush=True)
h = pkg.Hnsw(M, n, 16, efc, "DistL2")
t = time.time(); h.insert_flat(X); tb = time.time() - t
print(f"build {n}x{d}: {tb:.2f}s  {n/tb:.0f} inserts/s  max level {h.get_max_level_observed()}", flush=True)
h.enable_stats(True)
for it in range(4):
    t = time.time(); o, ds, ii, pid, cnt = h.search_flat(Q, k, ef); ts = time.time() - t
    st = h.get_stats()
    print(f"search_flat {nq} q: {ts*1e3:.2f} ms  {nq/ts:.0f} qps  evals/q {st['evals']/nq:.0f} exp/q {st['expansions']/nq:.1f} adj/q {st['adj_read']/nq:.0f}", flush=True)
t = time.time(); bi, bd = h.bruteforce(Q[:1000], k); print("bruteforce 1000 q", time.time() - t, flush=True)
rec = np.mean([len(set(ii[i, :cnt[i]].tolist()) & set(bi[i].tolist())) / k for i in range(1000)])
print("recall@10", rec, flush=True)
