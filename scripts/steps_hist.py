"""Distribution of expansions per query (one single-query launch per query, traversal counters read back)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("hnswlib-rs_b200")
n, d, nq, M, efc, ef, k = 1000000, 128, int(os.environ.get("NQ", 3000)), 16, 200, 64, 10
X = pkg.datagen.make("clustered", n, d, 1)
h = pkg.Hnsw(M, n, 16, efc, "DistL2"); h.insert_flat(X)
Q = torch.from_numpy(pkg.datagen.make("clustered", nq, d, 2)).cuda()
out = torch.empty((1, k, 16), dtype=torch.uint8, device="cuda"); cnt = torch.empty((1,), dtype=torch.int32, device="cuda")
h.enable_stats(True); h.get_stats()
ex = []; ev = []
for i in range(nq):
    h.search_device(Q[i:i+1].data_ptr(), 1, k, ef, out.data_ptr(), cnt.data_ptr(), True)
    st = h.get_stats(); ex.append(st["expansions"]); ev.append(st["evals"])
ex = np.array(ex); ev = np.array(ev)
print("expansions/query: mean %.1f p50 %d p90 %d p99 %d max %d" % (ex.mean(), *np.percentile(ex, [50, 90, 99]).astype(int), ex.max()))
print("evals/query:      mean %.1f p50 %d p90 %d p99 %d max %d" % (ev.mean(), *np.percentile(ev, [50, 90, 99]).astype(int), ev.max()))
