"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): batched insert, search,
filtered search, integer types — sizes kept tiny because the tools slow kernels down 10-100x."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("hnswlib-rs_b200")
X = pkg.datagen.clustered(3000, 128, 1)
h = pkg.Hnsw(16, 3000, 16, 64, "DistL2")
h.insert_flat(X)
Q = pkg.datagen.clustered(256, 128, 2)
r = h.search_flat(Q, 10, 64)
assert r[4].min() == 10
f = h.search_flat(Q[:32], 10, 32, filter=np.arange(0, 3000, 2))
assert np.all(f[0][f[4][:, None] > np.arange(10)[None, :]] % 2 == 0)
rng = np.random.default_rng(0)
U = rng.integers(0, 4, (1500, 40)).astype(np.uint8)
g = pkg.Hnsw(8, 1500, 16, 40, "DistHamming", dtype=np.uint8)
g.insert_flat(U)
g.search_flat(U[:128], 5, 32)
w = pkg.Hnsw(8, 500, 16, 40, "DistL2")
w.insert_flat(pkg.datagen.uniform(500, 784, 3))
w.search_flat(pkg.datagen.uniform(32, 784, 4), 5, 40)
print("sanitize_small ok")
