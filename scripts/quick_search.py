"""Scratch A/B measurement of the search kernel (not the bench): builds the C2 graph once, times launches."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("hnswlib-rs_b200")
n = int(os.environ.get("N", 1000000)); d = int(os.environ.get("D", 128)); nq = int(os.environ.get("NQ", 10000))
M = int(os.environ.get("M", 16)); efc = int(os.environ.get("EFC", 200)); ef = int(os.environ.get("EF", 64)); k = 10
kind = os.environ.get("KIND", "clustered"); metric = os.environ.get("METRIC", "DistL2")
X = pkg.datagen.make(kind, n, d, 1)
h = pkg.Hnsw(M, n, 16, efc, metric)
t = time.time(); h.insert_flat(X); print(f"build {time.time()-t:.2f}s", flush=True)
nqs = [int(x) for x in os.environ.get("NQS", str(nq)).split(",")]
for nq in nqs:
  qs = [torch.from_numpy(pkg.datagen.make(kind, nq, d, 2 + b)).cuda() for b in range(4)]
  out = torch.empty((nq, k, 16), dtype=torch.uint8, device="cuda"); cnt = torch.empty((nq,), dtype=torch.int32, device="cuda")
  h.enable_stats(True); h.search_device(qs[0].data_ptr(), nq, k, ef, out.data_ptr(), cnt.data_ptr(), True); st = h.get_stats(); h.enable_stats(False)
  E, A = st["evals"] / nq, st["adj_read"] / nq
  bpq = E * d * 4 + A * 4 + d * 4 + k * 16
  for i in range(5): h.search_device(qs[i % 4].data_ptr(), nq, k, ef, out.data_ptr(), cnt.data_ptr(), True)
  ms = [h.search_device(qs[i % 4].data_ptr(), nq, k, ef, out.data_ptr(), cnt.data_ptr(), True) for i in range(20)]
  m = float(np.mean(ms))
  print(f"RESULT n={n} d={d} ef={ef} M={M} nq={nq}: kernel {m:.3f} ms (min {min(ms):.3f})  {nq/m*1e3:.0f} qps  E={E:.0f} A={A:.0f}  {bpq*nq/m/1e6:.0f} GB/s  frac {bpq*nq/m/1e6/6573.8:.3f}", flush=True)
