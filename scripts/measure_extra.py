"""Throughput of the paths beside the headline one (profiles/README.md "other paths"):
   filtered search on C2 (sorted-id filter admitting 30 % / 3 % of the points), u8 Hamming and u8 L2 (SIFT-like bytes, lean kernel),
   tie mode std.  Host calls with pinned buffers, best of 5."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("hnswlib-rs_b200")


def best(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return min(ts)


what = sys.argv[1:] or ["filter", "u8"]
if "filter" in what:
    n, d, nq = 1000000, 128, 10000
    X = pkg.datagen.clustered(n, d, 1)
    h = pkg.Hnsw(16, n, 16, 200, "DistL2"); h.insert_flat(X)
    Q = torch.from_numpy(pkg.datagen.clustered(nq, d, 2)).pin_memory().numpy()
    t = best(lambda: h.search_flat(Q, 10, 64, with_pid=False))
    print(f"EXTRA unfiltered C2 search_flat: {t*1e3:.2f} ms  {nq/t:.0f} q/s", flush=True)
    rng = np.random.default_rng(0)
    for frac in (0.3, 0.03):
        allow = np.sort(rng.choice(n, int(n * frac), replace=False)).astype(np.uint64)
        h.enable_stats(True); h.get_stats()
        r = h.search_flat(Q, 10, 64, filter=allow, with_pid=False)
        st = h.get_stats(); h.enable_stats(False)
        t = best(lambda: h.search_flat(Q, 10, 64, filter=allow, with_pid=False), 3)
        print(f"EXTRA filtered C2 ({frac:.0%} admitted): {t*1e3:.1f} ms  {nq/t:.0f} q/s  evals/query {st['evals']/nq:.0f} expansions/query "
              f"{st['expansions']/nq:.0f}  mean hits {r[4].mean():.2f}", flush=True)
if "u8" in what:
    n, d, nq = 1000000, 128, 10000
    Xf = pkg.datagen.clustered(n, d, 1)
    X8 = np.clip(Xf, 0, 255).astype(np.uint8)
    Q8 = np.clip(pkg.datagen.clustered(nq, d, 2), 0, 255).astype(np.uint8)
    for metric, Xs, Qs in (("DistL2", X8, Q8), ("DistHamming", (X8 >> 6), (Q8 >> 6))):
        h = pkg.Hnsw(16, n, 16, 200, metric, dtype=np.uint8)
        t0 = time.perf_counter(); h.insert_flat(Xs); tb = time.perf_counter() - t0
        Qp = torch.from_numpy(Qs).pin_memory().numpy()
        h.enable_stats(True); h.get_stats(); h.search_flat(Qp, 10, 64, with_pid=False); st = h.get_stats(); h.enable_stats(False)
        t = best(lambda: h.search_flat(Qp, 10, 64, with_pid=False))
        E, A = st["evals"] / nq, st["adj_read"] / nq
        print(f"EXTRA u8 {metric} 1M x 128: build {tb:.2f} s  search_flat {t*1e3:.2f} ms  {nq/t:.0f} q/s  E={E:.0f} A={A:.0f}  "
              f"algorithmic {(E*d + A*4)*nq/t/1e9:.0f} GB/s", flush=True)
        if metric == "DistHamming":
            h.set_tie_mode(1)
            t = best(lambda: h.search_flat(Qp, 10, 64, with_pid=False), 3)
            print(f"EXTRA u8 DistHamming tie mode std: search_flat {t*1e3:.1f} ms  {nq/t:.0f} q/s", flush=True)
