#!/usr/bin/env python
"""bench.py — queries/s of HNSW search at matched recall@10 (BASELINE.json metric).

One "step" = one pass of the hot path (greedy descent + ef-bounded layer-0 expansion,
reference src/hnsw.rs:1487-1580 + 922-1064) over one batch of `--nq` synthetic queries against a
graph of `--n` points built on the GPU by this engine.

Default workload = BASELINE.json configs[1]: SIFT1M-shape synthetic, 1M x d=128 f32 L2, M=16,
ef_construction=200, 10k queries, k=10, ef=64, one GPU.

  value      queries/s, inputs resident in HBM, K launches timed with CUDA events on the launch stream
  e2e        same metric through the C-ABI call a user makes (hnsw_b200_search_flat) with HOST buffers:
             H2D of the queries and D2H of the answers inside the timed region
  roofline   algorithmic bytes (E*d*4 + A*4 + d*4 + k*16 per query, E/A counted by the kernel itself and
             equal to the oracle's counters, tests/test_gpu_search.py) / kernel time vs measured HBM peak
  cpu_baseline  the CPU restatement (oracle, MODE_STD + reference-shaped sums) of the same path on the SAME
             graph and queries, all host threads, bounded sample
  --impl reference   the CPU path alone (oracle-built graph with all host threads, then timed searches)

Multi-GPU (torchrun, one process per GPU): rank 0 builds, the frozen index is broadcast over NCCL, every
rank searches its own shard of nq queries per step (weak scaling), answers are all-gathered over NCCL
inside the timed region.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=1000000)
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--nq", type=int, default=10000)
    ap.add_argument("--M", type=int, default=16)
    ap.add_argument("--efc", type=int, default=200)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--ef", type=int, default=64)
    ap.add_argument("--metric", default="DistL2")
    ap.add_argument("--data", default="clustered", choices=["clustered", "uniform", "unit"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.stop = threading.Event()
        self.th = None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            hd = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(hd, nv.NVML_CLOCK_SM)
            bits = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}
            while not self.stop.is_set():
                sm = nv.nvmlDeviceGetClockInfo(hd, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(hd)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(hd)
                pw = nv.nvmlDeviceGetPowerUsage(hd) / 1000.0
                self.rows.append([str(sm), str(mx), str(pw)] + ["Active" if r & bits[k] else "Not Active" for k in
                                 ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])
                self.stop.wait(0.01)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        pw = max(float(r[2]) for r in self.rows)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": sorted(reasons),
                "samples": len(self.rows), "power_w_max": pw}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recall_stats(ids, dists, counts, t_ids, t_d):
    k = t_ids.shape[1]
    rid = rball = 0.0
    for i in range(t_ids.shape[0]):
        rid += len(set(ids[i, :counts[i]].tolist()) & set(t_ids[i].tolist())) / k
        rball += float(np.sum(dists[i, :counts[i]] <= t_d[i, k - 1])) / k   # the reference's recall definition
    return rid / t_ids.shape[0], rball / t_ids.shape[0]


def default_workload(a):
    return (a.n, a.d, a.nq, a.M, a.efc, a.k, a.ef, a.metric, a.data) == (1000000, 128, 10000, 16, 200, 10, 64, "DistL2", "clustered")


def workload_name(a):
    return (f"SIFT1M-shape synthetic ({a.data}): {a.n} x d={a.d} f32 {a.metric}, M={a.M} ef_c={a.efc}, "
            f"{a.nq} queries/step/GPU k={a.k} ef={a.ef}")


# ------------------------------------------------------------------------------------------- reference arm
def run_reference(a, rank, world):
    """The reference's own CPU implementation of the path.  The Rust crate cannot be built on this box
    (no cargo/rustc), so this times the CPU restatement (oracle/, kind "port") in its literal mode:
    racy parallel insert with every host thread (hnsw.rs:1224-1238), then parallel_search
    (hnsw.rs:1612-1635) of the same query batch, MODE_STD heaps + reference-shaped SIMD sums."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    pkg = importlib.import_module("hnswlib-rs_b200")
    po.build()
    cores = os.cpu_count() or 1
    X = pkg.datagen.make(a.data, a.n, a.d, 1)
    Q = pkg.datagen.make(a.data, a.nq, a.d, 2)
    o = po.Oracle(a.M, a.n, 16, a.efc, a.metric, a.d, mode=po.MODE_STD, order=po.ORDER_REF)
    t0 = time.perf_counter()
    o.insert_batch(X, nthreads=cores)
    build_s = time.perf_counter() - t0
    # all logical CPUs vs one thread per physical core: keep the faster for the timed steps
    best_t, best_q = cores, 0.0
    for nt in sorted({cores, max(1, cores // 2)}, reverse=True):
        o.search_batch(Q, a.k, a.ef, nthreads=nt)
        t0 = time.perf_counter()
        o.search_batch(Q, a.k, a.ef, nthreads=nt)
        q = a.nq / (time.perf_counter() - t0)
        if q > best_q:
            best_t, best_q = nt, q
    threads = best_t
    for _ in range(a.warmup):
        o.search_batch(Q, a.k, a.ef, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        res = o.search_batch(Q, a.k, a.ef, nthreads=threads)
    dt = time.perf_counter() - t0
    qps = a.steps * a.nq / dt
    nt = min(1000, a.nq)
    ti, td = po.bruteforce(X, Q[:nt], a.k, a.metric)
    # origin ids == row numbers of X; internal ids are NOT (a racy parallel insert numbers points in arrival order)
    rid, rball = recall_stats(res[0][:nt], res[1][:nt], res[4][:nt], ti, td)
    line = {
        "impl": "reference", "metric": "queries/sec @ recall@10", "value": qps, "unit": "queries/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "recall_at_10": rid, "recall_at_10_ball": rball,
                   "graph": "built by the CPU restatement (parallel insert, all host threads)", "build_s": build_s},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": f"{a.steps} x {a.nq} queries, full batch each step, {threads} threads "
                                   f"(best of all logical CPUs / half); graph built with {cores} threads"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------- our arm
def run_ours(a, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("hnswlib-rs_b200")
    replicate = importlib.import_module("hnswlib-rs_b200.replicate")
    dev = local_rank
    torch.cuda.set_device(dev)
    multi = world > 1
    t_setup = time.perf_counter()

    # ---- graph: rank 0 builds on its GPU; replicas receive the frozen arrays over NCCL
    h = pkg.Hnsw(a.M, a.n, 16, a.efc, a.metric, device=dev)
    build_s = 0.0
    X = None
    if rank == 0:
        X = pkg.datagen.make(a.data, a.n, a.d, 1)
        t0 = time.perf_counter()
        h.insert_flat(X)
        build_s = time.perf_counter() - t0
    bcast_s = 0.0
    if multi:
        t0 = time.perf_counter()
        replicate.broadcast_index(h, 0, f"cuda:{dev}")      # ncclBroadcast of the frozen arrays
        bcast_s = time.perf_counter() - t0

    # ---- queries: NB rotating batches per rank, seeded per rank; pinned host copies + device copies
    NB = 4
    q_host = [torch.from_numpy(pkg.datagen.make(a.data, a.nq, a.d, 2 + 1000 * rank + b)).pin_memory() for b in range(NB)]
    q_dev = [q.cuda(non_blocking=True) for q in q_host]
    out_dev = torch.empty((a.nq, a.k, 16), dtype=torch.uint8, device="cuda")        # Neighbour_api[nq][k]
    cnt_dev = torch.empty((a.nq,), dtype=torch.int32, device="cuda")
    gather_dev = torch.empty((world * a.nq, a.k, 16), dtype=torch.uint8, device="cuda") if multi else None
    torch.cuda.synchronize()
    # a dedicated (non-default) torch stream: the library's kernels, torch's events and the NCCL collectives all
    # go through it, so torch.cuda.Event brackets exactly the launches of the timed region
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    h.set_stream(stream.cuda_stream)

    def step_device(b, sync=False):
        ms = h.search_device(q_dev[b % NB].data_ptr(), a.nq, a.k, a.ef, out_dev.data_ptr(), cnt_dev.data_ptr(), sync=sync)
        if multi:
            dist.all_gather_into_tensor(gather_dev, out_dev)   # ncclAllGather of the answers
        return ms

    # ---- one instrumented pass: traversal counters (algorithmic bytes) and recall vs exact brute force
    h.enable_stats(True)
    step_device(0, sync=True)
    st = h.get_stats()
    h.enable_stats(False)
    E, A = st["evals"] / a.nq, st["adj_read"] / a.nq
    bytes_per_query = E * a.d * 4 + A * 4 + a.d * 4 + a.k * 16
    rid = rball = None
    if rank == 0:
        nt = min(1000, a.nq)
        qn = q_host[0][:nt].numpy()
        ti, td = h.bruteforce(qn, a.k)                      # exact ground truth (K5 kernel)
        o_, d_, i_, _, c_ = h.search_flat(qn, a.k, a.ef)
        rid, rball = recall_stats(i_, d_, c_, ti, td)

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident: W warm-up steps, then exactly K steps between CUDA events on the launch stream
    for i in range(a.warmup):
        step_device(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev) as clk:
        barrier()
        e0.record(stream)
        for i in range(a.steps):
            step_device(i)
        e1.record(stream)
        barrier()
    dev_ms = e0.elapsed_time(e1)
    if h.check_status() != 0:
        raise RuntimeError("visited table overflow during the timed region")
    clocks = clk.summary()
    # per-launch kernel duration (CUDA events inside the library, around the kernel alone)
    kms = [h.search_device(q_dev[i % NB].data_ptr(), a.nq, a.k, a.ef, out_dev.data_ptr(), cnt_dev.data_ptr(), sync=True)
           for i in range(min(a.steps, 10))]
    kernel_ms = float(np.mean(kms))

    # ---- steady-state reference point (not the metric): one launch of 10x the step's queries, where ramp-up and tail of
    # the 2-wave step no longer dominate (BASELINE configs[4] runs 125 000 queries per GPU, i.e. in this regime)
    steady = None
    if rank == 0 and not multi:
        nbig = 10 * a.nq
        qb = torch.from_numpy(pkg.datagen.make(a.data, nbig, a.d, 777)).cuda()
        ob = torch.empty((nbig, a.k, 16), dtype=torch.uint8, device="cuda")
        cb = torch.empty((nbig,), dtype=torch.int32, device="cuda")
        h.search_device(qb.data_ptr(), nbig, a.k, a.ef, ob.data_ptr(), cb.data_ptr(), sync=True)
        ms_big = min(h.search_device(qb.data_ptr(), nbig, a.k, a.ef, ob.data_ptr(), cb.data_ptr(), sync=True) for _ in range(3))
        steady = {"queries_per_launch": nbig, "kernel_ms": ms_big, "queries_per_s": nbig / ms_big * 1e3,
                  "algorithmic_GBps": bytes_per_query * nbig / ms_big / 1e6}
        del qb, ob, cb

    # ---- end to end through the C-ABI call with host buffers (H2D + kernel + D2H per step)
    gloo = dist.new_group(backend="gloo") if multi else None
    qh_np = [q.numpy() for q in q_host]
    h.set_stream(None)

    def step_e2e(b):
        res = h.search_flat(qh_np[b % NB], a.k, a.ef, with_internal=False, with_pid=False)   # ids + distances + counts
        if multi:   # answers to rank 0 (host side, gloo)
            t = torch.from_numpy(res[0].view(np.int64))   # gloo has no uint64
            gl = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
            dist.gather(t, gl, dst=0, group=gloo)
        return res

    for i in range(a.warmup):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(a.steps):
        step_e2e(i)
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- max over ranks
    if multi:
        t = torch.tensor([dev_ms, e2e_s * 1e3, kernel_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms, kernel_ms = t.tolist()
        e2e_s = e2e_ms / 1e3
    total_q = a.steps * a.nq * world
    value = total_q / (dev_ms / 1e3)
    e2e_v = total_q / e2e_s

    if rank != 0:
        return
    peak, peak_src = measured_peak_gbs()
    achieved = bytes_per_query * a.nq / (kernel_ms / 1e3) / 1e9
    line = {
        "metric": "queries/sec @ recall@10", "value": value, "unit": "queries/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": workload_name(a), "recall_at_10": rid, "recall_at_10_ball": rball,
            "graph": "built on the GPU by this engine (batched insert)", "build_s": build_s,
            "index_broadcast_s": bcast_s, "evals_per_query": E, "adj_ids_per_query": A,
            "l2": f"no flush: working set (point store + adjacency {a.n * (a.d * 4 + a.M * 8) / 1e6:.0f} MB) exceeds the 126 MB L2; "
                  f"{NB} query batches rotated",
            "parallelism": f"query-sharded x{world}, index replicated (ncclBroadcast), answers ncclAllGather" if multi else "1 GPU",
        },
        "clocks": clocks,
        "e2e": {"value": e2e_v, "unit": "queries/s", "h2d_bytes_per_step": a.nq * a.d * 4,
                "d2h_bytes_per_step": a.nq * a.k * 16 + a.nq * 4},
        "gpu_launches": a.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": TRAFFIC_BYTES_PER_LAUNCH if default_workload(a) else None, "peak_source": peak_src, "kernel": "search_kernel",
                     "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": bytes_per_query * a.nq},
    }
    if steady:
        steady["frac_of_peak"] = steady["algorithmic_GBps"] / peak
        line["config"]["steady_state"] = steady
    if world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(a, h, qh_np[0])
    line["config"]["setup_s"] = time.perf_counter() - t_setup
    emit(line)


# dram__bytes_read.sum + dram__bytes_write.sum of one search_kernel launch on the DEFAULT workload, from the
# ncu --set full capture summarised in profiles/ (reported only when the run uses the default workload)
TRAFFIC_BYTES_PER_LAUNCH = 6.407e9   # profiles/r1_final_search_kernel_ncu_full_selected.csv: 6.078 GB read + 0.329 GB written


def cpu_baseline(a, h, Q):
    """CPU restatement (oracle, kind "port") of the same path on the SAME graph and queries: literal reference
    mode (Rust-std heaps, AVX2-shaped sums), one query per task over all host threads like rayon par_iter."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    po.build()
    cores = os.cpu_count() or 1
    lv, rk, og, entry = h.export_points()
    # the graph is imported by ONE thread: interleave its pages over the NUMA nodes, as the reference's own
    # multi-threaded build would spread them by first touch (otherwise every search thread hammers one node)
    numa = po.numa_interleave(True)
    o = po.Oracle(a.M, a.n, 16, a.efc, a.metric, a.d, mode=po.MODE_STD, order=po.ORDER_REF)
    maxl = int(lv.max()) + 1 if len(lv) else 1
    o.import_graph(h.export_vectors(), og, lv, entry, {l: h.export_layer(l) for l in range(min(16, maxl + 1))})
    po.numa_interleave(False)
    # all logical CPUs vs one thread per physical core (hyper-threads share the load units): keep the faster
    best_t, best_q = cores, 0.0
    for nt in sorted({cores, max(1, cores // 2)}, reverse=True):
        o.search_batch(Q, a.k, a.ef, nthreads=nt)
        t0 = time.perf_counter()
        o.search_batch(Q, a.k, a.ef, nthreads=nt)
        q = a.nq / (time.perf_counter() - t0)
        if q > best_q:
            best_t, best_q = nt, q
    cores = best_t
    reps, t0 = 0, time.perf_counter()
    while True:
        o.search_batch(Q, a.k, a.ef, nthreads=cores)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= a.cpu_seconds or reps >= 2000:
            break
    return {"value": reps * a.nq / dt, "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": f"{reps} passes over the same {a.nq}-query batch on the GPU-built graph ({dt:.1f} s wall), "
                      f"{cores} threads (best of all logical CPUs / half), numa_interleave={'on' if numa == 0 else 'refused'}"}


class _OnlyJsonOnStdout:
    """Everything a library prints to stdout while the benchmark runs (NCCL's version banner, ...) goes to stderr, so
    that rank 0's stdout carries exactly one line: the JSON result."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def emit(line):
    """called inside _OnlyJsonOnStdout: write the JSON line to the REAL stdout"""
    sys.stdout.flush()
    os.write(_REAL_STDOUT[0], (json.dumps(line) + "\n").encode())


_REAL_STDOUT = [1]


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    guard = _OnlyJsonOnStdout()
    guard.__enter__()
    _REAL_STDOUT[0] = guard.saved
    try:
        _main_guarded(a, rank, world, local_rank)
    finally:
        guard.__exit__()


def _main_guarded(a, rank, world, local_rank):
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device(f"cuda:{local_rank}"))
    try:
        run_ours(a, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
