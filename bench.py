#!/usr/bin/env python
"""bench.py — queries/s of HNSW search at matched recall@10 (BASELINE.json metric).

One "step" = one pass of the hot path (greedy descent + ef-bounded layer-0 expansion,
reference src/hnsw.rs:1487-1580 + 922-1064) over one batch of synthetic queries against a graph built on the GPU by
this engine.  `--config` picks the workload (BASELINE.json configs[0..4]); the default is c2, the configuration the
metric is quoted on:

  c1  random.rs shape: 10 000 x d=25 f32 L2, M=16 ef_c=200, 1 000 queries k=10 ef=24
  c2  SIFT1M shape: 1 000 000 x d=128 f32 L2, M=16 ef_c=200, 10 000 queries k=10 ef=64   (default)
  c3  GloVe-25 shape: 1 183 514 x d=25 unit vectors, DistCosine (angular), M=24 ef_c=800, 10 000 queries k=10 ef=128
  c3dot  same with DistDot on the normalised vectors, what the reference's own example runs "to spare cpu"
  c4  MNIST-784 shape: 60 000 x d=784 f32 L2, M=32 ef_c=400, 10 000 queries k=10 ef=200
  c5  c2's graph, 1 000 000 queries per step query-sharded over the GPUs (strong scaling)

  value      queries/s, inputs resident in HBM, K launches timed with CUDA events on the launch stream
  e2e        the same metric through the C-ABI call a user makes with HOST buffers, H2D of the queries and D2H of the
             answers inside the timed region: hnsw_b200_search_flat_submit / _wait with two batches in flight from one
             host thread, per rank (e2e.sequential: one hnsw_b200_search_flat call at a time; e2e.one_process at N > 1: ONE
             process, rank 0's handle after hnsw_b200_replicate, the library sharding every batch over the N GPUs).
             e2e.row_pointers (N = 1): the reference's own entry point
             parallel_search_neighbours_f32 with pageable row pointers and malloc'ed answers
  roofline   algorithmic bytes (E*d*4 + A*4 + d*4 + k*16 per query, E/A counted by the kernel itself and
             equal to the oracle's counters, tests/test_gpu_search.py) / kernel time vs measured HBM peak
  cpu_baseline  the CPU restatement (oracle, MODE_STD + reference-shaped sums) of the same path on the SAME
             graph and queries, all host threads, bounded sample
  --impl reference   the CPU path alone (oracle-built graph with all host threads, then timed searches)

Multi-GPU (torchrun, one process per GPU): rank 0 builds; the library itself opens an NCCL communicator
(hnsw_b200_nccl_init, the unique id travels over torch.distributed) and broadcasts the frozen index
(hnsw_b200_nccl_broadcast_index); every rank searches its own shard (no data-path collective); the answers of step i are
all-gathered (hnsw_b200_nccl_allgather) on a second stream while step i+1 searches.  Every replica answers a shared probe
batch and is compared with rank 0 before anything is timed.
"""
import argparse
import ctypes
import importlib
import json
import os
import platform
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PRESETS = {
    "c1": dict(n=10000, d=25, nq=1000, M=16, efc=200, k=10, ef=24, metric="DistL2", data="uniform",
               name="C1 random.rs shape"),
    "c2": dict(n=1000000, d=128, nq=10000, M=16, efc=200, k=10, ef=64, metric="DistL2", data="clustered",
               name="C2 SIFT1M-shape synthetic"),
    "c3": dict(n=1183514, d=25, nq=10000, M=24, efc=800, k=10, ef=128, metric="DistCosine", data="unit",
               name="C3 GloVe-25-shape synthetic (angular)"),
    "c3dot": dict(n=1183514, d=25, nq=10000, M=24, efc=800, k=10, ef=128, metric="DistDot", data="unit",
                  name="C3 GloVe-25-shape synthetic (angular as DistDot on unit vectors)"),
    "c4": dict(n=60000, d=784, nq=10000, M=32, efc=400, k=10, ef=200, metric="DistL2", data="uniform",
               name="C4 MNIST-784-shape synthetic"),
    "c5": dict(n=1000000, d=128, nq=1000000, M=16, efc=200, k=10, ef=64, metric="DistL2", data="clustered",
               name="C5 SIFT1M-shape synthetic, 1M queries query-sharded"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(PRESETS))
    for name, typ in (("n", int), ("d", int), ("nq", int), ("M", int), ("efc", int), ("k", int), ("ef", int), ("metric", str),
                      ("data", str)):
        ap.add_argument("--" + name, type=typ, default=None, help="override the preset")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    preset = PRESETS[a.config]
    a.custom = False
    for key, val in preset.items():
        if key == "name":
            continue
        if getattr(a, key) is None:
            setattr(a, key, val)
        elif getattr(a, key) != val:
            a.custom = True
    a.preset_name = preset["name"]
    big = a.config == "c5"
    if a.steps is None:
        a.steps = 5 if big else 100
    if a.warmup is None:
        a.warmup = 3 if big else 10
    a.strong = big
    return a


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



def workload_name(a, world=1):
    per = a.nq // world if a.strong else a.nq
    q = (f"{a.nq} queries/step over {world} GPU(s) ({per}/GPU)" if a.strong else f"{a.nq} queries/step/GPU")
    return (f"{a.preset_name}{' (custom overrides)' if a.custom else ''} ({a.data}): {a.n} x d={a.d} f32 {a.metric}, M={a.M} "
            f"ef_c={a.efc}, {q} k={a.k} ef={a.ef}")


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor() or "unknown"


def committed_traffic(a):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the search kernel on the default workload, read
    from the ncu --set full summary committed under profiles/ (None for any other workload)."""
    if a.config != "c2" or a.custom:
        return None, None
    path = os.path.join(ROOT, "profiles", "r2_search_lean_kernel_ncu_full_selected.csv")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    try:
        tot = 0.0
        for ln in open(path):
            f = ln.strip().split(",")
            if f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(f[2]) * scale[f[1]]
        return (tot or None), os.path.relpath(path, ROOT)
    except (OSError, ValueError, KeyError, IndexError):
        return None, None


def best_threads(o, Q, a, cores):
    """all logical CPUs vs one thread per physical core (hyper-threads share the load units): keep the faster"""
    best_t, best_q = cores, 0.0
    for nt in sorted({cores, max(1, cores // 2)}, reverse=True):
        o.search_batch(Q, a.k, a.ef, nthreads=nt)
        t0 = time.perf_counter()
        o.search_batch(Q, a.k, a.ef, nthreads=nt)
        q = len(Q) / (time.perf_counter() - t0)
        if q > best_q:
            best_t, best_q = nt, q
    return best_t


# ------------------------------------------------------------------------------------------- reference arm
def run_reference(a, rank, world):
    """The reference's own CPU implementation of the path.  The Rust crate cannot be built on this box
    (no cargo/rustc), so this times the CPU restatement (oracle/, kind "port") in its literal mode:
    racy parallel insert with every host thread (hnsw.rs:1224-1238), then parallel_search
    (hnsw.rs:1612-1635) of a query batch, MODE_STD heaps + reference-shaped SIMD sums."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    pkg = importlib.import_module("hnswlib-rs_b200")
    po.build()
    cores = os.cpu_count() or 1
    X = pkg.datagen.make(a.data, a.n, a.d, 1)
    nq = min(a.nq, 10000)   # a bounded sample of the step (c5's step is 1M queries)
    Q = pkg.datagen.make(a.data, nq, a.d, 2)
    o = po.Oracle(a.M, a.n, 16, a.efc, a.metric, a.d, mode=po.MODE_STD, order=po.ORDER_REF)
    t0 = time.perf_counter()
    o.insert_batch(X, nthreads=cores)
    build_s = time.perf_counter() - t0
    threads = best_threads(o, Q, a, cores)
    for _ in range(a.warmup):
        o.search_batch(Q, a.k, a.ef, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        res = o.search_batch(Q, a.k, a.ef, nthreads=threads)
    dt = time.perf_counter() - t0
    qps = a.steps * nq / dt
    nt = min(1000, nq)
    ti, td = po.bruteforce(X, Q[:nt], a.k, a.metric)
    # origin ids == row numbers of X; internal ids are NOT (a racy parallel insert numbers points in arrival order)
    rid, rball = recall_stats(res[0][:nt], res[1][:nt], res[4][:nt], ti, td)
    line = {
        "impl": "reference", "metric": "queries/sec @ recall@10", "value": qps, "unit": "queries/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if a.strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a, a.gpus), "recall_at_10": rid, "recall_at_10_ball": rball,
                   "graph": "built by the CPU restatement (parallel insert, all host threads)", "build_s": build_s},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "threads": threads, "kind": "port",
                         "cpu_model": cpu_model(), "logical_cpus": cores, "numa_interleave": "first touch by the build threads",
                         "sample": f"{a.steps} x {nq} queries per step, {threads} threads (best of all logical CPUs / half); "
                                   f"graph built with {cores} threads"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------- our arm
def note(msg):
    """progress marks on stderr (BENCH_VERBOSE=1)"""
    if os.environ.get("BENCH_VERBOSE"):
        sys.stderr.write(f"[bench r{os.environ.get('RANK', '0')} {time.strftime('%H:%M:%S')}] {msg}\n")
        sys.stderr.flush()


def shard_bounds(n, rank, world):
    """contiguous shard [lo, hi) of n items for `rank` (sizes differ by at most one)"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def run_ours(a, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("hnswlib-rs_b200")
    dev = local_rank
    torch.cuda.set_device(dev)
    multi = world > 1
    t_setup = time.perf_counter()
    lo, hi = shard_bounds(a.nq, rank, world) if a.strong else (0, a.nq)
    nq = hi - lo                                   # this rank's queries per step
    total_per_step = a.nq if a.strong else a.nq * world

    # ---- graph: rank 0 builds on its GPU; the library broadcasts the frozen index over its own NCCL communicator
    h = pkg.Hnsw(a.M, a.n, 16, a.efc, a.metric, device=dev)
    build_s = bcast_s = 0.0
    if rank == 0:
        X = pkg.datagen.make(a.data, a.n, a.d, 1)
        t0 = time.perf_counter()
        h.insert_flat(X)
        build_s = time.perf_counter() - t0
        del X
    note('built')
    if multi:
        uid = torch.from_numpy(pkg.Hnsw.nccl_unique_id() if rank == 0 else np.zeros(128, np.uint8)).cuda()
        dist.broadcast(uid, 0)                     # 128 bytes of ncclUniqueId, by the host's own means
        t0 = time.perf_counter()
        note('uid exchanged')
        h.nccl_init(world, rank, uid.cpu().numpy())
        note('comm up')
        h.nccl_broadcast_index(0)                  # ncclBroadcast of header + 9 device arrays, inside the library
        bcast_s = time.perf_counter() - t0
        note('index broadcast')
        # every replica answers a shared probe batch; the answers must equal rank 0's
        probe = pkg.datagen.make(a.data, 512, a.d, 4242)
        ids = torch.from_numpy(h.search_flat(probe, a.k, a.ef, with_pid=False)[2].astype(np.int64)).cuda()
        ref = ids.clone()
        dist.broadcast(ref, 0)
        same = torch.tensor([int(torch.equal(ids, ref))], device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        if int(same.item()) != 1:
            raise RuntimeError("a replica's answers differ from rank 0's")

    note('replicas checked')
    # ---- queries: NB rotating batches per rank, seeded per rank; pinned host copies + device copies
    NB = 4 if nq <= 20000 else 1
    q_host = [torch.from_numpy(pkg.datagen.make(a.data, nq, a.d, 2 + 1000 * rank + b)).pin_memory() for b in range(NB)]
    q_dev = [q.cuda(non_blocking=True) for q in q_host]
    # Answers live in groups of GSTEP consecutive steps (one contiguous buffer per group, NGRP groups rotating): the answers of
    # a whole group are all-gathered by ONE collective on a second stream while the next groups search, so the GPUs
    # rendezvous once per GSTEP steps instead of every step.
    GSTEP, NGRP = 2, 3
    out_grp = [torch.empty((GSTEP, nq, a.k, 16), dtype=torch.uint8, device="cuda") for _ in range(NGRP)]   # Neighbour_api[nq][k] x GSTEP
    out_dev = [out_grp[0][0]]                       # kernel-only timings and the single-GPU path write here
    cnt_dev = torch.empty((nq,), dtype=torch.int32, device="cuda")
    same_shards = (not a.strong) or a.nq % world == 0
    gather_grp = [torch.empty((world, GSTEP, nq, a.k, 16), dtype=torch.uint8, device="cuda") for _ in range(NGRP)] if multi and same_shards else None
    gather_dev = gather_grp
    torch.cuda.synchronize()
    # a dedicated (non-default) torch stream: the library's kernels and torch's events go through it, so
    # torch.cuda.Event brackets exactly the launches of the timed region; the all-gathers run on a second stream
    stream = torch.cuda.Stream(device=dev)
    gstream = torch.cuda.Stream(device=dev, priority=-1)   # the gather's few CTAs go first when SM slots free up
    torch.cuda.set_stream(stream)
    h.set_stream(stream.cuda_stream)
    ev_gath = [torch.cuda.Event() for _ in range(NGRP)]
    open_group = [None]                            # group with launches not gathered yet

    def gather_group(g):
        h.nccl_allgather(out_grp[g].data_ptr(), gather_grp[g].data_ptr(), GSTEP * nq * a.k * 16, gstream.cuda_stream)
        ev_gath[g].record(gstream)
        open_group[0] = None

    def step_device(i, sync=False):
        g, j = (i // GSTEP) % NGRP, i % GSTEP
        if gather_grp is not None and j == 0 and i >= GSTEP * NGRP:
            stream.wait_event(ev_gath[g])          # the all-gather that read this group NGRP groups ago has finished
        ms = h.search_device(q_dev[i % NB].data_ptr(), nq, a.k, a.ef, out_grp[g][j].data_ptr(), cnt_dev.data_ptr(), sync=sync)
        if gather_grp is not None:
            h.stream_wait_last(gstream.cuda_stream)   # the gather stream waits for every launch of the group
            open_group[0] = g
            if j == GSTEP - 1:
                gather_group(g)                    # ncclAllGather of the group's answers, overlapped with the next searches
        return ms

    def drain():
        if gather_grp is not None and open_group[0] is not None:
            gather_group(open_group[0])            # a group left half-filled by an odd number of steps
        h.join()                                   # the launch stream waits for the launches in flight on the contexts
        if gather_grp is not None:
            stream.wait_stream(gstream)

    note('buffers ready')
    # ---- one instrumented pass: traversal counters (algorithmic bytes) and recall vs exact brute force
    h.enable_stats(True)
    step_device(0, sync=True)
    st = h.get_stats()
    h.enable_stats(False)
    E, A = st["evals"] / nq, st["adj_read"] / nq
    bytes_per_query = E * a.d * 4 + A * 4 + a.d * 4 + a.k * 16
    rid = rball = None
    if rank == 0:
        nt = min(1000, nq)
        qn = q_host[0][:nt].numpy()
        ti, td = h.bruteforce(qn, a.k)                      # exact ground truth (K5 kernel)
        o_, d_, i_, _, c_ = h.search_flat(qn, a.k, a.ef)
        rid, rball = recall_stats(i_, d_, c_, ti, td)

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    note('instrumented pass done')
    # ---- device-resident: W warm-up steps, then exactly K steps between CUDA events on the launch stream
    for i in range(a.warmup):
        step_device(i)
    drain()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev) as clk:
        barrier()
        e0.record(stream)
        for i in range(a.steps):
            step_device(i)
        drain()
        e1.record(stream)
        barrier()
    dev_ms = e0.elapsed_time(e1)
    note('device-timed loop done')
    if h.check_status() != 0:
        raise RuntimeError("visited table overflow during the timed region")
    clocks = clk.summary()
    # per-launch kernel duration (CUDA events inside the library, around the kernel alone)
    kms = [h.search_device(q_dev[i % NB].data_ptr(), nq, a.k, a.ef, out_dev[0].data_ptr(), cnt_dev.data_ptr(), sync=True)
           for i in range(min(a.steps, 10))]
    kernel_ms = float(np.mean(kms))

    # ---- steady-state reference point (not the metric): one launch of 10x the step's queries, where ramp-up and tail
    # of the 2-wave step no longer dominate (BASELINE configs[4] runs 125 000 queries per GPU, i.e. in this regime)
    steady = None
    if rank == 0 and not multi and nq <= 20000:
        nbig = 10 * nq
        qb = torch.from_numpy(pkg.datagen.make(a.data, nbig, a.d, 777)).cuda()
        ob = torch.empty((nbig, a.k, 16), dtype=torch.uint8, device="cuda")
        cb = torch.empty((nbig,), dtype=torch.int32, device="cuda")
        h.search_device(qb.data_ptr(), nbig, a.k, a.ef, ob.data_ptr(), cb.data_ptr(), sync=True)
        ms_big = min(h.search_device(qb.data_ptr(), nbig, a.k, a.ef, ob.data_ptr(), cb.data_ptr(), sync=True) for _ in range(3))
        steady = {"queries_per_launch": nbig, "kernel_ms": ms_big, "queries_per_s": nbig / ms_big * 1e3,
                  "algorithmic_GBps": bytes_per_query * nbig / ms_big / 1e6}
        del qb, ob, cb

    # ---- end to end through the C-ABI calls with host buffers (H2D + kernel + D2H per step)
    h.set_stream(None)
    torch.cuda.set_stream(torch.cuda.default_stream(dev))
    qh_np = [q.numpy() for q in q_host]

    gloo = dist.new_group(backend="gloo") if multi else None

    def host_barrier():
        # ranks > 0 must not park a spinning NCCL kernel on their GPU while rank 0's process drives that GPU too
        torch.cuda.synchronize()
        if multi:
            dist.barrier(group=gloo)

    def timed(fn, steps, warmup, bar=barrier):
        for i in range(warmup):
            fn(i)
        bar()
        t0 = time.perf_counter()
        for i in range(steps):
            fn(i)
        bar()
        return time.perf_counter() - t0

    # (1) every rank searches its own shard through the flat host call (ids + distances + counts back in host memory):
    # sequentially (one hnsw_b200_search_flat call at a time), and pipelined (hnsw_b200_search_flat_submit / _wait, batch
    # i+1 submitted before batch i's answers are collected: two batches in flight from ONE host thread)
    flat = lambda i: h.search_flat(qh_np[i % NB], a.k, a.ef, with_internal=False, with_pid=False)   # noqa: E731
    seq_s = timed(flat, a.steps, a.warmup)

    def pipelined(steps):
        prev = None
        for i in range(steps):
            t = h.submit_flat(qh_np[i % NB], a.k, a.ef, with_internal=False, with_pid=False)
            if prev is not None:
                h.wait_flat(prev)
            prev = t
        h.wait_flat(prev)

    pipelined(a.warmup)
    barrier()
    t0 = time.perf_counter()
    pipelined(a.steps)
    barrier()
    per_rank_s = time.perf_counter() - t0
    note('per-rank e2e done')
    # (2) N > 1: ONE process (rank 0) drives all the box's GPUs through its handle: hnsw_b200_replicate, then the same
    # submit / wait pipeline, every batch sharded over the N GPUs by the library
    one_call_s = None
    if multi:
        big = None
        host_barrier()
        if rank == 0:
            note('replicating in-process')
            h.replicate(list(range(world)))       # copies on the other GPUs, NCCL inside this process
            big = [torch.from_numpy(pkg.datagen.make(a.data, total_per_step, a.d, 9000 + b)).pin_memory().numpy()
                   for b in range(2 if total_per_step <= 200000 else 1)]

        def pipelined_big(steps):
            if rank != 0:
                return
            prev = None
            for i in range(steps):
                t = h.submit_flat(big[i % len(big)], a.k, a.ef, with_internal=False, with_pid=False)
                if prev is not None:
                    h.wait_flat(prev)
                prev = t
            h.wait_flat(prev)

        pipelined_big(a.warmup)
        host_barrier()
        t0 = time.perf_counter()
        pipelined_big(a.steps)
        host_barrier()
        one_call_s = time.perf_counter() - t0
        if rank == 0:
            h.replicate([dev])
        host_barrier()
    note('one-call e2e done')
    # (3) N = 1: the reference's own entry point, pageable row pointers, one malloc'ed Neighbourhood per query
    rowptr_s = None
    if not multi and nq <= 100000:
        rows = [np.array(r) for r in qh_np[0]]    # pageable copies, one allocation per query
        L, suf = h._L, h._suf
        ptrs = (ctypes.c_void_p * nq)(*[r.ctypes.data for r in rows])
        fn = getattr(L, "parallel_search_neighbours_" + suf)

        def call(_i):
            res = fn(h._h, nq, a.d, ptrs, a.k, a.ef)
            if not res:
                raise RuntimeError(pkg.last_error())
            L.hnsw_b200_free_vec_api(res)
        rowptr_s = timed(call, a.steps, min(a.warmup, 3))

    # ---- max over ranks
    times = [dev_ms, per_rank_s * 1e3, kernel_ms, (one_call_s or 0.0) * 1e3, seq_s * 1e3]
    if multi:
        t = torch.tensor(times, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times = t.tolist()
    dev_ms, per_rank_ms, kernel_ms, one_call_ms, seq_ms = times
    total_q = a.steps * total_per_step
    value = total_q / (dev_ms / 1e3)
    e2e_per_rank = total_q / (per_rank_ms / 1e3)
    e2e_seq = total_q / (seq_ms / 1e3)
    e2e_v = e2e_per_rank     # same definition at every N: each rank pipelines its own shard through submit / wait
    e2e_one = total_q / (one_call_ms / 1e3) if multi else None

    if rank != 0:
        return
    peak, peak_src = measured_peak_gbs()
    step_ms = dev_ms / a.steps                      # average launch duration over the timed region (launches overlap pairwise)
    achieved = bytes_per_query * nq / (step_ms / 1e3) / 1e9
    solo = bytes_per_query * nq / (kernel_ms / 1e3) / 1e9
    traffic, traffic_src = committed_traffic(a)
    e2e = {"value": e2e_v, "unit": "queries/s", "h2d_bytes_per_step": total_per_step * a.d * 4,
           "d2h_bytes_per_step": total_per_step * a.k * 16 + total_per_step * 4,
           "call": ("hnsw_b200_search_flat_submit / _wait, one host thread" + (" per rank" if multi else "") + ", batch i+1 "
                    "submitted before batch i is collected (two batches in flight); pinned host buffers read and written by the "
                    "kernel (zero-copy)"),
           "host_threads": 1,
           "sequential": {"value": e2e_seq, "unit": "queries/s",
                          "call": "every rank: one hnsw_b200_search_flat call at a time (no batches in flight together)"}}
    if multi:
        e2e["one_process"] = {"value": e2e_one, "unit": "queries/s",
                              "call": "ONE process: rank 0's handle after hnsw_b200_replicate, the same submit / wait pipeline, the "
                                      f"library shards every batch of {total_per_step} queries over the {world} GPUs"}
    if rowptr_s is not None:
        e2e["row_pointers"] = {"value": a.steps * nq / rowptr_s, "unit": "queries/s",
                               "call": f"parallel_search_neighbours_f32: {nq} pageable row pointers in, malloc'ed "
                                       "Vec_api<Neighbourhood_api> out and freed, per step"}
    line = {
        "metric": "queries/sec @ recall@10", "value": value, "unit": "queries/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dev_ms / a.steps, "higher_is_better": True,
        "scaling": "strong" if a.strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": workload_name(a, world), "preset": a.config, "recall_at_10": rid, "recall_at_10_ball": rball,
            "graph": "built on the GPU by this engine (batched insert)", "build_s": build_s,
            "index_broadcast_s": bcast_s, "evals_per_query": E, "adj_ids_per_query": A,
            "l2": f"no flush: working set (point store + adjacency {a.n * (a.d * 4 + a.M * 8) / 1e6:.0f} MB) "
                  f"{'exceeds' if a.n * (a.d * 4 + a.M * 8) > 126e6 else 'FITS IN'} the 126 MB L2; {NB} query batch(es) rotated",
            "overlap": "device-resident launches are asynchronous and alternate between two search contexts (streams, visited "
                       "tables, work counters) of the index: step i+1 starts while step i's last queries finish",
            "parallelism": (f"query-sharded x{world}, index replicated (ncclBroadcast inside the library), answers of every 2 steps "
                            "all-gathered by one ncclAllGather on a second stream, overlapped with the next steps; replicas checked "
                            "against rank 0") if multi else "1 GPU",
        },
        "clocks": clocks, "e2e": e2e, "gpu_launches": a.steps * world,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": "search_lean_kernel",
                     "kernel_ms": step_ms, "algorithmic_bytes_per_launch": bytes_per_query * nq,
                     "basis": "average launch duration over the timed region (CUDA events on the launch stream); consecutive "
                              "launches alternate between two contexts, so the last long searches of launch i run while launch "
                              "i+1 fills the SMs they left idle",
                     "solo": {"kernel_ms": kernel_ms, "achieved": solo, "frac": solo / peak,
                              "basis": "one launch alone on an idle GPU, CUDA events around the kernel"}},
    }
    if steady:
        steady["frac_of_peak"] = steady["algorithmic_GBps"] / peak
        line["config"]["steady_state"] = steady
    if world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(a, h, qh_np[0][:10000])
    line["config"]["setup_s"] = time.perf_counter() - t_setup
    emit(line)


def cpu_baseline(a, h, Q):
    """CPU restatement (oracle, kind "port") of the same path on the SAME graph and queries: literal reference
    mode (Rust-std heaps, AVX2-shaped sums), one query per task over all host threads like rayon par_iter."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    po.build()
    logical = os.cpu_count() or 1
    lv, rk, og, entry = h.export_points()
    # the graph is imported by ONE thread: interleave its pages over the NUMA nodes, as the reference's own
    # multi-threaded build would spread them by first touch (otherwise every search thread hammers one node)
    numa = po.numa_interleave(True)
    o = po.Oracle(a.M, a.n, 16, a.efc, a.metric, a.d, mode=po.MODE_STD, order=po.ORDER_REF)
    maxl = int(lv.max()) + 1 if len(lv) else 1
    o.import_graph(h.export_vectors(), og, lv, entry, {l: h.export_layer(l) for l in range(min(16, maxl + 1))})
    po.numa_interleave(False)
    threads = best_threads(o, Q, a, logical)
    reps, t0 = 0, time.perf_counter()
    while True:
        o.search_batch(Q, a.k, a.ef, nthreads=threads)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= a.cpu_seconds or reps >= 2000:
            break
    return {"value": reps * len(Q) / dt, "unit": "queries/s", "cores": threads, "threads": threads, "kind": "port",
            "cpu_model": cpu_model(), "logical_cpus": logical, "numa_interleave": "on" if numa == 0 else "refused",
            "sample": f"{reps} passes over the same {len(Q)}-query batch on the GPU-built graph ({dt:.1f} s wall), "
                      f"{threads} threads (best of all logical CPUs / half)"}


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
