"""Synthetic inputs of BASELINE.json's configs (no datasets on the box).  Seeded numpy PCG64.

uniform   — iid U[0,1) like /root/reference/examples/random.rs:19-24 and tests/serpar.rs:26
clustered — Gaussian mixture with low intrinsic dimension, SIFT-like value range [0,255] (iid uniform
            in 128-d is adversarial for any graph index: SURVEY.md §8d)
unit      — clustered, then l2-normalised (DistDot == angular, examples/ann-glove25-angular.rs:81-82)
"""
import numpy as np


def uniform(n, d, seed):
    return np.random.default_rng(seed).random((n, d), dtype=np.float32)


def clustered(n, d, seed, n_centres=1000, rank=16, s_low=0.25, s_iso=0.02, centre_seed=12345, lo=0.0, hi=255.0):
    """Gaussian mixture whose within-cluster spread lives in a shared `rank`-dimensional subspace plus a
    little isotropic noise: low intrinsic dimension like SIFT descriptors (recall@10 ~0.96 at ef=64, M=16,
    ef_c=200 on 100k points, vs ~0.5 for iid uniform 128-d)."""
    crng = np.random.default_rng(centre_seed)
    centres = crng.random((n_centres, d), dtype=np.float32)
    basis = crng.standard_normal((rank, d)).astype(np.float32) / np.float32(np.sqrt(rank))
    rng = np.random.default_rng(seed)
    out = np.empty((n, d), np.float32)
    step = 1 << 18
    for b in range(0, n, step):
        e = min(n, b + step)
        which = rng.integers(0, n_centres, e - b)
        z = rng.standard_normal((e - b, rank), dtype=np.float32)
        x = centres[which] + s_low * (z @ basis) + s_iso * rng.standard_normal((e - b, d), dtype=np.float32)
        out[b:e] = np.clip(x, 0.0, 1.0) * (hi - lo) + lo
    return out


def unit(n, d, seed, **kw):
    x = clustered(n, d, seed, lo=-1.0, hi=1.0, **kw)
    x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-30)
    return x.astype(np.float32)


def make(kind, n, d, seed):
    if kind == "uniform":
        return uniform(n, d, seed)
    if kind == "clustered":
        return clustered(n, d, seed)
    if kind == "unit":
        return unit(n, d, seed)
    raise ValueError(kind)
