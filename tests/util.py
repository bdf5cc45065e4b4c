"""Shared helpers for the parity tests."""
import numpy as np


def oracle_layers(o, nlayers=16):
    """Export an oracle graph as the `layers` list Hnsw.import_graph takes."""
    return [o.export_layer(l) for l in range(nlayers)]


def gpu_layers(h, nlayers=16):
    return [h.export_layer(l) for l in range(nlayers)]


def recall_ids(found, counts, truth):
    """mean over queries of |found[:count] ∩ truth| / k"""
    k = truth.shape[1]
    tot = 0.0
    for i in range(truth.shape[0]):
        tot += len(set(found[i, :counts[i]].tolist()) & set(truth[i].tolist())) / k
    return tot / truth.shape[0]


def recall_ball(dists, counts, truth_d):
    """the reference's recall: #{returned with d <= true k-th distance}/k
    (/root/reference/examples/ann-sift1m-128-euclidean.rs:172-186)"""
    k = truth_d.shape[1]
    tot = 0.0
    for i in range(truth_d.shape[0]):
        tot += float(np.sum(dists[i, :counts[i]] <= truth_d[i, k - 1])) / k
    return tot / truth_d.shape[0]


def csr_lists(off, ids):
    return [ids[int(off[i]):int(off[i + 1])].tolist() for i in range(len(off) - 1)]
