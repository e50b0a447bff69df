"""Exact (correctly rounded) weight sums of the nodes of a split tree, for the f64-weight tests.

Given the points, the f64 weights and the split positions of a trace (heap order), rebuilds the
membership of every node top-down with the reference's own rule (left = {x_f32 < split_pos},
recursive_bisection.rs:483-503, :613-616) and sums the weights with math.fsum.  TEST CODE ONLY."""
import math

import numpy as np


def exact_tree(points, weights, visited, split_pos, iter_count):
    pts = np.asarray(points, dtype=np.float64)
    n, dim = pts.shape
    xs = [pts[:, d].astype(np.float32) for d in range(dim)]
    w = np.asarray(weights, dtype=np.float64)
    m = (1 << iter_count) - 1
    wl = np.zeros(m)
    total = np.zeros(m)
    path = np.zeros(n, dtype=np.int64)
    members = {0: np.arange(n)}
    for node in range(m):
        sel = members.pop(node, None)
        if sel is None or len(sel) == 0:
            continue
        depth = (node + 1).bit_length() - 1
        assert visited[node], f"node {node} holds {len(sel)} points but was not visited"
        left = xs[depth % dim][sel] < split_pos[node]
        total[node] = math.fsum(w[sel])
        wl[node] = math.fsum(w[sel[left]])
        if depth + 1 < iter_count:
            members[2 * node + 1] = sel[left]
            members[2 * node + 2] = sel[~left]
        else:
            first = (1 << iter_count) - 1
            path[sel[left]] = 2 * node + 1 - first
            path[sel[~left]] = 2 * node + 2 - first
    ids = (path - path.min()).astype(np.uint64) if n else path.astype(np.uint64)
    return wl, total, ids
