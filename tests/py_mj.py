"""Independent pure-numpy restatement of coupe's Multi-Jagged (multi_jagged.rs:70-288) used to cross-check
oracle/mj_oracle.cpp: same pinned choices (stable sort, depth-first numbering, fixed fold chunks), written
from the reference text, not from the C++ oracle."""
import math

import numpy as np


def ulps_eq(a, b):
    if a == b:
        return True
    if math.isnan(a) or math.isnan(b):
        return False
    if abs(a - b) <= np.finfo(np.float64).eps:
        return True
    if math.copysign(1.0, a) != math.copysign(1.0, b):
        return False
    ia, ib = np.float64(a).view(np.int64), np.float64(b).view(np.int64)
    return abs(int(ia) - int(ib)) <= 4


def scheme(num_parts, max_iter):
    """nested (num_splits, modifiers, children) — multi_jagged.rs:70-98"""
    root = int(np.ceil(np.float32(num_parts) ** (np.float32(1.0) / np.float32(max_iter))))
    rem, quo = num_parts % root, num_parts // root
    sub = (root - rem) * quo + rem * (quo + 1)
    mods = [(quo + 1) / sub] * rem + [quo / sub] * (root - rem)
    if rem == 0 and max_iter == 0:
        return (root - 1, mods, None)
    kids = [scheme(quo + 1, max_iter - 1) for _ in range(rem)] + [scheme(quo, max_iter - 1) for _ in range(rem, root)]
    return (root - 1, mods, kids)


def split_positions(w, perm, mods, chunk):
    mods = mods[:-1]
    n = len(perm)
    chunk = chunk or max(n, 1)
    lows = list(range(0, n, chunk))
    sums = []
    for lo in lows:
        acc = 0.0
        for i in perm[lo:lo + chunk]:
            acc = acc + float(w[i])
        sums.append(acc)
    total = 0.0
    for s in sums:
        total = total + s
    thr, consumed = [], 0.0
    for m in mods:
        consumed += total * m
        thr.append(consumed)
    ret, cache, cur, it = [], [], 0.0, 0
    for t in thr:
        if cur > t:
            ret.append(ret[-1])
            cache.append(cache[-1])
            continue
        while True:
            low, s = lows[it], sums[it]  # IndexError = the reference's unwrap() panic
            it += 1
            if cur + s > t:
                ret.append(low)
                cache.append(cur)
                cur += s
                break
            cur += s
    out = []
    for idx, s, t in zip(ret, cache, thr):
        while s + float(w[perm[idx]]) < t or ulps_eq(t, s + float(w[perm[idx]])):
            s += float(w[perm[idx]])
            idx += 1
        out.append(idx)
    return out


def multi_jagged(points, weights, part_count, max_iter, chunk=0):
    pts, w = np.asarray(points, dtype=np.float64), np.asarray(weights, dtype=np.float64)
    part = np.zeros(len(pts), dtype=np.uint64)
    counter = [0]

    def rec(perm, coord, sch):
        nsplit, mods, kids = sch
        if nsplit:
            perm = perm[np.argsort(pts[perm, coord], kind="stable")]
            pos = [0] + split_positions(w, perm, mods, chunk) + [len(perm)]
            for c in range(len(pos) - 1):
                rec(perm[pos[c]:pos[c + 1]], (coord + 1) % pts.shape[1], kids[c])
        else:
            part[perm] = counter[0]
            counter[0] += 1

    rec(np.arange(len(pts)), 0, scheme(part_count, max_iter))
    return part
