"""Second, independent restatement of the reference RCB in plain numpy/Python
(no reordering, exact Python-int weight sums) used to cross-check the C++
oracle on small inputs.  Follows SURVEY.md §8(a) "exact behavioural spec":
recursive_bisection.rs:456-573 (split loop), :575-642 (recursion, ids, bbox
inheritance), :644-705 (narrowing, id renumbering).  TEST CODE ONLY."""
import numpy as np

f32 = np.float32


def split(x, w, sel, lo, hi, total, tol, to_f64=float):
    lo, hi = f32(lo), f32(hi)
    prev = None
    xs = x[sel]
    ws = [w[i] for i in sel]
    while True:
        st = f32((lo + hi) / f32(2.0))
        left = xs < st
        cnt = int(left.sum())
        wl = sum(wi for wi, l in zip(ws, left) if l)
        if cnt == len(xs):
            if prev == cnt:
                return sel, sel[:0], total, hi
            hi, prev = st, cnt
            continue
        pn = xs[~left].min()
        nd = f32(pn - st)
        ideal = to_f64(total) / 2.0
        with np.errstate(all="ignore"):
            imb = abs(np.float64(to_f64(wl) - ideal) / np.float64(ideal))
        if cnt == prev or hi <= f32(st + nd) or imb <= tol:
            return sel[left], sel[~left], wl, st
        prev = cnt
        if wl < total - wl:
            lo = st
        else:
            hi = st


def rcb(points, weights, iter_count, tol):
    pts = np.asarray(points, dtype=np.float64)
    n, dim = pts.shape
    part = np.zeros(n, dtype=np.int64)
    if n == 0:
        return part
    xs = [pts[:, d].astype(np.float32) for d in range(dim)]
    w = [weights[i].item() if hasattr(weights[i], "item") else weights[i] for i in range(n)]
    box = [(f32(pts[:, d].min()), f32(pts[:, d].max())) for d in range(dim)]

    def rec(sel, it, node, coord, total, box):
        if len(sel) == 0:
            return
        if it == 0:
            part[sel] = node
            return
        l, r, wl, sp = split(xs[coord], w, sel, box[coord][0], box[coord][1], total, tol)
        bl, br = list(box), list(box)
        bl[coord] = (box[coord][0], sp)
        br[coord] = (sp, box[coord][1])
        rec(l, it - 1, 2 * node + 1, (coord + 1) % dim, wl, bl)
        rec(r, it - 1, 2 * node + 2, (coord + 1) % dim, total - wl, br)

    rec(np.arange(n), iter_count, 0, 0, sum(w), box)
    return part - part.min()
