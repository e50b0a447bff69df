"""The cartesian-RCB oracle (oracle/grid_oracle.cpp, Grid::rcb) against the reference's own known answers
 and a direct numpy restatement."""
import numpy as np
import pytest


def py_weighted_median(w, total, threads):
    """coupe/src/cartesian/rcb.rs:52-99, written from the reference text (Python ints / floats)."""
    is_int = np.issubdtype(np.asarray(w).dtype, np.integer)
    conv = (lambda v: int(v)) if is_int else (lambda v: float(v))
    ideal = float(total) / 2.0
    lo_w, hi_w = conv(ideal * (1.0 - 0.01)), conv(ideal * (1.0 + 0.01))
    mn, mx, left = 0, len(w), conv(0)
    while True:
        chunk = max(1, (mx - mn) // threads)
        sums = []
        for s in range(mn, mx, chunk):
            acc = conv(0)
            for v in w[s:min(s + chunk, mx)]:
                acc = acc + conv(v)
            sums.append(acc)
        base, left0, prefix = mn, left, conv(0)
        for ci, cw in enumerate(sums):
            pos, pcw = base + ci * chunk, left0 + prefix
            prefix = prefix + cw
            if pcw < lo_w:
                mn, left = pos, pcw
            elif hi_w < pcw:
                mx = pos
                break
            else:
                return pos, pcw
        if mn + 1 >= mx:
            return mn, left


def py_grid_rcb(sizes, w, iters, threads):
    D = len(sizes)
    g = list(sizes) + [1] * (3 - D)
    W = np.asarray(w).reshape(g[2], g[1], g[0])  # [z][y][x]
    is_int = np.issubdtype(W.dtype, np.integer)
    conv = (lambda v: int(v)) if is_int else (lambda v: float(v))
    total = conv(0)
    for r in W.reshape(-1, g[0]):
        s = conv(0)
        for v in r:
            s = s + conv(v)
        total = total + s
    part = np.zeros(W.size, dtype=np.uint64)

    def rec(size, off, tot, it, coord, pid, cells):
        if size[coord] == 0 or it == 0:
            part[cells] = pid
            return
        outer, inner = (1 - coord, None) if D == 2 else ((coord + 1) % 3, (coord + 2) % 3)
        axis = []
        for a in range(size[coord]):
            pos = [0, 0, 0]
            pos[coord] = off[coord] + a
            s = conv(0)
            for o in range(size[outer]):
                pos[outer] = off[outer] + o
                if inner is None:
                    s = s + conv(W[0, pos[1], pos[0]])
                else:
                    for i in range(size[inner]):
                        pos[inner] = off[inner] + i
                        s = s + conv(W[pos[2], pos[1], pos[0]])
            axis.append(s)
        p, lw = py_weighted_median(np.array(axis, dtype=W.dtype), tot, threads)
        sp = p + off[coord]
        idx = np.arange(W.size)
        c = [idx % g[0], (idx // g[0]) % g[1], idx // g[0] // g[1]][coord]
        lo_s, hi_s, hi_o = list(size), list(size), list(off)
        lo_s[coord] = sp - off[coord]
        hi_s[coord] -= sp - off[coord]
        hi_o[coord] = sp
        rec(lo_s, off, lw, it - 1, (coord + 1) % D, 2 * pid, cells & (c < sp))
        rec(hi_s, hi_o, tot - lw, it - 1, (coord + 1) % D, 2 * pid + 1, cells & (c >= sp))

    rec(g[:], [0, 0, 0], total, iters, 1, 0, np.ones(W.size, dtype=bool))
    return part


def test_reference_doctest_and_test_3d(oracle):
    for threads in (2, 4, 16):
        p = oracle.grid_rcb((2, 2), np.ones(4), 2, threads)  # mod.rs:25-42
        assert sorted(p.tolist()) == [0, 1, 2, 3]
        p = oracle.grid_rcb((4, 4, 4), np.ones(64), 3, threads).reshape(4, 4, 4)  # rcb.rs:291-361
        assert len(set(p.ravel().tolist())) == 8
        for z in (0, 2):
            for y in (0, 2):
                for x in (0, 2):
                    assert len(set(p[z:z + 2, y:y + 2, x:x + 2].ravel().tolist())) == 1


@pytest.mark.parametrize("threads", [2, 3, 8, 64])
def test_weighted_median_property(oracle, threads):
    # rcb.rs:274-286: left_weight is the weight in front of the returned position
    rng = np.random.default_rng(threads)
    for _ in range(50):
        w = rng.integers(0, 1_000_000, int(rng.integers(2, 200))).astype(np.int64)
        pos, lw = oracle.grid_weighted_median(w, threads)
        assert lw == w[:pos].sum()
        assert (pos, lw) == py_weighted_median(w, int(w.sum()), threads)
    assert oracle.grid_weighted_median(np.ones(5), 1) is None  # a pool of one thread never returns


@pytest.mark.parametrize("sizes,iters", [((7, 5), 3), ((16, 16), 4), ((5, 4, 3), 4), ((8, 8, 8), 6), ((1, 9), 2), ((3, 1, 2), 5)])
@pytest.mark.parametrize("threads", [2, 5, 16])
def test_oracle_equals_numpy_restatement(oracle, sizes, iters, threads):
    rng = np.random.default_rng(sum(sizes) + iters)
    n = int(np.prod(sizes))
    for w in (rng.integers(0, 50, n).astype(np.int64), rng.uniform(0.0, 3.0, n), np.arange(n, dtype=np.float64)):
        assert np.array_equal(oracle.grid_rcb(sizes, w, iters, threads), py_grid_rcb(sizes, w, iters, threads))
