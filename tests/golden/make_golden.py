"""Writes the golden fixtures of tests/golden/ — run once, outputs committed.

The reference (Rust) cannot be built or imported in this image, so no fixture here is an
output of the reference binary.  What is pinned instead:
  * file-format fixtures: bytes spelled out from the ABNF of weight-gen(1) "WEIGHT FILE" and
    mesh-part(1) "PARTITION FILE" (tools/doc/*.scd) with Python's struct — independent of the
    library's own writer;
  * the reference's known-answer tests for the path (recursive_bisection.rs:1077-1115, doctests
    :739-768, :864-893, coupe-ffi/examples/rcb.c), transcribed with the ids the reference's
    assertions imply;
  * RCB fixtures on seeded inputs produced by the C++ oracle (marked "oracle-generated"): they
    guard the oracle and the CUDA path against regressions, not against the reference.
"""
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def mewe(is_int, rows):
    cc = len(rows[0]) if rows else 0
    out = b"MeWe" + bytes([1, 1 if is_int else 0]) + struct.pack("<H", cc) + struct.pack("<Q", len(rows))
    for r in rows:
        for v in r:
            out += struct.pack("<q" if is_int else "<d", v)
    return out


def mepe(ids):
    return b"MePe" + struct.pack("<Q", len(ids)) + b"".join(struct.pack("<Q", i) for i in ids)


def main():
    w = lambda name, data: open(os.path.join(HERE, name), "wb").write(data)
    w("weights_f64_1crit.mewe", mewe(False, [[0.5], [1.25], [-3.0], [1e300], [0.0]]))
    w("weights_i64_2crit.mewe", mewe(True, [[1, -2], [3, 4], [2**62, -2**63]]))
    w("weights_empty.mewe", b"MeWe" + bytes([1, 0]) + b"\0" * 10)  # weight.rs:147-152
    w("partition_6.mepe", mepe([0, 3, 1, 2, 2**40, 0]))
    w("partition_empty.mepe", mepe([]))

    kat = {
        "source": "reference tests, transcribed (ids implied by the reference's assertions)",
        "cases": [
            {"name": "test_rcb_basic recursive_bisection.rs:1077-1115", "dim": 2, "iter_count": 2, "tolerance": 0.05,
             "points": [[-1.3, 6.0], [2.0, -4.0], [1.0, 1.0], [-3.0, -2.5], [-1.3, -0.3], [2.0, 1.0], [-3.0, 1.0], [1.3, -2.0]],
             "weights": [1.0] * 8, "same_part": [[0, 6], [1, 7], [2, 5], [3, 4]]},
            {"name": "Rcb doctest recursive_bisection.rs:739-768", "dim": 2, "iter_count": 2, "tolerance": 0.05,
             "points": [[1.0, 1.0], [-1.0, 1.0], [1.0, -1.0], [-1.0, -1.0]], "weights": [1, 1, 1, 1],
             "all_distinct": True},
            {"name": "coupe-ffi/examples/rcb.c", "dim": 2, "iter_count": 1, "tolerance": 0.05,
             "points": [[0.0, 0.0], [0.0, 1.0], [1.0, 0.0], [1.0, 1.0]], "weights": [1.0] * 4,
             "same_part": [[0, 1], [2, 3]]},
        ],
    }
    json.dump(kat, open(os.path.join(HERE, "rcb_known_answers.json"), "w"), indent=1)

    # oracle-generated regression vectors (small: 4096 points each)
    from oracle import pyoracle

    pyoracle.build()
    rng = np.random.default_rng(20261017)
    vec = {"source": "oracle-generated (oracle/rcb_oracle.cpp, mode 0); regression guard only", "cases": []}
    for dim, wkind, iters, tol in ((2, "i64", 5, 0.05), (3, "i32", 6, 0.0), (3, "f64int", 4, 0.01)):
        n = 4096
        pts = np.round(rng.normal(size=(n, dim)) * 3, 3)
        if wkind == "i64":
            wt = rng.integers(1, 100, n).astype(np.int64)
        elif wkind == "i32":
            wt = rng.integers(1, 10, n).astype(np.int32)
        else:
            wt = rng.integers(1, 20, n).astype(np.float64)
        ids = pyoracle.rcb(pts, wt, iters, tol)
        vec["cases"].append({"dim": dim, "wkind": wkind, "iter_count": iters, "tolerance": tol,
                             "points": pts.tolist(), "weights": wt.tolist(), "ids": ids.astype(int).tolist()})
    json.dump(vec, open(os.path.join(HERE, "rcb_oracle_vectors.json"), "w"))


if __name__ == "__main__":
    main()
