#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REFERENCE itself (oracle/_ref/libraja_ref.so =
RAJA's seq_exec compiled from /root/reference; see oracle/Makefile).  Run in the build
container:  python tests/golden/generate.py
The fixtures are small on purpose (committed); inputs are seeded so they can be regenerated.
Cases follow the reference's own test data:
  reducers  rand()%100-style small integers cast to the type, segments [0,28) [3,642) [0,2057)
  scans     iota, N in {357, 4001} (the reference uses 357 and 32000), ops plus/minimum/maximum, non-identity exclusive offset
  sorts     uniform_int[-N, N] keys with duplicates, N up to 3000 (reference: 10000), ascending/descending, stable
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import Reference  # noqa: E402

TYPES = {"i32": np.int32, "u32": np.uint32, "i64": np.int64, "u64": np.uint64, "f32": np.float32, "f64": np.float64}


def main():
    ref = Reference(0)
    rng = np.random.default_rng(20251017)
    out = {}
    # reducers + forall bodies
    for name, dt in TYPES.items():
        for (lo, hi) in [(0, 28), (3, 642), (0, 2057)]:
            n = hi - lo
            x = rng.integers(0, 100, n).astype(dt)
            y = rng.integers(0, 10, n).astype(dt)
            key = f"red_{name}_{n}"
            out[key + "_x"], out[key + "_y"] = x, y
            mnv, mnl = ref.minloc(x)
            mxv, mxl = ref.maxloc(x)
            out[key + "_vals"] = np.array([ref.sum(x), ref.min(x), ref.max(x), mnv, mxv, ref.dot(x, y)], dtype=dt)
            out[key + "_locs"] = np.array([mnl, mxl], dtype=np.int64)
            out[key + "_triad"] = ref.triad(x, y, 3)
            out[key + "_daxpy"] = ref.daxpy(x, y, 3)
    # scans
    for name, dt in TYPES.items():
        for n in (357, 4001):
            x = np.arange(1, n + 1).astype(dt)
            if name in ("i32", "u32", "f32") and n == 4001:
                x = (x % 17).astype(dt)  # keep int32 sums from overflowing / f32 sums exact
            for op in ("plus", "minimum", "maximum"):
                xs = x if op == "plus" else rng.permutation(x)
                key = f"scan_{name}_{n}_{op}"
                out[key + "_in"] = xs
                out[key + "_incl"] = ref.inclusive_scan(xs, op)
                out[key + "_excl"] = ref.exclusive_scan(xs, op, 2)
    # sorts
    for name, dt in TYPES.items():
        for n in (10, 1000, 3000):
            lo = 0 if name.startswith("u") else -n
            keys = rng.integers(lo, n, n).astype(dt)
            vals = np.arange(n, dtype=np.uint64)
            for desc in (0, 1):
                k, v = ref.stable_sort_pairs(keys, vals, bool(desc))
                key = f"sort_{name}_{n}_{desc}"
                out[key + "_keys"] = keys
                out[key + "_sk"], out[key + "_sv"] = k, v
    # atomics
    keys = rng.integers(0, 256, 50000).astype(np.int32)
    out["hist_keys"] = keys
    out["hist_256"] = ref.histogram(keys, 256)
    out["pi_262144"] = np.array([ref.pi(512 * 512)])
    np.savez_compressed(os.path.join(HERE, "reference_seq_exec.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
