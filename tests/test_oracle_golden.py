"""oracle/oracle.c against the committed fixtures that were generated from the reference's own
seq_exec path (tests/golden/generate.py).  No GPU, no /root/reference needed."""
import os

import numpy as np
import pytest

from oracle_lib import Oracle

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_seq_exec.npz"))
TYPES = ["i32", "u32", "i64", "u64", "f32", "f64"]


@pytest.fixture(scope="module")
def orc():
    return Oracle()


@pytest.mark.parametrize("t", TYPES)
@pytest.mark.parametrize("n", [28, 639, 2057])
def test_reducers_and_forall(orc, t, n):
    k = f"red_{t}_{n}"
    x, y = G[k + "_x"], G[k + "_y"]
    mnv, mnl = orc.minloc(x)
    mxv, mxl = orc.maxloc(x)
    got = np.array([orc.sum(x), orc.min(x), orc.max(x), mnv, mxv, orc.dot(x, y)], dtype=x.dtype)
    assert np.array_equal(got, G[k + "_vals"])
    assert [mnl, mxl] == list(G[k + "_locs"])
    assert np.array_equal(orc.triad(x, y, 3), G[k + "_triad"])
    assert np.array_equal(orc.daxpy(x, y, 3), G[k + "_daxpy"])


@pytest.mark.parametrize("t", TYPES)
@pytest.mark.parametrize("n", [357, 4001])
@pytest.mark.parametrize("op", ["plus", "minimum", "maximum"])
def test_scans(orc, t, n, op):
    k = f"scan_{t}_{n}_{op}"
    x = G[k + "_in"]
    assert np.array_equal(orc.inclusive_scan(x, op), G[k + "_incl"])
    assert np.array_equal(orc.exclusive_scan(x, op, 2), G[k + "_excl"])
    assert np.array_equal(orc.inclusive_scan(x, op, inplace=True), G[k + "_incl"])


@pytest.mark.parametrize("t", TYPES)
@pytest.mark.parametrize("n", [10, 1000, 3000])
@pytest.mark.parametrize("desc", [0, 1])
def test_stable_sort_pairs(orc, t, n, desc):
    k = f"sort_{t}_{n}_{desc}"
    keys = G[k + "_keys"]
    sk, sv = orc.stable_sort_pairs(keys, np.arange(n, dtype=np.uint64), bool(desc))
    assert np.array_equal(sk, G[k + "_sk"]) and np.array_equal(sv, G[k + "_sv"])


def test_atomics(orc):
    assert np.array_equal(orc.histogram(G["hist_keys"], 256), G["hist_256"])
    assert orc.pi(512 * 512) == G["pi_262144"][0]
