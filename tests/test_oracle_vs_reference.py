"""Pins oracle/oracle.c (the plain-C restatement) to the reference itself:
oracle/_ref/libraja_ref.so is RAJA's own seq_exec / omp_parallel_for_exec code compiled from
/root/reference.  Runs without a GPU.  Skipped only when the reference build is absent
(it is built by __graft_entry__.build() whenever /root/reference exists)."""
import numpy as np
import pytest

from oracle_lib import Oracle, Reference

pytestmark = pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built")

TYPES = [np.int32, np.uint32, np.int64, np.uint64, np.float32, np.float64]
SIZES = [0, 1, 2, 15, 16, 17, 33, 357, 2057, 32000]


def data(dtype, n, seed, lo=-100, hi=100):
    rng = np.random.default_rng(seed)
    if np.issubdtype(dtype, np.floating):
        return rng.uniform(lo, hi, n).astype(dtype)
    if np.issubdtype(dtype, np.unsignedinteger):
        return rng.integers(0, hi, n).astype(dtype)
    return rng.integers(lo, hi, n).astype(dtype)


@pytest.fixture(scope="module")
def orc():
    return Oracle()


@pytest.fixture(scope="module", params=[0, 1], ids=["seq_exec", "omp_parallel_for_exec"])
def ref(request):
    return Reference(request.param)


@pytest.mark.parametrize("dtype", TYPES)
def test_forall_bodies_bit_exact(orc, ref, dtype):
    for n in SIZES:
        b, c, a0 = data(dtype, n, 1), data(dtype, n, 2), data(dtype, n, 3)
        assert np.array_equal(orc.triad(b, c, 3), ref.triad(b, c, 3))
        assert np.array_equal(orc.daxpy(a0, b, 3), ref.daxpy(a0, b, 3))


@pytest.mark.parametrize("dtype", TYPES)
def test_reducers(orc, ref, dtype):
    exact = ref.pol_args[0] == 0 or not np.issubdtype(dtype, np.floating)
    for n in SIZES:
        x, y = data(dtype, n, 4), data(dtype, n, 5, 0, 10)
        for name in ("min", "max"):
            assert getattr(orc, name)(x) == getattr(ref, name)(x), (name, n)
        for name in ("minloc", "maxloc"):
            (ov, ol), (rv, rl) = getattr(orc, name)(x), getattr(ref, name)(x)
            assert ov == rv, (name, n)
            if ref.pol_args[0] == 0:
                assert ol == rl, (name, n)      # seq_exec: first extremum
            elif n:
                assert x[rl] == rv              # omp: "may be any index" (reduction.rst:74-79)
        if exact:  # same left fold -> identical bits, also in floating point
            assert orc.sum(x) == ref.sum(x)
            assert orc.dot(x, y) == ref.dot(x, y)
        else:      # omp_reduce combines per-thread partials: n*eps tolerance
            tol = max(n, 1) * np.finfo(dtype).eps * float(np.abs(x).sum() + 1)
            assert abs(float(orc.sum(x)) - float(ref.sum(x))) <= tol


@pytest.mark.parametrize("dtype", [np.int32, np.float64])
def test_minloc_ties_first_index_wins(orc, ref, dtype):
    if ref.pol_args[0] != 0:
        pytest.skip("only seq_exec pins the tie rule; omp may return any tied index")
    x = np.array([5, 3, 9, 3, 9, 7], dtype=dtype)
    assert orc.minloc(x) == ref.minloc(x) == (dtype(3), 1)
    assert orc.maxloc(x) == ref.maxloc(x) == (dtype(9), 2)


@pytest.mark.parametrize("dtype", [np.int32, np.int64, np.float64])
def test_multi_param_reduction(orc, ref, dtype):
    if ref.pol_args[0] == 1 and np.issubdtype(dtype, np.floating):
        pytest.skip("omp float sums differ in the last bits by design")
    for n in [1, 33, 2057]:
        x = data(dtype, n, 6)
        o, r = orc.multi(x), ref.multi(x)
        assert o["sum"] == r["sum"] and o["min"] == r["min"] and o["max"] == r["max"]
        assert o["minloc"][0] == r["minloc"][0] and o["maxloc"][0] == r["maxloc"][0]
        if ref.pol_args[0] == 0:
            assert o["minloc"] == r["minloc"] and o["maxloc"] == r["maxloc"]


@pytest.mark.parametrize("dtype", TYPES)
@pytest.mark.parametrize("op", ["plus", "minimum", "maximum"])
def test_scans(orc, ref, dtype, op):
    fp_plus = op == "plus" and np.issubdtype(dtype, np.floating)
    for n in SIZES:
        x = data(dtype, n, 7, 0, 8) if op == "plus" else data(dtype, n, 7)
        for inplace in (False, True):
            a, b = orc.inclusive_scan(x, op, inplace), ref.inclusive_scan(x, op, inplace)
            c, d = orc.exclusive_scan(x, op, 2, inplace), ref.exclusive_scan(x, op, 2, inplace)
            if fp_plus and ref.pol_args[0] == 1:
                assert np.allclose(a, b, rtol=1e-4) and np.allclose(c, d, rtol=1e-4)
            else:
                assert np.array_equal(a, b), (n, inplace)
                assert np.array_equal(c, d), (n, inplace)


@pytest.mark.parametrize("dtype", TYPES)
@pytest.mark.parametrize("desc", [False, True])
def test_stable_sort_pairs(orc, ref, dtype, desc):
    for n in SIZES:
        keys = data(dtype, n, 8, -20, 20)  # many duplicates: stability is observable
        vals = np.arange(n, dtype=np.uint64)
        ok, ov = orc.stable_sort_pairs(keys, vals, desc)
        rk, rv = ref.stable_sort_pairs(keys, vals, desc)
        assert np.array_equal(ok, rk) and np.array_equal(ov, rv), n
        assert np.array_equal(orc.stable_sort(keys, desc), ref.stable_sort(keys, desc))
        # the reference's unstable sort agrees on keys (pairs may differ among equal keys)
        assert np.array_equal(ok, ref.sort(keys, desc))
        uk, uv = ref.sort_pairs(keys, vals, desc, stable=False)
        assert np.array_equal(uk, ok) and np.array_equal(np.sort(uv), np.sort(ov))


def test_signed_zero_keys_keep_input_order(orc, ref):
    keys = np.array([0.0, -0.0, 1.0, -0.0, 0.0, -1.0], dtype=np.float64)
    vals = np.arange(6, dtype=np.uint64)
    ok, ov = orc.stable_sort_pairs(keys, vals)
    rk, rv = ref.stable_sort_pairs(keys, vals)
    assert np.array_equal(ov, rv) and np.array_equal(np.signbit(ok), np.signbit(rk))
    assert list(ov) == [5, 0, 1, 3, 4, 2]


def test_atomics(orc, ref):
    rng = np.random.default_rng(9)
    keys = rng.integers(0, 256, 100000).astype(np.int32)
    assert np.array_equal(orc.histogram(keys, 256), ref.histogram(keys, 256))
    assert np.array_equal(orc.histogram(keys, 256), np.bincount(keys, minlength=256))
    if ref.pol_args[0] == 0:
        assert orc.pi(512 * 512) == ref.pi(512 * 512)
    assert abs(ref.pi(512 * 512) - np.pi) < 1e-10
