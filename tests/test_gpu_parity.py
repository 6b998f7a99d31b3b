"""Parity tests proper: the CUDA path (through the C ABI, raja_b200 -> libraja_b200.so) against
the oracle on the same seeded inputs, against the committed reference fixtures, and -- at the
BASELINE sizes (N = 2^28) -- through size-independent properties.

Bars: bit-exact for integer reductions/scans/atomics, elementwise forall bodies and stable
sorts; floating-point reductions and scans within n * eps * sum|x| (BASELINE.json north_star).
"""
import os

import numpy as np
import pytest
import torch

from oracle_lib import Oracle

pytestmark = pytest.mark.gpu

NP_TYPES = [np.int32, np.uint32, np.int64, np.uint64, np.float32, np.float64]
SIZES = [1, 2, 3, 31, 32, 33, 255, 1000, 4097, 30001, 262144 + 7, 1_000_003]
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_seq_exec.npz"))


@pytest.fixture(scope="module")
def rb():
    import raja_b200
    assert os.path.exists(raja_b200.LIB_PATH)
    return raja_b200


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def dev(a):
    return torch.from_numpy(a).cuda()


def host(t):
    return t.cpu().numpy()


def data(dtype, n, seed, lo=-100, hi=100):
    rng = np.random.default_rng(seed)
    if np.issubdtype(dtype, np.floating):
        return rng.uniform(lo, hi, n).astype(dtype)
    if np.issubdtype(dtype, np.unsignedinteger):
        return rng.integers(0, hi, n).astype(dtype)
    return rng.integers(lo, hi, n).astype(dtype)


def fp_tol(x, n):
    return max(n, 1) * np.finfo(x.dtype).eps * float(np.abs(x.astype(np.float64)).sum() + 1.0)


def unaligned(a):
    """same values at an address that is NOT 16-byte aligned (exercises the scalar kernels)"""
    buf = torch.empty(a.size + 4, dtype=dev(a[:1]).dtype, device="cuda")
    view = buf[1:1 + a.size]
    view.copy_(dev(a))
    assert view.data_ptr() % 16 != 0
    return view


# ------------------------------------------------------------------ forall bodies
@pytest.mark.parametrize("dtype", NP_TYPES)
def test_forall_triad_daxpy_bit_exact(rb, orc, dtype):
    for n in SIZES:
        b, c, a0 = data(dtype, n, 1), data(dtype, n, 2), data(dtype, n, 3)
        a = torch.empty(n, dtype=dev(b[:1]).dtype, device="cuda")
        rb.forall_triad(a, dev(b), dev(c), 3)
        assert np.array_equal(host(a), orc.triad(b, c, 3)), n
        a = dev(a0)
        rb.forall_daxpy(a, dev(b), 3)
        assert np.array_equal(host(a), orc.daxpy(a0, b, 3)), n
    # misaligned operands
    n = 70001
    b, c = data(dtype, n, 4), data(dtype, n, 5)
    a = unaligned(np.zeros(n, dtype))
    rb.forall_triad(a, unaligned(b), unaligned(c), 3)
    assert np.array_equal(host(a), orc.triad(b, c, 3))


def test_forall_empty_is_a_noop(rb):
    a = torch.empty(0, dtype=torch.float64, device="cuda")
    rb.forall_triad(a, a.clone(), a.clone(), 2.0)
    rb.forall_daxpy(a, a.clone(), 2.0)


# ------------------------------------------------------------------ reducers
@pytest.mark.parametrize("dtype", NP_TYPES)
def test_reducers_vs_oracle(rb, orc, dtype):
    fp = np.issubdtype(dtype, np.floating)
    for n in SIZES:
        x = data(dtype, n, 6)
        y = data(dtype, n, 7, 0, 10)
        d = dev(x)
        assert rb.reduce_min(d) == orc.min(x)
        assert rb.reduce_max(d) == orc.max(x)
        assert rb.reduce_minloc(d) == orc.minloc(x), n
        assert rb.reduce_maxloc(d) == orc.maxloc(x), n
        if fp:
            assert abs(rb.reduce_sum(d) - float(orc.sum(x))) <= fp_tol(x, n)
            assert abs(rb.reduce_dot(d, dev(y)) - float(orc.dot(x, y))) <= fp_tol(x * y, n)
        else:
            assert rb.reduce_sum(d) == orc.sum(x)
            assert rb.reduce_dot(d, dev(y)) == orc.dot(x, y)
        m, o = rb.reduce_multi(d), orc.multi(x)
        assert m["min"] == o["min"] and m["max"] == o["max"] and m["minloc"] == o["minloc"] and m["maxloc"] == o["maxloc"]
        if fp:
            assert abs(m["sum"] - float(o["sum"])) <= fp_tol(x, n)
        else:
            assert m["sum"] == o["sum"]


@pytest.mark.parametrize("dtype", [np.int32, np.float64, np.int64, np.float32])
def test_reducers_unaligned_and_empty(rb, orc, dtype):
    x = data(dtype, 50001, 8)
    d = unaligned(x)
    assert rb.reduce_minloc(d) == orc.minloc(x)
    assert rb.reduce_maxloc(d) == orc.maxloc(x)
    if np.issubdtype(dtype, np.floating):
        assert abs(rb.reduce_sum(d) - float(orc.sum(x))) <= fp_tol(x, x.size)
    else:
        assert rb.reduce_sum(d) == orc.sum(x)
    e = torch.empty(0, dtype=d.dtype, device="cuda")
    z = np.empty(0, dtype)
    assert rb.reduce_sum(e) == 0
    assert rb.reduce_min(e) == orc.min(z) and rb.reduce_max(e) == orc.max(z)
    assert rb.reduce_minloc(e) == orc.minloc(z) and rb.reduce_minloc(e)[1] == -1


@pytest.mark.parametrize("dtype", [np.int32, np.float64])
def test_minmaxloc_ties_resolve_to_first_index(rb, orc, dtype):
    # duplicate extrema far apart (different threads, warps, blocks)
    n = 3_000_000
    x = data(dtype, n, 9, -50, 50)
    x[[7, 1_000_000, 2_999_999]] = -1000
    x[[123_456, 123_457, 2_000_000]] = 1000
    d = dev(x)
    assert rb.reduce_minloc(d) == orc.minloc(x) == (dtype(-1000), 7)
    assert rb.reduce_maxloc(d) == orc.maxloc(x) == (dtype(1000), 123_456)
    m = rb.reduce_multi(d)
    assert m["minloc"] == (-1000, 7) and m["maxloc"] == (1000, 123_456)


def test_reduce_sum_is_run_to_run_deterministic(rb):
    x = dev(data(np.float64, 5_000_001, 10))
    r = {rb.reduce_sum(x) for _ in range(5)}
    assert len(r) == 1


def test_reduce_atomic_mode(rb, orc):
    x = data(np.int32, 1_000_003, 11)
    assert rb.reduce_sum(dev(x), deterministic=False) == orc.sum(x)
    assert rb.reduce_max(dev(x), deterministic=False) == orc.max(x)
    xf = data(np.float64, 1_000_003, 12)
    assert abs(rb.reduce_sum(dev(xf), deterministic=False) - float(orc.sum(xf))) <= fp_tol(xf, xf.size)
    with pytest.raises(rb.B200Error):
        rb.reduce_min(dev(xf), deterministic=False)  # no fp atomicMin: refuses, does not fall back


@pytest.mark.parametrize("t", ["i32", "u32", "i64", "u64", "f32", "f64"])
def test_reducers_and_forall_against_reference_fixtures(rb, t):
    for n in (28, 639, 2057):
        k = f"red_{t}_{n}"
        x, y = G[k + "_x"], G[k + "_y"]
        d = dev(x)
        mnv, mnl = rb.reduce_minloc(d)
        mxv, mxl = rb.reduce_maxloc(d)
        got = np.array([rb.reduce_sum(d), rb.reduce_min(d), rb.reduce_max(d), mnv, mxv, rb.reduce_dot(d, dev(y))], dtype=x.dtype)
        # the reference's test data are small integers: sums are exact in floating point too
        assert np.array_equal(got, G[k + "_vals"])
        assert [mnl, mxl] == list(G[k + "_locs"])
        a = torch.empty_like(d)
        rb.forall_triad(a, d, dev(y), 3)
        assert np.array_equal(host(a), G[k + "_triad"])
        a = d.clone()
        rb.forall_daxpy(a, dev(y), 3)
        assert np.array_equal(host(a), G[k + "_daxpy"])


# ------------------------------------------------------------------ scans
@pytest.mark.parametrize("dtype", NP_TYPES)
@pytest.mark.parametrize("op", ["plus", "minimum", "maximum"])
def test_scans_vs_oracle(rb, orc, dtype, op):
    rop = getattr(rb.operators, op)
    fp_plus = op == "plus" and np.issubdtype(dtype, np.floating)
    for n in SIZES:
        x = data(dtype, n, 13, 0, 8) if op == "plus" else data(dtype, n, 13)
        d = dev(x)
        out = torch.empty_like(d)
        rb.inclusive_scan(d, out, rop)
        ex = torch.empty_like(d)
        rb.exclusive_scan(d, ex, rop, init=2)
        inp = d.clone()
        rb.inclusive_scan_inplace(inp, rop)
        exi = d.clone()
        rb.exclusive_scan_inplace(exi, rop, init=2)
        oi, oe = orc.inclusive_scan(x, op), orc.exclusive_scan(x, op, 2)
        if fp_plus:
            tol = fp_tol(x, n)
            for got, want in ((out, oi), (inp, oi), (ex, oe), (exi, oe)):
                assert np.max(np.abs(host(got).astype(np.float64) - want.astype(np.float64))) <= tol, n
        else:
            assert np.array_equal(host(out), oi), n
            assert np.array_equal(host(inp), oi), n
            assert np.array_equal(host(ex), oe), n
            assert np.array_equal(host(exi), oe), n
        assert np.array_equal(host(d), x)  # out-of-place scans leave the input alone


@pytest.mark.parametrize("dtype", [np.int32, np.int64, np.float64])
def test_scans_unaligned_empty_and_default_init(rb, orc, dtype):
    x = data(dtype, 100_003, 14, 0, 8)
    d = unaligned(x)
    out = unaligned(np.zeros_like(x))
    rb.inclusive_scan(d, out)
    want = orc.inclusive_scan(x)
    if np.issubdtype(dtype, np.floating):
        assert np.max(np.abs(host(out) - want)) <= fp_tol(x, x.size)
    else:
        assert np.array_equal(host(out), want)
    ex = torch.empty_like(dev(x))
    rb.exclusive_scan(dev(x), ex)  # init defaults to the operator identity
    want = orc.exclusive_scan(x, "plus", 0)
    if np.issubdtype(dtype, np.floating):
        assert np.max(np.abs(host(ex) - want)) <= fp_tol(x, x.size)
    else:
        assert np.array_equal(host(ex), want)
    e = torch.empty(0, dtype=ex.dtype, device="cuda")
    rb.inclusive_scan(e, e.clone())
    rb.exclusive_scan_inplace(e, init=5)


@pytest.mark.parametrize("t", ["i32", "u32", "i64", "u64", "f32", "f64"])
def test_scans_against_reference_fixtures(rb, t):
    for n in (357, 4001):
        for op in ("plus", "minimum", "maximum"):
            k = f"scan_{t}_{n}_{op}"
            d = dev(G[k + "_in"])
            out = torch.empty_like(d)
            rb.inclusive_scan(d, out, getattr(rb.operators, op))
            assert np.array_equal(host(out), G[k + "_incl"])  # iota data: exact in fp as well
            rb.exclusive_scan(d, out, getattr(rb.operators, op), init=2)
            assert np.array_equal(host(out), G[k + "_excl"])


# ------------------------------------------------------------------ sorts
SORT_SIZES = [2, 3, 33, 1000, 7680, 7681, 100_003, 1_000_003]


@pytest.mark.parametrize("dtype", NP_TYPES)
@pytest.mark.parametrize("desc", [False, True])
def test_stable_sort_pairs_vs_oracle(rb, orc, dtype, desc):
    comp = rb.operators.greater if desc else rb.operators.less
    for n in SORT_SIZES:
        full = np.issubdtype(dtype, np.integer) and n > 1000
        keys = (np.random.default_rng(n).integers(np.iinfo(dtype).min, np.iinfo(dtype).max, n, dtype=dtype)
                if full else data(dtype, n, 15, -20, 20))  # full-range bits / many duplicates
        vals = np.arange(n, dtype=np.uint32)
        k, v = dev(keys), dev(vals)
        rb.stable_sort_pairs(k, v, comp)
        ok, ov = orc.stable_sort_pairs(keys, vals, desc)
        assert np.array_equal(host(k), ok), n
        assert np.array_equal(host(v), ov), n
    # 8-byte values, keys only
    keys = data(dtype, 50_001, 16, -300, 300)
    vals = np.arange(keys.size, dtype=np.int64) * 3
    k, v = dev(keys), dev(vals)
    rb.sort_pairs(k, v, comp)
    ok, ov = orc.stable_sort_pairs(keys, vals, desc)
    assert np.array_equal(host(k), ok) and np.array_equal(host(v), ov)
    k = dev(keys)
    rb.stable_sort(k, comp)
    assert np.array_equal(host(k), ok)
    k = dev(keys)
    rb.sort(k, comp)
    assert np.array_equal(host(k), ok)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_sort_float_special_values(rb, orc, dtype):
    rng = np.random.default_rng(17)
    keys = rng.uniform(-1e6, 1e6, 20_000).astype(dtype)
    keys[rng.integers(0, keys.size, 500)] = 0.0
    keys[rng.integers(0, keys.size, 500)] = -0.0
    keys[:4] = [np.inf, -np.inf, np.finfo(dtype).max, np.finfo(dtype).tiny]
    vals = np.arange(keys.size, dtype=np.uint32)
    for desc in (False, True):
        k, v = dev(keys), dev(vals)
        rb.stable_sort_pairs(k, v, rb.operators.greater if desc else rb.operators.less)
        ok, ov = orc.stable_sort_pairs(keys, vals, desc)
        assert np.array_equal(host(v), ov)                       # -0.0 == +0.0: input order kept
        assert np.array_equal(host(k).view(np.uint8), ok.view(np.uint8))  # stored keys keep their sign bit


@pytest.mark.parametrize("dtype", [np.uint32, np.int32, np.uint64, np.int64, np.float32, np.float64])
def test_sort_degenerate_digit_passes(rb, orc, dtype):
    """keys whose digits are constant in several passes (constant high bytes, constant low bytes, all keys equal): those passes are
    identity permutations and take the copy-through path of the sweep (radix_sort.cuh: identity_pass); floating-point keys are
    encoded / decoded on the way through exactly as in the general path"""
    rng = np.random.default_rng(23)
    n = 70_001
    if np.issubdtype(dtype, np.floating):
        cases = [(1.0 + rng.integers(0, 200, n) / 256.0).astype(dtype),            # only a few mantissa bits vary
                 (-(3.0 + rng.integers(0, 5, n))).astype(dtype),                   # negative, tiny alphabet
                 np.full(n, -0.0, dtype=dtype), np.where(rng.integers(0, 2, n) == 0, 0.0, -0.0).astype(dtype)]
    else:
        info = np.iinfo(dtype)
        cases = [(rng.integers(0, 200, n)).astype(dtype),                           # constant high bytes
                 ((rng.integers(0, 200, n).astype(np.uint64) << np.uint64(8 * (np.dtype(dtype).itemsize - 1) - 1))).astype(dtype),  # constant low bytes
                 np.full(n, info.max // 3, dtype=dtype)]                            # every pass degenerate
    vals = np.arange(n, dtype=np.uint32)
    for keys in cases:
        for desc in (False, True):
            k, v = dev(keys), dev(vals)
            rb.stable_sort_pairs(k, v, rb.operators.greater if desc else rb.operators.less)
            ok, ov = orc.stable_sort_pairs(keys, vals, desc)
            assert np.array_equal(host(v), ov)
            assert np.array_equal(host(k).view(np.uint8), ok.view(np.uint8))
            k2 = dev(keys)
            rb.stable_sort(k2, rb.operators.greater if desc else rb.operators.less)
            assert np.array_equal(host(k2).view(np.uint8), orc.stable_sort(keys, desc).view(np.uint8))


def test_sort_trivial_sizes_and_buffered_variant(rb, orc):
    for n in (0, 1):
        k = torch.arange(n, dtype=torch.int32, device="cuda")
        rb.sort(k)
        rb.sort_pairs(k, k.clone())
    keys = data(np.uint32, 300_000, 18, 0, 2**31)
    vals = np.arange(keys.size, dtype=np.uint32)
    k, v = dev(keys), dev(vals)
    rb.sort_pairs(k, v, keys_alt=torch.empty_like(k), vals_alt=torch.empty_like(v))
    ok, ov = orc.stable_sort_pairs(keys, vals)
    assert np.array_equal(host(k), ok) and np.array_equal(host(v), ov)


@pytest.mark.parametrize("t", ["i32", "u32", "i64", "u64", "f32", "f64"])
def test_sorts_against_reference_fixtures(rb, t):
    for n in (10, 1000, 3000):
        for desc in (0, 1):
            key = f"sort_{t}_{n}_{desc}"
            k, v = dev(G[key + "_keys"]), dev(np.arange(n, dtype=np.uint64))
            rb.stable_sort_pairs(k, v, desc)
            assert np.array_equal(host(k), G[key + "_sk"]) and np.array_equal(host(v), G[key + "_sv"])


# ------------------------------------------------------------------ atomics
def _np_partition(k, v, sp, desc):
    """stable split restated in numpy: bucket = number of splitters that sort at or before the key"""
    if len(sp):
        b = ((k[:, None] <= sp[None, :]) if desc else (k[:, None] >= sp[None, :])).sum(axis=1)
    else:
        b = np.zeros(k.size, dtype=np.int64)
    order = np.argsort(b, kind="stable")
    return k[order], (v[order] if v is not None else None), np.bincount(b, minlength=8)[:8]


@pytest.mark.parametrize("dtype", NP_TYPES)
@pytest.mark.parametrize("desc", [False, True])
def test_partition_pairs_is_the_stable_split(rb, dtype, desc):
    """rb200_partition_pairs (the exchange step of the multi-GPU sample sort): bit-exact against the numpy
    restatement, keys and values, for 0..7 splitters, ragged sizes, duplicated splitters and keys"""
    for n, nsplit, vdt in [(1, 1, np.uint32), (33, 0, np.uint32), (1000, 3, np.uint32), (7681, 7, np.uint64), (262144 + 7, 1, np.uint32),
                           (1_000_003, 7, np.uint32), (50_000, 4, None)]:
        k = data(dtype, n, 900 + n + nsplit, -50, 50)  # narrow range: many equal keys, keys equal to splitters
        if np.issubdtype(dtype, np.floating) and n > 4:
            k[1], k[3] = -0.0, 0.0
        sp = np.sort(np.random.default_rng(n).choice(k, size=nsplit, replace=True)) if nsplit else np.zeros(0, dtype=dtype)
        if desc:
            sp = sp[::-1].copy()
        v = np.arange(n).astype(vdt) if vdt is not None else None
        ko = torch.empty_like(dev(k))
        vo = torch.empty_like(dev(v)) if v is not None else None
        counts = torch.zeros(8, dtype=torch.int64, device="cuda")
        rb.partition_pairs(dev(k), ko, dev(v) if v is not None else None, vo, sp, counts, int(desc))
        ek, ev, ec = _np_partition(k, v, sp, desc)
        assert np.array_equal(host(counts), ec), (n, nsplit)
        assert np.array_equal(host(ko).view(np.uint8), ek.view(np.uint8)), (n, nsplit)  # bit patterns: -0.0 stays -0.0
        if v is not None:
            assert np.array_equal(host(vo), ev), (n, nsplit)


@pytest.mark.parametrize("dtype", [np.uint32, np.float64])
@pytest.mark.parametrize("shift", [1, 3, 17])
def test_partition_pairs_unaligned_pointers(rb, dtype, shift):
    """input and output pointers that are not 16-byte / 128-byte aligned: the line-aligned write-out of the splitter sweep
    (threads mapped onto the destination's 128-byte lines) and the 128-bit loads of the bucket count must handle the heads"""
    n, nsplit = 300_007, 5
    k = data(dtype, n, 4242 + shift, -1000, 1000)
    sp = np.sort(np.random.default_rng(shift).choice(k, size=nsplit, replace=False))
    v = np.arange(n).astype(np.uint32)
    kin = torch.empty(n + 64, dtype=dev(k).dtype, device="cuda")[shift:shift + n]
    kin.copy_(dev(k))
    vin = torch.empty(n + 64, dtype=torch.uint32, device="cuda")[shift + 1:shift + 1 + n]
    vin.copy_(dev(v))
    ko = torch.empty(n + 64, dtype=kin.dtype, device="cuda")[shift + 2:shift + 2 + n]
    vo = torch.empty(n + 64, dtype=torch.uint32, device="cuda")[shift:shift + n]
    counts = torch.zeros(8, dtype=torch.int64, device="cuda")
    rb.partition_pairs(kin, ko, vin, vo, sp, counts, 0)
    ek, ev, ec = _np_partition(k, v, sp, False)
    assert np.array_equal(host(counts), ec)
    assert np.array_equal(host(ko).view(np.uint8), ek.view(np.uint8))
    assert np.array_equal(host(vo), ev)


def test_partition_pairs_rejects_bad_arguments(rb):
    k = dev(np.arange(10, dtype=np.uint32))
    counts = torch.zeros(8, dtype=torch.int64, device="cuda")
    with pytest.raises(rb.B200Error):  # splitters out of order
        rb.partition_pairs(k, torch.empty_like(k), None, None, np.array([5, 3], dtype=np.uint32), counts)
    with pytest.raises(rb.B200Error):  # in place is not supported
        rb.partition_pairs(k, k, None, None, np.array([5], dtype=np.uint32), counts)
    with pytest.raises(rb.B200Error):  # more than 7 splitters
        rb.partition_pairs(k, torch.empty_like(k), None, None, np.arange(8, dtype=np.uint32), counts)


@pytest.mark.parametrize("aggregate", [False, True])
def test_atomic_histogram_and_pi(rb, orc, aggregate):
    rng = np.random.default_rng(19)
    for nbins, n in ((256, 1_000_003), (65536, 2_000_001), (3, 100_000)):
        keys = rng.integers(0, nbins, n).astype(np.int32)
        hist = torch.zeros(nbins, dtype=torch.int32, device="cuda")
        rb.atomic_histogram(hist, dev(keys), aggregate)
        assert np.array_equal(host(hist), orc.histogram(keys, nbins))
    # zipf-like skew, 64-bit and floating bins
    keys = np.minimum(rng.zipf(1.3, 500_000), 255).astype(np.int32)
    for dt in (torch.uint64, torch.float64, torch.float32):
        hist = torch.zeros(256, dtype=dt, device="cuda")
        rb.atomic_histogram(hist, dev(keys), aggregate)
        assert np.array_equal(host(hist).astype(np.int64), np.bincount(keys, minlength=256))
    hist = torch.zeros(256, dtype=torch.int32, device="cuda")
    rb.atomic_histogram(hist, dev(G["hist_keys"]), aggregate)
    assert np.array_equal(host(hist), G["hist_256"])
    nb = 512 * 512
    assert abs(rb.atomic_pi(nb, aggregate) - orc.pi(nb)) <= nb * np.finfo(np.float64).eps * 4


# ------------------------------------------------------------------ pools / streams
def test_pools_reuse_blocks_per_stream(rb):
    from raja_b200._lib import lib
    import ctypes
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    h1, h2 = ctypes.c_void_p(s1.cuda_stream), ctypes.c_void_p(s2.cuda_stream)
    p = lib.rb200_device_malloc(1 << 20, h1)
    assert p and lib.rb200_device_free(p, h1) == 0
    assert lib.rb200_device_malloc(1 << 20, h2) != p   # another stream never gets the cached block
    q = lib.rb200_device_malloc(1 << 20, h1)
    assert q == p                                      # same stream does (stream-ordered reuse)
    assert lib.rb200_device_free(q, h1) == 0
    assert lib.rb200_device_free(12345, h1) != 0      # foreign pointer is refused
    assert rb.pool_bytes_cached() > 0
    rb.pool_trim()
    assert rb.pool_bytes_cached() == 0


def test_async_policy_is_stream_ordered(rb, orc):
    s = torch.cuda.Stream()
    res = rb.Resource(s)
    x = data(np.int64, 2_000_000, 20, 0, 9)
    with torch.cuda.stream(s):
        d = dev(x)
        out = torch.empty_like(d)
        rb.inclusive_scan(d, out, pol=rb.b200_exec(async_=True), res=res)
        rb.inclusive_scan_inplace(out, pol=rb.b200_exec(async_=True), res=res)  # enqueued behind the first
    res.wait()
    assert np.array_equal(host(out), orc.inclusive_scan(orc.inclusive_scan(x)))


# ------------------------------------------------------------------ BASELINE size (2^28): properties
N_FULL = 1 << 28


def test_full_size_reduce_and_scan_properties(rb):
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randint(0, 8, (N_FULL,), dtype=torch.int32, device="cuda", generator=g)
    total = rb.reduce_sum(x)
    assert total == int(x.sum(dtype=torch.int64))          # no overflow: 7 * 2^28 < 2^31
    out = torch.empty_like(x)
    rb.inclusive_scan(x, out)
    assert int(out[-1]) == total                             # last inclusive element = reduction
    assert bool((out[1:] - out[:-1] == x[1:]).all())         # first difference recovers the input
    assert int(out[0]) == int(x[0])
    rb.exclusive_scan(x, out, init=5)
    assert int(out[0]) == 5 and int(out[-1]) + int(x[-1]) == total + 5
    mn, mnl = rb.reduce_minloc(x)
    assert mn == 0 and int(x[mnl]) == 0 and (mnl == 0 or not bool((x[:mnl] == 0).any()))
    del out
    xf = torch.rand(N_FULL, dtype=torch.float64, device="cuda", generator=g) + 0.5
    s = rb.reduce_sum(xf)
    ref = float(xf.sum())
    assert abs(s - ref) <= N_FULL * np.finfo(np.float64).eps * ref
    assert abs(rb.reduce_dot(xf, xf) - float((xf * xf).sum())) <= N_FULL * np.finfo(np.float64).eps * 4 * ref


def test_full_size_sort_pairs_properties(rb):
    g = torch.Generator(device="cuda").manual_seed(42)
    # 2^20 distinct keys -> ~256 duplicates each: stability is observable
    keys = torch.randint(0, 1 << 20, (N_FULL,), dtype=torch.int32, device="cuda", generator=g).view(torch.uint32)
    orig = keys.clone()
    vals = torch.arange(N_FULL, dtype=torch.int32, device="cuda").view(torch.uint32)
    rb.sort_pairs(keys, vals)
    k = keys.view(torch.int32)  # < 2^20: safe to compare as signed
    v = vals.view(torch.int32)
    assert bool((k[1:] >= k[:-1]).all())                                  # sorted
    same = k[1:] == k[:-1]
    assert bool((v[1:][same] > v[:-1][same]).all())                       # stable: ties keep input order
    assert bool((orig.view(torch.int32)[v.long()] == k).all())            # every pair still matches its key
    assert int(v.sum(dtype=torch.int64)) == N_FULL * (N_FULL - 1) // 2    # values are a permutation (checksum)


def test_sort_beyond_2_30_elements_uses_64_bit_lookback():
    """n >= 2^30 switches the look-back words to 64 bits (radix_sort.cuh: C = unsigned long long); size-independent properties at
    2^30 + 12345 keys: sortedness, checksum, values = the matching permutation, stability (tools/micro/big_sort_check.py)"""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if torch.cuda.get_device_properties(0).total_memory < 60 * 2**30:
        pytest.skip("needs ~40 GB of device memory")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "micro", "big_sort_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "BIG SORT OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_scan_beyond_2_31_elements():
    """64-bit lengths (the reference's CUDA scans take int): inclusive / exclusive scan of 2^31 + 7 int32 elements, checked element by
    element through out[i] - out[i-1] == in[i] and an independent 64-bit total (tools/micro/big_scan_check.py)"""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if torch.cuda.get_device_properties(0).total_memory < 60 * 2**30:
        pytest.skip("needs ~30 GB of device memory")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "micro", "big_scan_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "BIG SCAN OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_forall_and_reducers_beyond_2_31_elements():
    """64-bit index spaces: ReduceSum / MinLoc / MaxLoc with the extrema (and their duplicates) planted past index 2^31, triad over
    2^31 + 5 floats (tools/micro/big_reduce_check.py)"""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if torch.cuda.get_device_properties(0).total_memory < 60 * 2**30:
        pytest.skip("needs ~40 GB of device memory")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "micro", "big_reduce_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "BIG REDUCE OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
