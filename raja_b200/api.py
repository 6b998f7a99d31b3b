"""Front-end mirror: same names and argument meaning as the reference's pattern layer
(/root/reference/include/RAJA/pattern/{forall,reduce,scan,sort,atomic}.hpp), executing on
the b200 back-end through the C ABI.

Mapping of concepts
  RAJA::b200_exec<BLOCK, Async>      -> b200_exec(async_=False)
  camp::resources::Cuda              -> Resource (wraps a torch CUDA stream)
  RAJA::operators::plus/minimum/...  -> operators.plus / .minimum / .maximum
  RAJA::forall + named body          -> forall_triad / forall_daxpy
  RAJA::ReduceSum/Min/Max/MinLoc/..  -> reduce_sum / reduce_min / ... (one forall each)
  RAJA::expt::Reduce multi-param     -> reduce_multi
  RAJA::inclusive_scan[_inplace] ... -> inclusive_scan[_inplace] / exclusive_scan[_inplace]
  RAJA::sort / stable_sort[_pairs]   -> sort / stable_sort / sort_pairs / stable_sort_pairs
  RAJA::atomicAdd in a forall        -> atomic_histogram / atomic_pi

Every tensor argument must live on a CUDA device; host tensors raise (the reference
equally requires device-accessible pointers for cuda_exec).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check, lib

__all__ = [
    "operators", "b200_exec", "Resource", "synchronize",
    "forall_triad", "forall_daxpy",
    "reduce_sum", "reduce_min", "reduce_max", "reduce_minloc", "reduce_maxloc", "reduce_dot", "reduce_multi",
    "inclusive_scan", "inclusive_scan_inplace", "exclusive_scan", "exclusive_scan_inplace",
    "sort", "stable_sort", "sort_pairs", "stable_sort_pairs", "partition_pairs",
    "atomic_histogram", "atomic_pi", "pool_trim", "pool_bytes_cached", "sm_count", "PinnedBuffer",
]


class operators:  # RAJA::operators
    plus = _lib.PLUS
    minimum = _lib.MINIMUM
    maximum = _lib.MAXIMUM
    less = 0      # ascending comparator for sorts
    greater = 1   # descending


_DTYPES = {
    torch.int32: _lib.I32, torch.uint32: _lib.U32, torch.int64: _lib.I64, torch.uint64: _lib.U64,
    torch.float32: _lib.F32, torch.float64: _lib.F64,
}
_NP = {_lib.I32: np.int32, _lib.U32: np.uint32, _lib.I64: np.int64, _lib.U64: np.uint64, _lib.F32: np.float32, _lib.F64: np.float64}


def _dt(t):
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError(f"raja_b200: unsupported dtype {t.dtype}") from None


def _dev(*tensors):
    for t in tensors:
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise ValueError("raja_b200: arguments must be CUDA tensors (no CPU fallback)")
        if not t.is_contiguous():
            raise ValueError("raja_b200: arguments must be contiguous")


class Resource:
    """camp::resources::Cuda: work is enqueued on its stream; wait() synchronises it."""

    def __init__(self, stream=None):
        self.stream = stream if stream is not None else torch.cuda.current_stream()

    @property
    def handle(self):
        return ctypes.c_void_p(self.stream.cuda_stream)

    def wait(self):
        check(lib.rb200_stream_synchronize(self.handle), "stream synchronize")


class b200_exec:
    """Execution policy: synchronous unless async_ (cf. cuda_exec / cuda_exec_async,
    policy/cuda/policy.hpp:1308-1322)."""

    def __init__(self, async_=False):
        self.async_ = async_


def _res(res):
    return res if res is not None else Resource()


def _finish(pol, res):
    if pol is None or not pol.async_:
        res.wait()


def synchronize():
    """RAJA::synchronize<b200_synchronize>()"""
    check(lib.rb200_device_synchronize(), "device synchronize")


def sm_count():
    return lib.rb200_sm_count()


def pool_trim():
    check(lib.rb200_pool_trim(), "pool trim")


def pool_bytes_cached():
    return lib.rb200_pool_bytes_cached()


class PinnedBuffer:
    """A block from the pinned pool viewed as a numpy array (device-writable result slot)."""

    def __init__(self, nbytes):
        self.nbytes = nbytes
        self.ptr = lib.rb200_pinned_malloc(nbytes)
        if not self.ptr:
            raise _lib.B200Error("pinned allocation failed: " + lib.rb200_last_error().decode())
        self._buf = (ctypes.c_uint8 * nbytes).from_address(self.ptr)
        self.bytes = np.frombuffer(self._buf, dtype=np.uint8)

    def view(self, dtype, offset=0, count=1):
        return self.bytes[offset:offset + np.dtype(dtype).itemsize * count].view(dtype)

    def free(self):
        if self.ptr:
            lib.rb200_pinned_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ---------------------------------------------------------------------------- forall
def forall_triad(a, b, c, s, pol=None, res=None):
    """RAJA::forall<b200_exec>(RangeSegment(0,N), [=](i){ a[i] = b[i] + s*c[i]; })"""
    _dev(a, b, c)
    res = _res(res)
    n = a.numel()
    if b.numel() != n or c.numel() != n or b.dtype != a.dtype or c.dtype != a.dtype:
        raise ValueError("forall_triad: a, b, c must match in length and dtype")
    check(lib.rb200_forall_triad(_dt(a), a.data_ptr(), b.data_ptr(), c.data_ptr(), float(s), n, res.handle), "forall_triad")
    _finish(pol, res)


def forall_daxpy(a, b, s, pol=None, res=None):
    """RAJA::forall<b200_exec>(RangeSegment(0,N), [=](i){ a[i] += b[i]*s; })  (examples/tut_daxpy.cpp)"""
    _dev(a, b)
    res = _res(res)
    n = a.numel()
    if b.numel() != n or b.dtype != a.dtype:
        raise ValueError("forall_daxpy: a, b must match in length and dtype")
    check(lib.rb200_forall_daxpy(_dt(a), a.data_ptr(), b.data_ptr(), float(s), n, res.handle), "forall_daxpy")
    _finish(pol, res)


# ---------------------------------------------------------------------------- reducers
def _reduce(kind, x, res, deterministic=True):
    _dev(x)
    res = _res(res)
    dt = _dt(x)
    out = PinnedBuffer(16)
    check(lib.rb200_reduce(kind, dt, x.data_ptr(), x.numel(), out.ptr, out.ptr + 8, 1 if deterministic else 0, res.handle),
          "reduce")
    res.wait()  # get(): the reference's reducers synchronise on read (policy/cuda/reduce.hpp:1128-1149)
    val = out.view(_NP[dt])[0].item()
    loc = int(out.view(np.int64, 8)[0])
    out.free()
    return val, loc


def reduce_sum(x, res=None, deterministic=True):
    """ReduceSum<b200_reduce, T> s(0); forall(..., s += x[i]); s.get()"""
    return _reduce(_lib.REDUCE_SUM, x, res, deterministic)[0]


def reduce_min(x, res=None, deterministic=True):
    return _reduce(_lib.REDUCE_MIN, x, res, deterministic)[0]


def reduce_max(x, res=None, deterministic=True):
    return _reduce(_lib.REDUCE_MAX, x, res, deterministic)[0]


def reduce_minloc(x, res=None):
    """ReduceMinLoc: (value, index of the first minimum)"""
    return _reduce(_lib.REDUCE_MINLOC, x, res)


def reduce_maxloc(x, res=None):
    return _reduce(_lib.REDUCE_MAXLOC, x, res)


def reduce_dot(x, y, res=None):
    """exercises/dot-product: ReduceSum over x[i]*y[i]"""
    _dev(x, y)
    if x.numel() != y.numel() or x.dtype != y.dtype:
        raise ValueError("reduce_dot: x, y must match in length and dtype")
    res = _res(res)
    dt = _dt(x)
    out = PinnedBuffer(16)
    check(lib.rb200_reduce_dot(dt, x.data_ptr(), y.data_ptr(), x.numel(), out.ptr, res.handle), "reduce_dot")
    res.wait()
    val = out.view(_NP[dt])[0].item()
    out.free()
    return val


def reduce_multi(x, res=None):
    """examples/forall-param-reductions.cpp: five expt::Reduce params in one forall."""
    _dev(x)
    res = _res(res)
    dt = _dt(x)
    nbytes = lib.rb200_reduce_multi_result_bytes(dt)
    loc_off = lib.rb200_reduce_multi_loc_offset(dt)
    out = PinnedBuffer(nbytes)
    check(lib.rb200_reduce_multi(dt, x.data_ptr(), x.numel(), out.ptr, res.handle), "reduce_multi")
    res.wait()
    vals = out.view(_NP[dt], 0, 5)
    locs = out.view(np.int64, loc_off, 2)
    r = {"sum": vals[0].item(), "min": vals[1].item(), "max": vals[2].item(),
         "minloc": (vals[3].item(), int(locs[0])), "maxloc": (vals[4].item(), int(locs[1]))}
    out.free()
    return r


# ---------------------------------------------------------------------------- scans
def _scan(inclusive, src, dst, op, init, pol, res):
    _dev(src, dst)
    if src.numel() != dst.numel() or src.dtype != dst.dtype:
        raise ValueError("scan: input and output must match in length and dtype")
    res = _res(res)
    dt = _dt(src)
    init_p = None
    if init is not None:
        init_arr = np.array([init], dtype=_NP[dt])
        init_p = init_arr.ctypes.data_as(ctypes.c_void_p)
    check(lib.rb200_scan(1 if inclusive else 0, op, dt, src.data_ptr(), dst.data_ptr(), src.numel(), init_p, res.handle), "scan")
    _finish(pol, res)


def inclusive_scan(src, dst, op=operators.plus, pol=None, res=None):
    _scan(True, src, dst, op, None, pol, res)


def inclusive_scan_inplace(x, op=operators.plus, pol=None, res=None):
    _scan(True, x, x, op, None, pol, res)


def exclusive_scan(src, dst, op=operators.plus, init=None, pol=None, res=None):
    _scan(False, src, dst, op, init, pol, res)


def exclusive_scan_inplace(x, op=operators.plus, init=None, pol=None, res=None):
    _scan(False, x, x, op, init, pol, res)


# ---------------------------------------------------------------------------- sorts
def stable_sort(keys, comp=operators.less, pol=None, res=None):
    _dev(keys)
    res = _res(res)
    check(lib.rb200_sort(_dt(keys), keys.data_ptr(), keys.numel(), int(comp), res.handle), "sort")
    _finish(pol, res)


def sort(keys, comp=operators.less, pol=None, res=None):
    """RAJA::sort: unstable in the contract; this back-end's sort is always stable
    (policy/cuda/sort.hpp:285-321 does the same)."""
    stable_sort(keys, comp, pol, res)


def stable_sort_pairs(keys, vals, comp=operators.less, pol=None, res=None, keys_alt=None, vals_alt=None):
    _dev(keys, vals)
    if keys.numel() != vals.numel():
        raise ValueError("sort_pairs: keys and vals must have the same length")
    vb = vals.element_size()
    if vb not in (4, 8):
        raise TypeError("sort_pairs: values must be 4 or 8 bytes wide")
    res = _res(res)
    if keys_alt is not None:
        _dev(keys_alt, vals_alt)
        check(lib.rb200_sort_pairs_buffered(_dt(keys), keys.data_ptr(), keys_alt.data_ptr(), vals.data_ptr(), vals_alt.data_ptr(),
                                            vb, keys.numel(), int(comp), res.handle), "sort_pairs")
    else:
        check(lib.rb200_sort_pairs(_dt(keys), keys.data_ptr(), vals.data_ptr(), vb, keys.numel(), int(comp), res.handle),
              "sort_pairs")
    _finish(pol, res)


def sort_pairs(keys, vals, comp=operators.less, pol=None, res=None, keys_alt=None, vals_alt=None):
    stable_sort_pairs(keys, vals, comp, pol, res, keys_alt, vals_alt)


# ---------------------------------------------------------------------------- atomics
def partition_pairs(keys_in, keys_out, vals_in, vals_out, splitters, counts, comp=operators.less, pol=None, res=None):
    """Stable multi-way split by <= 7 splitters (rb200_partition_pairs; the exchange step of the
    multi-GPU sample sort).  `splitters`: host sequence / numpy array of keys in sort order;
    `counts`: int64/uint64 device tensor of >= 8 elements, receives the bucket sizes."""
    _dev(keys_in, keys_out, counts)
    vb = 0
    if vals_in is not None:
        _dev(vals_in, vals_out)
        vb = vals_in.element_size()
    if counts.numel() < 8 or counts.element_size() != 8:
        raise ValueError("partition_pairs: counts must hold 8 64-bit integers")
    dt = _dt(keys_in)
    sp = np.ascontiguousarray(np.asarray(splitters, dtype=_NP[dt]))
    res = _res(res)
    check(lib.rb200_partition_pairs(dt, keys_in.data_ptr(), keys_out.data_ptr(), vals_in.data_ptr() if vb else None,
                                    vals_out.data_ptr() if vb else None, vb, keys_in.numel(),
                                    sp.ctypes.data_as(ctypes.c_void_p) if sp.size else None, int(sp.size), int(comp),
                                    counts.data_ptr(), res.handle), "partition_pairs")
    _finish(pol, res)


def atomic_histogram(hist, keys, aggregate=False, pol=None, res=None):
    """forall(i){ RAJA::atomicAdd<b200_atomic>(&hist[keys[i]], 1); }"""
    _dev(hist, keys)
    if keys.dtype != torch.int32:
        raise TypeError("atomic_histogram: keys must be int32")
    res = _res(res)
    check(lib.rb200_atomic_histogram(_dt(hist), hist.data_ptr(), keys.data_ptr(), keys.numel(), 1 if aggregate else 0, res.handle),
          "atomic_histogram")
    _finish(pol, res)


def atomic_pi(num_bins, aggregate=False, res=None):
    """examples/pi-reduce_vs_atomic.cpp atomic variant; returns pi."""
    res = _res(res)
    acc = torch.zeros(1, dtype=torch.float64, device="cuda")
    check(lib.rb200_atomic_pi(acc.data_ptr(), int(num_bins), 1 if aggregate else 0, res.handle), "atomic_pi")
    res.wait()
    return 4.0 * acc.item()
