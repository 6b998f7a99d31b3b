"""ctypes doors onto the test oracles (TEST INFRASTRUCTURE):
  Oracle     oracle/liboracle.so          plain-C restatement of the reference's seq_exec path
  Reference  oracle/_ref/libraja_ref.so   the unmodified reference compiled from its own headers
Both expose the same numpy-level methods so tests can swap one for the other.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libraja_ref.so")

SUFFIX = {np.dtype(np.int32): "i32", np.dtype(np.uint32): "u32", np.dtype(np.int64): "i64", np.dtype(np.uint64): "u64",
          np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}
CT = {"i32": ctypes.c_int32, "u32": ctypes.c_uint32, "i64": ctypes.c_int64, "u64": ctypes.c_uint64, "f32": ctypes.c_float,
      "f64": ctypes.c_double}
OPS = {"plus": 0, "minimum": 1, "maximum": 2}


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def ensure_oracle_built():
    if not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])


class _Base:
    prefix = None
    pol_args = ()

    def _fn(self, name, restype, argtypes):
        f = getattr(self.lib, f"{self.prefix}_{name}")
        f.restype = restype
        f.argtypes = [ctypes.c_int] * len(self.pol_args) + list(argtypes)
        return lambda *a: f(*self.pol_args, *a)

    # -- forall
    def triad(self, b, c, s):
        sfx = SUFFIX[b.dtype]
        a = np.empty_like(b)
        self._fn(f"triad_{sfx}", None, [ctypes.c_void_p] * 3 + [CT[sfx], ctypes.c_int64])(_ptr(a), _ptr(b), _ptr(c), CT[sfx](s), b.size)
        return a

    def daxpy(self, a, b, s):
        sfx = SUFFIX[a.dtype]
        a = a.copy()
        self._fn(f"daxpy_{sfx}", None, [ctypes.c_void_p] * 2 + [CT[sfx], ctypes.c_int64])(_ptr(a), _ptr(b), CT[sfx](s), a.size)
        return a

    # -- reducers
    def _val(self, name, x):
        sfx = SUFFIX[x.dtype]
        r = self._fn(f"{name}_{sfx}", CT[sfx], [ctypes.c_void_p, ctypes.c_int64])(_ptr(x), x.size)
        return x.dtype.type(r)

    def sum(self, x):
        return self._val("sum", x)

    def min(self, x):
        return self._val("min", x)

    def max(self, x):
        return self._val("max", x)

    def dot(self, x, y):
        sfx = SUFFIX[x.dtype]
        r = self._fn(f"dot_{sfx}", CT[sfx], [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64])(_ptr(x), _ptr(y), x.size)
        return x.dtype.type(r)

    def _loc(self, name, x):
        sfx = SUFFIX[x.dtype]
        loc = ctypes.c_int64(-7)
        r = self._fn(f"{name}_{sfx}", CT[sfx], [ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)])(
            _ptr(x), x.size, ctypes.byref(loc))
        return x.dtype.type(r), int(loc.value)

    def minloc(self, x):
        return self._loc("minloc", x)

    def maxloc(self, x):
        return self._loc("maxloc", x)

    def multi(self, x):
        mn, ml = self.minloc(x)
        mx, xl = self.maxloc(x)
        return {"sum": self.sum(x), "min": self.min(x), "max": self.max(x), "minloc": (mn, ml), "maxloc": (mx, xl)}

    # -- scans
    def inclusive_scan(self, x, op="plus", inplace=False):
        sfx = SUFFIX[x.dtype]
        src = x.copy()
        out = src if inplace else np.empty_like(x)
        self._fn(f"inclusive_scan_{sfx}", None, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64])(
            OPS[op], _ptr(src), _ptr(out), x.size)
        return out

    def exclusive_scan(self, x, op="plus", init=0, inplace=False):
        sfx = SUFFIX[x.dtype]
        src = x.copy()
        out = src if inplace else np.empty_like(x)
        self._fn(f"exclusive_scan_{sfx}", None, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, CT[sfx]])(
            OPS[op], _ptr(src), _ptr(out), x.size, CT[sfx](init))
        return out

    # -- atomics
    def histogram(self, keys, nbins):
        hist = np.zeros(nbins, dtype=np.int32)
        self._fn("histogram_i32", None, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64])(_ptr(hist), _ptr(keys), keys.size)
        return hist

    def pi(self, num_bins):
        return 4.0 * self._fn("pi_atomic", ctypes.c_double, [ctypes.c_int64])(num_bins)


class Oracle(_Base):
    """oracle/liboracle.so"""
    prefix = "orc"
    kind = "port"

    def __init__(self):
        ensure_oracle_built()
        self.lib = ctypes.CDLL(ORACLE_SO)
        self.cores = 1

    def stable_sort_pairs(self, keys, vals=None, descending=False):
        sfx = SUFFIX[keys.dtype]
        k = keys.copy()
        v = None if vals is None else vals.astype(np.uint64)
        rc = self._fn(f"stable_sort_pairs_{sfx}", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int])(
            _ptr(k), None if v is None else _ptr(v), k.size, int(descending))
        assert rc == 0
        return (k, None) if vals is None else (k, v.astype(vals.dtype))

    def stable_sort(self, keys, descending=False):
        return self.stable_sort_pairs(keys, None, descending)[0]


class Reference(_Base):
    """oracle/_ref/libraja_ref.so: RAJA's own seq_exec (pol=0) / omp_parallel_for_exec (pol=1)."""
    prefix = "ref"
    kind = "reference"

    def __init__(self, pol=0, threads=None):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        self.lib = ctypes.CDLL(REF_SO)
        self.pol_args = (pol,)
        self.lib.ref_omp_max_threads.restype = ctypes.c_int
        if threads and pol != 0:
            # explicit: launchers (torchrun) export OMP_NUM_THREADS=1 to their children
            self.lib.ref_omp_set_num_threads.argtypes = [ctypes.c_int]
            self.lib.ref_omp_set_num_threads(int(threads))
        self.cores = 1 if pol == 0 else self.lib.ref_omp_max_threads()

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def multi(self, x):
        sfx = SUFFIX[x.dtype]
        T = CT[sfx]
        s, mn, mx, lv, hv = T(), T(), T(), T(), T()
        ll, hl = ctypes.c_int64(), ctypes.c_int64()
        P = ctypes.POINTER
        self._fn(f"multi_{sfx}", None, [ctypes.c_void_p, ctypes.c_int64, P(T), P(T), P(T), P(T), P(ctypes.c_int64), P(T), P(ctypes.c_int64)])(
            _ptr(x), x.size, *(ctypes.byref(o) for o in (s, mn, mx, lv, ll, hv, hl)))
        t = x.dtype.type
        return {"sum": t(s.value), "min": t(mn.value), "max": t(mx.value), "minloc": (t(lv.value), ll.value), "maxloc": (t(hv.value), hl.value)}

    def sort_pairs(self, keys, vals, descending=False, stable=True):
        sfx = SUFFIX[keys.dtype]
        k = keys.copy()
        if vals.dtype.itemsize == 4:
            v = vals.copy().view(np.uint32)
            name = f"sort_pairs32_{sfx}"
        else:
            v = vals.copy().view(np.uint64)
            name = f"sort_pairs_{sfx}"
        self._fn(name, None, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int])(
            int(stable), _ptr(k), _ptr(v), k.size, int(descending))
        return k, v.view(vals.dtype)

    def stable_sort_pairs(self, keys, vals=None, descending=False):
        if vals is None:
            return self.stable_sort(keys, descending), None
        return self.sort_pairs(keys, vals, descending, stable=True)

    def sort(self, keys, descending=False, stable=False):
        sfx = SUFFIX[keys.dtype]
        k = keys.copy()
        self._fn(f"sort_{sfx}", None, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int])(int(stable), _ptr(k), k.size, int(descending))
        return k

    def stable_sort(self, keys, descending=False):
        return self.sort(keys, descending, stable=True)
