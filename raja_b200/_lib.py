"""ctypes binding of libraja_b200.so (the C ABI declared in include/raja_b200.h).

There is no fallback: if the CUDA library is missing or fails to load, importing
this module raises, and every product entry point raises with it.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libraja_b200.so")

# enums of include/raja_b200.h
I32, U32, I64, U64, F32, F64 = range(6)
PLUS, MINIMUM, MAXIMUM = range(3)
REDUCE_SUM, REDUCE_MIN, REDUCE_MAX, REDUCE_MINLOC, REDUCE_MAXLOC = range(5)

c_void_p, c_int, c_int64, c_size_t, c_double = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t, ctypes.c_double)

# name -> (restype, argtypes); kept in lock-step with include/raja_b200.h
# (tests/test_cabi_symbols.py parses the header and checks every prototype is here)
SIGNATURES = {
    "rb200_version": (c_int, []),
    "rb200_last_error": (ctypes.c_char_p, []),
    "rb200_sm_count": (c_int, []),
    "rb200_stream_synchronize": (c_int, [c_void_p]),
    "rb200_device_synchronize": (c_int, []),
    "rb200_stream_mark_launch": (None, [c_void_p]),
    "rb200_stream_mark_synced": (None, [c_void_p]),
    "rb200_synchronize_streams": (c_int, []),
    "rb200_launch_begin": (None, [ctypes.c_char_p, c_void_p]),
    "rb200_launch_end": (None, [ctypes.c_char_p, c_void_p]),
    "rb200_range_push": (None, [ctypes.c_char_p]),
    "rb200_range_pop": (None, []),
    "rb200_launch_count": (ctypes.c_longlong, []),
    "rb200_profile_enable": (c_int, [c_int]),
    "rb200_profile_reset": (c_int, []),
    "rb200_profile_query": (c_int, [ctypes.c_char_p, ctypes.POINTER(c_double), ctypes.POINTER(ctypes.c_longlong)]),
    "rb200_device_malloc": (c_void_p, [c_size_t, c_void_p]),
    "rb200_device_zeroed_malloc": (c_void_p, [c_size_t, c_void_p]),
    "rb200_pinned_malloc": (c_void_p, [c_size_t]),
    "rb200_device_free": (c_int, [c_void_p, c_void_p]),
    "rb200_device_zeroed_free": (c_int, [c_void_p, c_void_p]),
    "rb200_pinned_free": (c_int, [c_void_p]),
    "rb200_pool_trim": (c_int, []),
    "rb200_pool_bytes_cached": (c_size_t, []),
    "rb200_forall_triad": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_double, c_int64, c_void_p]),
    "rb200_forall_daxpy": (c_int, [c_int, c_void_p, c_void_p, c_double, c_int64, c_void_p]),
    "rb200_reduce": (c_int, [c_int, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p]),
    "rb200_reduce_dot": (c_int, [c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "rb200_reduce_multi": (c_int, [c_int, c_void_p, c_int64, c_void_p, c_void_p]),
    "rb200_reduce_multi_result_bytes": (c_size_t, [c_int]),
    "rb200_reduce_multi_loc_offset": (c_size_t, [c_int]),
    "rb200_scan": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "rb200_scan_carry": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "rb200_sort": (c_int, [c_int, c_void_p, c_int64, c_int, c_void_p]),
    "rb200_sort_pairs": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p]),
    "rb200_sort_pairs_buffered": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p]),
    "rb200_partition_pairs": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "rb200_comm_create": (c_int, [c_int, c_int, ctypes.POINTER(c_void_p), c_size_t, ctypes.POINTER(c_void_p)]),
    "rb200_comm_destroy": (c_int, [c_void_p]),
    "rb200_comm_header_bytes": (c_size_t, []),
    "rb200_comm_allgather_small": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "rb200_comm_barrier": (c_int, [c_void_p, c_void_p]),
    "rb200_comm_exchange_pairs": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_int, c_int, c_size_t, c_size_t, c_int64,
                                          ctypes.POINTER(c_int64), c_void_p]),
    "rb200_comm_sample_exchange_pairs": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_size_t, c_size_t, c_int64,
                                                 ctypes.POINTER(c_int64), c_void_p]),
    "rb200_comm_allreduce_small": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "rb200_atomic_histogram": (c_int, [c_int, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "rb200_atomic_pi": (c_int, [c_void_p, c_int64, c_int, c_void_p]),
}


class B200Error(RuntimeError):
    """A C-ABI call failed (mirrors the reference's std::runtime_error on CUDA errors,
    tpl/camp/include/camp/defines.hpp:206-233)."""


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"raja_b200: {LIB_PATH} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C raja_b200/csrc`. There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc, what):
    if rc != 0:
        msg = lib.rb200_last_error()
        raise B200Error(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")
