"""The C-ABI library loads without a GPU and exports every symbol include/raja_b200.h declares;
the ctypes table in raja_b200/_lib.py covers exactly the same set (no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "raja_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rb200_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = header_functions()
    for must in ["rb200_forall_triad", "rb200_reduce", "rb200_reduce_multi", "rb200_scan", "rb200_sort_pairs",
                 "rb200_atomic_histogram", "rb200_device_malloc", "rb200_pinned_malloc"]:
        assert must in names


def test_library_exports_every_declared_symbol():
    from raja_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), f"{name} declared in include/raja_b200.h but not exported"


def test_ctypes_table_matches_header():
    from raja_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_functions()


def test_version_and_argument_errors_without_gpu():
    from raja_b200 import _lib
    assert _lib.lib.rb200_version() == 100
    # argument validation happens before any CUDA call
    assert _lib.lib.rb200_scan(1, 0, 0, None, None, -1, None, None) == -1
    assert b"rb200_scan" in _lib.lib.rb200_last_error()
    assert _lib.lib.rb200_sort_pairs(0, None, None, 3, 10, 0, None) == -1
    assert _lib.lib.rb200_reduce(9, 0, None, 0, ctypes.c_void_p(8), None, 1, None) == -1


def test_host_tensors_are_rejected_not_computed_on_cpu():
    import pytest
    import torch
    import raja_b200 as rb
    x = torch.arange(8, dtype=torch.int32)
    with pytest.raises(ValueError):
        rb.reduce_sum(x)
    with pytest.raises(ValueError):
        rb.inclusive_scan(x, x.clone())
