"""The reference's OWN typed gtest suites, instantiated for the b200 back-end (tests/refsuite/suites.py does what the
reference's cmake does with its *.cpp.in templates, with BACKEND = B200; test logic, data and assertions are the
reference's, only the policy lists in tests/refsuite/b200_lists.hpp are ours).  Built in the build container by
__graft_entry__.build(); the binaries travel to the GPU box."""
import os
import re
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "refsuite", "_build")
LIST = os.path.join(BUILD, "suites.txt")
HAVE_REF = os.path.exists("/root/reference/test/include/RAJA_test-base.hpp")


def _suites():
    if os.path.exists(LIST):
        return [ln.strip() for ln in open(LIST) if ln.strip()]
    return []


def test_refsuite_manifest_and_lists_present():
    assert os.path.exists(os.path.join(HERE, "refsuite", "suites.py"))
    assert os.path.exists(os.path.join(HERE, "refsuite", "b200_lists.hpp"))


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present")
def test_refsuite_binaries_were_built():
    names = _suites()
    assert len(names) >= 100, "run __graft_entry__.build()"
    missing = [n for n in names if not os.path.exists(os.path.join(BUILD, "bin", n))]
    assert not missing, f"not built: {missing}"


@pytest.mark.gpu
@pytest.mark.parametrize("suite", _suites() or ["<none>"])
def test_reference_suite_on_gpu(suite):
    assert suite != "<none>", "tests/refsuite/_build/suites.txt missing (built by __graft_entry__.build() in the build container)"
    exe = os.path.join(BUILD, "bin", suite)
    assert os.path.exists(exe), f"{exe} missing"
    r = subprocess.run([exe, "--gtest_brief=1", "--gtest_color=no"], capture_output=True, text=True, timeout=900)
    tail = r.stdout[-6000:] + r.stderr[-2000:]
    assert r.returncode == 0, tail
    m = re.search(r"\[  PASSED  \] (\d+) test", r.stdout)
    assert m and int(m.group(1)) > 0, tail
    assert "FAILED" not in r.stdout, tail
