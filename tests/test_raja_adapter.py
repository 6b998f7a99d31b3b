"""The RAJA policy adapter (include/RAJA_b200/): C++ test programs that run the SAME lambdas under the
reference's seq_exec and under RAJA::b200_exec / b200_reduce / b200_atomic and compare (tests/cpp/*.cu).
They are compiled in the build container against the reference headers (tests/cpp/Makefile, called by
__graft_entry__.build()); the binaries travel to the GPU box."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "cpp", "_build")
PROGRAMS = ["test_forall_reduce", "test_scan_sort_atomic", "test_multi_reduce", "test_launch_reduce", "test_workgroup", "test_multi"]


def test_adapter_headers_present():
    inc = os.path.join(os.path.dirname(HERE), "include", "RAJA_b200")
    for h in ("RAJA.hpp", "policy.hpp", "pools.hpp", "forall.hpp", "reduce.hpp", "reduce_classic.hpp", "scan.hpp", "sort.hpp", "atomic.hpp", "multi_reduce.hpp", "launch.hpp", "workgroup.hpp", "kernel.hpp", "indexset.hpp", "multi.hpp"):
        assert os.path.exists(os.path.join(inc, h)), h


@pytest.mark.skipif(not os.path.exists("/root/reference/include/RAJA/RAJA.hpp"), reason="reference tree not present")
def test_adapter_programs_were_built():
    for p in PROGRAMS:
        assert os.path.exists(os.path.join(BUILD, p)), f"{p} missing: run __graft_entry__.build()"


@pytest.mark.gpu
@pytest.mark.parametrize("prog", PROGRAMS)
def test_adapter_parity_on_gpu(prog):
    exe = os.path.join(BUILD, prog)
    assert os.path.exists(exe), f"{exe} missing (built by __graft_entry__.build() in the build container)"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert ": OK" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("world", [3, 8])
def test_multi_policy_world_logic_on_one_gpu(world):
    """The C++ multi-GPU policy with MORE ranks than devices (ranks are host threads, rank r on device r % ndev): the world-3 / world-8
    logic -- seven splitters, shards of different lengths, EMPTY shards (5 elements over 8 ranks), peer exchanges -- runs on whatever
    the box has; on real 8 GPUs the same program runs one rank per device (profiles/r02_test_multi_cpp.txt)."""
    exe = os.path.join(BUILD, "test_multi")
    assert os.path.exists(exe), f"{exe} missing (built by __graft_entry__.build() in the build container)"
    env = dict(os.environ, RB200_OVERSUBSCRIBE="1", RB200_WORLD=str(world))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert f"OK (world = {world})" in r.stdout
