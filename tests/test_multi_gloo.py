"""world_size-2 (and 3) runs of the sharded-range exchange logic on CPU (gloo); the local kernels are
replaced by an oracle-backed double (tests/multi_worker.py), the collectives and the carry / splitter /
rank-order logic are the product code in raja_b200/multi.py."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world, mode=None, timeout=300):
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "multi_worker.py"), str(r), str(world), str(port)] + ([mode] if mode else []),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{out[-3000:]}"
        assert "multi OK" in out


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_paths_gloo(world):
    _run(world)


@pytest.mark.gpu
def test_sharded_paths_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, "cuda")
