"""Worker for tests/test_multi_gloo.py: one rank of a world_size-N gloo job that drives
raja_b200.multi with a CPU double of the local ops (built on the test oracle), and compares every
path with the oracle run on the whole (unsharded) input.  TEST INFRASTRUCTURE.

    python tests/multi_worker.py <rank> <world> <port> [cuda]
With `cuda` the local ops are the real b200 library (nccl backend) -- used by the -m gpu tests.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle_lib import Oracle  # noqa: E402

NP2T = {np.dtype(np.int32): torch.int32, np.dtype(np.uint32): torch.uint32, np.dtype(np.int64): torch.int64,
        np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32}


class OracleOps:
    """CPU stand-in for raja_b200.multi.B200Ops: same methods, computed by the oracle."""

    def __init__(self):
        self.orc = Oracle()

    @staticmethod
    def _np(t):
        return t.numpy() if t.dtype not in (torch.uint32,) else t.view(torch.int32).numpy().view(np.uint32)

    def reduce(self, kind, x):
        a = self._np(x)
        loc = -1
        if kind == 0:
            v = self.orc.sum(a)
        elif kind == 1:
            v = self.orc.min(a)
        elif kind == 2:
            v = self.orc.max(a)
        elif kind == 3:
            v, loc = self.orc.minloc(a)
        else:
            v, loc = self.orc.maxloc(a)
        return torch.from_numpy(np.array([v], dtype=a.dtype)), torch.tensor([loc], dtype=torch.int64)

    def dot(self, x, y):
        return torch.from_numpy(np.array([self.orc.dot(self._np(x), self._np(y))]))

    def scan(self, inclusive, op, src, dst, init, carry):
        name = ["plus", "minimum", "maximum"][op]
        a = self._np(src)
        f = {"plus": np.add, "minimum": np.minimum, "maximum": np.maximum}[name]
        if inclusive:
            out = self.orc.inclusive_scan(a, name)
            if carry is not None:
                out = f(out, carry.numpy()[0]).astype(a.dtype)
        else:
            ident = {"plus": 0, "minimum": (np.finfo if a.dtype.kind == "f" else np.iinfo)(a.dtype).max,
                     "maximum": (np.finfo if a.dtype.kind == "f" else np.iinfo)(a.dtype).min}[name]
            seed = a.dtype.type(ident if init is None else init)
            if carry is not None:
                seed = f(carry.numpy()[0], seed).astype(a.dtype)
            out = self.orc.exclusive_scan(a, name, seed)
        dst.copy_(torch.from_numpy(out))

    def sort_pairs(self, keys, vals, descending, keys_alt=None, vals_alt=None):
        k = self._np(keys)
        if vals is None:
            ok = self.orc.stable_sort(k, descending)
            self._np(keys)[:] = ok
        else:
            ok, ov = self.orc.stable_sort_pairs(k, self._np(vals), descending)
            self._np(keys)[:] = ok
            self._np(vals)[:] = ov

    def partition_pairs(self, keys, vals, splitters, descending, keys_out=None, vals_out=None):
        """same contract as rb200_partition_pairs: bucket = number of splitters that sort at or before the key, stable"""
        k = self._np(keys)
        sp = np.asarray(splitters, dtype=k.dtype)
        if sp.size:
            b = ((k[:, None] <= sp[None, :]) if descending else (k[:, None] >= sp[None, :])).sum(axis=1)
        else:
            b = np.zeros(k.size, dtype=np.int64)
        order = np.argsort(b, kind="stable")
        counts = torch.zeros(8, dtype=torch.int64)
        counts[:sp.size + 1] = torch.from_numpy(np.bincount(b, minlength=sp.size + 1).astype(np.int64))
        ko = keys_out if keys_out is not None else torch.empty_like(keys)
        self._np(ko)[:] = k[order]
        vo = None
        if vals is not None:
            vo = vals_out if vals_out is not None else torch.empty_like(vals)
            self._np(vo)[:] = self._np(vals)[order]
        return ko, vo, counts

    def histogram(self, hist, keys):
        hist += torch.from_numpy(self.orc.histogram(keys.numpy(), hist.numel()))

    def triad(self, a, b, c, s):
        a.copy_(torch.from_numpy(self.orc.triad(b.numpy(), c.numpy(), s)))

    def daxpy(self, a, b, s):
        a.copy_(torch.from_numpy(self.orc.daxpy(a.numpy(), b.numpy(), s)))


def main():
    rank, world, port = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    cuda = len(sys.argv) > 4 and sys.argv[4] == "cuda"
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    if cuda:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        dev = torch.device("cuda", rank)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        dev = torch.device("cpu")
    from raja_b200 import multi
    orc = Oracle()
    ops = multi.B200Ops() if cuda else OracleOps()

    def T(a):  # numpy -> tensor on this rank's device
        t = torch.from_numpy(a.view(np.int32)).view(torch.uint32) if a.dtype == np.uint32 else torch.from_numpy(a)
        return t.to(dev)

    def N(t):  # tensor -> numpy
        t = t.cpu()
        return t.view(torch.int32).numpy().view(np.uint32) if t.dtype == torch.uint32 else t.numpy()

    checks = 0
    # on GPUs every path runs twice: exchanges by NCCL collectives, then over the NVLink peer-memory communicator
    # (the third communicator's receive buffers are too small for the largest size: the sort must fall back, on every rank
    # alike, to the collective all-to-all)
    comms = [None, multi.PeerComm(exchange_bytes=32 << 20), multi.PeerComm(exchange_bytes=2 << 20)] if cuda else [None]
    for comm, n_global in [(c, n) for c in comms for n in ((0, 1, 5, 1000, 20011) if not cuda else (5, 1000, 1 << 20))]:
        rng_np = np.random.default_rng(1234 + n_global)  # same stream on every rank
        sr = multi.ShardedRange.split(n_global, world, rank)
        sr.comm = comm
        lo, hi = sr.offset, sr.offset + sr.n_local
        # ---- reductions (int: bit-exact; float: n*eps)
        xi = rng_np.integers(-100, 100, n_global).astype(np.int32)
        if n_global:
            xi[rng_np.integers(0, n_global, 3)] = -100  # duplicate extrema across shards: first index must win
            xi[rng_np.integers(0, n_global, 3)] = 99
        assert multi.reduce_sum(sr, T(xi[lo:hi]), ops=ops) == orc.sum(xi)
        assert multi.reduce_min(sr, T(xi[lo:hi]), ops=ops) == orc.min(xi)
        assert multi.reduce_max(sr, T(xi[lo:hi]), ops=ops) == orc.max(xi)
        if n_global:
            assert multi.reduce_minloc(sr, T(xi[lo:hi]), ops=ops) == tuple(int(v) for v in orc.minloc(xi)), n_global
            assert multi.reduce_maxloc(sr, T(xi[lo:hi]), ops=ops) == tuple(int(v) for v in orc.maxloc(xi)), n_global
        xf = rng_np.uniform(0.5, 1.5, n_global)
        yf = rng_np.uniform(0.5, 1.5, n_global)
        d = multi.reduce_dot(sr, T(xf[lo:hi]), T(yf[lo:hi]), ops=ops)
        assert abs(d - float(orc.dot(xf, yf))) <= max(n_global, 1) * np.finfo(np.float64).eps * float((xf * yf).sum() + 1)
        # ---- scans: carry exchange
        for op, name in ((0, "plus"), (1, "minimum"), (2, "maximum")):
            src = T(xi[lo:hi].copy())
            dst = torch.empty_like(src)
            multi.inclusive_scan(sr, src, dst, op, ops=ops)
            assert np.array_equal(N(dst), orc.inclusive_scan(xi, name)[lo:hi]), (n_global, name)
            multi.exclusive_scan(sr, src, dst, op, init=7, ops=ops)
            assert np.array_equal(N(dst), orc.exclusive_scan(xi, name, 7)[lo:hi]), (n_global, name)
        x64 = rng_np.integers(-1000, 1000, n_global).astype(np.int64)
        src = T(x64[lo:hi].copy())
        multi.inclusive_scan(sr, src, src, 0, ops=ops)  # in place
        assert np.array_equal(N(src), orc.inclusive_scan(x64)[lo:hi])
        # ---- sample sort: the concatenation over ranks must equal the oracle's stable sort
        for keys in (rng_np.integers(0, 2**32, n_global, dtype=np.uint64).astype(np.uint32),
                     (rng_np.integers(0, 2**32, n_global, dtype=np.uint64) & 0x7).astype(np.uint32),   # heavy duplicates
                     rng_np.integers(-50, 50, n_global).astype(np.int32),
                     (rng_np.integers(-2000, 2000, n_global) / 8.0).astype(np.float64),                 # floating keys, duplicates, both signs
                     rng_np.standard_normal(n_global).astype(np.float32)):
            vals = np.arange(n_global, dtype=np.uint32)
            for desc in (False, True):
                ko, vo = multi.sort_pairs(sr, T(keys[lo:hi].copy()), T(vals[lo:hi].copy()), descending=desc, ops=ops)
                ko, vo = N(ko), N(vo)
                sizes = [None] * world
                dist.all_gather_object(sizes, int(ko.size))
                start = sum(sizes[:rank])
                ek, ev = orc.stable_sort_pairs(keys, vals, desc)
                assert sum(sizes) == n_global
                assert np.array_equal(ko, ek[start:start + ko.size]), (n_global, desc)
                assert np.array_equal(vo, ev[start:start + vo.size]), (n_global, desc)
        kf = rng_np.uniform(-1e3, 1e3, n_global)
        ko = N(multi.sort(sr, T(kf[lo:hi].copy()), ops=ops))
        sizes = [None] * world
        dist.all_gather_object(sizes, int(ko.size))
        start = sum(sizes[:rank])
        assert np.array_equal(ko, orc.stable_sort(kf)[start:start + ko.size])
        # ---- histogram: private bins, all_reduce
        hk = rng_np.integers(0, 256, n_global).astype(np.int32)
        hist = torch.zeros(256, dtype=torch.int32, device=dev)
        multi.atomic_histogram(sr, hist, T(hk[lo:hi]), ops=ops)
        assert np.array_equal(N(hist), orc.histogram(hk, 256))
        # ---- elementwise forall: no exchange
        b, c = rng_np.uniform(0, 1, n_global), rng_np.uniform(0, 1, n_global)
        a = torch.zeros(hi - lo, dtype=torch.float64, device=dev)
        multi.forall_triad(sr, a, T(b[lo:hi]), T(c[lo:hi]), 3.0, ops=ops)
        if cuda:
            torch.cuda.synchronize()
        assert np.array_equal(N(a), orc.triad(b, c, 3.0)[lo:hi])
        checks += 1
    dist.barrier()
    for c in comms:
        if c is not None:
            c.close()
    dist.destroy_process_group()
    print(f"rank {rank}: multi OK ({checks} sizes)")


if __name__ == "__main__":
    main()
