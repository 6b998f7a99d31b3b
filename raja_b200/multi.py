"""Sharded-range execution across the GPUs of one node: one process per GPU, contiguous
index shards, collectives only where the path has a real exchange step.

    path              local work (b200 kernels)                 exchange (torch.distributed: NCCL on GPUs)
    forall            on the shard, lambda sees global indices   none
    reduce sum/min/max  grid reduce -> device scalar             all_reduce (1 element)
    reduce min/maxloc   grid reduce -> (val, global loc)         all_gather, fold in rank order (first index wins)
    scan              shard total -> carry -> seeded scan        all_gather of world_size totals
    sort[_pairs]      regular sampling of the UNSORTED shard,    all_gather (samples), all_to_all_single (counts),
                      one stable splitter partition (a single    all_to_all_single (keys, values)
                      onesweep pass), ONE local onesweep sort
                      of what arrives
    atomic histogram  private bins per GPU                       all_reduce over the bins

RAJA has no multi-GPU policy (the reference runs one device per process and leaves the rest
to MPI); the surface here is the minimal one: a ShardedRange says which contiguous piece of the
global iteration space this rank owns, and every function takes the LOCAL shard tensors.

The local work goes through an `ops` object.  The default, B200Ops, is the CUDA library and
nothing else (it raises without a GPU).  tests/ inject a CPU double built on the oracle to
exercise the exchange logic with the gloo backend.
"""
import torch
import torch.distributed as dist

__all__ = ["ShardedRange", "B200Ops", "PeerComm", "reduce_sum", "reduce_min", "reduce_max", "reduce_dot", "reduce_minloc", "reduce_maxloc",
           "inclusive_scan", "exclusive_scan", "sort", "sort_pairs", "atomic_histogram", "forall_triad", "forall_daxpy"]


class ShardedRange:
    """Rank `rank` of `world` owns global indices [offset, offset + n_local)."""

    def __init__(self, world, rank, n_local=None, offset=None, group=None, comm=None):
        self.world, self.rank, self.group = world, rank, group
        self.comm = comm  # PeerComm: exchanges go over NVLink peer memory instead of torch.distributed collectives
        self.n_local = n_local
        self.offset = offset if offset is not None else (rank * n_local if n_local is not None else None)

    def for_length(self, n_local):
        """Same range, shard length filled in (equal shards) when it was left open."""
        if self.n_local == n_local and self.offset is not None:
            return self
        return ShardedRange(self.world, self.rank, n_local, self.offset if self.offset is not None else self.rank * n_local, self.group, self.comm)

    @staticmethod
    def split(n_global, world, rank, group=None):
        """Contiguous near-equal shards of a global range (the first n % world ranks get one more)."""
        base, rem = divmod(n_global, world)
        n_local = base + (1 if rank < rem else 0)
        offset = rank * base + min(rank, rem)
        return ShardedRange(world, rank, n_local, offset, group)


class PeerComm:
    """NVLink / NVSwitch peer-memory communicator (rb200_comm_* in include/raja_b200.h).

    Every rank owns one symmetric workspace that all processes of the node have mapped (torch
    symmetric memory does the handle exchange: plumbing); the library's kernels store straight into
    the peers' workspaces.  Used for the small exchanges of the sharded paths (one 1-CTA kernel
    instead of a NCCL collective) and for the sort, whose splitter sweep writes every bucket directly
    into its destination GPU (fused partition + all-to-all).  `exchange_bytes`: room for the sort's
    receive buffers (keys and values of ~1.25x a shard); 0 = small exchanges only."""

    def __init__(self, group=None, exchange_bytes=0, device=None):
        import ctypes
        import torch.distributed._symmetric_memory as symm
        from ._lib import lib, check
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.header = int(lib.rb200_comm_header_bytes())
        self.nbytes = self.header + ((int(exchange_bytes) + 63) // 64) * 64
        self.buf = symm.empty(self.nbytes, dtype=torch.uint8, device=self.device)
        self.hdl = symm.rendezvous(self.buf, self.group)
        ptrs = (ctypes.c_void_p * self.world)(*[int(p) for p in self.hdl.buffer_ptrs])
        handle = ctypes.c_void_p()
        check(lib.rb200_comm_create(self.rank, self.world, ptrs, self.nbytes, ctypes.byref(handle)), "comm_create")
        self.handle = handle
        torch.cuda.synchronize(self.device)
        self.hdl.barrier()  # every header is zeroed before anyone stores into it

    def close(self):
        from ._lib import lib
        if getattr(self, "handle", None):
            lib.rb200_comm_destroy(self.handle)
            self.handle = None

    def allgather_small(self, t, res):
        """t: small contiguous device tensor -> [world, *t.shape] in rank order (stream ordered, no host sync)"""
        from ._lib import lib, check
        nb = t.numel() * t.element_size()
        pad = (-nb) % 8
        src = t.contiguous().view(torch.uint8).reshape(-1)
        if pad:
            src = torch.cat([src, torch.zeros(pad, dtype=torch.uint8, device=t.device)])
        out = torch.empty(self.world, nb + pad, dtype=torch.uint8, device=t.device)
        check(lib.rb200_comm_allgather_small(self.handle, src.data_ptr(), nb + pad, out.data_ptr(), res.handle), "comm_allgather_small")
        return out[:, :nb].contiguous().view(t.dtype).reshape(self.world, *t.shape)

    def allreduce_small(self, t, op, res, exclusive_prefix=False):
        """t: small contiguous device tensor (<= 504 bytes) -> same shape: the fold with rb200_op `op` over all ranks, in rank
        order (exclusive_prefix: over the ranks below this one; the identity on rank 0).  One peer-memory kernel, no host sync."""
        from ._lib import lib, check
        from . import api
        src = t.contiguous()
        if src.data_ptr() % 8:
            src = src.clone()
        out = torch.empty_like(src)
        check(lib.rb200_comm_allreduce_small(self.handle, int(op), api._dt(src), src.data_ptr(), out.data_ptr(), src.numel(), int(bool(exclusive_prefix)),
                                             res.handle), "comm_allreduce_small")
        return out

    def barrier(self, res):
        from ._lib import lib, check
        check(lib.rb200_comm_barrier(self.handle, res.handle), "comm_barrier")

    def capacity(self, key_bytes, val_bytes):
        """elements the receive buffers of one exchange can take"""
        return max((self.nbytes - self.header - 128) // (key_bytes + val_bytes), 0)

    def exchange_pairs(self, keys, vals, splitters, descending, res):
        """fused partition + all-to-all -> (keys_recv, vals_recv): views of this rank's receive buffers,
        valid until the next exchange on this communicator"""
        import ctypes
        import numpy as np
        from ._lib import lib, check
        from . import api
        ks, vb = keys.element_size(), (vals.element_size() if vals is not None else 0)
        cap = self.capacity(ks, vb)
        koff = self.header
        voff = koff + ((cap * ks + 127) // 128) * 128  # 128-byte aligned: the split sweep writes whole lines
        dt = api._dt(keys)
        sp = np.ascontiguousarray(np.asarray(splitters, dtype=api._NP[dt]))
        got = ctypes.c_int64(0)
        check(lib.rb200_comm_exchange_pairs(self.handle, dt, keys.data_ptr(), vals.data_ptr() if vb else None, vb, keys.numel(),
                                            sp.ctypes.data_as(ctypes.c_void_p) if sp.size else None, int(sp.size), int(descending),
                                            koff, voff, cap, ctypes.byref(got), res.handle), "comm_exchange_pairs")
        n_out = int(got.value)
        kout = self.buf[koff:koff + n_out * ks].view(keys.dtype)
        vout = self.buf[voff:voff + n_out * vb].view(vals.dtype) if vb else None
        return kout, vout


def _peer_sample_exchange(comm, keys, vals, samples, descending, res):
    """sampling, splitter selection (on the device), fused partition + all-to-all -> (keys_recv, vals_recv) views of this
    rank's receive buffers, or None when some rank's share does not fit (decided identically on every rank)"""
    import ctypes
    from ._lib import lib, check, B200Error
    from . import api
    ks, vb = keys.element_size(), (vals.element_size() if vals is not None else 0)
    cap = comm.capacity(ks, vb)
    if cap <= 0:
        return None
    koff = comm.header
    voff = koff + ((cap * ks + 127) // 128) * 128
    got = ctypes.c_int64(0)
    rc = lib.rb200_comm_sample_exchange_pairs(comm.handle, api._dt(keys), keys.data_ptr(), vals.data_ptr() if vb else None, vb, keys.numel(),
                                              int(samples), int(descending), koff, voff, cap, ctypes.byref(got), res.handle)
    if rc == -3:  # RB200_ENOMEM on every rank: fall back to the collective path
        return None
    check(rc, "comm_sample_exchange_pairs")
    n_out = int(got.value)
    kout = comm.buf[koff:koff + n_out * ks].view(keys.dtype)
    vout = comm.buf[voff:voff + n_out * vb].view(vals.dtype) if vb else None
    return kout, vout


class B200Ops:
    """Local work on the b200 back-end (CUDA tensors; no fallback)."""

    def __init__(self, res=None):
        from . import api
        self.api = api
        self.res = res if res is not None else api.Resource()
        self.apol = api.b200_exec(async_=True)

    # every method enqueues on self.res.stream and returns device tensors (no host sync)
    def reduce(self, kind, x):
        """-> (value tensor [1] of x.dtype, loc tensor [1] int64) on the device"""
        from ._lib import lib, check
        val = torch.empty(1, dtype=x.dtype, device=x.device)
        loc = torch.empty(1, dtype=torch.int64, device=x.device)
        check(lib.rb200_reduce(kind, self.api._dt(x), x.data_ptr(), x.numel(), val.data_ptr(), loc.data_ptr(), 1, self.res.handle), "reduce")
        return val, loc

    def dot(self, x, y):
        from ._lib import lib, check
        val = torch.empty(1, dtype=x.dtype, device=x.device)
        check(lib.rb200_reduce_dot(self.api._dt(x), x.data_ptr(), y.data_ptr(), x.numel(), val.data_ptr(), self.res.handle), "dot")
        return val

    def scan(self, inclusive, op, src, dst, init, carry):
        from ._lib import lib, check
        import ctypes
        import numpy as np
        dt = self.api._dt(src)
        init_p = None
        if init is not None:
            init_arr = np.array([init], dtype=self.api._NP[dt])
            init_p = init_arr.ctypes.data_as(ctypes.c_void_p)
        check(lib.rb200_scan_carry(1 if inclusive else 0, op, dt, src.data_ptr(), dst.data_ptr(), src.numel(), init_p,
                                   carry.data_ptr() if carry is not None else None, self.res.handle), "scan")

    def sort_pairs(self, keys, vals, descending, keys_alt=None, vals_alt=None):
        if vals is None:
            self.api.stable_sort(keys, int(descending), pol=self.apol, res=self.res)
        else:
            self.api.stable_sort_pairs(keys, vals, int(descending), pol=self.apol, res=self.res, keys_alt=keys_alt, vals_alt=vals_alt)

    def partition_pairs(self, keys, vals, splitters, descending, keys_out=None, vals_out=None):
        """stable split of (keys, vals) by the host `splitters` -> (keys_out, vals_out, counts int64[8] on the device)"""
        keys_out = keys_out if keys_out is not None else torch.empty_like(keys)
        if vals is not None and vals_out is None:
            vals_out = torch.empty_like(vals)
        counts = torch.zeros(8, dtype=torch.int64, device=keys.device)
        self.api.partition_pairs(keys, keys_out, vals, vals_out if vals is not None else None, splitters, counts, int(descending),
                                 pol=self.apol, res=self.res)
        return keys_out, (vals_out if vals is not None else None), counts

    def histogram(self, hist, keys):
        self.api.atomic_histogram(hist, keys, pol=self.apol, res=self.res)

    def triad(self, a, b, c, s):
        self.api.forall_triad(a, b, c, s, pol=self.apol, res=self.res)

    def daxpy(self, a, b, s):
        self.api.forall_daxpy(a, b, s, pol=self.apol, res=self.res)


_KIND = {"sum": 0, "min": 1, "max": 2, "minloc": 3, "maxloc": 4}
_OPS = {0: "plus", 1: "minimum", 2: "maximum"}  # rb200_op values


def _ops(ops):
    return ops if ops is not None else B200Ops()


def _identity(op, dtype):
    if op == 0:
        return 0
    info = torch.finfo(dtype) if dtype.is_floating_point else torch.iinfo(dtype)
    # RAJA::operators::limits: minimum -> max(), maximum -> min() (-FLT_MAX, not -inf)
    return info.max if op == 1 else info.min


def _fold(op, t, dim=0):
    """fold a small tensor with a rb200_op (plumbing on world_size elements)"""
    if op == 0:
        return t.sum(dim=dim, dtype=t.dtype)
    return t.amin(dim=dim) if op == 1 else t.amax(dim=dim)


# ----------------------------------------------------------------------------------------- forall
def forall_triad(rng, a, b, c, s, ops=None):
    """elementwise: every rank runs the loop body over its own shard, no exchange"""
    _ops(ops).triad(a, b, c, s)


def forall_daxpy(rng, a, b, s, ops=None):
    _ops(ops).daxpy(a, b, s)


# ----------------------------------------------------------------------------------------- reductions
def _allreduce(rng, val, red, res=None):
    if rng.world > 1:
        if rng.comm is not None and res is not None:
            # one peer-memory kernel: all-gather + fold in rank order (deterministic for floating point)
            op = {dist.ReduceOp.SUM: 0, dist.ReduceOp.MIN: 1, dist.ReduceOp.MAX: 2}[red]
            val = rng.comm.allreduce_small(val, op, res)
        else:
            dist.all_reduce(val, op=red, group=rng.group)
    return val


def reduce_sum(rng, x, ops=None, as_tensor=False):
    ops = _ops(ops)
    v = _allreduce(rng, ops.reduce(_KIND["sum"], x)[0], dist.ReduceOp.SUM, getattr(ops, "res", None))
    return v if as_tensor else v.item()


def reduce_min(rng, x, ops=None, as_tensor=False):
    ops = _ops(ops)
    v = _allreduce(rng, ops.reduce(_KIND["min"], x)[0], dist.ReduceOp.MIN, getattr(ops, "res", None))
    return v if as_tensor else v.item()


def reduce_max(rng, x, ops=None, as_tensor=False):
    ops = _ops(ops)
    v = _allreduce(rng, ops.reduce(_KIND["max"], x)[0], dist.ReduceOp.MAX, getattr(ops, "res", None))
    return v if as_tensor else v.item()


def reduce_dot(rng, x, y, ops=None, as_tensor=False, res=None):
    ops = _ops(ops) if ops is not None or res is None else B200Ops(res)
    v = _allreduce(rng, ops.dot(x, y), dist.ReduceOp.SUM, getattr(ops, "res", None))
    return v if as_tensor else v.item()


def _reduce_loc(rng, x, kind, ops):
    rng = rng.for_length(x.numel())
    val, loc = _ops(ops).reduce(_KIND[kind], x)
    loc = torch.where(loc >= 0, loc + rng.offset, loc)  # global index; -1 stays "empty shard"
    if rng.world == 1:
        return val.item(), int(loc.item())
    vals = [torch.empty_like(val) for _ in range(rng.world)]
    locs = [torch.empty_like(loc) for _ in range(rng.world)]
    dist.all_gather(vals, val, group=rng.group)
    dist.all_gather(locs, loc, group=rng.group)
    vals, locs = torch.cat(vals).cpu(), torch.cat(locs).cpu()
    # fold in rank order; ranks own ascending index ranges, so "strictly better wins" keeps the
    # FIRST extremum, which is the sequential answer (util/Operators.hpp:555-594)
    best_v, best_l = None, -1
    for v, l in zip(vals.tolist(), locs.tolist()):
        if l < 0:
            continue
        if best_l < 0 or (v < best_v if kind == "minloc" else v > best_v):
            best_v, best_l = v, l
    if best_l < 0:
        best_v = vals[0].item()
    return best_v, best_l


def reduce_minloc(rng, x, ops=None):
    return _reduce_loc(rng, x, "minloc", ops)


def reduce_maxloc(rng, x, ops=None):
    return _reduce_loc(rng, x, "maxloc", ops)


# ----------------------------------------------------------------------------------------- scans
def _scan(rng, inclusive, src, dst, op, init, ops, res):
    ops = ops if ops is not None else B200Ops(res)
    carry = None
    if rng.world > 1:
        # shard total -> all_gather -> fold of the lower ranks' totals, all on the device
        total = ops.reduce(op, src)[0]  # rb200_reduce_kind SUM/MIN/MAX == rb200_op PLUS/MINIMUM/MAXIMUM
        if rng.comm is not None and getattr(ops, "res", None) is not None:
            # the fold of the lower ranks' totals comes out of the exchange kernel itself
            carry = rng.comm.allreduce_small(total, op, ops.res, exclusive_prefix=True)
            if rng.rank == 0:
                carry = None
        else:
            totals = torch.empty(rng.world, dtype=src.dtype, device=src.device)
            dist.all_gather_into_tensor(totals, total, group=rng.group)
            lower = totals[:rng.rank]
            carry = None if rng.rank == 0 else _fold(op, lower).reshape(1).contiguous()
    ops.scan(inclusive, op, src, dst, init, carry)


def inclusive_scan(rng, src, dst, op=0, ops=None, res=None):
    _scan(rng, True, src, dst, op, None, ops, res)


def exclusive_scan(rng, src, dst, op=0, init=None, ops=None, res=None):
    _scan(rng, False, src, dst, op, init, ops, res)


# ----------------------------------------------------------------------------------------- sorts
_SIGNED = {torch.uint32: torch.int32, torch.uint64: torch.int64}
_SAMPLE_POS = {}  # (shard length, samples, device) -> sample positions (constant across calls)


def _signed_view(t):
    """same bits as a signed tensor: torch's indexing ops and the collectives do not take uint32/uint64"""
    return t.view(_SIGNED[t.dtype]) if t.dtype in _SIGNED else t


def _wide(small, orig_dtype):
    """order-preserving int64/float64 copy of a TINY tensor taken from a signed view (samples, probes)"""
    if orig_dtype == torch.uint32:
        return small.to(torch.int64) & 0xFFFFFFFF
    if orig_dtype == torch.uint64:
        return small ^ torch.iinfo(torch.int64).min
    if orig_dtype.is_floating_point:
        return small.to(torch.float64)
    return small.to(torch.int64)


def _bounds(kview, orig_dtype, n, targets, descending):
    """For each target: how many keys of the locally sorted shard sort strictly before it
    (ascending: key < t; descending: key >= t ... i.e. the first index whose key sorts at/after t
    for ascending, after t for descending).  Vectorised binary search, log2(n) tiny gathers."""
    m = targets.numel()
    if n == 0:
        return torch.zeros(m, dtype=torch.int64, device=targets.device)
    lo = torch.zeros(m, dtype=torch.int64, device=targets.device)
    hi = torch.full((m,), n, dtype=torch.int64, device=targets.device)
    for _ in range(max(n, 1).bit_length() + 1):
        active = lo < hi
        mid = (lo + hi) // 2
        v = _wide(kview[mid.clamp(max=max(n - 1, 0))], orig_dtype)
        right = (v >= targets) if descending else (v < targets)
        lo = torch.where(active & right, mid + 1, lo)
        hi = torch.where(active & ~right, mid, hi)
    return lo


def sort_pairs(rng, keys, vals=None, descending=False, ops=None, res=None, keys_alt=None, vals_alt=None, oversample=512):
    """Stable sample sort.  keys/vals: this rank's shard.  Returns (keys_out, vals_out): this rank's
    piece of the globally sorted sequence -- rank r holds keys that sort before rank r+1's; piece
    sizes differ with the sampling error (oversample samples per rank and destination).

    One rank: the local sort, in place.  Several: splitters from a regular sample of the UNSORTED
    shards, one stable splitter partition per rank (a single onesweep pass, rb200_partition_pairs),
    the all-to-all, and ONE local stable sort of what arrives.  Equal keys always meet on one rank
    and arrive ordered by (source rank, position in the source shard), so the result is the
    stable sort of the concatenated shards.  keys_alt / vals_alt: optional scratch of the shard's
    size (the partition output); keys / vals are left untouched when world > 1."""
    ops = ops if ops is not None else B200Ops(res)
    if rng.world == 1:
        ops.sort_pairs(keys, vals, descending, keys_alt, vals_alt)
        return keys, vals
    W, n = rng.world, keys.numel()
    if W - 1 > 7:
        raise ValueError("sort_pairs: at most 8 ranks (one node)")
    dev, kdt = keys.device, keys.dtype
    comm = rng.comm if getattr(ops, "res", None) is not None else None
    if comm is not None:
        # everything up to the local sort inside the library, on the stream: samples all-gathered through peer memory,
        # splitters selected on the device, fused partition + all-to-all; one host sync (the received count)
        got = _peer_sample_exchange(comm, keys, vals, min(W * oversample, 4096), descending, ops.res)
        if got is not None:
            keys_out, vals_out = got
            ops.sort_pairs(keys_out, vals_out, descending, keys_alt if keys_alt is not None and keys_alt.numel() >= keys_out.numel() else None,
                           vals_alt if vals_alt is not None and vals_out is not None and vals_alt.numel() >= vals_out.numel() else None)
            return keys_out, vals_out
        # some rank's share exceeds its receive buffer (the decision is the same on every rank): collective path below
    kview = _signed_view(keys)
    # regular sample of the unsorted shard: [count, samples ...] travels in ONE all-gather (count <= W * oversample is exact in
    # every key type), invalid slots sort behind everything through a sentinel, so the only host sync is the splitters' D2H
    s = W * oversample
    n_samples = min(s, n)
    packed = torch.zeros(s + 1, dtype=kview.dtype, device=dev)
    packed[0] = n_samples
    if n_samples > 0:
        key = (n, n_samples, str(dev))
        pos = _SAMPLE_POS.get(key)
        if pos is None:
            pos = ((torch.arange(n_samples, device=dev, dtype=torch.int64) * 2 + 1) * n) // (2 * n_samples)
            if len(_SAMPLE_POS) > 64:
                _SAMPLE_POS.clear()
            _SAMPLE_POS[key] = pos
        packed[1:1 + n_samples] = kview[pos]
    gathered2 = torch.empty(W * (s + 1), dtype=kview.dtype, device=dev)
    dist.all_gather_into_tensor(gathered2, packed, group=rng.group)
    gathered2 = gathered2.view(W, s + 1)
    all_counts = gathered2[:, 0].to(torch.int64)
    raw = gathered2[:, 1:].reshape(-1)
    valid = (torch.arange(s, device=dev).unsqueeze(0) < all_counts.unsqueeze(1)).reshape(-1)
    wide = _wide(raw, kdt)
    if wide.dtype.is_floating_point:
        sentinel = float("-inf") if descending else float("inf")
    else:
        sentinel = torch.iinfo(torch.int64).min if descending else torch.iinfo(torch.int64).max
    wide = torch.where(valid, wide, torch.full_like(wide, sentinel))
    order = torch.sort(wide, descending=descending, stable=True).indices  # <= W*s elements: plumbing
    m = all_counts.sum()
    pick = (torch.arange(1, W, device=dev, dtype=torch.int64) * m) // W
    splitters = raw[order[pick]].cpu().numpy()  # tiny D2H: the splitters are kernel arguments (all shards empty: zeros, unused)
    if kdt in _SIGNED:
        splitters = splitters.view({torch.uint32: "uint32", torch.uint64: "uint64"}[kdt])
    # a key goes to the rank given by the number of splitters that sort at or before it
    kpart, vpart, bucket_counts = ops.partition_pairs(keys, vals, splitters, descending, keys_alt, vals_alt)
    send_counts = bucket_counts[:W].to(torch.int64).contiguous()
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=rng.group)
    send_l, recv_l = send_counts.tolist(), recv_counts.tolist()  # split sizes are host-side arguments of the collective
    n_out = sum(recv_l)
    kpv = _signed_view(kpart)
    kout = torch.empty(n_out, dtype=kpv.dtype, device=dev)
    dist.all_to_all_single(kout, kpv, output_split_sizes=recv_l, input_split_sizes=send_l, group=rng.group)
    keys_out = kout.view(kdt)
    vals_out = None
    if vals is not None:
        vpv = _signed_view(vpart)
        vout = torch.empty(n_out, dtype=vpv.dtype, device=dev)
        dist.all_to_all_single(vout, vpv, output_split_sizes=recv_l, input_split_sizes=send_l, group=rng.group)
        vals_out = vout.view(vals.dtype)
    # the runs arrive concatenated in source-rank order, each in its source order: one stable sort finishes the job
    ops.sort_pairs(keys_out, vals_out, descending, None, None)
    return keys_out, vals_out


def sort(rng, keys, descending=False, ops=None, res=None):
    return sort_pairs(rng, keys, None, descending, ops, res)[0]


# ----------------------------------------------------------------------------------------- atomics
def atomic_histogram(rng, hist, keys, ops=None):
    """private bins per GPU, summed across GPUs"""
    _ops(ops).histogram(hist, keys)
    if rng.world > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=rng.group)
    return hist
