#!/usr/bin/env python
"""bench.py -- headline benchmark of the b200 back-end (BASELINE.json metric:
"GB/s & % HBM peak: forall+ReduceSum, inclusive_scan, sort_pairs @2^28, 1-8 GPU").

A "step" is one pass of the hot path over one batch of synthetic input, N = 2^28 per GPU:
    1. forall + ReduceSum   dot product, fp64            16 B/element  (4.29 GB)
    2. inclusive_scan       int32, operators::plus        8 B/element  (2.15 GB)
    3. sort_pairs           uint32 keys + uint32 values  68 B/element  (18.25 GB, onesweep model:
                            K + P*2*(K+V), P = 4 passes; the 2*(K+V) = 16 B lower bound is reported too)
`value` = algorithmic bytes of the step / device time of the step, summed over ranks, in GB/s.
Inputs are resident in HBM before the timed region (every step sorts its own fresh unsorted
buffers, so nothing is restored inside the region) and each is >= 1 GiB, far larger than the
126 MB L2.  `e2e` is the same step through the public API from pinned HOST buffers (H2D of all
inputs, D2H of all results inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--log2n 28] [--impl b200|reference]
For N > 1 launch with torch.distributed.run (one rank per GPU); rank 0 prints one JSON line.
--impl reference times the reference's own CPU implementation (oracle/_ref: RAJA
omp_parallel_for_exec compiled from the reference headers; falls back to the oracle port) on the
host cores on a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "hot-path throughput: forall+ReduceSum(dot f64) + inclusive_scan(i32) + sort_pairs(u32,u32) @2^28 per GPU"
BYTES_DOT, BYTES_SCAN, BYTES_SORT_MODEL, BYTES_SORT_LOWER = 16, 8, 68, 16
BYTES_ONESWEEP_PASS = 16  # one onesweep launch: read K+V, write K+V
NCU_ONESWEEP_TRAFFIC = 4.289e9  # bytes per launch measured by ncu at N=2^28 (2.169 GB read + 2.120 GB written, profiles/r01_onesweep_v4_ncu.txt)


def workload_config(log2n):
    """the same dict on both arms (b200 and --impl reference), so the two lines describe one configuration"""
    return {"workload": f"dot(f64)+inclusive_scan(i32)+sort_pairs(u32,u32), N=2^{log2n} elements per path per GPU (rank)",
            "l2": "inputs (1-2 GiB each) larger than the 126 MB L2; every step sorts fresh unsorted buffers"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled DURING the timed region.  The region is tens of
    milliseconds, so the sampler reads NVML in-process every ~2 ms (nvidia-smi -lms cannot go
    below ~100 ms per line); if NVML is not importable it falls back to one nvidia-smi query."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.mask, self.mx, self.stop_flag, self.how = index, [], 0, None, False, "nvml"
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None
            self.how = "nvidia-smi"

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                return int(ids[index])
        return index

    def _smi_once(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0]
            f = [x.strip() for x in out.split(",")]
            self.sm.append(float(f[0]))
            self.mx = float(f[1])
            for (name, bit), v in zip(self.REASONS, f[2:6]):
                if v.lower().startswith("active"):
                    self.mask |= bit
        except Exception:
            pass

    def run(self):
        if self.nv is None:
            self._smi_once()
            return
        while not self.stop_flag:
            try:
                self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                try:
                    self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                except Exception:
                    pass
            time.sleep(0.002)

    def finish(self):
        self.stop_flag = True
        self.join(timeout=15)
        reasons = sorted(name for name, bit in self.REASONS if self.mask & bit)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.mx, "reasons": reasons,
                "samples": len(self.sm), "how": self.how}


# --------------------------------------------------------------------------------------- reference arm
def cpu_workload(n, seed=12345):
    import numpy as np
    rng = np.random.default_rng(seed)
    x = rng.uniform(0.5, 1.5, n)
    y = rng.uniform(0.5, 1.5, n)
    s = rng.integers(0, 8, n).astype(np.int32)
    k = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    v = np.arange(n, dtype=np.uint32)
    return x, y, s, k, v


def cpu_step(ref, data):
    """one step of the hot path on the host; returns seconds per path"""
    import numpy as np
    x, y, s, k, v = data
    t0 = time.perf_counter()
    ref.dot(x, y)
    t_dot = time.perf_counter() - t0
    scan_fn = ref._fn("inclusive_scan_i32", None, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64])
    out = np.empty_like(s)
    t1 = time.perf_counter()
    scan_fn(0, s.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), s.size)
    t2 = time.perf_counter()
    kk, vv = k.copy(), v.copy()
    if ref.kind == "reference":
        f = ref._fn("sort_pairs32_u32", None, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int])
        t3 = time.perf_counter()
        f(0, kk.ctypes.data_as(ctypes.c_void_p), vv.ctypes.data_as(ctypes.c_void_p), kk.size, 0)  # RAJA::sort_pairs (unstable)
        t4 = time.perf_counter()
    else:
        t3 = time.perf_counter()
        ref.stable_sort_pairs(k, v)
        t4 = time.perf_counter()
    return {"dot": t_dot, "scan": t2 - t1, "sort": t4 - t3}


def host_threads():
    """cores this process may run on (torchrun exports OMP_NUM_THREADS=1, so the OpenMP default is not it)"""
    try:
        return max(len(os.sched_getaffinity(0)), 1)
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference(log2n, steps, warmup):
    from oracle_lib import Oracle, Reference
    if Reference.available():
        ref = Reference(1, threads=host_threads())  # omp_parallel_for_exec with every host thread, set explicitly
        what = "RAJA omp_parallel_for_exec/omp_reduce from oracle/_ref (reference headers, g++ -O2 -fopenmp)"
    else:
        ref = Oracle()
        what = "oracle/liboracle.so (plain-C port, single thread)"
    n = 1 << log2n
    data = cpu_workload(n)
    for _ in range(warmup):
        cpu_step(ref, data)
    acc = {"dot": 0.0, "scan": 0.0, "sort": 0.0}
    for _ in range(steps):
        for kk, vv in cpu_step(ref, data).items():
            acc[kk] += vv
    per = {k: v / steps for k, v in acc.items()}
    total_bytes = n * (BYTES_DOT + BYTES_SCAN + BYTES_SORT_MODEL)
    total_s = sum(per.values())
    paths = {"dot_f64": {"gbs": n * BYTES_DOT / per["dot"] / 1e9, "ms": per["dot"] * 1e3},
             "inclusive_scan_i32": {"gbs": n * BYTES_SCAN / per["scan"] / 1e9, "ms": per["scan"] * 1e3},
             "sort_pairs_u32": {"gbs_model": n * BYTES_SORT_MODEL / per["sort"] / 1e9, "ms": per["sort"] * 1e3,
                                "mpairs_s": n / per["sort"] / 1e6}}
    return {"value": total_bytes / total_s / 1e9, "ms_per_step": total_s * 1e3, "cores": ref.cores, "kind": ref.kind,
            "sample": f"same 3-path step at N=2^{log2n} per path ({what}); GB/s uses the same algorithmic bytes/element",
            "paths": paths}




# --------------------------------------------------------------------------------------- how one rank runs the three paths
POLICY_SO = os.path.join(ROOT, "bench", "_build", "libpolicy_bench.so")


class PolicyLib:
    """bench/_build/libpolicy_bench.so: RAJA::forall / inclusive_scan / sort_pairs instantiated with user-style lambdas for
    the b200 policies (the public API a RAJA user calls) and for the reference's own CUDA back-end (cuda_exec + CUB)."""

    def __init__(self):
        self.lib = ctypes.CDLL(POLICY_SO)
        self.lib.pb_last_error.restype = ctypes.c_char_p
        P, L, D = ctypes.c_void_p, ctypes.c_long, ctypes.c_double
        for name in ("pb_dot_expt", "pb_dot_classic"):
            for arm in ("b200", "cuda"):
                getattr(self.lib, f"{name}_{arm}").argtypes = [P, P, L, ctypes.POINTER(D), P]
        for arm in ("b200", "cuda"):
            getattr(self.lib, f"pb_inclusive_scan_{arm}").argtypes = [P, P, L, P]
            getattr(self.lib, f"pb_sort_pairs_{arm}").argtypes = [P, P, L, P]
            getattr(self.lib, f"pb_triad_{arm}").argtypes = [P, P, P, D, L, P]
            getattr(self.lib, f"pb_stable_sort_f64_{arm}").argtypes = [P, L, P]
        for name in ("pb_histogram_b200", "pb_histogram_aggregated_b200", "pb_histogram_cuda"):
            getattr(self.lib, name).argtypes = [P, P, L, P]

    @staticmethod
    def available():
        return os.path.exists(POLICY_SO)

    def call(self, name, *args):
        if getattr(self.lib, name)(*args) != 0:
            raise RuntimeError(f"{name}: {self.lib.pb_last_error().decode()}")


class Paths:
    """dot / scan / sort_pairs of one rank, by one of three routes:
       policy   RAJA front-ends + b200 policies (user lambdas, pools, resource: the public C++ API)
       cabi     the C ABI of libraja_b200.so (named loop bodies of the same engine)
       cuda     RAJA front-ends + the reference's cuda_exec / cuda_reduce / CUB ("the kernels to beat")"""

    def __init__(self, kind, rb, res, pb=None):
        self.kind, self.rb, self.res, self.pb = kind, rb, res, pb
        self.out = ctypes.c_double(0.0)
        self.pin = rb.PinnedBuffer(16) if kind == "cabi" else None
        self.apol = rb.b200_exec(async_=True)

    def dot(self, x, y):
        n = x.numel()
        if self.kind == "cabi":
            from raja_b200._lib import lib
            self.rb._lib.check(lib.rb200_reduce_dot(self.rb._lib.F64, x.data_ptr(), y.data_ptr(), n, self.pin.ptr, self.res.handle), "dot")
        else:
            arm = "b200" if self.kind == "policy" else "cuda"
            self.pb.call(f"pb_dot_expt_{arm}", x.data_ptr(), y.data_ptr(), n, ctypes.byref(self.out), self.res.handle)

    def dot_value(self):
        """host value of the last dot (cabi: after a stream sync)"""
        if self.kind == "cabi":
            self.res.wait()
            return float(self.pin.view("float64")[0])
        return float(self.out.value)

    def scan(self, src, dst):
        if self.kind == "cabi":
            self.rb.inclusive_scan(src, dst, pol=self.apol, res=self.res)
        else:
            arm = "b200" if self.kind == "policy" else "cuda"
            self.pb.call(f"pb_inclusive_scan_{arm}", src.data_ptr(), dst.data_ptr(), src.numel(), self.res.handle)

    def sort_pairs(self, k, v, k_alt=None, v_alt=None):
        if self.kind == "cabi":
            self.rb.sort_pairs(k, v, pol=self.apol, res=self.res, keys_alt=k_alt, vals_alt=v_alt)
        else:
            arm = "b200" if self.kind == "policy" else "cuda"
            self.pb.call(f"pb_sort_pairs_{arm}", k.data_ptr(), v.data_ptr(), k.numel(), self.res.handle)


# --------------------------------------------------------------------------------------- verification (outside the timed region)
def verify_step(rb, rbm, sharded, paths, n, dev, res, world, rank, seed=987):
    """One more step on fresh inputs whose results are CHECKED on the device (torch ops here are the checker, not
    the product): every rank's piece of the sort is sorted and stable, pieces are ordered across ranks, nothing
    was lost or duplicated (count, pair checksum), the scan equals an int64 cumsum seeded with the lower ranks'
    totals, the dot agrees with an fp64 all-reduce.  Returns a dict of booleans; bench.py exits non-zero on any False."""
    import torch
    import torch.distributed as dist
    from raja_b200._lib import lib
    g = torch.Generator(device=dev).manual_seed(seed + rank)
    apol = rb.b200_exec(async_=True)
    out = {}

    def allsum(t):
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    # ---- dot
    x = torch.rand(n, dtype=torch.float64, device=dev, generator=g) + 0.5
    y = torch.rand(n, dtype=torch.float64, device=dev, generator=g) + 0.5
    if sharded:
        got = rbm.reduce_dot(sharded, x, y, res=res, as_tensor=True).reshape(1).clone()
    else:
        paths.dot(x, y)
        got = torch.tensor([paths.dot_value()], dtype=torch.float64, device=dev)
    want = allsum(torch.dot(x, y).reshape(1))  # all positive: sum |x y| == the dot itself
    tol = world * n * torch.finfo(torch.float64).eps * want
    out["dot"] = bool((got - want).abs() <= tol)
    del x, y

    # ---- scan: int64 cumsum of the shard + the lower ranks' totals, wrapped to int32 like the device arithmetic
    s_in = torch.randint(0, 8, (n,), dtype=torch.int32, device=dev, generator=g)
    s_out = torch.empty_like(s_in)
    if sharded:
        rbm.inclusive_scan(sharded, s_in, s_out, res=res)
    else:
        paths.scan(s_in, s_out)
    tot = s_in.sum(dtype=torch.int64).reshape(1)
    totals = torch.zeros(world, dtype=torch.int64, device=dev)
    totals[rank] = tot[0]
    allsum(totals)
    carry = totals[:rank].sum()
    want = (torch.cumsum(s_in, 0, dtype=torch.int64) + carry).to(torch.int32)
    out["scan"] = bool(torch.equal(want, s_out))
    del s_in, s_out, want

    # ---- sort_pairs: values carry the GLOBAL original index, so stability is checkable across ranks
    k = torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
    k[::5] &= 0xFFFF  # duplicates, so that stability is exercised
    v = (torch.arange(n, dtype=torch.int64, device=dev) + rank * n).to(torch.int32)
    k64 = k.to(torch.int64) & 0xFFFFFFFF
    v64 = v.to(torch.int64) & 0xFFFFFFFF
    chk_in = allsum(torch.stack([(k64 * (v64 + 1)).sum(), k64.sum(), v64.sum()]))  # int64 arithmetic wraps: a checksum
    del k64, v64
    ka, va = torch.empty_like(k), torch.empty_like(v)
    ku, vu = k.view(torch.uint32), v.view(torch.uint32)
    if sharded:
        ko, vo = rbm.sort_pairs(sharded, ku, vu, res=res, keys_alt=ka.view(torch.uint32), vals_alt=va.view(torch.uint32))
    else:
        paths.sort_pairs(ku, vu, ka.view(torch.uint32), va.view(torch.uint32))
        ko, vo = ku, vu
    torch.cuda.synchronize()
    m = ko.numel()
    ko64 = ko.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    vo64 = vo.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    chk_out = allsum(torch.stack([(ko64 * (vo64 + 1)).sum(), ko64.sum(), vo64.sum()]))
    count = allsum(torch.tensor([m], dtype=torch.int64, device=dev))
    ok_sorted = bool((ko64[1:] >= ko64[:-1]).all()) if m > 1 else True
    ok_stable = bool(((ko64[1:] != ko64[:-1]) | (vo64[1:] > vo64[:-1])).all()) if m > 1 else True
    # piece boundaries: (last key, last value) of rank r sorts before (first key, first value) of the next non-empty rank
    edge = torch.zeros(world, 5, dtype=torch.int64, device=dev)
    if m > 0:
        edge[rank] = torch.stack([ko64[0], vo64[0], ko64[-1], vo64[-1], torch.tensor(m, device=dev)])
    allsum(edge)
    e = [r for r in edge.tolist() if r[4] > 0]
    ok_edges = all((a[2], a[3]) < (b[0], b[1]) for a, b in zip(e, e[1:]))
    flags = allsum(torch.tensor([int(not ok_sorted), int(not ok_stable)], dtype=torch.int64, device=dev))
    out["sort_sorted"] = int(flags[0]) == 0
    out["sort_stable"] = int(flags[1]) == 0
    out["sort_boundaries"] = bool(ok_edges)
    out["sort_count"] = int(count[0]) == world * n
    out["sort_checksum"] = bool(torch.equal(chk_in, chk_out))
    out["sort_piece_sizes"] = [int(r[4]) for r in edge.tolist()]
    return out

# --------------------------------------------------------------------------------------- b200 arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 back-end has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import raja_b200 as rb
    from raja_b200._lib import lib
    rbm = None
    if world > 1:
        from raja_b200 import multi as rbm

    n = 1 << args.log2n
    K, W = args.steps, max(args.warmup, 3)
    dev = torch.device("cuda", local)
    g = torch.Generator(device=dev).manual_seed(12345 + rank)

    # ---- inputs resident in HBM before the timed region
    x = torch.rand(n, dtype=torch.float64, device=dev, generator=g) + 0.5
    y = torch.rand(n, dtype=torch.float64, device=dev, generator=g) + 0.5
    s_in = torch.randint(0, 8, (n,), dtype=torch.int32, device=dev, generator=g)
    s_out = torch.empty_like(s_in)
    nsets = K + W
    sort_sets = []
    for i in range(nsets):
        k = torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device=dev, generator=g).to(torch.int32).view(torch.uint32)
        v = torch.arange(n, dtype=torch.int32, device=dev).view(torch.uint32)
        sort_sets.append((k, v))
    k_alt = torch.empty(n, dtype=torch.int32, device=dev).view(torch.uint32)
    v_alt = torch.empty(n, dtype=torch.int32, device=dev).view(torch.uint32)
    sharded = None
    if world > 1:
        # exchanges over NVLink peer memory (rb200_comm_*); receive buffers for 1.25x a shard of keys + values
        comm = None if args.nccl_exchange else rbm.PeerComm(exchange_bytes=(n + n // 4 + 4096) * 8 + 128)
        sharded = rbm.ShardedRange(world, rank, comm=comm)

    stream = torch.cuda.current_stream()
    res = rb.Resource(stream)
    apol = rb.b200_exec(async_=True)
    # one GPU: the three paths go through the RAJA front-ends with the b200 policies (user lambdas in a C++ translation
    # unit: bench/policy_bench.cu) -- the public API; --route cabi keeps the C-ABI bodies.  Several GPUs: the sharded layer.
    pb = PolicyLib() if PolicyLib.available() else None
    route = args.route or ("policy" if pb is not None else "cabi")
    if route == "policy" and pb is None:
        raise SystemExit("bench.py: --route policy needs bench/_build/libpolicy_bench.so (built by __graft_entry__.build())")
    paths = Paths(route, rb, res, pb)

    def run_step(pth, k, v, ev=None):
        if ev: ev[0].record(stream)
        if sharded:
            rbm.reduce_dot(sharded, x, y, res=res, as_tensor=True)
        else:
            pth.dot(x, y)
        if ev: ev[1].record(stream)
        if sharded:
            rbm.inclusive_scan(sharded, s_in, s_out, res=res)
        else:
            pth.scan(s_in, s_out)
        if ev: ev[2].record(stream)
        if sharded:
            rbm.sort_pairs(sharded, k, v, res=res, keys_alt=k_alt, vals_alt=v_alt)
        else:
            pth.sort_pairs(k, v, k_alt, v_alt)
        if ev: ev[3].record(stream)

    def step(i, ev=None):
        run_step(paths, *sort_sets[i], ev=ev)

    def timed_arm(pth, steps, warm):
        """a few steps of another route over re-randomised sort buffers (single GPU, after the headline region)"""
        for i in range(steps + warm):
            k, v = sort_sets[i]
            k.view(torch.int32).copy_(torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device=dev, generator=g).to(torch.int32))
            v.view(torch.int32).copy_(torch.arange(n, dtype=torch.int32, device=dev))
        for i in range(warm):
            run_step(pth, *sort_sets[i])
        torch.cuda.synchronize()
        es = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
        for i in range(steps):
            run_step(pth, *sort_sets[warm + i], ev=es[i])
        torch.cuda.synchronize()
        t = [sum(e[j].elapsed_time(e[j + 1]) for e in es) / steps for j in range(3)]
        return {"dot_ms": t[0], "scan_ms": t[1], "sort_pairs_ms": t[2], "step_ms": sum(t), "steps": steps,
                "value_gbs": n * (BYTES_DOT + BYTES_SCAN + BYTES_SORT_MODEL) / sum(t) / 1e6}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # untimed pre-warm (clocks and memory to steady state) in front of the W warm-up steps: 0.3 s of dot + scan,
    # which leave the per-step sort buffers untouched
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < 0.3:
        paths.dot(x, y)
        paths.scan(s_in, s_out)
        torch.cuda.synchronize()
    for i in range(W):
        step(i)
    barrier()

    lib.rb200_profile_reset()
    lib.rb200_profile_enable(1)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler: sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(K):
        step(W + i, evs[i])
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.finish() if sampler else None
    launches = lib.rb200_launch_count()
    lib.rb200_profile_enable(0)

    ms = {"dot": 0.0, "scan": 0.0, "sort": 0.0}
    for e in evs:
        ms["dot"] += e[0].elapsed_time(e[1])
        ms["scan"] += e[1].elapsed_time(e[2])
        ms["sort"] += e[2].elapsed_time(e[3])
    ms = {k: v / K for k, v in ms.items()}
    step_ms = sum(ms.values())
    t = torch.tensor([step_ms, ms["dot"], ms["scan"], ms["sort"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, dot_ms, scan_ms, sort_ms = t.tolist()

    # dominant kernel (onesweep pass) timed live with CUDA events on the launching stream
    tot, cnt = ctypes.c_double(), ctypes.c_longlong()
    lib.rb200_profile_query(b"onesweep_kernel", ctypes.byref(tot), ctypes.byref(cnt))
    pass_ms = tot.value / max(cnt.value, 1)
    kern = {}
    for name in (b"forall_reduce_kernel", b"forall_kernel", b"scan_kernel", b"histogram_kernel", b"scan_bins_kernel", b"onesweep_kernel", b"bucket_count_kernel",
                 b"peer_allgather_kernel", b"peer_sample_allgather_kernel", b"pick_splitters_kernel"):
        lib.rb200_profile_query(name, ctypes.byref(tot), ctypes.byref(cnt))
        kern[name.decode()] = {"launches": cnt.value, "avg_ms": tot.value / max(cnt.value, 1), "share_of_step": tot.value / K / step_ms if step_ms else None}

    # ---- one GPU: the same step by the other routes, same process, same buffers: the reference's own CUDA back-end
    #      (cuda_exec + cuda_reduce + CUB: "the kernels to beat") and the C ABI
    reference_gpu = cabi = None
    if world == 1 and pb is not None and not args.no_reference_gpu:
        reference_gpu = timed_arm(Paths("cuda", rb, res, pb), min(K, 5), 2)
        reference_gpu["what"] = ("RAJA::forall<cuda_exec_async<256>> + expt::Reduce<plus>, RAJA::inclusive_scan<cuda_exec> (cub::DeviceScan), "
                                 "RAJA::sort_pairs<cuda_exec> (cub::DeviceRadixSort), unmodified reference headers, same lambdas and buffers")
        other = "cabi" if route == "policy" else "policy"
        cabi = timed_arm(Paths(other, rb, res, pb), min(K, 5), 2)
        cabi["route"] = other

    # ---- e2e: same step from pinned host buffers through the public API
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(rb, rbm, sharded, paths, n, dev, stream, res, world, barrier, min(K, args.e2e_steps))

    # ---- results checked on the device, every rank (outside the timed region)
    verified = None
    if not args.no_verify:
        del sort_sets[:]
        torch.cuda.empty_cache()
        verified = verify_step(rb, rbm, sharded, paths, n, dev, res, world, rank)
        verified["all"] = all(v for k, v in verified.items() if isinstance(v, bool))

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if verified is not None and not verified["all"]:
        print(f"bench.py: rank {rank}: VERIFICATION FAILED: {verified}", file=sys.stderr, flush=True)
        if rank != 0:
            raise SystemExit(3)
    if rank != 0:
        return
    peak, peak_src = peaks()
    bytes_step = n * (BYTES_DOT + BYTES_SCAN + BYTES_SORT_MODEL)
    value = world * bytes_step / (step_ms * 1e-3) / 1e9
    paths = {
        "forall_reducesum_dot_f64": {"ms": dot_ms, "gbs": n * BYTES_DOT / dot_ms / 1e6, "frac_hbm": n * BYTES_DOT / dot_ms / 1e6 / peak},
        "inclusive_scan_i32": {"ms": scan_ms, "gbs": n * BYTES_SCAN / scan_ms / 1e6, "frac_hbm": n * BYTES_SCAN / scan_ms / 1e6 / peak},
        "sort_pairs_u32_u32": {"ms": sort_ms, "gbs_model68": n * BYTES_SORT_MODEL / sort_ms / 1e6, "frac_hbm_model68": n * BYTES_SORT_MODEL / sort_ms / 1e6 / peak,
                               "gbs_lower16": n * BYTES_SORT_LOWER / sort_ms / 1e6, "gpairs_s": n / sort_ms / 1e6},
    }
    achieved = n * BYTES_ONESWEEP_PASS / pass_ms / 1e6 if pass_ms else None
    cpu = None
    if world == 1 and not args.no_cpu:
        c = cpu_reference(args.cpu_log2n, 1, 1)
        cpu = {"value": c["value"], "unit": "GB/s", "cores": c["cores"], "kind": c["kind"], "sample": c["sample"], "paths": c["paths"]}
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+i32+u32", "data": "synthetic",
        "config": workload_config(args.log2n),
        "run": {"route": ("sharded layer (raja_b200/multi.py over the C ABI)" if world > 1 else
                          {"policy": "RAJA::forall<b200_exec_async<256>> + expt::Reduce<plus> lambda, RAJA::inclusive_scan<b200_exec>, RAJA::sort_pairs<b200_exec> "
                                     "(bench/policy_bench.cu: the public C++ API, pools and resource included)",
                           "cabi": "C ABI of libraja_b200.so (rb200_reduce_dot / rb200_scan / rb200_sort_pairs_buffered)"}[route]),
                "prewarm": "0.3 s of untimed dot+scan in front of the W warm-up steps",
                "sharding": "none" if world == 1 else ("contiguous shards; allreduce(dot), carry allgather(scan), sample-sort all-to-all(sort); exchanges by " +
                                                        ("NCCL collectives" if args.nccl_exchange else "the NVLink peer-memory communicator (sort: fused partition + all-to-all sweep)"))},
        "verified": verified["all"] if verified is not None else None, "verification": verified,
        "paths": paths,
        "roofline": {"bound": "hbm", "kernel": "onesweep_kernel (one radix pass, u32 keys + u32 values)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak if achieved else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch at N=2^28 (ncu --set full, profiles/r01_onesweep_v4_ncu.txt)
                     "traffic": NCU_ONESWEEP_TRAFFIC if args.log2n == 28 else None, "traffic_source": "from_profile (ncu --set full of this kernel at this N, not re-measured in this run)",
                     "peak_source": peak_src,
                     "bytes_per_launch": n * BYTES_ONESWEEP_PASS, "avg_launch_ms": pass_ms},
        "kernels": kern,
        "reference_gpu": reference_gpu,
        "other_route": cabi,
        "cpu_baseline": cpu,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "wall_ms_per_step": t_wall * 1e3 / K,
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if verified is not None and not verified["all"]:
        raise SystemExit(3)


def run_e2e(rb, rbm, sharded, paths, n, dev, stream, res, world, barrier, steps):
    """The same step through the public API from pinned HOST buffers: every step copies its inputs
    host -> device and its results device -> host inside the timed region.  Three streams (H2D, compute,
    D2H) with events between them, so the two copy engines and the kernels overlap within a step and
    across steps; a device buffer is overwritten only after its previous consumer has finished."""
    import torch
    import torch.distributed as dist
    from raja_b200._lib import lib
    g = torch.Generator().manual_seed(7)
    cap = n + n // 4 + 4096 if sharded else n  # sharded sorts return a piece of data-dependent size
    hx = torch.empty(n, dtype=torch.float64).pin_memory()
    hy = torch.empty(n, dtype=torch.float64).pin_memory()
    hs = torch.empty(n, dtype=torch.int32).pin_memory()
    hk = torch.empty(n, dtype=torch.int32).pin_memory()
    hv = torch.empty(n, dtype=torch.int32).pin_memory()
    hx.uniform_(0.5, 1.5, generator=g); hy.uniform_(0.5, 1.5, generator=g)
    hs.random_(0, 8, generator=g); hk.random_(-2**31, 2**31 - 1, generator=g)
    hv.copy_(torch.arange(n, dtype=torch.int32))
    ho_s = torch.empty(n, dtype=torch.int32).pin_memory()
    ho_k = torch.empty(cap, dtype=torch.int32).pin_memory()
    ho_v = torch.empty(cap, dtype=torch.int32).pin_memory()
    ho_dot = torch.empty(1, dtype=torch.float64).pin_memory()
    dot_pin = rb.PinnedBuffer(16)
    dx = torch.empty(n, dtype=torch.float64, device=dev); dy = torch.empty_like(dx)
    ds = torch.empty(n, dtype=torch.int32, device=dev); dso = torch.empty_like(ds)
    dk = torch.empty(n, dtype=torch.int32, device=dev); dv = torch.empty_like(dk)
    ka = torch.empty_like(dk); va = torch.empty_like(dk)
    apol = rb.b200_exec(async_=True)
    s_c = stream
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    last = {}  # events of the previous step: c_dot, c_scan (compute done), o_scan, o_sort (D2H done)

    def ev(st):
        e = torch.cuda.Event()
        e.record(st)
        return e

    def one():
        with torch.cuda.stream(s_in):
            if "c_dot" in last: s_in.wait_event(last["c_dot"])
            dx.copy_(hx, non_blocking=True); dy.copy_(hy, non_blocking=True)
            e_xy = ev(s_in)
            if "c_scan" in last: s_in.wait_event(last["c_scan"])
            ds.copy_(hs, non_blocking=True)
            e_s = ev(s_in)
            if "o_sort" in last: s_in.wait_event(last["o_sort"])
            dk.copy_(hk, non_blocking=True); dv.copy_(hv, non_blocking=True)
            e_kv = ev(s_in)
        s_c.wait_event(e_xy)
        if sharded:
            dot_t = rbm.reduce_dot(sharded, dx, dy, res=res, as_tensor=True)
        else:
            paths.dot(dx, dy)
        last["c_dot"] = ev(s_c)
        s_c.wait_event(e_s)
        if "o_scan" in last: s_c.wait_event(last["o_scan"])
        if sharded: rbm.inclusive_scan(sharded, ds, dso, res=res)
        else: paths.scan(ds, dso)
        last["c_scan"] = ev(s_c)
        s_c.wait_event(e_kv)
        if "o_sort" in last: s_c.wait_event(last["o_sort"])  # sharded: nobody writes into a receive buffer that is still being read
        ku, vu = dk.view(torch.uint32), dv.view(torch.uint32)
        if sharded:
            ko, vo = rbm.sort_pairs(sharded, ku, vu, res=res, keys_alt=ka.view(torch.uint32), vals_alt=va.view(torch.uint32))
        else:
            paths.sort_pairs(ku, vu, ka.view(torch.uint32), va.view(torch.uint32))
            ko, vo = ku, vu
        c_sort = ev(s_c)
        m = ko.numel()
        with torch.cuda.stream(s_out):
            s_out.wait_event(last["c_dot"])
            if sharded: ho_dot.copy_(dot_t.reshape(1), non_blocking=True)
            s_out.wait_event(last["c_scan"])
            ho_s.copy_(dso, non_blocking=True)
            last["o_scan"] = ev(s_out)
            s_out.wait_event(c_sort)
            ho_k[:m].copy_(ko.view(torch.int32), non_blocking=True); ho_v[:m].copy_(vo.view(torch.int32), non_blocking=True)
            last["o_sort"] = ev(s_out)
        return m

    one()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s_c)
    for _ in range(steps):
        m = one()
    s_c.wait_stream(s_in); s_c.wait_stream(s_out)
    e1.record(s_c)
    barrier()
    dot = float(ho_dot[0]) if sharded else paths.dot_value()
    if not dot > 0:
        raise SystemExit("bench.py: e2e dot product did not arrive on the host")
    ms = e0.elapsed_time(e1) / steps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    h2d = n * (8 + 8 + 4 + 4 + 4)
    d2h = n * (4 + 4 + 4) + 8
    return {"value": world * n * (BYTES_DOT + BYTES_SCAN + BYTES_SORT_MODEL) / (ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms,
            "steps": steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "note": "pinned host -> device copies of x,y,scan input,keys,vals and device -> host of scan output, sorted keys+vals, dot scalar, all inside "
                    "the timed region; H2D, kernels and D2H on three streams (events in between), so copies in both directions overlap the kernels"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    c = cpu_reference(args.cpu_log2n, max(args.steps, 1), max(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": c["value"], "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": c["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+i32+u32",
            "data": "synthetic",
            "config": workload_config(args.cpu_log2n),
            "run": {"sample": f"each step = one pass over N=2^{args.cpu_log2n} per path on the host cores ({c['cores']} OpenMP threads, set explicitly); rank 0 only"},
            "paths": c["paths"],
            "cpu_baseline": {"value": c["value"], "unit": "GB/s", "cores": c["cores"], "kind": c["kind"], "sample": c["sample"]},
            "e2e": {"value": c["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--cpu-log2n", type=int, default=None, help="N of the CPU arm (default: --log2n, the GPU arm's N)")
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the checked step after the timed region")
    ap.add_argument("--route", default=None, choices=["policy", "cabi"], help="one GPU: RAJA front-ends + b200 policies (default when built) or the C ABI")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip the cuda_exec / CUB arm and the other route")
    ap.add_argument("--nccl-exchange", action="store_true", help="N > 1: exchange with NCCL collectives instead of the peer-memory communicator")
    args = ap.parse_args()
    if args.cpu_log2n is None:
        args.cpu_log2n = args.log2n
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
