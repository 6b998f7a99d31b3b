#!/usr/bin/env python
"""Runs one path of the hot path a few times at N = 2^log2n (for ncu / quick timing).
   python tools/run_path.py --path scan --log2n 28 --reps 3"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import raja_b200 as rb

ap = argparse.ArgumentParser()
ap.add_argument("--path", required=True)
ap.add_argument("--log2n", type=int, default=28)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
n = 1 << a.log2n
g = torch.Generator(device="cuda").manual_seed(1)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.reps + 1)]
pol = rb.b200_exec(async_=True)


def timed(fn, bytes_per_elem):
    fn(0)
    torch.cuda.synchronize()
    for i in range(a.reps):
        ev[i].record()
        fn(i + 1)
    ev[a.reps].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.reps)]
    best = min(ts)
    print(f"{a.path}: best {best:.4f} ms  median {sorted(ts)[len(ts)//2]:.4f} ms  {n*bytes_per_elem/best/1e6:.1f} GB/s")


if a.path in ("dot", "sum", "multi", "minloc"):
    x = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) + 0.5
    if a.path == "dot":
        y = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) + 0.5
        timed(lambda i: rb.reduce_dot(x, y), 16)
    elif a.path == "sum":
        timed(lambda i: rb.reduce_sum(x), 8)
    elif a.path == "multi":
        timed(lambda i: rb.reduce_multi(x), 8)
    else:
        timed(lambda i: rb.reduce_minloc(x), 8)
elif a.path in ("triad", "daxpy"):
    b = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
    c = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
    aa = torch.zeros_like(b)
    if a.path == "triad":
        timed(lambda i: rb.forall_triad(aa, b, c, 3.0, pol=pol), 24)
    else:
        timed(lambda i: rb.forall_daxpy(aa, b, 3.0, pol=pol), 24)
elif a.path.startswith("scan"):
    dt = {"scan": torch.int32, "scan64": torch.int64, "scanf64": torch.float64}[a.path]
    if dt == torch.float64:
        x = torch.rand(n, dtype=dt, device="cuda", generator=g)
    else:
        x = torch.randint(0, 8, (n,), dtype=dt, device="cuda", generator=g)
    out = torch.empty_like(x)
    timed(lambda i: rb.inclusive_scan(x, out, pol=pol), 2 * x.element_size())
elif a.path in ("sort", "sort_lowent", "sortf64", "sortu64", "sortf64u"):
    sets = []
    for i in range(a.reps + 1):
        if a.path == "sortf64":
            k = (torch.rand(n, dtype=torch.float64, device="cuda", generator=g) - 0.5) * 2e6
            sets.append((k, None))
        elif a.path == "sortu64":  # every byte uniform
            k = torch.randint(-2**63, 2**63 - 1, (n,), dtype=torch.int64, device="cuda", generator=g)
            sets.append((k.view(torch.uint64), None))
        elif a.path == "sortf64u":  # fp64 keys whose bit patterns are uniform (finite, both signs, every exponent)
            k = torch.randint(-2**63, 2**63 - 1, (n,), dtype=torch.int64, device="cuda", generator=g)
            k &= ~(1 << 62)  # exponent < 1024: no inf / nan
            sets.append((k.view(torch.float64), None))
        else:
            k = torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
            if a.path == "sort_lowent":
                k &= 0xFF
            sets.append((k.view(torch.uint32), torch.arange(n, dtype=torch.int32, device="cuda").view(torch.uint32)))
    ka = torch.empty_like(sets[0][0])
    if sets[0][1] is None:
        timed(lambda i: rb.stable_sort(sets[i][0], pol=pol), 136)
    else:
        va = torch.empty_like(sets[0][1])
        timed(lambda i: rb.sort_pairs(sets[i][0], sets[i][1], pol=pol, keys_alt=ka, vals_alt=va), 68)
elif a.path in ("hist256", "hist64k", "hist256agg", "hist64kagg"):
    nb = 256 if "256" in a.path else 65536
    keys = torch.randint(0, nb, (n,), dtype=torch.int32, device="cuda", generator=g)
    hist = torch.zeros(nb, dtype=torch.int32, device="cuda")
    timed(lambda i: rb.atomic_histogram(hist, keys, aggregate=a.path.endswith("agg"), pol=pol), 4)
elif a.path in ("dotf32", "sumi32", "minloc64"):
    if a.path == "dotf32":
        x = torch.rand(n, dtype=torch.float32, device="cuda", generator=g) + 0.5
        y = torch.rand(n, dtype=torch.float32, device="cuda", generator=g) + 0.5
        timed(lambda i: rb.reduce_dot(x, y), 8)
    elif a.path == "sumi32":
        x = torch.randint(-100, 100, (n,), dtype=torch.int32, device="cuda", generator=g)
        timed(lambda i: rb.reduce_sum(x), 4)
    else:
        x = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
        timed(lambda i: rb.reduce_maxloc(x), 8)
elif a.path in ("triadf32", "daxpyf32"):
    b = torch.rand(n, dtype=torch.float32, device="cuda", generator=g)
    c = torch.rand(n, dtype=torch.float32, device="cuda", generator=g)
    aa = torch.zeros_like(b)
    if a.path == "triadf32":
        timed(lambda i: rb.forall_triad(aa, b, c, 3.0, pol=pol), 12)
    else:
        timed(lambda i: rb.forall_daxpy(aa, b, 3.0, pol=pol), 12)
elif a.path in ("exscan", "inplacescan64", "maxscanf32", "exscanf64"):
    if a.path == "exscan":
        x = torch.randint(0, 8, (n,), dtype=torch.int32, device="cuda", generator=g)
        out = torch.empty_like(x)
        timed(lambda i: rb.exclusive_scan(x, out, init=5, pol=pol), 8)
    elif a.path == "inplacescan64":
        x = torch.randint(-1000, 1000, (n,), dtype=torch.int64, device="cuda", generator=g)
        timed(lambda i: rb.inclusive_scan_inplace(x, pol=pol), 16)
    elif a.path == "maxscanf32":
        x = torch.rand(n, dtype=torch.float32, device="cuda", generator=g)
        out = torch.empty_like(x)
        timed(lambda i: rb.inclusive_scan(x, out, op=rb.operators.maximum, pol=pol), 8)
    else:
        x = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
        out = torch.empty_like(x)
        timed(lambda i: rb.exclusive_scan(x, out, pol=pol), 16)
elif a.path in ("sortkeys", "sortdesc", "sorti64pairs"):
    sets = []
    for i in range(a.reps + 1):
        if a.path == "sorti64pairs":
            k = torch.randint(-2**62, 2**62, (n,), dtype=torch.int64, device="cuda", generator=g)
            sets.append((k, torch.arange(n, dtype=torch.int32, device="cuda").view(torch.uint32)))
        else:
            k = torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32).view(torch.uint32)
            sets.append((k, torch.arange(n, dtype=torch.int32, device="cuda").view(torch.uint32)))
    if a.path == "sortkeys":
        timed(lambda i: rb.stable_sort(sets[i][0], pol=pol), 36)
    elif a.path == "sortdesc":
        timed(lambda i: rb.sort_pairs(sets[i][0], sets[i][1], rb.operators.greater, pol=pol), 68)
    else:
        timed(lambda i: rb.stable_sort_pairs(sets[i][0], sets[i][1], pol=pol), 200)
elif a.path in ("histzipf", "pi"):
    if a.path == "pi":
        # examples/pi-reduce_vs_atomic.cpp with num_bins = n: one atomicAdd on ONE address per iteration (no memory traffic: Gupdates/s)
        import time
        rb.atomic_pi(1 << 20)
        t0 = time.perf_counter(); v = rb.atomic_pi(n); dt = time.perf_counter() - t0
        print(f"pi: {dt*1e3:.3f} ms  {n/dt/1e9:.2f} Gupdates/s  value {v:.12f}")
    else:
        # Zipf-like skew over 65536 bins: bin = floor(65536 * u^4)
        u = torch.rand(n, dtype=torch.float32, device="cuda", generator=g)
        keys = (65536 * u ** 4).to(torch.int32).clamp_(0, 65535)
        hist = torch.zeros(65536, dtype=torch.int32, device="cuda")
        timed(lambda i: rb.atomic_histogram(hist, keys, pol=pol), 4)
else:
    raise SystemExit("unknown path")
