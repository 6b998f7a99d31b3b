#!/usr/bin/env python
"""How fast do pinned host <-> device copies go on this box, and does CPU / NUMA affinity of the pinned allocation matter?
   python tools/micro/h2d_probe.py   (one GPU)"""
import os, subprocess, time
import torch

def bw(dst, src, stream, reps=3):
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream); dst.copy_(src, non_blocking=True); e1.record(stream)
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return src.numel() * src.element_size() / best / 1e6

print(subprocess.run("lscpu | grep -i -E 'numa|^CPU\\(s\\)|model name'; nvidia-smi topo -m | head -8; cat /proc/self/status | grep -i cpus_allowed_list", shell=True, capture_output=True, text=True).stdout)
n = 1 << 28  # 1 GiB of int32
dev = torch.device("cuda:0")
d = torch.empty(n, dtype=torch.int32, device=dev); d2 = torch.empty_like(d)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def probe(tag):
    h = torch.empty(n, dtype=torch.int32).pin_memory(); h.fill_(1)
    h2 = torch.empty(n, dtype=torch.int32).pin_memory(); h2.fill_(2)
    print(f"{tag}: H2D {bw(d, h, s1):.1f} GB/s   D2H {bw(h2, d2, s2):.1f} GB/s", end="   ")
    # both directions at once
    torch.cuda.synchronize()
    e0, e1, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s1.wait_event(e0); s2.wait_event(e0)
    with torch.cuda.stream(s1):
        for _ in range(3): d.copy_(h, non_blocking=True)
        e1.record(s1)
    with torch.cuda.stream(s2):
        for _ in range(3): h2.copy_(d2, non_blocking=True)
        f1.record(s2)
    torch.cuda.synchronize()
    print(f"concurrent: H2D {3*n*4/e0.elapsed_time(e1)/1e6:.1f} GB/s  D2H {3*n*4/e0.elapsed_time(f1)/1e6:.1f} GB/s")
probe("default affinity ")
try:
    import pynvml
    pynvml.nvmlInit()
    hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
    print("ideal cpu affinity words:", list(pynvml.nvmlDeviceGetCpuAffinity(hnd, 4)))
    pynvml.nvmlDeviceSetCpuAffinity(hnd)
    print("affinity now:", sorted(os.sched_getaffinity(0)))
    probe("nvml cpu affinity")
except Exception as e:
    print("nvml affinity unavailable:", e)
