"""Where does the sharded sort_pairs spend its time?  Re-runs the steps of raja_b200.multi.sort_pairs (peer
communicator path) with a device synchronize after each one.  torchrun --nproc-per-node N tools/micro/multi_sort_breakdown.py"""
import os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import raja_b200 as rb
from raja_b200 import multi

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << int(os.environ.get("LOG2N", "28"))
dev = torch.device("cuda", local)
comm = multi.PeerComm(exchange_bytes=(n + n // 4 + 4096) * 8 + 128)
rng = multi.ShardedRange(world, rank, comm=comm)
res = rb.Resource(torch.cuda.current_stream())
ops = multi.B200Ops(res)
g = torch.Generator(device=dev).manual_seed(1 + rank)
ka = torch.empty(n, dtype=torch.int32, device=dev).view(torch.uint32); va = torch.empty_like(ka)

def stamp(t, name, acc):
    torch.cuda.synchronize()
    now = time.perf_counter()
    acc[name] = acc.get(name, 0.0) + (now - t) * 1e3
    return now

acc = {}
for it in range(4):
    keys = torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device=dev, generator=g).to(torch.int32).view(torch.uint32)
    vals = torch.arange(n, dtype=torch.int32, device=dev).view(torch.uint32)
    dist.barrier(); torch.cuda.synchronize()
    if it == 0:
        multi.sort_pairs(rng, keys, vals, ops=ops, keys_alt=ka, vals_alt=va); continue
    a = {} if it == 1 else acc
    t = time.perf_counter()
    W = world; kdt = keys.dtype; kview = multi._signed_view(keys); s = W * 512
    pos = ((torch.arange(s, device=dev, dtype=torch.int64) * 2 + 1) * n) // (2 * s)
    padded = kview[pos]
    t = stamp(t, "1 sample gather", a)
    all_counts = torch.empty(W, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_counts, torch.tensor([s], device=dev, dtype=torch.int64))
    gathered = torch.empty(W * s, dtype=kview.dtype, device=dev)
    dist.all_gather_into_tensor(gathered, padded)
    t = stamp(t, "2 nccl all_gathers", a)
    order = torch.sort(multi._wide(gathered, kdt), stable=True).indices
    pick = (torch.arange(1, W, device=dev, dtype=torch.int64) * gathered.numel()) // W
    splitters = gathered[order[pick]].cpu().numpy().view("uint32")
    t = stamp(t, "3 sort samples + D2H splitters", a)
    ko, vo = comm.exchange_pairs(keys, vals, splitters, False, res)
    t = stamp(t, "4 exchange_pairs (count, allgather, sync, sweep, barrier)", a)
    ops.sort_pairs(ko, vo, False, ka if ka.numel() >= ko.numel() else None, va if va.numel() >= ko.numel() else None)
    t = stamp(t, "5 local sort of %d" % ko.numel(), a)
    t0 = time.perf_counter()
    keys2 = keys.clone(); vals2 = vals.clone(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    multi.sort_pairs(rng, keys2, vals2, ops=ops, keys_alt=ka, vals_alt=va)
    stamp(t0, "whole multi.sort_pairs (no intermediate syncs)", a)
if rank == 0:
    for k, v in acc.items():
        print(f"{k:60s} {v / 2:8.3f} ms")
# per-kernel times of one more whole sort (event-bracketed launches: rb200_profile_*)
import ctypes
from raja_b200._lib import lib
keys2 = keys.clone(); vals2 = vals.clone(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
lib.rb200_profile_reset(); lib.rb200_profile_enable(1)
multi.sort_pairs(rng, keys2, vals2, ops=ops, keys_alt=ka, vals_alt=va)
torch.cuda.synchronize()
lib.rb200_profile_enable(0)
if rank == 0:
    tot, cnt = ctypes.c_double(), ctypes.c_longlong()
    for name in (b"bucket_count_kernel", b"peer_allgather_kernel", b"scan_bins_kernel", b"histogram_kernel", b"onesweep_kernel"):
        lib.rb200_profile_query(name, ctypes.byref(tot), ctypes.byref(cnt))
        print(f"kernel {name.decode():28s} launches {cnt.value:3d} total {tot.value:8.3f} ms")
comm.close()
dist.destroy_process_group()
