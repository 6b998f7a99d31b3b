"""Probe: does torch symmetric memory (CUDA IPC / fabric handles) work between the ranks of one box here?
   torchrun --nproc-per-node 2 tools/micro/symm_probe.py"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = symm.empty(1 << 20, dtype=torch.int32, device=f"cuda:{local}")
t.fill_(rank + 100)
hdl = symm.rendezvous(t, dist.group.WORLD)
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal", hex(hdl.signal_pad_ptrs[0]), flush=True)
hdl.barrier()
peer = (rank + 1) % world
remote = hdl.get_buffer(peer, (16,), torch.int32)
print(rank, "reads from", peer, remote[:4].tolist(), flush=True)
remote[8:12] = rank + 500  # write into the peer's buffer
hdl.barrier()
torch.cuda.synchronize()
print(rank, "own buffer after the peer wrote", t[8:12].tolist(), flush=True)
dist.destroy_process_group()
