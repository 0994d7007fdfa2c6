"""Probe: torch symmetric memory on this box (peer pointers, barrier, a kernel-free signal)."""
import os, time
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 16
buf = symm.empty(2 * world * n, dtype=torch.int64, device=f"cuda:{local}")
buf.zero_()
hdl = symm.rendezvous(buf, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], flush=True)
root = hdl.get_buffer(0, (2 * world * n,), torch.int64)
mine = root[rank * n:(rank + 1) * n]
src = torch.full((n,), rank + 1, dtype=torch.int64, device="cuda")
torch.cuda.synchronize(); dist.barrier()
mine.copy_(src)                       # peer write into rank 0's memory
hdl.barrier(channel=0)
torch.cuda.synchronize()
if rank == 0:
    print("rank0 sees", [int(buf[r * n].item()) for r in range(world)], flush=True)
# timing of barrier launches
t0 = time.perf_counter()
for i in range(1000):
    hdl.barrier(channel=i & 1)
torch.cuda.synchronize()
print(rank, "barrier us", (time.perf_counter() - t0) * 1e3, flush=True)
dist.destroy_process_group()
