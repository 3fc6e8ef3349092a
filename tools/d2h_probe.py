"""Aggregate device->host bandwidth with one process per GPU (pinned host buffers): the ceiling of the e2e number.
torchrun --nproc-per-node N tools/d2h_probe.py"""
import os, time, torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 1 << 30
src = torch.empty(n, dtype=torch.uint8, device="cuda")
dst = torch.empty(n, dtype=torch.uint8, pin_memory=True)
for _ in range(2): dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
if world > 1: dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(8): dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"rank {rank}: D2H {8 * n / dt / 1e9:.1f} GB/s with {world} rank(s) copying at once", flush=True)
if world > 1: dist.destroy_process_group()
