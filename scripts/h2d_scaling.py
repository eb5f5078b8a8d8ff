"""Pure host->device copy bandwidth per rank with N ranks copying at once (pinned 2 GiB buffers, no kernels): separates the
platform's host-feed limit from anything the library does.  Launch under torchrun with N = 1, 2, 4, 8."""
import os, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 2 << 30
host = torch.empty(n, dtype=torch.uint8).pin_memory()
host.fill_(rank + 1)
d = torch.empty(n, dtype=torch.uint8, device=dev)
d.copy_(host, non_blocking=True); torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(8):
    d.copy_(host, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
gbs = torch.tensor([8 * n / dt / 1e9], device=dev, dtype=torch.float64)
if world > 1:
    all_g = [torch.zeros_like(gbs) for _ in range(world)]
    dist.all_gather(all_g, gbs)
    vals = [float(x.item()) for x in all_g]
else:
    vals = [float(gbs.item())]
if rank == 0:
    print(f"N={world}: per-rank H2D GB/s {['%.1f' % v for v in vals]}  aggregate {sum(vals):.1f} GB/s", flush=True)
if world > 1:
    dist.destroy_process_group()
