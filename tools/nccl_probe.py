"""Timing of the collectives used by the multi-GPU dist step (run under torchrun)."""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
buf = torch.empty(82_000_000, dtype=torch.uint8, device=dev)
small = torch.empty(1 << 20, dtype=torch.uint8, device=dev)
recv = torch.empty(world << 20, dtype=torch.uint8, device=dev)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n
t_b = timeit(lambda: dist.broadcast(buf, src=0))
t_g = timeit(lambda: dist.all_gather_into_tensor(recv, small))
t_ar = timeit(lambda: dist.all_reduce(small[:8].view(torch.int64)))
plane = torch.empty(41_000_000 // world, dtype=torch.uint8, device=dev)
plane_all = torch.empty(plane.numel() * world, dtype=torch.uint8, device=dev)
t_ag = timeit(lambda: dist.all_gather_into_tensor(plane_all, plane))
if rank == 0:
    print("world %d: broadcast 82 MB %.3f ms (%.1f GB/s), all_gather 1 MiB/rank %.3f ms, all_reduce 8 B %.3f ms" %
          (world, t_b, 82e6 / t_b / 1e6, t_g, t_ar))
    print("all_gather of a 41 MB s8 plane (%.1f MB per rank): %.3f ms" % (plane.numel() / 1e6, t_ag))
dist.destroy_process_group()
