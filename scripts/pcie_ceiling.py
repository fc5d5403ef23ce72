"""Concurrent pinned-copy ceiling of the box: every rank copies the e2e step's bytes (81.7 MB host -> device,
22.0 MB device -> host, full duplex on two streams) back to back; reports per-rank and aggregate GB/s and the
images/s that bandwidth would allow (32 images per step and GPU).  torchrun --nproc-per-node N scripts/pcie_ceiling.py"""
import os, time, json
import torch
import torch.distributed as dist
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
h2d_bytes, d2h_bytes = 81748992, 22009344
hin = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
hout = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
din = torch.empty(h2d_bytes, dtype=torch.uint8, device="cuda")
dout = torch.empty(d2h_bytes, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(n):
    for _ in range(n):
        with torch.cuda.stream(s1):
            din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(dout, non_blocking=True)
    torch.cuda.synchronize()
run(5)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
run(50)
dt = time.perf_counter() - t0
if world > 1:
    t = torch.tensor([dt], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
if rank == 0:
    step = dt / 50
    print(json.dumps({"n_gpus": world, "ms_per_step_copies_only": step * 1e3, "h2d_GBs_per_gpu": h2d_bytes / step / 1e9,
                      "d2h_GBs_per_gpu": d2h_bytes / step / 1e9, "aggregate_GBs": world * (h2d_bytes + d2h_bytes) / step / 1e9,
                      "images_per_s_ceiling": 32 * world / step}))
if world > 1:
    dist.destroy_process_group()
