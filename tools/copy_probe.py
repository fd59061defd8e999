"""What does the box's host<->device path deliver per rank when N ranks copy at once?  Regular pinned vs write-combined
pinned input buffers; H2D alone, D2H alone, both directions.  torchrun --nproc-per-node N tools/copy_probe.py"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gseg = importlib.import_module("graph-algorithm-image-segmentation-gpgpu_b200")
NB = 1 << 30
dev_in = torch.empty(NB, dtype=torch.uint8, device="cuda")
dev_out = torch.empty(NB // 3, dtype=torch.uint8, device="cuda")
hosts = {"pinned": gseg.HostBuffer((NB,)), "write-combined": gseg.HostBuffer((NB,), write_combined=True)}
hout = gseg.HostBuffer((NB // 3,))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
CH = 6220800  # one 1080p image
def run(kind, h2d, d2h):
    src = torch.from_numpy(hosts[kind].array)
    dst = torch.from_numpy(hout.array)
    def once():
        if h2d:
            with torch.cuda.stream(s1):
                for o in range(0, NB - CH, CH):
                    dev_in[o:o + CH].copy_(src[o:o + CH], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                for o in range(0, NB // 3 - CH // 3, CH // 3):
                    dst[o:o + CH // 3].copy_(dev_out[o:o + CH // 3], non_blocking=True)
    once(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
for kind in ("pinned", "write-combined"):
    for name, h2d, d2h in (("H2D only", 1, 0), ("D2H only", 0, 1), ("both", 1, 1)):
        dt = run(kind, h2d, d2h)
        gb = (NB * h2d + NB // 3 * d2h) / 1e9
        if rank == 0:
            print("%d ranks, %-15s %-9s: %.1f GB/s per rank, %.1f GB/s aggregate" % (world, kind, name, gb / dt, world * gb / dt), flush=True)
if world > 1:
    dist.destroy_process_group()
