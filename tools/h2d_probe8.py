"""Concurrent host->device bandwidth of all ranks from pinned memory, default CPU affinity vs the GPU's NVML-local CPUs
(bound BEFORE the pinned allocation, so first touch places the pages on the local NUMA node).  Development aid:

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_probe8.py [local]
"""
import os, subprocess, sys
import torch, torch.distributed as dist

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
mode = sys.argv[1] if len(sys.argv) > 1 else "default"
allowed = sorted(os.sched_getaffinity(0))
note = f"{len(allowed)} cpus allowed"
if mode == "local":
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(local)
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
    loc = [c for c in cpus if c in allowed]
    if loc:
        os.sched_setaffinity(0, loc)
    note = f"nvml-local cpus {loc[:4]}..{loc[-1:]} ({len(loc)})"
if rank == 0:
    print(subprocess.run("lscpu | grep -i -E 'numa|socket'; nvidia-smi topo -m | head -14", shell=True, capture_output=True, text=True).stdout, flush=True)
h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
h.fill_(1)
d = torch.empty_like(h, device=dev)
for _ in range(3):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    d.copy_(h, non_blocking=True)
e1.record()
torch.cuda.synchronize()
bw = 20 * h.numel() / e0.elapsed_time(e1) / 1e6
t = torch.tensor([bw], device=dev)
g = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(g, t)
if rank == 0:
    vals = [round(float(v), 1) for v in g]
    print(f"mode={mode}: per-rank H2D GB/s {vals}, aggregate {sum(vals):.0f} GB/s", flush=True)
print(f"[rank {rank}] {note}", flush=True)
dist.destroy_process_group()
