"""Host->device bandwidth from pinned memory vs CPU affinity / NUMA placement (development aid)."""
import os, subprocess, sys, time
import torch
print("affinity", sorted(os.sched_getaffinity(0)))
print(subprocess.run("lscpu | grep -i -E 'numa|socket|model name'; nvidia-smi topo -m | head -8; cat /sys/bus/pci/devices/$(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader | head -1 | cut -c5- | tr A-Z a-z)/numa_node 2>/dev/null", shell=True, capture_output=True, text=True).stdout)
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
    print("nvml cpu affinity of GPU 0:", cpus[:8], "...", len(cpus))
except Exception as e:
    cpus = []
    print("nvml affinity unavailable:", e)
dev = torch.device("cuda", 0)
def bw(tag):
    h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    d = torch.empty_like(h, device=dev)
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(f"{tag}: H2D {10 * h.numel() / e0.elapsed_time(e1) / 1e6:.1f} GB/s", flush=True)
bw("default affinity")
allowed = sorted(os.sched_getaffinity(0))
for part in (allowed[: len(allowed) // 2], allowed[len(allowed) // 2:]):
    os.sched_setaffinity(0, part)
    bw(f"cpus {part[0]}-{part[-1]}")
local = [c for c in cpus if c in allowed]
if local:
    os.sched_setaffinity(0, local)
    bw(f"nvml-local cpus ({len(local)})")
