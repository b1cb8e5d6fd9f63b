"""Developer aid: kernel timeline (start offset, duration, stream) of one graph-replayed step, via torch.profiler (CUPTI).

    python tools/timeline.py [config] [batch]
"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

import einx

synth = importlib.import_module("ei-nexus_official_b200.synth")
bench = importlib.import_module("bench")
name = sys.argv[1] if len(sys.argv) > 1 else "c2_ec_superpoint"
B = int(sys.argv[2]) if len(sys.argv) > 2 else bench.DEFAULT_BATCH[name]
prec = sys.argv[3] if len(sys.argv) > 3 else "fp16x3"
c = synth.CONFIGS[name]
dev = torch.device("cuda", 0)
cfg = einx.PathConfig(bins=c["bins"], height=c["H"], width=c["W"], top_k=c["top_k"], descriptor_mode=c["kind"],
                      descriptor_scale=c["scale"], precision=prec)
pipe = einx.ExtractMatchPipeline(cfg)
sets = []
for s in range(3):
    evs, s0, r0, s1, r1 = bench.make_batch(synth, name, B, s * B)
    ev = tuple(t.to(dev) for t in einx.pack_events(evs))
    sets.append((ev, [torch.from_numpy(a).to(dev) for a in (s0, r0, s1, r1)]))
caps = [pipe.capture(ev, *m) for ev, m in sets]
for i in range(6):
    caps[i % 3].replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(6):
        caps[i % 3].replay()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# one step = from one nms_kernel start to the next (a step launches it once)
marks = [i for i, e in enumerate(evs) if "nms_kernel" in e.name or "detect" in e.name]
a, b = marks[3], marks[4]
# kernels that started before the marker but belong to the same step (voxel memset / scatter on the side stream)
while a > 0 and evs[a - 1].time_range.start > evs[marks[2]].time_range.start and "mnn" not in evs[a - 1].name and "memset" in evs[a - 1].name:
    a -= 1
t0 = evs[a].time_range.start
for e in evs[a:b + 3]:
    print(f"{e.time_range.start - t0:9.1f} us  +{e.time_range.end - e.time_range.start:8.1f} us  {e.name[:100]}")
print(f"step period: {evs[marks[4]].time_range.start - evs[marks[3]].time_range.start:.1f} us")
