"""Developer aid: phase timeline of the detect kernel's CTA 0 (EINX_DETECT_TRACE=1) and launch times per shape."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import einx

synth = importlib.import_module("ei-nexus_official_b200.synth")
det = importlib.import_module("ei-nexus_official_b200.detection")
rng = np.random.default_rng(0)
ctx = einx.context_for("cuda:0")
for (B, Hp, Wp, K) in [(64, 184, 240, 1024), (32, 260, 346, 2048), (32, 264, 352, 2048), (1, 184, 240, 1024), (1, 720, 1280, 8192)]:
    s = torch.from_numpy(synth.score_map(rng, B, Hp, Wp)).cuda()
    copies = [s.clone() for _ in range(8)]
    for c in copies[:2]:
        det.detect(c, 1.0, 4, 4, K, kcap=K)
    torch.cuda.synchronize()
    if os.environ.get("EINX_DETECT_TRACE"):
        continue
    ctx.profile(True)
    ts = []
    for c in copies[2:]:
        det.detect(c, 1.0, 4, 4, K, kcap=K)
        ts.append(ctx.profile_read()[1])
    ctx.profile(False)
    print(f"detect B={B} {Hp}x{Wp} k={K}: kernel {np.median(ts) * 1e3:.1f} us (min {min(ts) * 1e3:.1f})", flush=True)
    if B > 1:
        ctx.profile(True)
        ts = []
        for i in range(3):
            det.detect_pair(copies[2 * i].copy_(s), copies[2 * i + 1].copy_(s), 1.0, 4, 4, K, kcap=K)
            ts.append(ctx.profile_read()[1])
        ctx.profile(False)
        print(f"detect_pair 2x{B} {Hp}x{Wp} k={K}: kernel {np.median(ts) * 1e3:.1f} us (min {min(ts) * 1e3:.1f})", flush=True)
