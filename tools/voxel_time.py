"""Developer aid: time of the whole einx_voxelize entry (memset + scatter + normalise) per config, CUDA events."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import einx
synth = importlib.import_module("ei-nexus_official_b200.synth")
DEV = torch.device("cuda", 0)
for name, B in (("c2_ec_superpoint", 64), ("c3_mvsec_silk_b256", 32)):
    c = synth.CONFIGS[name]
    rng = np.random.default_rng(0)
    sets = []
    for s in range(3):
        evs = [synth.events(rng, c["events"], c["H"], c["W"], c["style"], c["dt"]) for _ in range(B)]
        sets.append(tuple(t.to(DEV) for t in einx.pack_events(evs)))
    f = lambda i: einx.voxelize_device(*sets[i % 3], (c["bins"], c["H"], c["W"]), True)
    for i in range(5): f(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30): f(i)
    e1.record(); torch.cuda.synchronize()
    print(f"{name} B={B}: voxelize entry {e0.elapsed_time(e1) / 30 * 1e3:.1f} us", flush=True)
