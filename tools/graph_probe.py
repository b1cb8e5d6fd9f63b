"""Probe: does a CUDA-graph capture slow down later eager steps? (development aid)"""
import dataclasses, importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import einx
synth = importlib.import_module("ei-nexus_official_b200.synth")
DEV = torch.device("cuda", 0)
c = synth.CONFIGS["c2_ec_superpoint"]; B = 64
evs, s0, r0, s1, r1 = [], [], [], [], []
for i in range(B):
    ev, sides = synth.pair_inputs("c2_ec_superpoint", i, None)
    evs.append(ev); s0.append(sides[0][0]); r0.append(sides[0][1]); s1.append(sides[1][0]); r1.append(sides[1][1])
ev = tuple(t.to(DEV) for t in einx.pack_events(evs))
s0, r0, s1, r1 = (torch.from_numpy(np.concatenate(a)).to(DEV) for a in (s0, r0, s1, r1))
cfg = einx.PathConfig(bins=c["bins"], height=c["H"], width=c["W"], top_k=c["top_k"], descriptor_mode=c["kind"], descriptor_scale=c["scale"], precision="tf32x3")
def t(fn, iters=50):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, (time.perf_counter() - t0) / iters * 1e3
def stats():
    m = torch.cuda.memory_stats(DEV)
    return m["num_device_alloc"], m["num_alloc_retries"], m["reserved_bytes.all.current"] >> 20
conc = einx.ExtractMatchPipeline(cfg); ser = einx.ExtractMatchPipeline(dataclasses.replace(cfg, concurrent=False))
print("conc eager", t(lambda: conc(ev, s0, r0, s1, r1)), stats(), flush=True)
print("ser eager", t(lambda: ser(ev, s0, r0, s1, r1)), stats(), flush=True)
g = ser.capture(ev, s0, r0, s1, r1)
print("ser graph", t(g.replay), stats(), flush=True)
print("conc eager after capture", t(lambda: conc(ev, s0, r0, s1, r1)), stats(), flush=True)
print("ser eager after capture", t(lambda: ser(ev, s0, r0, s1, r1)), stats(), flush=True)
g2 = conc.capture(ev, s0, r0, s1, r1)
print("conc graph", t(g2.replay), stats(), flush=True)
print("conc eager after 2 captures", t(lambda: conc(ev, s0, r0, s1, r1)), stats(), flush=True)
conc2 = einx.ExtractMatchPipeline(cfg)
print("new conc eager", t(lambda: conc2(ev, s0, r0, s1, r1)), stats(), flush=True)
