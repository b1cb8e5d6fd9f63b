"""Host-side cost of one step: a tiny workload whose GPU time is negligible (development aid)."""
import dataclasses, importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import einx
synth = importlib.import_module("ei-nexus_official_b200.synth")
DEV = torch.device("cuda", 0)
rng = np.random.default_rng(0)
H, W, B, K, D = 64, 64, 2, 64, 64
evs = [synth.events(rng, 2000, H, W, "ec") for _ in range(B)]
ev = tuple(t.to(DEV) for t in einx.pack_events(evs))
sc = [torch.from_numpy(synth.score_map(rng, B, H, W)).to(DEV) for _ in range(2)]
rw = [torch.from_numpy(synth.descriptor_map(rng, B, D, H // 8, W // 8)).to(DEV) for _ in range(2)]
cfg = einx.PathConfig(bins=5, height=H, width=W, top_k=K, descriptor_mode="bilinear", precision="tf32x3")
for conc in (False, True):
    pipe = einx.ExtractMatchPipeline(dataclasses.replace(cfg, concurrent=conc))
    for _ in range(20): pipe(ev, sc[0], rw[0], sc[1], rw[1])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(300): pipe(ev, sc[0], rw[0], sc[1], rw[1])
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"concurrent={conc}: issue {1e6 * (t1 - t0) / 300:.0f} us/step, with drain {1e6 * (t2 - t0) / 300:.0f} us/step", flush=True)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(300): pipe(ev, sc[0], rw[0], sc[1], rw[1])
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
