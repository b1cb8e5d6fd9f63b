"""Developer aid: sampler launch time (library events) at the C2 shape, with and without the fp16 operand output."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import einx
synth = importlib.import_module("ei-nexus_official_b200.synth")
det = importlib.import_module("ei-nexus_official_b200.detection")
desc = importlib.import_module("ei-nexus_official_b200.describe")
DEV = torch.device("cuda", 0)
rng = np.random.default_rng(0)
B, Hp, Wp, K, D = 64, 184, 240, 1024, 256
sc = torch.from_numpy(synth.score_map(rng, B, Hp, Wp)).to(DEV)
raws = [torch.randn((B, D, Hp // 8, Wp // 8), device=DEV) for _ in range(3)]
_, kp, cn = det.detect(sc, 1.0, 4, 4, K, kcap=K)
ctx = einx.context_for(DEV)
for split in (False, True):
    for i in range(3): desc.sample(raws[i], kp, cn, desc.BILINEAR, (Hp, Wp), 1.0, True, split=split)
    ctx.profile(True); ts = []
    for i in range(9):
        desc.sample(raws[i % 3], kp, cn, desc.BILINEAR, (Hp, Wp), 1.0, True, split=split); ts.append(ctx.profile_read()[2])
    ctx.profile(False)
    print(f"sample C2 split={split}: {np.median(ts) * 1e3:.1f} us (min {min(ts) * 1e3:.1f})", flush=True)
