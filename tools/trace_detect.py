import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch, importlib
import einx
synth = importlib.import_module("ei-nexus_official_b200.synth")
det = importlib.import_module("ei-nexus_official_b200.detection")
rng = np.random.default_rng(0)
for (B, Hp, Wp, K) in [(64, 184, 240, 1024), (32, 260, 346, 2048)]:
    s = torch.from_numpy(synth.score_map(rng, B, Hp, Wp)).cuda()
    for i in range(2):
        det.detect(s.clone(), 1.0, 4, 4, K, kcap=K)
    torch.cuda.synchronize()
