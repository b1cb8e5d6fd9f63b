"""Developer aid: a few detect_pair launches at the C4 shape (2 x 720x1280, top-8192), for ncu."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, einx
synth = importlib.import_module("ei-nexus_official_b200.synth")
det = importlib.import_module("ei-nexus_official_b200.detection")
rng = np.random.default_rng(0)
s = torch.from_numpy(synth.score_map(rng, 1, 720, 1280)).cuda()
for _ in range(4):
    det.detect_pair(s.clone(), s.clone(), 1.0, 4, 4, 8192, kcap=8192)
torch.cuda.synchronize()
