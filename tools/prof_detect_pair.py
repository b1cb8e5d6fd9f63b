"""Developer aid: a few detect_pair launches at a bench shape, for ncu (`-k regex:nms_kernel`)."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import einx

synth = importlib.import_module("ei-nexus_official_b200.synth")
det = importlib.import_module("ei-nexus_official_b200.detection")
B, Hp, Wp, K = (int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (64, 184, 240, 1024)))
rng = np.random.default_rng(0)
s = torch.from_numpy(synth.score_map(rng, B, Hp, Wp)).cuda()
for _ in range(4):
    det.detect_pair(s.clone(), s.clone(), 1.0, 4, 4, K, kcap=K)
torch.cuda.synchronize()
