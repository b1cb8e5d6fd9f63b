#!/usr/bin/env python
"""Run one stage of a config a few times (for `ncu -k regex:<kernel> -c 1 python tools/prof_stage.py voxel`)."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import einx  # noqa: E402

synth = importlib.import_module("ei-nexus_official_b200.synth")
det, desc, mt = (importlib.import_module(f"ei-nexus_official_b200.{m}") for m in ("detection", "describe", "match"))
what = sys.argv[1]
config = sys.argv[2] if len(sys.argv) > 2 else "c2_ec_superpoint"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
prec = sys.argv[4] if len(sys.argv) > 4 else "tf32x3"
DEV = torch.device("cuda", 0)
c = synth.CONFIGS[config]
Hp, Wp, _ = synth.padded_size(c["H"], c["W"], c["cell"])
evs, s0, r0 = [], [], []
for i in range(B):
    ev, sides = synth.pair_inputs(config, i, None)
    evs.append(ev); s0.append(sides[0][0]); r0.append(sides[0][1])
cfg = einx.PathConfig(bins=c["bins"], height=c["H"], width=c["W"], top_k=c["top_k"], descriptor_mode=c["kind"],
                      descriptor_scale=c["scale"], precision=prec)
pipe = einx.ExtractMatchPipeline(cfg)
mode = desc.BILINEAR if cfg.descriptor_mode == "bilinear" else desc.GATHER
for rep in range(3):
    if what == "voxel":
        ev = tuple(t.to(DEV) for t in einx.pack_events(evs))
        pipe.voxelize(*ev)
    else:
        sc = torch.from_numpy(np.concatenate(s0)).to(DEV)
        rw = torch.from_numpy(np.concatenate(r0)).to(DEV)
        _, kp, cn = det.detect(sc, cfg.detection_threshold, cfg.nms_radius, cfg.remove_borders, cfg.top_k, kcap=cfg.top_k)
        if what in ("sample", "mnn"):
            d = desc.sample(rw, kp, cn, mode, (Hp, Wp), cfg.descriptor_scale, True)
        if what == "mnn":
            mt.mnn(d, d.flip(0).contiguous(), cn, cn, kp, kp, None, None, True, prec)
    torch.cuda.synchronize()
