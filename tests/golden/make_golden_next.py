"""Golden fixtures for the rows adjacent to the hot path (SURVEY.md section 8 f), from the REAL reference.

Runs only in the build container (reference mounted at /root/reference); writes tests/golden/next.npz.
Loaded by file path with stub parents, like make_golden.py; ``matplotlib`` and ``omegaconf`` (imported
at module level by datasets/visualize.py and core/modules/matchers/lightglue.py, not installed here,
not used by the functions called) are stubbed with empty modules.

    python tests/golden/make_golden_next.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import REF, OUT, load_reference, synth_events  # noqa: E402


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    torch.set_num_threads(1)
    det, desc, util, mnn, rep = load_reference()
    for name in ("matplotlib", "matplotlib.pyplot", "omegaconf"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["omegaconf"].OmegaConf = object
    vis = load("einx_ref_visualize", f"{REF}/datasets/visualize.py")
    lg = load("core.modules.matchers.lightglue", f"{REF}/core/modules/matchers/lightglue.py")
    rng = np.random.default_rng(20241018)
    g = {}

    # ---- event accumulation image + detector mask ---------------------------------------------- #
    cases = [("mvsec", 5000, 48, 64, 8), ("ec", 3000, 30, 40, 8), ("ec", 40, 12, 16, 1), ("mvsec", 2000, 36, 52, 8)]
    for ci, (style, n, H, W, cell) in enumerate(cases):
        ev = synth_events(rng, n, H, W, style)
        if ci == 1:  # a hot pixel: most other pixels fall below 1/255 of the maximum and leave the mask
            ev["x"][:1500] = 7.0
            ev["y"][:1500] = 5.0
        img = vis.draw_events_accumulation_image(ev, (W, H))
        g[f"img{ci}_x"], g[f"img{ci}_y"] = ev["x"], ev["y"]
        g[f"img{ci}_shape"] = np.array([H, W, cell])
        g[f"img{ci}_out"] = img
        # the extractor's mask: events_image > 0 (train_extractor.py:225), Padder.pad (bool -> constant),
        # then the 3x3 box filter and `> 0` of EventExtractors.py:357-363 (the method itself needs the
        # whole network, so the same torch ops are issued here on the reference's own Padder output)
        m = torch.from_numpy(img)[None, None] > 0
        padder = util.Padder(m.shape, cell)
        m = padder.pad(m)[0].float()
        k = torch.ones((1, 1, 3, 3)) / 9.0
        g[f"img{ci}_mask"] = (torch.nn.functional.conv2d(m, k, padding=1) > 0)[0, 0].numpy()
    g["img_ncases"] = np.array(len(cases))

    # ---- detector head: logits_to_prob + depth_to_space ---------------------------------------- #
    lo = (3.0 * rng.standard_normal((2, 65, 6, 9))).astype(np.float32)
    prob = det.logits_to_prob(torch.from_numpy(lo), channel_dim=1)
    g["head_logits65"], g["head_prob65"] = lo, prob.numpy()
    g["head_score65"] = det.depth_to_space(prob, cell_size=8).numpy()
    l1 = (4.0 * rng.standard_normal((2, 1, 20, 28))).astype(np.float32)
    p1 = det.logits_to_prob(torch.from_numpy(l1), channel_dim=1)
    g["head_logits1"], g["head_prob1"] = l1, p1.numpy()
    g["head_score1"] = det.depth_to_space(p1, cell_size=1).numpy()
    l17 = rng.standard_normal((1, 17, 5, 7)).astype(np.float32)
    p17 = det.logits_to_prob(torch.from_numpy(l17), channel_dim=1)
    g["head_logits17"], g["head_score17"] = l17, det.depth_to_space(p17, cell_size=4).numpy()

    # ---- LightGlue filter_matches --------------------------------------------------------------- #
    fcases = [(2, 70, 90, 0.1), (1, 128, 96, 0.0), (3, 33, 31, 0.2), (1, 1, 5, 0.0)]
    for ci, (B, M, N, th) in enumerate(fcases):
        sim = torch.from_numpy((4.0 * rng.standard_normal((B, M, N))).astype(np.float32))
        if ci == 1:  # ties: duplicated columns / rows must resolve to the first index
            sim[:, :, 7] = sim[:, :, 3]
            sim[:, 11, :] = sim[:, 2, :]
        z0 = torch.from_numpy(rng.standard_normal((B, M, 1)).astype(np.float32))
        z1 = torch.from_numpy(rng.standard_normal((B, N, 1)).astype(np.float32))
        scores = lg.sigmoid_log_double_softmax(sim, z0, z1)
        m0, m1, s0, s1 = lg.filter_matches(scores, th)
        g[f"fm{ci}_scores"], g[f"fm{ci}_th"] = scores.numpy(), np.array(th)
        g[f"fm{ci}_m0"], g[f"fm{ci}_m1"], g[f"fm{ci}_s0"], g[f"fm{ci}_s1"] = m0.numpy(), m1.numpy(), s0.numpy(), s1.numpy()
    g["fm_ncases"] = np.array(len(fcases))
    np.savez_compressed(f"{OUT}/next.npz", **g)
    print("next", os.path.getsize(f"{OUT}/next.npz") // 1024, "KiB")


if __name__ == "__main__":
    main()
