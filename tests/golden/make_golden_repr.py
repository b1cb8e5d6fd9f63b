"""Golden fixtures for the reference's other scatter representations (SURVEY.md section 8 f, row 4):
events_to_event_stack and events_to_time_surface of datasets/representations.py, run in the build container.

    python tests/golden/make_golden_repr.py        # writes tests/golden/repr.npz
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, load_reference, synth_events  # noqa: E402


def main():
    torch.set_num_threads(1)
    _, _, _, _, rep = load_reference()
    rng = np.random.default_rng(20241019)
    g = {}
    cases = [("ec", 4000, 6, 30, 40), ("ec", 600, 4, 12, 16), ("mvsec01", 3000, 10, 36, 52), ("ec", 50, 2, 8, 8)]
    # (the distance-map fixture of case 3 has sparse bins: distances of tens of pixels)
    for ci, (style, n, bins, H, W) in enumerate(cases):
        ev = synth_events(rng, n, H, W, "ec" if style == "ec" else "mvsec", dt=0.04)
        if style == "mvsec01":
            ev["p"] = (ev["p"] > 0).astype(np.float64)  # sub-pixel coordinates with 0/1 polarity
        if ci == 1:  # events exactly on bin boundaries: they belong to two bins
            t = ev["t"]
            span = t[-1] - t[0] + 1e-8
            for k, frac in ((100, 0.25), (300, 0.5), (450, 0.75)):
                t[k] = t[0] + frac * span
            ev["t"] = np.sort(t)
        for k in "xytp":
            g[f"c{ci}_{k}"] = ev[k]
        g[f"c{ci}_shape"] = np.array([bins, H, W])
        g[f"c{ci}_stack"] = rep.events_to_event_stack({k: v.copy() for k, v in ev.items()}, (bins, H, W)).numpy()
        g[f"c{ci}_surface"] = rep.events_to_time_surface({k: v.copy() for k, v in ev.items()}, (bins, H, W)).numpy()
        # events_to_distance_map (:215-248; cv.distanceTransform of this container's opencv, an IPP build)
        g[f"c{ci}_distance"] = rep.events_to_distance_map({k: v.copy() for k, v in ev.items()}, (bins, H, W)).numpy()
    # distance map only: a sparse window -- distances of tens of pixels and bins without any event (FLT_MAX everywhere)
    ev = synth_events(rng, 24, 40, 56, "ec", dt=0.04)
    ev["t"][8:] += 0.02  # a gap in time: the middle bins stay empty
    for k in "xytp":
        g[f"sparse_{k}"] = ev[k]
    g["sparse_shape"] = np.array([8, 40, 56])
    g["sparse_distance"] = rep.events_to_distance_map({k: v.copy() for k, v in ev.items()}, (8, 40, 56)).numpy()
    g["ncases"] = np.array(len(cases))
    np.savez_compressed(f"{OUT}/repr.npz", **g)
    print("repr", os.path.getsize(f"{OUT}/repr.npz") // 1024, "KiB")


if __name__ == "__main__":
    main()
