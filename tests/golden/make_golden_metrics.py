"""Golden fixtures for the metric-side pairwise reduction (SURVEY.md section 8 f, row 4), from the REAL reference:
Repeatability.update_one of core/metrics/keypoints_metrics.py:57-128 (keep_true_points / warp_points of
core/metrics/util.py, the N x M distance matrix of :110-113 and its two min reductions).

    python tests/golden/make_golden_metrics.py        # writes tests/golden/metrics.npz
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import REF, OUT, load_reference  # noqa: E402
from make_golden_next import load  # noqa: E402


def main():
    torch.set_num_threads(1)
    load_reference()
    import types

    pkg = types.ModuleType("core.metrics")
    pkg.__path__ = [f"{REF}/core/metrics"]
    sys.modules["core.metrics"] = pkg
    mu = load("core.metrics.util", f"{REF}/core/metrics/util.py")
    km = load("core.metrics.keypoints_metrics", f"{REF}/core/metrics/keypoints_metrics.py")
    rng = np.random.default_rng(20241020)
    g = {}
    # (N, M, H, W, ordering, threshold); the last two cases: no point survives on one side / on both sides
    cases = [(300, 280, 180, 240, "yx", 3), (1024, 1000, 260, 346, "xy", 3), (64, 50, 60, 80, "yx", 1), (5, 0, 60, 80, "xy", 3),
             (0, 0, 60, 80, "yx", 3)]
    for ci, (N, M, H, W, ordering, thr) in enumerate(cases):
        a = 0.05 * rng.standard_normal()
        hom = np.array([[np.cos(a), -np.sin(a), 4.0 * rng.standard_normal()], [np.sin(a), np.cos(a), 3.0 * rng.standard_normal()],
                        [1e-4 * rng.standard_normal(), 1e-4 * rng.standard_normal(), 1.0]], dtype=np.float32)
        p1 = np.stack([rng.uniform(0, W, N), rng.uniform(0, H, N), rng.random(N)], 1).astype(np.float32)  # (x, y, prob)
        # half of side 2: side 1 warped by the homography plus up to two pixels of noise (true repeats)
        k = min(N, M) // 2
        w = (hom.astype(np.float64) @ np.concatenate([p1[:k, :2].T.astype(np.float64), np.ones((1, k))]))
        rep = (w[:2] / w[2]).T + rng.uniform(-2, 2, (k, 2))
        p2 = np.concatenate([np.concatenate([rep, rng.random((k, 1))], 1),
                             np.stack([rng.uniform(0, W, M - k), rng.uniform(0, H, M - k), rng.random(M - k)], 1)]).astype(np.float32)
        if ordering == "yx":
            p1, p2 = p1[:, [1, 0, 2]], p2[:, [1, 0, 2]]
        metric = km.Repeatability("repeatability", distance_thresh=thr, ordering=ordering)
        metric.device = torch.device("cpu")
        out = metric.update_one(torch.from_numpy(p1), torch.from_numpy(p2), (H, W), (H, W), torch.from_numpy(hom))
        g[f"c{ci}_p1"], g[f"c{ci}_p2"], g[f"c{ci}_hom"] = p1, p2, hom
        g[f"c{ci}_cfg"] = np.array([H, W, thr, 1 if ordering == "xy" else 0])
        g[f"c{ci}_value"] = np.array(out.get("repeatability", np.nan), dtype=np.float64)
        # the two min reductions themselves, through the reference's own helpers (same lines as update_one)
        sel = [0, 1] if ordering == "xy" else [1, 0]
        q1, q2 = torch.from_numpy(p1).T[sel], torch.from_numpy(p2).T[sel]
        h = torch.from_numpy(hom).float()
        q2, _ = mu.keep_true_points(q2, torch.linalg.inv(h), (H, W))
        q1, _ = mu.keep_true_points(q1, h, (H, W))
        wp = mu.warp_points(q1, h).T
        q2 = q2.T
        norm = torch.linalg.norm(wp[:, :2].unsqueeze(1) - q2[:, :2].unsqueeze(0), dim=2)
        g[f"c{ci}_min_over_1"] = torch.min(norm, 0).values.numpy() if wp.shape[0] else np.zeros(0, np.float32)
        g[f"c{ci}_min_over_2"] = torch.min(norm, 1).values.numpy() if q2.shape[0] else np.zeros(0, np.float32)
    g["ncases"] = np.array(len(cases))
    np.savez_compressed(f"{OUT}/metrics.npz", **g)
    print("metrics", os.path.getsize(f"{OUT}/metrics.npz") // 1024, "KiB")


if __name__ == "__main__":
    main()
