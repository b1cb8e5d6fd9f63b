"""Seeded synthetic inputs of the shapes BASELINE.json names (SURVEY.md section 8 d).

``seed = 1234 + 1000 * config + sample_index`` with numpy ``default_rng``; no dataset or checkpoint
is needed (there is no network on the build or GPU boxes).
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

# name -> sensor (H, W), events per window, bins, window seconds, descriptor type, top-k, D, scale
CONFIGS = {
    # configs[0]: MVSEC 346x260, ~200k events, 5 bins, SiLK-MNN (full-res gather, D=128, scale 1.41)
    "c1_mvsec_silk": dict(idx=1, H=260, W=346, events=200_000, bins=5, dt=0.4, style="mvsec", kind="gather",
                          cell=1, top_k=2048, D=128, scale=1.41),
    # configs[1]: EC 240x180, SuperPoint-MNN, 1024 keypoints, 256-d, batch 64 on one GPU
    "c2_ec_superpoint": dict(idx=2, H=180, W=240, events=60_000, bins=5, dt=0.04, style="ec", kind="bilinear",
                             cell=8, top_k=1024, D=256, scale=1.0),
    # configs[2]: MVSEC SiLK-MNN, 2048 keypoints, 128-d, batch 256 over 8 GPUs
    "c3_mvsec_silk_b256": dict(idx=3, H=260, W=346, events=200_000, bins=5, dt=0.4, style="mvsec", kind="gather",
                               cell=1, top_k=2048, D=128, scale=1.41),
    # configs[3]: 1280x720, ~5M events, 10 bins, 8192-keypoint MNN
    "c4_hires": dict(idx=4, H=720, W=1280, events=5_000_000, bins=10, dt=0.4, style="mvsec", kind="gather",
                     cell=1, top_k=8192, D=128, scale=1.41),
}


def seed_for(config_idx: int, sample: int) -> int:
    return 1234 + 1000 * config_idx + sample


def padded_size(h: int, w: int, cell: int) -> Tuple[int, int, Tuple[int, int, int, int]]:
    """Padder of core/modules/utils/util.py:9-15 -> (Hp, Wp, (w0, w1, h0, h1))."""
    hp = (((h // cell) + 1) * cell - h) % cell
    wp = (((w // cell) + 1) * cell - w) % cell
    pad = (wp // 2, wp - wp // 2, hp // 2, hp - hp // 2)
    return h + hp, w + wp, pad


def events(rng, n: int, H: int, W: int, style: str = "mvsec", dt: float = 0.4, T0: float = 1.5e9,
           clustered: bool = False) -> Dict[str, np.ndarray]:
    """fp64 event arrays: sub-pixel (MVSEC-rectified style) or integer-pixel, 0/1 polarity (EC style);
    epoch-scale sorted timestamps (exposes fp32 cancellation if t were cast too early)."""
    if clustered:  # 80 % of the events on 50 Gaussian blobs (atomic-contention stress)
        nb = int(0.8 * n)
        c = rng.integers(0, 50, nb)
        cx, cy = rng.uniform(8, W - 9, 50), rng.uniform(8, H - 9, 50)
        x = np.concatenate([np.clip(cx[c] + 3 * rng.standard_normal(nb), 0, W - 1.001), rng.uniform(0, W - 1, n - nb)])
        y = np.concatenate([np.clip(cy[c] + 3 * rng.standard_normal(nb), 0, H - 1.001), rng.uniform(0, H - 1, n - nb)])
        perm = rng.permutation(n)
        x, y = x[perm], y[perm]
    else:
        x = rng.uniform(0, W - 1, n)
        y = rng.uniform(0, H - 1, n)
    if style == "ec":
        x, y = np.floor(x), np.floor(y)
        p = rng.integers(0, 2, n).astype(np.float64)
    else:
        p = rng.integers(0, 2, n).astype(np.float64) * 2 - 1
    t = np.sort(rng.uniform(T0, T0 + dt, n))
    return {"x": x, "y": y, "t": t, "p": p}


def score_map(rng, B: int, Hp: int, Wp: int, kind: str = "uniform") -> np.ndarray:
    """(B, 1, Hp, Wp) fp32 in [0, 1]: i.i.d. uniform (worst-case survivor count) or 1/8-quantised (tie stress)."""
    v = rng.random((B, 1, Hp, Wp), dtype=np.float32)
    if kind == "ties":
        v = (np.round(v * 8) / 8).astype(np.float32)
    return v


def descriptor_map(rng, B: int, C: int, Hd: int, Wd: int) -> np.ndarray:
    return rng.standard_normal((B, C, Hd, Wd), dtype=np.float32)


def descriptor_pair(rng, n: int, m: int, d: int, scale: float, planted: float = 0.5, dups: int = 0):
    """L2-normalised rows x scale; `planted` of side-1 rows are noisy copies of side-0 rows."""
    a = rng.standard_normal((n, d))
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b = rng.standard_normal((m, d))
    k = int(min(n, m) * planted)
    if k:
        perm = rng.permutation(m)[:k]
        b[perm] = a[:k] + 0.2 * rng.standard_normal((k, d))
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    a, b = (scale * a).astype(np.float32), (scale * b).astype(np.float32)
    for q in range(dups):  # exact duplicate rows: first-index tie-breaking
        b[(7 * q + 3) % m] = b[(11 * q + 5) % m]
        a[(5 * q + 2) % n] = a[(13 * q + 7) % n]
    return a, b


def pair_inputs(cfg_name: str, sample: int, n_events: int | None = None):
    """Everything one pair needs: events + (score, raw) maps for the two sides, as numpy arrays."""
    c = CONFIGS[cfg_name]
    rng = np.random.default_rng(seed_for(c["idx"], sample))
    Hp, Wp, _ = padded_size(c["H"], c["W"], c["cell"])
    ev = events(rng, n_events or c["events"], c["H"], c["W"], c["style"], c["dt"])
    Hd, Wd = (Hp // c["cell"], Wp // c["cell"])
    sides = []
    for _ in range(2):
        sides.append((score_map(rng, 1, Hp, Wp), descriptor_map(rng, 1, c["D"], Hd, Wd)))
    return ev, sides
