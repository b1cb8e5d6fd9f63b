"""The whole hot path for a batch of event-image pairs, device-resident end to end:

    voxelise(events) -> [detect -> sample] for both sides -> MNN

Nothing synchronises with the host between the stages: keypoints stay in padded (B, K, 3) buffers
with per-image counts, exactly the layout the next kernel consumes.  This is the unit
``bench.py`` times ("pairs/sec") and ``__graft_entry__.smoke()`` runs.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import describe, detection as _detect, match, voxel


@dataclass
class PathConfig:
    """Parameters pinned by the reference configs."""
    bins: int = 5
    height: int = 260          # sensor H (voxel grid)
    width: int = 346           # sensor W
    nms_radius: int = 4        # configs/model/SiLK_MNN.yaml:15
    remove_borders: int = 4    # :18
    detection_threshold: float = 1.0  # :16  -> min(topk_thr, 1.0) selects the top-k threshold
    top_k: int = 1024          # :17
    descriptor_mode: str = "bilinear"  # "bilinear" = SuperPoint type (cell 8), "gather" = SiLK type (cell 1)
    descriptor_scale: float = 1.0      # 1.0 (SP, D=256) / 1.41 (SiLK, D=128)
    precision: str = "fp32"    # MNN arithmetic: fp32 | tf32x3 | bf16
    normalize_voxels: bool = True


class ExtractMatchPipeline:
    """Batched drop-in for the post-backbone part of ``EIM.forward`` (core/modules/EIM.py:89-93)."""

    def __init__(self, cfg: PathConfig):
        self.cfg = cfg

    @torch.no_grad()
    def voxelize(self, x, y, t, p, offsets) -> torch.Tensor:
        c = self.cfg
        return voxel.voxelize_device(x, y, t, p, offsets, (c.bins, c.height, c.width), c.normalize_voxels)

    @torch.no_grad()
    def extract(self, score: torch.Tensor, raw: torch.Tensor, mask: Optional[torch.Tensor] = None):
        """(B,1,Hp,Wp) score + (B,C,Hd,Wd) raw descriptors -> padded keypoints, counts, descriptors."""
        c = self.cfg
        _, kpts, counts = _detect.detect(score, c.detection_threshold, c.nms_radius, c.remove_borders, c.top_k,
                                         mask=mask, want_map=False)
        mode = describe.BILINEAR if c.descriptor_mode == "bilinear" else describe.GATHER
        desc = describe.sample(raw, kpts, counts, mode, score.shape[-2:], c.descriptor_scale, True)
        return kpts, counts, desc

    @torch.no_grad()
    def __call__(self, events, score0, raw0, score1, raw1, mask0=None, mask1=None) -> Dict[str, torch.Tensor]:
        """events = (x, y, t, p, offsets) on the device; score/raw maps of both sides on the device."""
        grid = self.voxelize(*events)
        k0, c0, d0 = self.extract(score0, raw0, mask0)
        k1, c1, d1 = self.extract(score1, raw1, mask1)
        out = match.mnn(d0, d1, c0, c1, k0, k1, None, None, True, self.cfg.precision)
        out.update(voxel_grid=grid, keypoints0=k0, keypoints1=k1, counts0=c0, counts1=c1,
                   descriptors0=d0, descriptors1=d1)
        return out
