"""Metric-side N x M reductions (SURVEY.md section 8 f, row 4): the pairwise distance matrices of the reference's
evaluation / ground-truth code, reduced along both axes on the device without ever being stored.

* ``Repeatability`` mirrors ``core/metrics/keypoints_metrics.py:52-128`` (same constructor, ``update_one`` returns
  the same dict); the warp / in-bounds filtering of ``core/metrics/util.py`` is a handful of torch ops on the
  device, the N x M part is ``einx_pairwise_min_dist``.
* ``gt_assign`` is the block of ``gt_matches_from_pose_depth`` between ``project`` and the epipolar pass
  (``core/geometry/gt_generation.py:96-126``): ``einx_gt_assign``.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _lib


def _need_cuda(t: torch.Tensor, who: str):
    if not t.is_cuda:
        raise _lib.EinxError(f"{who}: expected CUDA tensors (there is no CPU fallback)")


@torch.no_grad()
def pairwise_min_dist(a: torch.Tensor, b: torch.Tensor, na: Optional[torch.Tensor] = None,
                      nb: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(B, N, 2), (B, M, 2) fp32 -> min_j ||a_i - b_j|| (B, N) and min_i ||a_i - b_j|| (B, M)."""
    _need_cuda(a, "pairwise_min_dist")
    a = a.float().contiguous()
    b = b.float().contiguous()
    B, N, M = a.shape[0], a.shape[1], b.shape[1]
    if a.shape[-1] != 2 or b.shape[-1] != 2 or b.shape[0] != B:
        raise ValueError("pairwise_min_dist: expected (B, N, 2) and (B, M, 2)")
    ctx = _lib.context_for(a.device)
    rowmin = torch.empty((B, N), dtype=torch.float32, device=a.device)
    colmin = torch.empty((B, M), dtype=torch.float32, device=a.device)
    for c in (na, nb):
        if c is not None and (c.dtype != torch.int32 or not c.is_cuda):
            raise ValueError("pairwise_min_dist: counts must be int32 CUDA tensors")
    rc = ctx.lib.einx_pairwise_min_dist(ctx.handle, _lib.ptr(a), _lib.ptr(b), _lib.ptr(na) if na is not None else None,
                                        _lib.ptr(nb) if nb is not None else None, B, N, M, _lib.ptr(rowmin), _lib.ptr(colmin),
                                        ctx.stream)
    ctx.check(rc, "einx_pairwise_min_dist")
    return rowmin, colmin


def _warp(points: torch.Tensor, hom: torch.Tensor) -> torch.Tensor:
    """core/metrics/util.py:5-39 for (2, n) points: homogeneous product, then the division."""
    p = torch.vstack((points[:2], torch.ones(1, points.shape[1], device=points.device)))
    q = torch.mm(hom, p)
    return torch.vstack((q[0] / q[2], q[1] / q[2]))


def _keep(points: torch.Tensor, hom: torch.Tensor, shape) -> torch.Tensor:
    """core/metrics/util.py:72-104: the points whose warp lands inside (H, W)."""
    w = _warp(points, hom)
    mask = (w[0] >= 0) & (w[0] < shape[1]) & (w[1] >= 0) & (w[1] < shape[0])
    return points[:, mask]


class Repeatability:
    """Drop-in for ``core/metrics/keypoints_metrics.py:52-128``."""

    def __init__(self, name, distance_thresh=3, ordering="xy", device="cuda") -> None:
        self.distance_thresh = distance_thresh
        self.metric_name = name
        self.ordering = ordering
        assert self.ordering in ["xy", "yx"]
        self.device = torch.device(device)

    @torch.no_grad()
    def update_one(self, points1, points2, img1_shape, img2_shape, homography) -> Dict:
        out_dict = {}
        points1 = points1.to(self.device).float()
        points2 = points2.to(self.device).float()
        assert homography.shape == (3, 3)
        sel = [0, 1] if self.ordering == "xy" else [1, 0]
        points1 = points1.T[sel]
        points2 = points2.T[sel]
        homography = homography.to(self.device).float()
        points2 = _keep(points2, torch.linalg.inv(homography), img1_shape)
        points1 = _keep(points1, homography, img2_shape)
        warped = _warp(points1, homography).T.contiguous()
        points2 = points2.T.contiguous()
        original_num, warped_num = warped.shape[0], points2.shape[0]
        # min2 = torch.min(norm, 1) per warped side-1 point, min1 = torch.min(norm, 0) per side-2 point (:117-121)
        min2, min1 = pairwise_min_dist(warped[None], points2[None])
        count1 = count2 = 0
        if original_num != 0:
            count1 = torch.sum(min1[0] <= self.distance_thresh)
        if warped_num != 0:
            count2 = torch.sum(min2[0] <= self.distance_thresh)
        if original_num + warped_num > 0:
            out_dict[self.metric_name] = float(count1 + count2) / (original_num + warped_num)
        return out_dict

    @torch.no_grad()
    def update_batch(self, points1, points2, img1_shape, img2_shape, homography) -> Dict:
        """keypoints_metrics.py:130-157: mean of the per-sample values that exist."""
        assert len(points1) == len(points2) == len(homography)
        values = []
        for i in range(len(points1)):
            one = self.update_one(points1[i], points2[i], img1_shape, img2_shape, homography[i])
            if self.metric_name in one:
                values.append(one[self.metric_name])
        return {self.metric_name: torch.tensor(values).mean().item()}

    @torch.no_grad()
    def min_distances(self, warped_points1: torch.Tensor, points2: torch.Tensor):
        """The two reductions themselves: (min over side 1 per side-2 point, min over side 2 per side-1 point)."""
        min2, min1 = pairwise_min_dist(warped_points1[None].to(self.device), points2[None].to(self.device))
        return min1[0], min2[0]


IGNORE_FEATURE = -2
UNMATCHED_FEATURE = -1


@torch.no_grad()
def gt_assign(kp0, kp1, kp0_1, kp1_0, visible0, visible1, valid0, valid1, pos_th=3, neg_th=5, dense: bool = True):
    """``core/geometry/gt_generation.py:96-126``: returns (assignment | None, m0, m1).

    ``kp0`` (B, N, 2) / ``kp1`` (B, M, 2) are the keypoints in the order the reference indexes them after its
    ``ordering`` flip, ``kp0_1`` / ``kp1_0`` their reprojections (``project``), ``visible*`` / ``valid*`` its masks.
    ``assignment`` is the (B, N, M) bool matrix of positives (scattered from the argmins; ``dense=False`` skips it)."""
    _need_cuda(kp0, "gt_assign")
    dev = kp0.device
    B, N, M = kp0.shape[0], kp0.shape[1], kp1.shape[1]
    if N == 0 or M == 0:  # :63-71
        assignment = torch.zeros(B, N, M, dtype=torch.bool, device=dev)
        return (assignment if dense else None, -torch.ones((B, N), dtype=torch.long, device=dev),
                -torch.ones((B, M), dtype=torch.long, device=dev))
    f = lambda t: t.to(dev).float().contiguous()
    u = lambda t: t.to(dev).to(torch.uint8).contiguous()
    kp0, kp1, kp0_1, kp1_0 = f(kp0), f(kp1), f(kp0_1), f(kp1_0)
    visible0, visible1, valid0, valid1 = u(visible0), u(visible1), u(valid0), u(valid1)
    m0 = torch.empty((B, N), dtype=torch.int64, device=dev)
    m1 = torch.empty((B, M), dtype=torch.int64, device=dev)
    min0 = torch.empty((B, N), dtype=torch.int32, device=dev)
    min1 = torch.empty((B, M), dtype=torch.int32, device=dev)
    ctx = _lib.context_for(dev)
    rc = ctx.lib.einx_gt_assign(ctx.handle, _lib.ptr(kp0), _lib.ptr(kp1), _lib.ptr(kp0_1), _lib.ptr(kp1_0), _lib.ptr(visible0),
                                _lib.ptr(visible1), _lib.ptr(valid0), _lib.ptr(valid1), B, N, M, float(pos_th), float(neg_th),
                                _lib.ptr(m0), _lib.ptr(m1), _lib.ptr(min0), _lib.ptr(min1), ctx.stream)
    ctx.check(rc, "einx_gt_assign")
    assignment = None
    if dense:
        # positive[b, i, j] <=> j == min0[i], i == min1[j], dist < pos_th^2: row i has at most one, at column min0[i];
        # it is set where the pre-negative value of m0 was a match, i.e. where column min0[i] points back with m1 >= 0
        # or was only overridden by a negative -- recomputed here from the argmins, not from m0
        j = min0.long()
        mutual = torch.gather(min1.long(), 1, j) == torch.arange(N, device=dev)[None]
        close = _gt_close(kp0, kp1, kp0_1, kp1_0, j, pos_th) & torch.gather(visible1.bool(), 1, j) & visible0.bool()
        assignment = torch.zeros((B, N, M), dtype=torch.bool, device=dev)
        assignment.scatter_(2, j[..., None], (mutual & close)[..., None])
    return assignment, m0, m1


def _gt_close(kp0, kp1, kp0_1, kp1_0, j, pos_th):
    """dist[b, i, j_i] < pos_th ** 2 for the one candidate column of every row (O(N) work)."""
    idx = j[..., None].expand(-1, -1, 2)
    d0 = ((kp0_1 - torch.gather(kp1, 1, idx)) ** 2).sum(-1)
    d1 = ((kp0 - torch.gather(kp1_0, 1, idx)) ** 2).sum(-1)
    return torch.max(d0, d1) < pos_th ** 2
